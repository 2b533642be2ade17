"""update() time (CUDA events, 4 repetitions after one warm-up) of the bench-sized storage for several work-item plans of the fused chain
kernel (number of one-tile items per program at the tail of the large launches (0 = two-tile items only, -1 = the planner's own choice for
the penalty given).  One process, one JSON line per setting.   python tools/chain_split_sweep.py [precision ...]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from dwbc_b200 import _lib as L
from dwbc_b200.actor_critic import FlatActorCritic
from dwbc_b200.ppo import FusedPPO
N, T = 4096, 40
lib = L.lib()
lib.dwbc_debug_set_chain_singles.argtypes = [C.c_int]
lib.dwbc_debug_set_chain_single_penalty.argtypes = [C.c_double]
lib.dwbc_debug_set_wgrad_snake.argtypes = [C.c_int]
lib.dwbc_debug_set_wgrad_items.argtypes = [C.c_int]
lib.dwbc_debug_set_wgrad_reverse.argtypes = [C.c_int]
lib.dwbc_debug_set_chain_bwd_reverse.argtypes = [C.c_int]
for prec in (sys.argv[1:] or ["tf32x3", "tf32"]):
    ac = FlatActorCritic(device="cuda:0", seed=0, init_std=[[0.8, 1.0, 1.0] * 4 + [1.0] * 6], num_priv=24, num_hist=10, num_prop=76)
    alg = FusedPPO(ac, device="cuda:0", precision=prec, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95,
                   learning_rate=2e-4, mixing_schedule=[1.0, 0, 1], priv_reg_coef_schedual=[0, 1, 1000, 1000])
    alg.init_storage(N, T, [860], [None], [18]); alg.counter = 1500
    s = alg.storage
    s._obs_all.normal_(); s.actions.normal_(); s.values.normal_(); s.returns.normal_(); s.advantages.normal_(); s.actions_log_prob.normal_().sub_(20)
    for singles, pen, snake, wgi, rev, brev in ((-1, 1.35, 0, 4, 0, 0), (-1, 1.35, 0, 4, 1, 0), (-1, 1.35, 0, 4, 0, 1), (-1, 1.35, 0, 4, 1, 1),
                                                (-1, 1.35, 0, 4, 0, 0), (-1, 1.35, 0, 4, 1, 0), (-1, 1.35, 0, 4, 0, 1), (-1, 1.35, 0, 4, 1, 1)):
        lib.dwbc_debug_set_wgrad_reverse(rev); lib.dwbc_debug_set_chain_bwd_reverse(brev)
        lib.dwbc_debug_set_chain_singles(singles); lib.dwbc_debug_set_chain_single_penalty(pen); lib.dwbc_debug_set_wgrad_snake(snake)
        lib.dwbc_debug_set_wgrad_items(wgi)
        alg.update(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            alg.update()
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"precision": prec, "singles_per_program": singles, "penalty": pen, "wgrad_snake": snake, "wgrad_items_per_cta": wgi, "wgrad_reverse": rev, "bwd_chain_reverse": brev, "update_ms": round(e0.elapsed_time(e1) / 4, 3)}), flush=True)
    lib.dwbc_debug_set_chain_singles(-1); lib.dwbc_debug_set_chain_single_penalty(1.35); lib.dwbc_debug_set_wgrad_snake(0); lib.dwbc_debug_set_wgrad_items(4); lib.dwbc_debug_set_wgrad_reverse(0); lib.dwbc_debug_set_chain_bwd_reverse(1)
