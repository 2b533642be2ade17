"""clock64 stamps of the weight-gradient kernel (wgrad_group.cuh), second work item of a CTA, relative to the kernel start; per chunk c:
P = producer thread 0 has issued the copies of chunk c, S = (3xTF32) it has split chunk c, F = the MMA warp sees chunk c complete,
M = its MMAs are issued, B = the bias-gradient warp has summed chunk c."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from dwbc_b200 import _lib as L
from dwbc_b200.actor_critic import FlatActorCritic
from dwbc_b200.ppo import FusedPPO
N, T = 4096, 40
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
ac = FlatActorCritic(device="cuda:0", seed=0, init_std=[[0.8, 1.0, 1.0] * 4 + [1.0] * 6], num_priv=24, num_hist=10, num_prop=76)
alg = FusedPPO(ac, device="cuda:0", precision=prec, num_learning_epochs=1, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, learning_rate=2e-4,
               mixing_schedule=[1.0, 0, 1], priv_reg_coef_schedual=[0, 1, 1000, 1000])
alg.init_storage(N, T, [860], [None], [18]); alg.counter = 1500
s = alg.storage
s._obs_all.normal_(); s.actions.normal_(); s.values.normal_(); s.returns.normal_(); s.advantages.normal_(); s.actions_log_prob.normal_().sub_(20)
alg.update(); torch.cuda.synchronize()
lib = L.lib()
buf = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
lib.dwbc_debug_set_wg_cycle_buffer.argtypes = [C.c_void_p]
lib.dwbc_debug_set_wg_cycle_buffer(buf.data_ptr())
alg.update(); torch.cuda.synchronize()
lib.dwbc_debug_set_wg_cycle_buffer(None)
c = buf.view(148, 64).cpu()
for b in (0, 77, 147):
    row = c[b]; t0 = int(row[62])
    rel = lambda k: int(row[k]) - t0 if int(row[k]) > 0 else None
    print("CTA", b, "end", rel(63), "item1 start", rel(60))
    for n in range(12):
        print(f"  chunk {n:2d}  P {rel(n)}  S {rel(12 + n)}  F {rel(24 + n)}  M {rel(36 + n)}  B {rel(48 + n)}", flush=True)
