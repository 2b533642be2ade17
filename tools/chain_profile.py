"""clock64 stamps of the fused chain kernel (mlp_chain2.cuh), first work item of a CTA; per op n, relative to the CTA start:
[0] weights landed, [1] slot X ready (MMAs issued next), [2] slot Y ready, [3] all MMAs of the op retired, [4] a worker sees X's accumulator,
[5] the worker has finished Y's epilogue; then the CTA's end"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from dwbc_b200 import _lib as L
from dwbc_b200.actor_critic import FlatActorCritic
from dwbc_b200.ppo import FusedPPO
N, T = 4096, 40
ac = FlatActorCritic(device="cuda:0", seed=0, init_std=[[0.8, 1.0, 1.0] * 4 + [1.0] * 6], num_priv=24, num_hist=10, num_prop=76)
alg = FusedPPO(ac, device="cuda:0", precision=(sys.argv[2] if len(sys.argv) > 2 else "tf32"), num_learning_epochs=1, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, learning_rate=2e-4,
               mixing_schedule=[1.0, 0, 1], priv_reg_coef_schedual=[0, 1, 1000, 1000])
alg.init_storage(N, T, [860], [None], [18]); alg.counter = 1500
s = alg.storage
s._obs_all.normal_(); s.actions.normal_(); s.values.normal_(); s.returns.normal_(); s.advantages.normal_(); s.actions_log_prob.normal_().sub_(20)
alg.update(); torch.cuda.synchronize()
lib = L.lib()
buf = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
lib.dwbc_debug_set_tc_cycle_buffer.argtypes = [C.c_void_p]
lib.dwbc_debug_set_tc_cycle_buffer(buf.data_ptr())
mode = sys.argv[1] if len(sys.argv) > 1 else "update"
if mode == "update":
    alg.update()
else:
    big = s._obs_all.view(-1, 860)[:40960]
    alg.act(big, big, False)
    torch.cuda.synchronize(); buf.zero_()
    alg.act(big, big, False)
torch.cuda.synchronize()
lib.dwbc_debug_set_tc_cycle_buffer(None)
c = buf.view(148, 64).cpu()
for b in (0, 77):
    row = c[b]; t0 = int(row[62])
    print("CTA", b, "end", int(row[63]) - t0)
    for n in range(10):
        print("  op", n, [int(row[6 * n + k]) - t0 if int(row[6 * n + k]) > 0 else None for k in range(6)])
