"""Per parameter tensor ||g - g_fp32|| / ||g_fp32|| of one mini-batch gradient: tensor-core modes against the exact-fp32 path of the library."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from dwbc_b200 import _lib as L
from test_gpu_ppo import make_alg, golden_params, G
g = np.load(os.path.join(G, "ppo.npz"))
P = golden_params(g, int(g["meta"][2]))
for N in [int(a) for a in sys.argv[2:]] or [8192, 148 * 128 * 2 + 333]:
    gen = torch.Generator(device="cuda").manual_seed(9)
    alg = make_alg(N, 1, P, num_mini_batches=1, num_learning_epochs=1); alg.counter = 1500
    s = alg.storage
    s._obs_all.normal_(generator=gen)
    for k in ("actions", "values", "returns", "advantages"):
        getattr(s, k).normal_(generator=gen)
    s.actions_log_prob.normal_(generator=gen).sub_(20.0)
    idx = torch.randperm(N, device="cuda", generator=gen)
    ac = alg.actor_critic
    grads = {}
    for prec in ("fp32", sys.argv[1] if len(sys.argv) > 1 else "tf32x3"):
        alg.precision = prec
        h = alg._fill_hp(); alg._losses.zero_()
        L.check(L.lib().dwbc_ppo_minibatch_grad(C.addressof(ac.net_cfg), L.ptr(ac.flat), s.c_struct_ptr(), L.ptr(idx), N, C.addressof(h),
                                                L.ptr(alg.grad), L.ptr(alg._losses), L.ptr(alg._workspace(N)), L.stream_ptr()), "grad")
        grads[prec] = {k: v.clone() for k, v in ac.unflat(alg.grad).items()}
    print("rows", N)
    a_, b_ = grads.values()
    for k in a_:
        a, b = a_[k].double(), b_[k].double()
        print(f"  {k:34s} {tuple(a.shape)!s:14s} |g| {float(a.norm()):.3e}  rel {float((a - b).norm()) / max(float(a.norm()), 1e-12):.3e}")
