import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import dwbc_b200
from dwbc_b200 import _lib as L
lib = L.lib()
lib.dwbc_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
torch.set_printoptions(linewidth=250, precision=0, sci_mode=False)
def probe(M, N, K):
    G = torch.zeros(M, K, device="cuda")
    for m in range(M): G[m, m % K] = 1.0
    W = (torch.arange(K, device="cuda").float()[:, None] * 100 + torch.arange(N, device="cuda").float()[None, :]).contiguous()
    dX = torch.zeros(M, N, device="cuda")
    L.check(lib.dwbc_debug_gemm(1, 1, G.data_ptr(), K, W.data_ptr(), N, dX.data_ptr(), N, None, None, M, N, K, 0, L.stream_ptr()), "g")
    torch.cuda.synchronize()
    print(f"M={M} N={N} K={K}: expected row m = (m%K)*100 + n")
    print(dX[:min(M, 18)].cpu())
probe(16, 16, 8)
probe(16, 32, 16)
