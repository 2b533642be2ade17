"""GPU: the GEMM building blocks (fp32 CUDA-core tile kernel and TF32 tcgen05 kernel) against torch fp32/fp64
matmul, all three operand modes, including the odd widowGo1 layer shapes (K = 24 / 76 / 96 / 100, N = 1 / 6 / 12 / 20)."""
import ctypes as C

import numpy as np
import pytest
import torch

from dwbc_b200 import _lib as L, synth

pytestmark = pytest.mark.gpu

SHAPES = [(40960, 128, 128), (4096, 128, 96), (4096, 128, 100), (4096, 64, 24), (4096, 20, 64), (1000, 12, 128), (777, 6, 128), (333, 1, 128),
          (4096, 30, 76), (64, 128, 128), (129, 20, 36)]


def _lib():
    lib = L.lib()
    lib.dwbc_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.dwbc_debug_gemm.restype = C.c_int
    return lib


def _rand(seed, stream, shape):
    return torch.from_numpy(synth.normal(seed, stream, shape)).cuda()


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_forward(tc, M, N, K):
    lib = _lib()
    X, W, b = _rand(1, 1, (M, K)), _rand(1, 2, (N, K)) / np.sqrt(K), _rand(1, 3, (N,))
    Y = torch.full((M, N), float("nan"), device="cuda")
    for act, fn in ((0, lambda t: t), (1, torch.nn.functional.elu), (2, torch.tanh)):
        L.check(lib.dwbc_debug_gemm(0, tc, X.data_ptr(), K, W.data_ptr(), K, Y.data_ptr(), N, b.data_ptr(), None, M, N, K, act, L.stream_ptr()), "gemm")
        ref = fn((X.double() @ W.double().T + b.double())).float()
        tol = 2e-5 if tc == 0 else 6e-3          # TF32: 10-bit mantissa inputs, |y| ~ 1
        torch.testing.assert_close(Y, ref, rtol=tol, atol=tol)


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_backward_data(tc, M, N, K):
    """dX[M, N_in] = G[M, K_out] W[K_out, N_in]"""
    lib = _lib()
    G, W = _rand(2, 1, (M, K)), _rand(2, 2, (K, N)) / np.sqrt(K)
    dX = torch.full((M, N), float("nan"), device="cuda")
    L.check(lib.dwbc_debug_gemm(1, tc, G.data_ptr(), K, W.data_ptr(), N, dX.data_ptr(), N, None, None, M, N, K, 0, L.stream_ptr()), "gemm")
    ref = (G.double() @ W.double()).float()
    tol = 2e-5 if tc == 0 else 6e-3
    torch.testing.assert_close(dX, ref, rtol=tol, atol=tol)


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("R,N,K", SHAPES)
def test_linear_backward_weight(tc, R, N, K):
    """dW[N_out, K_in] += G[R, N_out]^T X[R, K_in]; db += colsum(G)"""
    lib = _lib()
    G, X = _rand(3, 1, (R, N)) / np.sqrt(R), _rand(3, 2, (R, K))
    dW = torch.zeros(N, K, device="cuda")
    db = torch.zeros(N, device="cuda")
    L.check(lib.dwbc_debug_gemm(2, tc, G.data_ptr(), N, X.data_ptr(), K, dW.data_ptr(), K, None, db.data_ptr(), N, K, R, 0, L.stream_ptr()), "gemm")
    ref = (G.double().T @ X.double()).float()
    tol = 3e-5 if tc == 0 else 6e-3
    torch.testing.assert_close(dW, ref, rtol=tol, atol=tol)
    torch.testing.assert_close(db, G.double().sum(0).float(), rtol=1e-4, atol=1e-4)
