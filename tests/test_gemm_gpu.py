"""K3 building blocks through the C ABI against torch fp32/fp64: `azg_gemm_f32` over every operand orientation,
ragged tails, unaligned leading dimensions (generic kernel), split-K shapes, the matrix-vector and rank-1 dispatches."""
import numpy as np
import pytest
import torch

from azgnn_b200 import _lib
from azgnn_b200._lib import ptr, stream

pytestmark = pytest.mark.gpu


def _gemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta):
    _lib.check(_lib.lib().azg_gemm_f32(int(ta), int(tb), M, N, K, ptr(A), lda, ptr(B), ldb, ptr(C), ldc, float(beta), stream()))


SHAPES = [(64, 3136, 3136), (64, 8, 3136), (1, 3136, 6272), (1, 128, 3136), (63, 128, 3136), (3136, 288, 64), (128, 3136, 63),
          (64, 288, 3136), (1, 3136, 63), (3136, 6272, 1), (128, 3136, 1), (5, 7, 9), (130, 257, 33), (1, 63, 3136), (200, 64, 72)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_gemm_f32_matches_torch(M, N, K, ta, tb, beta):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    pad_a, pad_b, pad_c = (0, 4, 1)[(M + N) % 3], (4, 0, 3)[(N + K) % 3], (0, 8)[K % 2]
    a_shape = (K, M + pad_a) if ta else (M, K + pad_a)
    b_shape = (N, K + pad_b) if tb else (K, N + pad_b)
    A = torch.randn(a_shape, device="cuda", generator=g)
    B = torch.randn(b_shape, device="cuda", generator=g)
    C = torch.randn(M, N + pad_c, device="cuda", generator=g)
    C0 = C.clone()
    opA = (A[:, :M].t() if ta else A[:, :K]).double()
    opB = (B[:, :K].t() if tb else B[:, :N]).double()
    want = opA @ opB + beta * C0[:, :N].double()
    _gemm(ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, C.shape[1], beta)
    got = C[:, :N].double()
    scale = (opA.abs() @ opB.abs()).max().item() + 1.0
    assert (got - want).abs().max().item() <= 2e-6 * scale
    if pad_c:
        assert torch.equal(C[:, N:], C0[:, N:])  # nothing written beyond the N columns


def test_linear_backward_pieces():
    """dX = g W, dW = g^T X, db = colsum(g) with and without the ReLU gate, small and split batch sizes"""
    lib = _lib.lib()
    for M, N, K, relu in [(64, 3136, 3136, 1), (1, 3136, 6272, 0), (9000, 64, 128, 1), (64, 8, 3136, 0)]:
        g = torch.Generator(device="cuda").manual_seed(M + N)
        X = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
        b = torch.randn(N, device="cuda", generator=g)
        Y = torch.relu(X @ W.t() + b) if relu else X @ W.t() + b
        dY = torch.randn(M, N, device="cuda", generator=g)
        dX, dW, db = torch.empty_like(X), torch.empty_like(W), torch.empty_like(b)
        scratch = torch.empty(M * N, device="cuda")
        _lib.check(lib.azg_linear_backward(ptr(dY), ptr(X), ptr(W), ptr(Y), M, N, K, relu, ptr(dX), ptr(dW), ptr(db), ptr(scratch),
                                           stream()))
        gd = (dY * (Y > 0)).double() if relu else dY.double()
        for got, want in ((dX, gd @ W.double()), (dW, gd.t() @ X.double()), (db, gd.sum(0))):
            assert (got.double() - want).abs().max().item() <= 1e-5 * (want.abs().max().item() + 1.0)
