"""GPU probe of the tcgen05 GEMM (azg_tc_linear) with error diagnostics -- run under `timeout`."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from azgnn_b200 import _lib


def run(M, F, prec, relu, seed=0):
    lib = _lib.lib()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, F, generator=g) * 0.5
    W = (torch.rand(F, F, generator=g) * 2 - 1) / F ** 0.5
    b = torch.randn(F, generator=g) * 0.1
    Ad, Wd, bd = A.cuda(), W.cuda(), b.cuda()
    C = torch.full((M, F), float("nan"), device="cuda")
    Mp = (M + 255) // 256 * 256
    nbytes = 2 * (Mp + F) * F * 2 + 1024 if prec == _lib.PREC_BF16X3 else (Mp + F) * F * 2 + 1024
    scratch = torch.zeros(nbytes + 1024, dtype=torch.uint8, device="cuda")
    _lib.check(lib.azg_tc_linear(_lib.ptr(Ad), _lib.ptr(Wd), _lib.ptr(bd), _lib.ptr(C), M, F, prec, relu,
                                 _lib.ptr(scratch), scratch.numel(), _lib.stream()))
    torch.cuda.synchronize()
    ref64 = Ad.double() @ Wd.double().t() + bd.double()
    if prec == _lib.PREC_BF16:
        refb = Ad.bfloat16().double() @ Wd.bfloat16().double().t() + bd.double()
    else:
        refb = ref64
    if relu:
        ref64, refb = ref64.clamp(min=0), refb.clamp(min=0)
    err = (C.double() - refb).abs()
    err64 = (C.double() - ref64).abs()
    bad = (err > 1e-3) | torch.isnan(C)
    print(f"M={M} F={F} prec={prec} relu={relu}: max_err_vs_model={err.max().item():.3e} max_err_vs_fp64={err64.max().item():.3e} "
          f"mean={err.mean().item():.3e} bad_frac={bad.float().mean().item():.4f} nan={torch.isnan(C).sum().item()}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten()[:16].tolist()
        cols = bad.any(0).nonzero().flatten()[:16].tolist()
        print("   bad rows:", rows, " bad cols:", cols)
        print("   bad by row%8:", [round(bad[i::8].float().mean().item(), 3) for i in range(8)])
        print("   bad by col%64 (first 8 of 64):", [round(bad[:, i::64].float().mean().item(), 3) for i in range(0, 64, 8)])
        print("   sample got/ref:", C[0, :6].tolist(), refb[0, :6].tolist())
    return err.max().item()


if __name__ == "__main__":
    torch.manual_seed(0)
    for (M, F) in [(128, 1024), (200, 3136), (1024, 3136), (256, 1600)]:
        for prec in (_lib.PREC_BF16, _lib.PREC_BF16X3):
            run(M, F, prec, 0)
    run(4096, 3136, _lib.PREC_BF16X3, 1)
