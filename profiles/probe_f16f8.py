"""GPU probe of AZG_PREC_F16F8 (fp16 product + block-scaled FP8 correction product in one accumulator).
Run once per AZG_F8_TERMS in {unset, 1, 2}: with a single term the GEMM output is compared with the exact emulation of
that term (operand roundings in torch, fp64 accumulation), which pins the instruction/scale-factor semantics;
with both terms: error against fp64, Connect4 forward errors against the oracle, and timing against bf16x3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from azgnn_b200 import _lib

SA, LO = 3, 11


def q8(t):
    return t.clamp(-448, 448).to(torch.float8_e4m3fn).double()


def w_exp(W):
    m = float(W.abs().max())
    e = int(np.floor(np.log2(448.0 / m)))
    while m * 2.0 ** e > 448.0:
        e -= 1
    while m * 2.0 ** (e + 1) <= 448.0:
        e += 1
    return e


def emulate(A, W, terms):
    A, W = A.double().cpu(), W.double().cpu()
    Ah, Wh = A.half().double(), W.half().double()
    out = torch.zeros(A.shape[0], W.shape[0], dtype=torch.float64)
    if terms & 1:
        out += Ah @ Wh.t()
    if terms & 2:
        sw = w_exp(W)
        A8, Al8 = q8(A * 2.0 ** SA), q8((A - Ah) * 2.0 ** (SA + LO))
        W8, Wl8 = q8(W * 2.0 ** sw), q8((W - Wh) * 2.0 ** (sw + LO))
        out += (A8 @ Wl8.t() + Al8 @ W8.t()) * 2.0 ** -(SA + sw + LO)
    return out


def linear(M, F, relu, terms, seed=0, wscale=1.0):
    lib = _lib.lib()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, F, generator=g) * 0.5
    W = (torch.rand(F, F, generator=g) * 2 - 1) / F ** 0.5 * wscale
    b = torch.randn(F, generator=g) * 0.1
    Ad, Wd, bd = A.cuda(), W.cuda(), b.cuda()
    C = torch.full((M, F), float("nan"), device="cuda")
    Mp = (M + 255) // 256 * 256
    scratch = torch.zeros(2 * (Mp + F) * F * 2 + 4096, dtype=torch.uint8, device="cuda")
    _lib.check(lib.azg_tc_linear(_lib.ptr(Ad), _lib.ptr(Wd), _lib.ptr(bd), _lib.ptr(C), M, F, _lib.PREC_F16F8, relu,
                                 _lib.ptr(scratch), scratch.numel(), _lib.stream()))
    torch.cuda.synchronize()
    model = emulate(A, W, terms) + b.double()
    exact = A.double() @ W.double().t() + b.double()
    if relu:
        model, exact = model.clamp(min=0), exact.clamp(min=0)
    got = C.double().cpu()
    e_model = (got - model).abs()
    e_exact = (got - exact).abs()
    nan = int(torch.isnan(C).sum())
    print(f"linear M={M} F={F} relu={relu} terms={terms} wscale={wscale}: max|got-emulation|={e_model.max().item():.3e} "
          f"max|got-fp64|={e_exact.max().item():.3e} (emulation vs fp64 {float((model - exact).abs().max()):.3e}) nan={nan} "
          f"rms out={float(exact.pow(2).mean().sqrt()):.3f}")
    if terms == 2:
        corr = emulate(A, W, 2)
        print(f"    correction term alone: rms={float(corr.pow(2).mean().sqrt()):.3e}; got-bias rms={float((got - b.double()).pow(2).mean().sqrt()):.3e}; "
              f"ratio got/emul (median)={float(((got - b.double()) / corr).median()):.4f}")
    bad = e_model > max(1e-3, 10 * float(e_model.median()))
    if bad.any() or nan:
        print("    bad fraction", float(bad.float().mean()), "by row%8", [round(float(bad[i::8].float().mean()), 3) for i in range(8)],
              "by col%32 (step 4)", [round(float(bad[:, i::32].float().mean()), 3) for i in range(0, 32, 4)])
        print("    sample got", got[0, :6].tolist(), "emul", model[0, :6].tolist())
    return float(e_model.max())


def c4(n, B, fold=False):
    from oracle import nets as onets
    from oracle import rules as orules
    from azgnn_b200.nets import B200Connect4GNNWrapper

    class Args(dict):
        __getattr__ = dict.__getitem__
    torch.manual_seed(0)
    w = B200Connect4GNNWrapper(orules.Connect4Rules(n), Args(lr=1e-3, dropout=0.3, epochs=1, batch_size=64, gnn_layers=2, use_gnn=True,
                                                             b200_precision="f16f8"))
    w.fold_heads = fold
    boards = np.random.default_rng(B + n).integers(-1, 2, size=(B, n, n)).astype(np.int64)
    p = {k: v.detach().cpu() for k, v in w.nnet.state_dict().items()}
    q = {k: v.detach().cpu() for k, v in w.gnn.state_dict().items()}
    with torch.no_grad():
        spi, sv = onets.c4_predict(p, onets.boards_to_tensor(boards), n)
        gpi, gv = onets.c4_predict_with_gnn(p, q, onets.boards_to_tensor(boards), n)
    states = w.states_from_boards(boards)
    res = []
    for prec in (_lib.PREC_F16F8, _lib.PREC_BF16X3):
        o = w.forward_states(states, _lib.EVAL_STD | _lib.EVAL_GNN, precision=prec)
        res.append(tuple(float(np.abs(o[k].cpu().numpy().reshape(t.shape) - t.numpy()).max())
                         for k, t in (("pi", spi), ("v", sv), ("pi_gnn", gpi), ("v_gnn", gv))))
    print(f"c4 n={n} B={B} fold={fold}: f16f8 |dpi|,|dv|,|dpi_gnn|,|dv_gnn| = " + " ".join(f"{e:.2e}" for e in res[0]) +
          "   bf16x3 = " + " ".join(f"{e:.2e}" for e in res[1]))
    return w


def timing(w, B=65536, iters=20):
    rng = np.random.default_rng(1)
    states = w.states_from_boards(rng.integers(-1, 2, size=(B, 7, 7)).astype(np.int8))
    for name, prec in (("bf16x3", _lib.PREC_BF16X3), ("f16f8", _lib.PREC_F16F8), ("bf16", _lib.PREC_BF16)):
        for fold in (False, True):
            w.fold_heads = fold
            for _ in range(5):
                w.forward_states(states, _lib.EVAL_STD | _lib.EVAL_GNN, precision=prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                w.forward_states(states, _lib.EVAL_STD | _lib.EVAL_GNN, precision=prec)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"timing B={B} {name} fold={fold}: {ms:.3f} ms/step = {B / ms / 1e3:.2f} M leaf evals/s")
    w.fold_heads = False


if __name__ == "__main__":
    terms = int(os.environ.get("AZG_F8_TERMS", "3")) & 3 or 3
    torch.manual_seed(0)
    for (M, F) in [(256, 1024), (200, 3136), (1000, 1600), (4096, 3136)]:
        linear(M, F, 0, terms)
    linear(300, 3136, 0, terms, wscale=37.0)
    linear(300, 3136, 1, terms, wscale=0.01)
    if terms == 3:
        w7 = None
        for n, B in [(7, 777), (7, 1), (5, 300), (4, 300), (6, 300), (8, 300)]:
            w = c4(n, B)
            if n == 7 and B == 777:
                w7 = w
        c4(7, 777, fold=True)
        timing(w7)
