"""K1/K2 parity on the GPU: CUDA forward of the three games' networks vs the torch-fp32 oracle
(oracle/nets.py) and vs the committed golden outputs of the reference itself.

Tolerances (north_star): 1e-5 absolute on pi and v for the fp32 path; the tcgen05 paths state
their own below."""
import numpy as np
import pytest
import torch

from oracle import nets as onets
from oracle import rules as orules

from azgnn_b200 import _lib
from azgnn_b200.mcts import pack_states
from azgnn_b200.nets import (B200Connect4GNNWrapper, B200FrozenLakeNet, B200TicTacToeGNNWrapper)
from helpers import dotdict, golden

pytestmark = pytest.mark.gpu
TOL = dict(rtol=0, atol=1e-5)


def _args(**kw):
    # the fp32 CUDA-core path unless a test asks otherwise (the wrappers' own default is bf16x3; the tensor-core tests
    # pass `precision=` per call and test_default_precision_is_the_tensor_core_path checks the default)
    return dotdict(dict(dict(lr=1e-3, dropout=0.3, epochs=1, batch_size=64, gnn_layers=2, use_gnn=True, b200_precision="fp32"), **kw))


def _wrapper(kind, n, **kw):
    game = (orules.Connect4Rules if kind == "c4" else orules.TicTacToeRules)(n)
    torch.manual_seed(0)  # same construction order as the reference -> the golden weights
    return (B200Connect4GNNWrapper if kind == "c4" else B200TicTacToeGNNWrapper)(game, _args(**kw))


def _cpu_sd(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("kind,n", [("c4", 5), ("c4", 7), ("ttt", 3), ("ttt", 4)])
def test_golden_outputs_of_the_reference(kind, n):
    g = golden(f"nets_{kind}_{n}")
    w = _wrapper(kind, n)
    out = w.predict_batch(g["boards"])
    np.testing.assert_allclose(out["pi"], g["pi"], **TOL)
    np.testing.assert_allclose(out["v"], g["v"], **TOL)
    np.testing.assert_allclose(out["pi_gnn"], g["gnn_pi"], **TOL)
    np.testing.assert_allclose(out["v_gnn"], g["gnn_v"], **TOL)
    # reference-facing single-position calls return the reference's types
    pi, v = w.predict(g["boards"][0])
    assert pi.dtype == np.float32 and pi.shape == (w.action_size,) and isinstance(v, np.float32)
    np.testing.assert_allclose(pi, g["pi"][0], **TOL)
    gpi, gv = w.predict_with_gnn(g["boards"][0])
    np.testing.assert_allclose(gpi, g["gnn_pi"][0], **TOL)
    assert abs(float(gv) - float(g["gnn_v"][0])) <= 1e-5


@pytest.mark.parametrize("kind,n,B", [("c4", 7, 1), ("c4", 7, 333), ("c4", 7, 4096), ("c4", 4, 130), ("c4", 8, 257),
                                      ("ttt", 3, 1000), ("ttt", 4, 1000), ("ttt", 5, 129)])
def test_forward_matches_oracle(kind, n, B):
    w = _wrapper(kind, n)
    rng = np.random.default_rng(B + n)
    boards = rng.integers(-1, 2, size=(B, n, n)).astype(np.int64)
    out = w.predict_batch(boards)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    bt = onets.boards_to_tensor(boards)
    with torch.no_grad():
        if kind == "c4":
            pi, v = onets.c4_predict(p, bt, n)
            gpi, gv = onets.c4_predict_with_gnn(p, q, bt, n)
        else:
            pi, v = onets.ttt_predict(p, bt, n)
            gpi, gv = onets.ttt_predict_with_gnn(p, q, bt, n)
    np.testing.assert_allclose(out["pi"], pi.numpy(), **TOL)
    np.testing.assert_allclose(out["v"], v.numpy(), **TOL)
    np.testing.assert_allclose(out["pi_gnn"], gpi.numpy(), **TOL)
    np.testing.assert_allclose(out["v_gnn"], gv.numpy(), **TOL)
    assert np.allclose(out["pi"].sum(1), 1, atol=1e-5)


def test_input_dtypes_and_empty_batch():
    w = _wrapper("c4", 7)
    rng = np.random.default_rng(1)
    boards = rng.integers(-1, 2, size=(64, 7, 7))
    ref = w.predict_batch(boards.astype(np.int64))
    for dt in (np.int8, np.float32, np.float64):
        out = w.predict_batch(boards.astype(dt))
        assert np.array_equal(out["pi_gnn"], ref["pi_gnn"])
    pinned = torch.from_numpy(boards.astype(np.int8)).pin_memory()
    assert np.array_equal(w.predict_batch(pinned)["v_gnn"], ref["v_gnn"])
    empty = w.predict_batch(np.zeros((0, 7, 7), dtype=np.int64))
    assert empty["pi"].shape == (0, 8)


def test_encode_planes_and_pack_round_trip():
    """K1: states -> planes equals the reference's float tensor of the board (Connect4GNN.py:71-72)."""
    lib = _lib.lib()
    rng = np.random.default_rng(3)
    for n in (3, 4, 7, 8):
        B = 1000
        boards = rng.integers(-1, 2, size=(B, n, n)).astype(np.int64)
        dev = torch.device("cuda")
        cells = torch.from_numpy(boards).to(dev)
        states = torch.empty(B, 2, dtype=torch.int64, device=dev)
        _lib.check(lib.azg_pack_boards(_lib.ptr(cells), _lib.CELL_I64, n, B, _lib.ptr(states), _lib.stream()))
        assert np.array_equal(states.cpu().numpy(), pack_states("connect4", boards))
        planes = torch.empty(B, n * n, dtype=torch.float32, device=dev)
        _lib.check(lib.azg_encode_planes(_lib.ptr(states), n, B, _lib.ptr(planes), _lib.stream()))
        assert np.array_equal(planes.cpu().numpy().reshape(B, n, n), boards.astype(np.float32))


@pytest.mark.parametrize("n,layers", [(4, 2), (4, 3), (8, 2), (8, 3)])
def test_frozenlake_forward(n, layers):
    g = golden(f"nets_fl_{n}_L{layers}")
    game = orules.FrozenLakeRules(n)
    torch.manual_seed(0)
    w = B200FrozenLakeNet(game, dotdict(dict(lr=1e-3, embedding_dim=128, gnn_layers=layers)))
    cells = g["cells"]
    states = np.zeros((len(cells), 2), dtype=np.int64)
    states[:, 0] = cells
    out = w.forward_states(torch.from_numpy(states).cuda())
    np.testing.assert_allclose(out["pi"].cpu().numpy(), g["pi"], **TOL)
    np.testing.assert_allclose(out["v"].cpu().numpy(), g["v"][:, 0], **TOL)
    b = np.zeros((n, n)); b[0, 0] = 1
    pi, v = w.predict(b)
    assert v.shape == (1,) and v.dtype == np.float32  # FrozenLakeNet.py:226
    # K1 graph build (FrozenLakeNet.py:197-213)
    lib = _lib.lib()
    st = torch.from_numpy(states).cuda()
    nodes = torch.empty(len(cells), 5, n * n, dtype=torch.float32, device="cuda")
    counts = torch.empty(len(cells), dtype=torch.int32, device="cuda")
    _lib.check(lib.azg_fl_encode_graph(_lib.ptr(st), n, len(cells), _lib.ptr(nodes), _lib.ptr(counts), _lib.stream()))
    for i, cell in enumerate(cells):
        want = onets.fl_node_cells(int(cell), n)
        assert int(counts[i]) == len(want)
        got = nodes[i].cpu().numpy()
        for j, c in enumerate(want):
            assert got[j].argmax() == c and got[j].sum() == 1
        assert got[len(want):].sum() == 0


def test_checkpoint_round_trip(tmp_path):
    w = _wrapper("c4", 5)
    w.save_checkpoint(str(tmp_path), "best_gnn.pth.tar")
    ck = torch.load(str(tmp_path / "best_gnn.pth.tar"), map_location="cpu")
    assert set(ck.keys()) == {"state_dict", "gnn"}  # Connect4GNN.py:205-208
    assert "conv1.weight" in ck["state_dict"] and "output_transform.0.weight" in ck["gnn"]
    b = np.random.default_rng(0).integers(-1, 2, size=(8, 5, 5))
    before = w.predict_batch(b)
    torch.manual_seed(123)
    w2 = B200Connect4GNNWrapper(orules.Connect4Rules(5), _args())
    w2.load_checkpoint(str(tmp_path), "best_gnn.pth.tar")
    after = w2.predict_batch(b)
    assert np.array_equal(before["pi_gnn"], after["pi_gnn"])


# ------------------------------------------------------------------------------------ tcgen05 path
def _tc_linear(A, W, b, prec, relu):
    lib = _lib.lib()
    M, F = A.shape
    Mp = (M + 255) // 256 * 256
    nbytes = (1 if prec == _lib.PREC_BF16 else 2) * (Mp + F) * F * 2 + 4096
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    C = torch.full((M, F), float("nan"), device="cuda")
    _lib.check(lib.azg_tc_linear(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(C), M, F, prec, relu, _lib.ptr(scratch),
                                 scratch.numel(), _lib.stream()))
    return C


@pytest.mark.parametrize("M,F", [(1, 1024), (128, 1024), (200, 3136), (1000, 1600), (4096, 3136)])
@pytest.mark.parametrize("relu", [0, 1])
def test_tc_linear(M, F, relu):
    """The tensor-core dense layer vs a float64 reference.  bf16x3 (3-term split, fp32 accumulate)
    keeps ~16 mantissa bits per operand: |err| <= 1e-4 at K = 3136 with O(1) outputs; plain bf16 is
    compared against the same product of bf16-rounded operands (fp32 accumulation error only)."""
    g = torch.Generator().manual_seed(M + F)
    A = (torch.randn(M, F, generator=g) * 0.5).cuda()
    W = ((torch.rand(F, F, generator=g) * 2 - 1) / F ** 0.5).cuda()
    b = (torch.randn(F, generator=g) * 0.1).cuda()
    ref = A.double() @ W.double().t() + b.double()
    refb = A.bfloat16().double() @ W.bfloat16().double().t() + b.double()
    if relu:
        ref, refb = ref.clamp(min=0), refb.clamp(min=0)
    c3 = _tc_linear(A, W, b, _lib.PREC_BF16X3, relu)
    assert (c3.double() - ref).abs().max().item() <= 1e-4
    c1 = _tc_linear(A, W, b, _lib.PREC_BF16, relu)
    assert (c1.double() - refb).abs().max().item() <= 5e-5
    # fp16 product + block-scaled FP8 correction product (one accumulator): ~15 mantissa bits per operand
    c8 = _tc_linear(A, W, b, _lib.PREC_F16F8, relu)
    e8 = (c8.double() - ref).abs().max().item()
    print(f"f16f8 M={M} F={F}: max |err| vs fp64 = {e8:.2e}")
    assert e8 <= 1e-4


def _e4m3(t):
    return t.clamp(-448, 448).to(torch.float8_e4m3fn).double()


@pytest.mark.parametrize("wscale", [1.0, 37.0, 0.01])
def test_f16f8_matches_its_exact_emulation(wscale):
    """AZG_PREC_F16F8 pinned term by term: fp16(x) fp16(w) + [e4m3(2^3 x) | e4m3(2^14 x_lo)] . [e4m3(2^(s+11) w_lo) | e4m3(2^s w)]
    * 2^-(s+14) with s = the largest exponent keeping 2^s max|w| <= 448, operands rounded in torch and accumulated in fp64.
    The kernel (tcgen05 kind::f16 + kind::mxf8f6f4.block_scale with constant UE8M0 scale factors, fp32 accumulation in TMEM)
    must agree to fp32 accumulation noise -- this fixes the scale-factor semantics, the K-concatenation order and the
    per-tensor weight scale, not just "small error"."""
    g = torch.Generator().manual_seed(17)
    M, F = 300, 3136
    A = torch.randn(M, F, generator=g) * 0.5
    W = (torch.rand(F, F, generator=g) * 2 - 1) / F ** 0.5 * wscale
    b = torch.randn(F, generator=g) * 0.1
    got = _tc_linear(A.cuda(), W.cuda(), b.cuda(), _lib.PREC_F16F8, 0).double().cpu()
    Ad, Wd = A.double(), W.double()
    Ah, Wh = A.half().double(), W.half().double()
    m = float(W.abs().max())
    s_w = int(np.floor(np.log2(448.0 / m)))
    while m * 2.0 ** s_w > 448.0:
        s_w -= 1
    while m * 2.0 ** (s_w + 1) <= 448.0:
        s_w += 1
    corr = (_e4m3(Ad * 8.0) @ _e4m3((Wd - Wh) * 2.0 ** (s_w + 11)).t() + _e4m3((Ad - Ah) * 2.0 ** 14) @ _e4m3(Wd * 2.0 ** s_w).t()) * 2.0 ** -(s_w + 14)
    model = Ah @ Wh.t() + corr + b.double()
    scale = float(model.abs().max())
    e_model = (got - model).abs().max().item()
    e_exact = (got - (Ad @ Wd.t() + b.double())).abs().max().item()
    print(f"wscale {wscale}: max |out| {scale:.2f}, |got - emulation| {e_model:.2e}, |got - fp64| {e_exact:.2e}")
    assert e_model <= 2e-5 * max(scale, 1.0)  # the tensor core's fp32 accumulation over K = 3136 (+ 6272 correction terms)
    assert e_exact <= 4e-5 * max(scale, 1.0)
    # the correction product is really there: without it the fp16 rounding of the operands alone costs ~4e-4
    assert e_exact < 0.2 * (Ah @ Wh.t() + b.double() - (Ad @ Wd.t() + b.double())).abs().max().item()
    # AZG_PREC_F16F8_KS: the same operands, K accumulated in four launches with round-to-nearest fp32 adds in between.
    # What separates the kernel from its emulation is the tensor core's truncating accumulation, linear in the number of
    # k-steps: a quarter of K per accumulator must bring it down (measured ~3x; the final adds round too)
    got_ks = _tc_linear(A.cuda(), W.cuda(), b.cuda(), _lib.PREC_F16F8_KS, 0).double().cpu()
    e_model_ks = (got_ks - model).abs().max().item()
    mean_ks, mean_1 = (got_ks - model).abs().mean().item(), (got - model).abs().mean().item()
    print(f"   K-split: |got - emulation| max {e_model_ks:.2e} (one accumulator {e_model:.2e}), mean {mean_ks:.2e} ({mean_1:.2e})")
    assert e_model_ks <= 0.6 * e_model and mean_ks <= 0.5 * mean_1


@pytest.mark.parametrize("n,B", [(7, 1), (7, 777), (7, 4096), (4, 300), (5, 300), (6, 300), (8, 300)])
def test_forward_tensor_core_precisions(n, B):
    """Connect4 GNN head on tcgen05: bf16x3 stays inside the fp32 contract (1e-5 on pi and v);
    plain bf16 is the throughput mode with a stated tolerance of 5e-3 (random-init weights)."""
    w = _wrapper("c4", n)
    rng = np.random.default_rng(B + n)
    boards = rng.integers(-1, 2, size=(B, n, n)).astype(np.int64)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    with torch.no_grad():
        gpi, gv = onets.c4_predict_with_gnn(p, q, onets.boards_to_tensor(boards), n)
    with torch.no_grad():
        spi, sv = onets.c4_predict(p, onets.boards_to_tensor(boards), n)
    states = w.states_from_boards(boards)
    both = _lib.EVAL_STD | _lib.EVAL_GNN
    for prec in (_lib.PREC_F16F8, _lib.PREC_F16F8_KS, _lib.PREC_BF16X3, _lib.PREC_BF16X3_KS):  # the splits hold the fp32 contract
        o3 = w.forward_states(states, both, precision=prec)
        np.testing.assert_allclose(o3["pi_gnn"].cpu().numpy(), gpi.numpy(), rtol=0, atol=1e-5)
        np.testing.assert_allclose(o3["v_gnn"].cpu().numpy(), gv.numpy(), rtol=0, atol=1e-5)
        # the trunk runs on the tensor cores too (conv2 as an implicit GEMM): std heads see the same features
        np.testing.assert_allclose(o3["pi"].cpu().numpy(), spi.numpy(), rtol=0, atol=1e-5)
        np.testing.assert_allclose(o3["v"].cpu().numpy(), sv.numpy(), rtol=0, atol=1e-5)
        only_std = w.forward_states(states, _lib.EVAL_STD, precision=prec)
        assert torch.equal(only_std["pi"], o3["pi"])
    o1 = w.forward_states(states, both, precision=_lib.PREC_BF16)
    assert np.abs(o1["pi"].cpu().numpy() - spi.numpy()).max() <= 5e-3 and np.abs(o1["v"].cpu().numpy() - sv.numpy()).max() <= 5e-3
    e_pi = np.abs(o1["pi_gnn"].cpu().numpy() - gpi.numpy()).max()
    e_v = np.abs(o1["v_gnn"].cpu().numpy() - gv.numpy()).max()
    print(f"bf16 n={n} B={B}: max |dpi| = {e_pi:.2e}, max |dv| = {e_v:.2e}")
    assert e_pi <= 5e-3 and e_v <= 5e-3
    # re-tiling follows weight updates
    with torch.no_grad():
        w.gnn.output_transform[2].bias.add_(0.25)
    w.weights_changed()
    o3b = w.forward_states(states, _lib.EVAL_GNN, precision=_lib.PREC_BF16X3)
    assert not np.allclose(o3b["v_gnn"].cpu().numpy(), o3["v_gnn"].cpu().numpy())


@pytest.mark.parametrize("prec", ["f16f8", "f16f8ks", "bf16x3ks"])
@pytest.mark.parametrize("fold", [False, True])
def test_device_row_count_and_fold_per_precision(prec, fold):
    """forward_states(count=device scalar), the compacted leaf batches of the search: the first `count` rows equal a plain
    call on those rows bit for bit (also through the K-split launches, which all read the count on the device), with and
    without the head fold, and stay within 1e-5 of the oracle."""
    w = _wrapper("c4", 7)
    rng = np.random.default_rng(21)
    boards = rng.integers(-1, 2, size=(700, 7, 7)).astype(np.int64)
    states = w.states_from_boards(boards)
    p_, live = _lib.PRECISIONS[prec], 333
    mask = _lib.EVAL_STD | _lib.EVAL_GNN | (_lib.EVAL_FOLD if fold else 0)
    count = torch.tensor([live], dtype=torch.int32, device="cuda")
    out = {k: torch.full_like(v, 7.0) for k, v in w._outputs(700, mask).items()}
    w.forward_states(states, mask, precision=p_, count=count, out=out)
    plain = w.forward_states(states[:live].contiguous(), mask, precision=p_)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    with torch.no_grad():
        gpi, gv = onets.c4_predict_with_gnn(p, q, onets.boards_to_tensor(boards[:live]), 7)
    for k in plain:
        assert torch.equal(out[k][:live], plain[k]), k
        assert bool((out[k][live:] == 7.0).all()), k  # rows beyond the count are left untouched
    assert np.abs(plain["pi_gnn"].cpu().numpy() - gpi.numpy()).max() <= 1e-5
    assert np.abs(plain["v_gnn"].cpu().numpy() - gv.numpy()).max() <= 1e-5


_KS_DUMP = """
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
import azgnn_b200
from azgnn_b200 import _lib, games
from azgnn_b200.nets import B200Connect4GNNWrapper
out = {{}}
for n, B in ((7, 2500), (5, 700), (4, 300)):
    torch.manual_seed(n)
    w = B200Connect4GNNWrapper(games.Connect4Game(n), dict(lr=1e-3, dropout=0.3, gnn_layers=2, use_gnn=True, b200_precision="f16f8ks"))
    states = w.states_from_boards(np.random.default_rng(n).integers(-1, 2, size=(B, n, n)).astype(np.int8))
    for fold in (0, _lib.EVAL_FOLD):
        o = w.forward_states(states, _lib.EVAL_STD | _lib.EVAL_GNN | fold)
        for k, v in o.items():
            out[f"{{n}}_{{fold}}_{{k}}"] = v.cpu().numpy()
    o = w.forward_states(states, _lib.EVAL_STD)
    out[f"{{n}}_std_only_v"] = o["v"].cpu().numpy()
np.savez({path!r}, **out)
"""


def test_in_kernel_k_split_equals_the_split_by_launches(tmp_path):
    """AZG_PREC_F16F8_KS inside one launch (four work items per tile on alternating TMEM accumulators, partial sums in a per-CTA
    scratch tile that stays in L2) performs the same fp32 adds in the same order as four launches over a quarter of K each
    (AZG_KSPLIT=launches, partial sums through an [M, F] matrix in HBM): bit-identical outputs, 7x7 / 5x5 / 4x4, ragged
    last tiles, with and without the head fold, std-only calls (side tiles only).  Env read once per process -> subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    got = {}
    for mode in ("kernel", "launches"):
        path = str(tmp_path / f"ks_{mode}.npz")
        env = dict(os.environ)
        env.pop("AZG_KSPLIT", None)
        if mode == "launches":
            env["AZG_KSPLIT"] = "launches"
        r = subprocess.run([sys.executable, "-c", _KS_DUMP.format(root=root, path=path)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        got[mode] = dict(np.load(path))
    assert set(got["kernel"]) == set(got["launches"]) and len(got["kernel"]) >= 27
    for k in got["kernel"]:
        assert np.array_equal(got["kernel"][k], got["launches"][k]), k


@pytest.mark.parametrize("scale", [0.3, 3.0])
def test_bf16x3_margin_under_weight_scale(scale):
    """The 3-term split keeps ~16 mantissa bits per operand whatever the weight magnitude; with every
    weight matrix scaled (sharper or flatter policies than random init) pi and v still track the fp32
    oracle to 1e-5 / 2e-5."""
    w = _wrapper("c4", 7)
    with torch.no_grad():
        for p_ in list(w.nnet.parameters()) + list(w.gnn.output_transform.parameters()):
            if p_.dim() > 1:
                p_.mul_(scale ** 0.5)
    w.weights_changed()
    rng = np.random.default_rng(11)
    boards = rng.integers(-1, 2, size=(512, 7, 7)).astype(np.int64)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    with torch.no_grad():
        gpi, gv = onets.c4_predict_with_gnn(p, q, onets.boards_to_tensor(boards), 7)
    o3 = w.forward_states(w.states_from_boards(boards), _lib.EVAL_GNN, precision=_lib.PREC_BF16X3)
    e_pi = np.abs(o3["pi_gnn"].cpu().numpy() - gpi.numpy()).max()
    e_v = np.abs(o3["v_gnn"].cpu().numpy() - gv.numpy()).max()
    print(f"scale {scale}: max |dpi| = {e_pi:.2e}, max |dv| = {e_v:.2e}, max pi = {gpi.max().item():.3f}")
    assert e_pi <= 1e-5 and e_v <= 2e-5


@pytest.mark.parametrize("scale", [0.3, 1.0, 3.0, 10.0])
def test_auto_precision_guard(scale):
    """`b200_precision: auto` (the default): per weight version the wrapper evaluates a probe batch in f16f8, then f16f8ks (the
    K-split accumulation), then bf16x3, against its own fp32 CUDA-core path and keeps the first mode within AUTO_TOL; a weight scale that breaks a split makes it
    fall back (down to fp32) instead of silently leaving the 1e-5 contract.  Whatever it chose must hold 1e-5 against the
    oracle on other positions."""
    w = _wrapper("c4", 7, b200_precision="auto")
    with torch.no_grad():
        for p_ in list(w.nnet.parameters()) + list(w.gnn.output_transform.parameters()):
            if p_.dim() > 1:
                p_.mul_(scale ** 0.5)
    w.weights_changed()
    boards = np.random.default_rng(5).integers(-1, 2, size=(384, 7, 7)).astype(np.int64)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    with torch.no_grad():
        spi, sv = onets.c4_predict(p, onets.boards_to_tensor(boards), 7)
        gpi, gv = onets.c4_predict_with_gnn(p, q, onets.boards_to_tensor(boards), 7)
    out = w.predict_batch(boards)
    choice = _lib.PRECISION_NAMES[w.active_precision()]
    errs = {k: float(np.abs(out[k] - t.numpy().reshape(out[k].shape)).max()) for k, t in (("pi", spi), ("v", sv), ("pi_gnn", gpi), ("v_gnn", gv))}
    print(f"scale {scale}: auto -> {choice}; probe {w.precision_report}; errors vs oracle {errs}")
    if scale <= 1.0:
        assert choice == "f16f8"
    assert max(errs.values()) <= (1e-5 if scale < 10 else 4e-5)  # at x10 the fp32 paths themselves differ by rounding order
    # the choice follows the weights
    with torch.no_grad():
        for p_ in list(w.nnet.parameters()) + list(w.gnn.output_transform.parameters()):
            if p_.dim() > 1:
                p_.mul_(scale ** -0.5)
    w.weights_changed()
    w.predict_batch(boards[:4])
    assert _lib.PRECISION_NAMES[w.active_precision()] == "f16f8"


@pytest.mark.parametrize("n,B", [(7, 1), (7, 777), (5, 300), (8, 300)])
@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_folded_heads_match_reference(n, B, scale):
    """EVAL_FOLD: heads(W2 h + b2) evaluated as ([Wp; Wv] W2) h + const in GEMM-1's epilogue (no second F x F
    contraction).  Same pi / v as the reference module within the fp32 contract, also with scaled weights, and it
    follows weight updates."""
    w = _wrapper("c4", n)
    if scale != 1.0:
        with torch.no_grad():
            for p_ in list(w.nnet.parameters()) + list(w.gnn.output_transform.parameters()):
                if p_.dim() > 1:
                    p_.mul_(scale ** 0.5)
        w.weights_changed()
    rng = np.random.default_rng(B + n)
    boards = rng.integers(-1, 2, size=(B, n, n)).astype(np.int64)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    with torch.no_grad():
        gpi, gv = onets.c4_predict_with_gnn(p, q, onets.boards_to_tensor(boards), n)
    states = w.states_from_boards(boards)
    w.fold_heads = True
    o = w.forward_states(states, _lib.EVAL_STD | _lib.EVAL_GNN, precision=_lib.PREC_BF16X3)
    e_pi = np.abs(o["pi_gnn"].cpu().numpy() - gpi.numpy()).max()
    e_v = np.abs(o["v_gnn"].cpu().numpy() - gv.numpy()).max()
    print(f"folded n={n} B={B} scale={scale}: max |dpi| = {e_pi:.2e}, max |dv| = {e_v:.2e}")
    assert e_pi <= 1e-5 and e_v <= 2e-5
    w.fold_heads = False
    ref = w.forward_states(states, _lib.EVAL_STD | _lib.EVAL_GNN, precision=_lib.PREC_BF16X3)
    assert torch.equal(ref["pi"], o["pi"])  # the standard prediction is untouched
    with torch.no_grad():
        w.gnn.output_transform[2].bias.add_(0.25)
    w.weights_changed()
    w.fold_heads = True
    o2 = w.forward_states(states, _lib.EVAL_GNN, precision=_lib.PREC_BF16X3)
    assert not np.allclose(o2["v_gnn"].cpu().numpy(), o["v_gnn"].cpu().numpy())


@pytest.mark.parametrize("n,B", [(3, 1), (3, 1000), (4, 1), (4, 777), (4, 5000), (5, 300), (8, 130)])
def test_tictactoe_tensor_core_forward(n, B):
    """TicTacToe with conv2 / conv3 / fc / output_transform on tcgen05: bf16x3 within the fp32 contract (1e-5 on pi, v),
    bf16 within its stated 5e-3; std and GNN predictions; follows weight updates."""
    w = _wrapper("ttt", n)
    rng = np.random.default_rng(B + n)
    boards = rng.integers(-1, 2, size=(B, n, n)).astype(np.int64)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    bt = onets.boards_to_tensor(boards)
    with torch.no_grad():
        pi, v = onets.ttt_predict(p, bt, n)
        gpi, gv = onets.ttt_predict_with_gnn(p, q, bt, n)
    states = w.states_from_boards(boards)
    both = _lib.EVAL_STD | _lib.EVAL_GNN
    o3 = w.forward_states(states, both, precision=_lib.PREC_BF16X3)
    for got, want in ((o3["pi"], pi), (o3["v"], v), (o3["pi_gnn"], gpi), (o3["v_gnn"], gv)):
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=0, atol=1e-5)
    o1 = w.forward_states(states, both, precision=_lib.PREC_BF16)
    for got, want in ((o1["pi"], pi), (o1["v"], v), (o1["pi_gnn"], gpi), (o1["v_gnn"], gv)):
        assert np.abs(got.cpu().numpy() - want.numpy()).max() <= 5e-3
    std_only = w.forward_states(states, _lib.EVAL_STD, precision=_lib.PREC_BF16X3)
    assert torch.equal(std_only["pi"], o3["pi"])
    with torch.no_grad():
        w.nnet.fc2.bias.add_(0.5)
    w.weights_changed()
    o3b = w.forward_states(states, _lib.EVAL_STD, precision=_lib.PREC_BF16X3)
    assert not np.allclose(o3b["v"].cpu().numpy(), o3["v"].cpu().numpy())


def test_predict_batches_pipeline_equals_predict_batch():
    """the pipelined host-facing call returns, in order, exactly what predict_batch returns per batch"""
    w = _wrapper("c4", 7)
    rng = np.random.default_rng(0)
    host = [torch.from_numpy(rng.integers(-1, 2, size=(257, 7, 7)).astype(np.int8)).pin_memory() for _ in range(5)]
    want = [w.predict_batch(h) for h in host]
    got = []
    for res in w.predict_batches(iter(host)):
        got.append({k: v.clone().numpy() for k, v in res.items()})
    assert len(got) == len(want)
    for g, e in zip(got, want):
        for k in e:
            assert np.array_equal(g[k], e[k]), k


@pytest.mark.parametrize("kind,n", [("c4", 7), ("ttt", 4)])
def test_default_precision_is_the_tensor_core_path(kind, n):
    """a config without `b200_precision` (the reference's config.yaml) runs the guarded tensor-core path (Connect4: f16f8,
    TicTacToe: bf16x3 on random-init weights) and still reproduces the reference module's outputs within 1e-5"""
    game = (orules.Connect4Rules if kind == "c4" else orules.TicTacToeRules)(n)
    torch.manual_seed(0)
    W = B200Connect4GNNWrapper if kind == "c4" else B200TicTacToeGNNWrapper
    w = W(game, dotdict(dict(lr=1e-3, dropout=0.3, epochs=1, batch_size=64, gnn_layers=2, use_gnn=True)))
    assert w.precision == _lib.PREC_AUTO
    assert w.active_precision() == (_lib.PREC_F16F8 if kind == "c4" else _lib.PREC_BF16X3), w.precision_report
    rng = np.random.default_rng(n)
    boards = rng.integers(-1, 2, size=(200, n, n)).astype(np.int64)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    bt = onets.boards_to_tensor(boards)
    with torch.no_grad():
        pi, v = (onets.c4_predict if kind == "c4" else onets.ttt_predict)(p, bt, n)
        gpi, gv = (onets.c4_predict_with_gnn if kind == "c4" else onets.ttt_predict_with_gnn)(p, q, bt, n)
    out = w.predict_batch(boards)
    for got, want in ((out["pi"], pi), (out["v"], v), (out["pi_gnn"], gpi), (out["v_gnn"], gv)):
        np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=1e-5)
    one_pi, one_v = w.predict_with_gnn(boards[0])
    assert abs(one_v - gv[0].item()) <= 1e-5 and np.abs(one_pi - gpi[0].numpy()).max() <= 1e-5


@pytest.mark.parametrize("prec", ["f16f8", "f16f8ks", "bf16x3", "bf16x3ks", "bf16"])
def test_tensor_core_forward_is_run_to_run_identical(prec):
    """The fused trunk (conv1 slot queued inside the previous tile's conv2 k-blocks, one TMEM result region) and GEMM-1's
    side tile are synchronised by mbarriers only: repeated launches over many tiles per CTA (20,011 positions = 68
    tiles per SM, ragged tail) must be bit-identical, and a std-only call must equal the std half of the combined call."""
    w = _wrapper("c4", 7)
    rng = np.random.default_rng(5)
    states = w.states_from_boards(rng.integers(-1, 2, size=(20011, 7, 7)).astype(np.int8))
    both = _lib.EVAL_STD | _lib.EVAL_GNN
    p = _lib.PRECISIONS[prec]
    first = {k: v.clone() for k, v in w.forward_states(states, both, precision=p).items()}
    for _ in range(4):
        again = w.forward_states(states, both, precision=p)
        for k in first:
            assert torch.equal(first[k], again[k]), k
    only_std = w.forward_states(states, _lib.EVAL_STD, precision=p)
    assert torch.equal(only_std["pi"], first["pi"]) and torch.equal(only_std["v"], first["v"])
    # rows do not depend on their neighbours or on the batch they arrive in
    part = w.forward_states(states[1000:1300].contiguous(), both, precision=p)
    for k in first:
        assert torch.equal(part[k], first[k][1000:1300]), k


def test_full_size_batch_properties():
    """BASELINE configs[1] at its full size (65,536 positions, 7x7, std + GNN predictions, the bench's f16f8), checked through
    size-independent properties: a random sample of rows against the oracle at 1e-5; rows evaluated alone or in
    another order give bit-identical outputs (no dependence on tile position or neighbours); duplicated positions
    give duplicated outputs; policies are distributions, values lie in [-1, 1]."""
    w = _wrapper("c4", 7, b200_precision="f16f8")
    B = 65536
    rng = np.random.default_rng(65536)
    boards = rng.integers(-1, 2, size=(B, 7, 7)).astype(np.int8)
    boards[B // 2:B // 2 + 100] = boards[:100]  # duplicates
    both = _lib.EVAL_STD | _lib.EVAL_GNN
    states = w.states_from_boards(boards)
    full = {k: v.clone() for k, v in w.forward_states(states, both).items()}
    # (1) sample vs oracle
    idx = np.sort(rng.choice(B, size=384, replace=False))
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    bt = onets.boards_to_tensor(boards[idx].astype(np.int64))
    with torch.no_grad():
        spi, sv = onets.c4_predict(p, bt, 7)
        gpi, gv = onets.c4_predict_with_gnn(p, q, bt, 7)
    tidx = torch.from_numpy(idx).to(full["pi"].device)
    for name, want in (("pi", spi), ("v", sv), ("pi_gnn", gpi), ("v_gnn", gv)):
        got = full[name][tidx].cpu().numpy().reshape(want.shape)
        np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=1e-5, err_msg=name)
    # (2) the same rows evaluated alone, and the whole batch in another order
    alone = w.forward_states(states[tidx].contiguous(), both)
    for name in full:
        assert torch.equal(alone[name], full[name][tidx]), name
    perm = torch.from_numpy(rng.permutation(B)).to(states.device)
    shuffled = w.forward_states(states[perm].contiguous(), both)
    for name in full:
        assert torch.equal(shuffled[name], full[name][perm]), name
    # (3) duplicates, (4) ranges
    for name in full:
        assert torch.equal(full[name][B // 2:B // 2 + 100], full[name][:100]), name
    for name in ("pi", "pi_gnn"):
        s = full[name].sum(dim=1)
        assert float((s - 1).abs().max()) <= 1e-5 and float(full[name].min()) >= 0.0
    for name in ("v", "v_gnn"):
        assert float(full[name].abs().max()) <= 1.0


def test_search_mode_defaults():
    """Leaf evaluations made for the tree search (`forward_states(search=True)`, what BatchedMCTS calls) run the fastest
    tensor-core mode with the exact head fold by default -- unguarded, stated tolerance 5e-5 -- while the reference-facing calls
    keep the guard and both contractions; explicit settings are followed by both."""
    w = _wrapper("c4", 7, b200_precision="auto")
    assert w.search_precision() == _lib.PREC_F16F8 and w.fold_search and not w.fold_heads
    rng = np.random.default_rng(3)
    boards = rng.integers(-1, 2, size=(300, 7, 7)).astype(np.int64)
    states = w.states_from_boards(boards)
    p, q = _cpu_sd(w.nnet), _cpu_sd(w.gnn)
    with torch.no_grad():
        gpi, gv = onets.c4_predict_with_gnn(p, q, onets.boards_to_tensor(boards), 7)
    s = w.forward_states(states, _lib.EVAL_GNN, search=True)
    folded = w.forward_states(states, _lib.EVAL_GNN | _lib.EVAL_FOLD, precision=_lib.PREC_F16F8)
    assert torch.equal(s["pi_gnn"], folded["pi_gnn"]) and torch.equal(s["v_gnn"], folded["v_gnn"])
    assert np.abs(s["pi_gnn"].cpu().numpy() - gpi.numpy()).max() <= 1e-5 and np.abs(s["v_gnn"].cpu().numpy() - gv.numpy()).max() <= 1e-5
    api = w.forward_states(states, _lib.EVAL_GNN)
    plain = w.forward_states(states, _lib.EVAL_GNN, precision=w.active_precision())
    assert torch.equal(api["v_gnn"], plain["v_gnn"])
    for kw, prec, fold in ((dict(b200_precision="fp32"), _lib.PREC_FP32, True), (dict(b200_precision="bf16x3", b200_fold_heads=False), _lib.PREC_BF16X3, False),
                           (dict(b200_search_precision="bf16x3"), _lib.PREC_BF16X3, True)):
        w2 = _wrapper("c4", 5, **{**dict(b200_precision="auto"), **kw})
        assert w2.search_precision() == prec and w2.fold_search == fold
    assert _wrapper("ttt", 3, b200_precision="auto").search_precision() == _lib.PREC_BF16X3
