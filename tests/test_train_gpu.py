"""K3 parity on the GPU: forward/backward of the training step in libazgnn_b200.so vs the
reference's autograd (golden gradient summaries) and vs the oracle graph on random batches.
Tolerances (north_star: 1e-5 in fp32): losses/outputs 1e-5; every gradient entry within 1e-5 ABSOLUTE of the reference's
(asserted and printed per parameter), and additionally within rtol 1e-3 / atol 2e-6 so that small gradients are held
relatively too.  The one documented exception: a hidden unit whose pre-activation is within rounding of 0 may land on the
other side of a ReLU than in the reference's summation order, which flips one row of a weight gradient -- at most two such
rows per parameter are tolerated in the full-tensor comparison, none in the golden samples."""
import numpy as np
import pytest
import torch

from azgnn_b200 import games, training
from azgnn_b200.nets import B200Connect4GNNWrapper, B200TicTacToeGNNWrapper
from helpers import dotdict, golden, sample_index
from train_helpers import OracleOps

pytestmark = pytest.mark.gpu
# the comparison graph runs on torch's CUDA ops: keep them in true fp32 (cuDNN convs default to TF32)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
GRAD_ATOL = 1e-5


def _check_samples(tag, name, got, want):
    err = float(np.abs(got - want).max())
    print(f"{tag} {name}: max |d grad| on the golden samples = {err:.2e} (max |grad| {float(np.abs(want).max()):.2e})")
    assert err <= GRAD_ATOL, (tag, name, err)
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-6)


def _wrapper(kind, n, dropout=0.0, **kw):
    game = games.Connect4Game(n) if kind == "c4" else games.TicTacToeGame(n)
    args = dotdict(dict(dict(lr=1e-3, dropout=dropout, epochs=2, batch_size=64, gnn_layers=2, use_gnn=True), **kw))
    torch.manual_seed(0)
    return (B200Connect4GNNWrapper if kind == "c4" else B200TicTacToeGNNWrapper)(game, args)


def _zero(w):
    for p in list(w.nnet.parameters()) + list(w.gnn.parameters()):
        p.grad = None


@pytest.mark.parametrize("kind,n", [("c4", 5), ("c4", 7), ("ttt", 3), ("ttt", 4)])
def test_gradients_match_reference_golden(kind, n):
    g = golden(f"nets_{kind}_{n}")
    w = _wrapper(kind, n)
    dev = w.device
    boards = torch.FloatTensor(g["train_boards"].astype(np.float64)).to(dev)
    tpi, tv = torch.tensor(g["train_pi"]).to(dev), torch.tensor(g["train_v"]).to(dev)
    _zero(w)
    loss = training.std_step(training.CudaOps, w, boards, tpi, tv)
    assert abs(loss.item() - float(g["std_loss"])) < 1e-5
    loss.backward()
    named = dict(w.nnet.named_parameters())
    for name, row, samp in zip(g["std_grad_names"], g["std_grad_rows"], g["std_grad_samples"]):
        gr = named[str(name)].grad.double().flatten().cpu()
        _check_samples(f"{kind}{n} std", str(name), gr[sample_index(gr.numel(), 64)].numpy(), samp)
        assert abs(gr.norm().item() - row[2]) <= 1e-3 * max(1e-3, row[2])
    _zero(w)
    loss = training.gnn_step(training.CudaOps, w, boards, tpi, tv)
    assert abs(loss.item() - float(g["gnn_loss"])) < 1e-5
    loss.backward()
    named = dict(w.gnn.named_parameters())
    for name, row, samp in zip(g["gnn_grad_names"], g["gnn_grad_rows"], g["gnn_grad_samples"]):
        gr = named[str(name)].grad.double().flatten().cpu()
        _check_samples(f"{kind}{n} gnn", str(name), gr[sample_index(gr.numel(), 64)].numpy(), samp)
        assert abs(gr.norm().item() - row[2]) <= 1e-3 * max(1e-3, row[2])
    assert all(p.grad is None for p in w.nnet.parameters())  # the GNN step leaves the trunk alone


@pytest.mark.parametrize("kind,n,B", [("c4", 5, 2), ("c4", 5, 33), ("c4", 6, 64), ("ttt", 3, 17), ("ttt", 5, 64)])
def test_gradients_match_oracle_graph(kind, n, B):
    w = _wrapper(kind, n)
    dev = w.device
    rng = np.random.default_rng(B * n)
    A = w.action_size
    boards = torch.FloatTensor(rng.integers(-1, 2, size=(B, n, n)).astype(np.float64)).to(dev)
    tpi = torch.FloatTensor(rng.dirichlet(np.ones(A), size=B)).to(dev)
    tv = torch.FloatTensor(rng.uniform(-1, 1, B)).to(dev)
    for step, params in ((training.std_step, lambda: w.nnet), (training.gnn_step, lambda: w.gnn)):
        _zero(w)
        l_cuda = step(training.CudaOps, w, boards, tpi, tv)
        l_cuda.backward()
        got = {k: p.grad.clone() for k, p in params().named_parameters()}
        _zero(w)
        l_ref = step(OracleOps, w, boards, tpi, tv)  # same graph on torch's own CUDA ops
        l_ref.backward()
        assert abs(l_cuda.item() - l_ref.item()) < 1e-5
        for k, p in params().named_parameters():
            ref = p.grad
            diff = (got[k] - ref).abs()
            bad = (diff > 1e-3 * ref.abs().max() + 2e-6) | (diff > GRAD_ATOL)
            # A hidden unit whose pre-activation is within rounding of 0 can land on either side of the
            # ReLU in the two implementations; that flips one row of the weight gradient (and one bias
            # entry).  Anything beyond two such rows is a real mismatch.
            row_bad = bad.reshape(bad.shape[0], -1).any(dim=1)
            rows = row_bad.sum().item()
            rest = diff.reshape(bad.shape[0], -1)[~row_bad]
            print(f"{kind}{n} B={B} {step.__name__} {k}: max |d grad| = {float(rest.max()) if rest.numel() else 0.0:.2e} over {int((~row_bad).sum())} rows "
                  f"({rows} ReLU-boundary rows excluded), max |grad| {float(ref.abs().max()):.2e}")
            assert rows <= 2, (k, rows, diff.max().item(), ref.abs().max().item())


def test_gnn_layers_are_identity_at_batch_one():
    w = _wrapper("c4", 5)
    b = torch.zeros(1, 5, 5, device=w.device)
    feats = training.trunk_features(training.CudaOps, w, b, training=False)
    enh = training.gnn_enhance(training.CudaOps, w.gnn, feats)
    ref = training.gnn_enhance(OracleOps, w.gnn, feats)
    assert torch.allclose(enh, ref, atol=1e-5)


def test_train_runs_and_learns():
    """Wrapper.train with the reference's example formats: repeated minibatches reduce the losses."""
    w = _wrapper("c4", 5, dropout=0.3, epochs=30, batch_size=16)
    rng = np.random.default_rng(0)
    A = w.action_size
    boards = rng.integers(-1, 2, size=(16, 5, 5)).astype(np.int64)
    pis = rng.dirichlet(np.ones(A), size=16)
    vs = rng.choice([-1, 1], size=16)
    examples = [(boards[i], list(pis[i]), int(vs[i])) for i in range(16)]
    gnn_examples = [(boards[i], 1, pis[i], np.float32(0.1), pis[i], np.float32(vs[i] * 0.5), int(vs[i])) for i in range(16)]
    bt = torch.FloatTensor(boards.astype(np.float64)).to(w.device)
    tpi, tv = torch.FloatTensor(pis).to(w.device), torch.FloatTensor(vs.astype(np.float64)).to(w.device)

    def losses():
        w.nnet.dropout, keep = 0.0, w.nnet.dropout
        with torch.no_grad():
            a = training.std_step(training.CudaOps, w, bt, tpi, tv).item()
            b = training.gnn_step(training.CudaOps, w, bt, tpi, tv * 0.5).item()
        w.nnet.dropout = keep
        return a, b
    before = losses()
    out0 = w.predict_batch(boards)
    np.random.seed(0)
    w.train(examples, gnn_examples)
    after = losses()
    assert after[0] < before[0] and after[1] < before[1], (before, after)
    assert not np.allclose(out0["pi_gnn"], w.predict_batch(boards)["pi_gnn"])


@pytest.mark.parametrize("kind,n", [("c4", 5), ("ttt", 3)])
def test_captured_steps_equal_eager_steps(kind, n, monkeypatch):
    """train() replays CUDA-graph captures of its optimizer steps (Adam state zeroed per call = the reference's
    per-call optimizers).  Two calls of train() must leave the same weights as the eager loop (dropout off)."""
    rng = np.random.default_rng(3)
    ws = {}
    for mode in ("eager", "graph"):
        if mode == "eager":
            monkeypatch.setenv("AZG_TRAIN_EAGER", "1")
        else:
            monkeypatch.delenv("AZG_TRAIN_EAGER", raising=False)
        w = _wrapper(kind, n, dropout=0.0, epochs=4, batch_size=8)
        A = w.action_size
        r = np.random.default_rng(3)
        boards = r.integers(-1, 2, size=(24, n, n)).astype(np.int64)
        pis = r.dirichlet(np.ones(A), size=24)
        vs = r.choice([-1, 1], size=24)
        examples = [(boards[i], list(pis[i]), int(vs[i])) for i in range(24)]
        gnn_examples = [(boards[i], 1, pis[i], np.float32(0.1), pis[i], np.float32(vs[i] * 0.5), int(vs[i])) for i in range(24)]
        np.random.seed(5)
        w.train(examples, gnn_examples)
        w.train(examples[:6], gnn_examples[:6])  # a second call, and a different minibatch shape (6 < batch_size)
        w.train(examples, gnn_examples)
        ws[mode] = w
    for mod in ("nnet", "gnn"):
        a, b = dict(getattr(ws["eager"], mod).named_parameters()), dict(getattr(ws["graph"], mod).named_parameters())
        for k in a:
            assert torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-6), (mod, k, (a[k] - b[k]).abs().max().item())
    o1, o2 = ws["eager"].predict_batch(boards), ws["graph"].predict_batch(boards)
    assert np.allclose(o1["pi_gnn"], o2["pi_gnn"], atol=1e-5) and np.allclose(o1["v"], o2["v"], atol=1e-5)


# ------------------------------------------------------------------------------------ FrozenLake training
@pytest.mark.parametrize("n,layers", [(4, 3), (8, 2)])
def test_frozenlake_training_step(n, layers):
    """One minibatch of FrozenLakeNet.train (FrozenLakeNet.py:105-166): loss, outputs and every parameter
    gradient vs the reference's autograd (golden), and vs the oracle graph on the same GPU."""
    from azgnn_b200.nets import B200FrozenLakeNet
    g = golden(f"train_fl_{n}_L{layers}")
    torch.manual_seed(0)
    w = B200FrozenLakeNet(games.FrozenLakeGame(n), dotdict(dict(lr=1e-3, epochs=2, batch_size=32, embedding_dim=128, gnn_layers=layers)))
    states = torch.zeros(len(g["cells"]), 2, dtype=torch.int64, device=w.device)
    states[:, 0] = torch.as_tensor(g["cells"]).to(w.device)
    tpi, tv = torch.tensor(g["train_pi"]).to(w.device), torch.tensor(g["train_v"]).to(w.device)
    for p in w.nnet.parameters():
        p.grad = None
    loss = training.fl_step(training.CudaOps, w, states, tpi, tv)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    loss.backward()
    named = dict(w.nnet.named_parameters())
    for name, row, samp in zip(g["grad_names"], g["grad_rows"], g["grad_samples"]):
        gr = named[str(name)].grad.double().flatten().cpu()
        _check_samples(f"fl{n} L{layers}", str(name), gr[sample_index(gr.numel(), 64)].numpy(), samp)
        assert abs(gr.norm().item() - row[2]) <= 1e-3 * max(1e-3, row[2])
    got = {k: p.grad.clone() for k, p in w.nnet.named_parameters()}
    for p in w.nnet.parameters():
        p.grad = None
    l2 = training.fl_step(OracleOps, w, states, tpi, tv)
    l2.backward()
    assert abs(l2.item() - loss.item()) < 1e-5
    for k, p in w.nnet.named_parameters():
        bad = (got[k] - p.grad).abs() > 1e-3 * p.grad.abs().max() + 2e-6
        assert bad.reshape(bad.shape[0], -1).any(dim=1).sum().item() <= 2, k


def test_frozenlake_train_runs():
    from azgnn_b200.nets import B200FrozenLakeNet
    game = games.FrozenLakeGame(4)
    torch.manual_seed(0)
    w = B200FrozenLakeNet(game, dotdict(dict(lr=1e-2, epochs=15, batch_size=8, embedding_dim=128, gnn_layers=2)))
    rng = np.random.default_rng(0)
    examples = []
    for _ in range(16):
        cell = int(rng.choice([0, 1, 2, 4, 6, 8, 9, 10, 13, 14]))
        b = np.zeros((4, 4)); b[cell // 4, cell % 4] = 1
        pi = np.zeros(4); pi[2] = 1.0
        examples.append((b, list(pi), 1.0))
    before = w.predict(examples[0][0])
    np.random.seed(0)
    w.train(examples)
    after = w.predict(examples[0][0])
    assert after[0][2] > before[0][2] and after[1][0] > before[1][0]  # learns "down" and a positive value
