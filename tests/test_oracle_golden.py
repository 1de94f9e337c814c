"""Pin the oracle (oracle/) to the reference's own outputs (tests/golden/*.npz).

The golden files were produced by tests/golden/make_golden.py importing the unmodified
reference.  These tests run anywhere (CPU only, no /root/reference needed).
"""
import numpy as np
import pytest
import torch

from oracle import nets as onets
from oracle import rules as orules
from oracle.mcts import FakeNet, OracleMCTS

from helpers import (assert_tables_equal, checksum_rows, dotdict, golden, golden_as_tables, mcts_as_tables,
                     sample_index, seeded_fl_module, seeded_two_player_modules)

torch.set_num_threads(4)


# ------------------------------------------------------------------------------------ rules
def _game(name):
    kind, n = name.split("_")
    n = int(n)
    return {"c4": orules.Connect4Rules, "ttt": orules.TicTacToeRules, "fl": orules.FrozenLakeRules}[kind](n)


@pytest.mark.parametrize("name", ["c4_7", "c4_5", "c4_4", "ttt_3", "ttt_4", "fl_4", "fl_8"])
def test_rules_match_reference(name):
    g = golden("rules_" + name)
    game = _game(name)
    is_fl = name.startswith("fl")
    for i in range(g["boards"].shape[0]):
        b = g["boards"][i] if is_fl else g["boards"][i].astype(np.int64)
        player = int(g["players"][i])
        valids = np.asarray(game.getValidMoves(b, player))
        assert np.array_equal(valids.astype(np.int64), g["valids"][i])
        e = game.getGameEnded(b, player)
        assert float(e) == g["ended"][i]
        assert np.array_equal(np.asarray(game.getCanonicalForm(b, player), dtype=np.float64), g["canonical"][i])
        for a in range(game.getActionSize()):
            if valids[a] and e == 0:
                nb, npl = game.getNextState(np.array(b, copy=True), player, a)
                assert np.array_equal(np.asarray(nb, dtype=np.float64), g["next"][i, a])
                assert npl == g["next_player"][i, a]


def test_rules_types_and_keys():
    c4 = orules.Connect4Rules(7)
    b = c4.getInitBoard()
    assert b.dtype == np.int64 and len(c4.stringRepresentation(b)) == 392
    assert c4.getGameEnded(b, 1) == 0 and isinstance(c4.getGameEnded(b, 1), int)
    fl = orules.FrozenLakeRules(4)
    assert fl.stringRepresentation(fl.getInitBoard()) == "0,0"
    assert fl.getValidMoves(fl.getInitBoard(), 1).dtype == np.int8


# ------------------------------------------------------------------------------------ nets
@pytest.mark.parametrize("kind,n", [("c4", 5), ("c4", 7), ("ttt", 3), ("ttt", 4)])
def test_two_player_nets_match_reference(kind, n):
    g = golden(f"nets_{kind}_{n}")
    nnet, gnn = seeded_two_player_modules(kind, n)
    # identical random-init weights as the reference classes under the same seed
    names, rows = checksum_rows(nnet.state_dict())
    assert list(names) == list(g["nnet_names"]) and np.array_equal(rows, g["nnet_checksum"])
    names, rows = checksum_rows(gnn.state_dict())
    assert list(names) == list(g["gnn_names"]) and np.array_equal(rows, g["gnn_checksum"])
    p, q = dict(nnet.state_dict()), dict(gnn.state_dict())
    boards = onets.boards_to_tensor(g["boards"])
    with torch.no_grad():
        if kind == "c4":
            pi, v = onets.c4_predict(p, boards, n)
            gpi, gv = onets.c4_predict_with_gnn(p, q, boards, n)
        else:
            pi, v = onets.ttt_predict(p, boards, n)
            gpi, gv = onets.ttt_predict_with_gnn(p, q, boards, n)
    tol = dict(rtol=0, atol=2e-6)  # batched vs B=1 GEMM blocking only
    np.testing.assert_allclose(pi.numpy(), g["pi"], **tol)
    np.testing.assert_allclose(v.numpy(), g["v"], **tol)
    np.testing.assert_allclose(gpi.numpy(), g["gnn_pi"], **tol)
    np.testing.assert_allclose(gv.numpy(), g["gnn_v"], **tol)


@pytest.mark.parametrize("kind,n", [("c4", 5), ("c4", 7), ("ttt", 3), ("ttt", 4)])
def test_training_step_matches_reference(kind, n):
    """Std step and GNN step (B>1: GNNLayer couples every row to row 0, gnn_utils.py:38-74):
    losses, outputs and gradient summaries vs the reference's autograd."""
    g = golden(f"nets_{kind}_{n}")
    nnet, gnn = seeded_two_player_modules(kind, n)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in nnet.state_dict().items()}
    q = {k: v.detach().clone().requires_grad_(True) for k, v in gnn.state_dict().items()}
    boards = onets.boards_to_tensor(g["train_boards"])
    tpi, tv = torch.tensor(g["train_pi"]), torch.tensor(g["train_v"])
    feats_fn = (lambda: onets.c4_features(p, boards, n)) if kind == "c4" else (lambda: onets.ttt_features(p, boards, n))
    heads_fn = onets.c4_heads if kind == "c4" else onets.ttt_heads

    lp, v = heads_fn(p, feats_fn())
    loss = onets.policy_value_loss(lp, v, tpi, tv)
    assert abs(loss.item() - float(g["std_loss"])) < 1e-5
    np.testing.assert_allclose(lp.detach().numpy(), g["std_out_logpi"], rtol=0, atol=5e-6)
    loss.backward()
    for name, row, samp in zip(g["std_grad_names"], g["std_grad_rows"], g["std_grad_samples"]):
        gr = p[str(name)].grad.double().flatten()
        np.testing.assert_allclose(gr[sample_index(gr.numel(), 64)].numpy(), samp, rtol=1e-4, atol=1e-6)
        assert abs(gr.norm().item() - row[2]) <= 1e-4 * max(1.0, row[2])
    for t in p.values():
        t.grad = None

    enh = onets.gnn_forward(q, feats_fn(), 2)
    idx = sample_index(enh.shape[1], 32)
    np.testing.assert_allclose(enh[0, idx].detach().numpy(), g["gnn_enh_row0"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(enh[1, idx].detach().numpy(), g["gnn_enh_row1"], rtol=0, atol=1e-5)
    lp, v = heads_fn(p, enh)
    loss = onets.policy_value_loss(lp, v, tpi, tv)
    assert abs(loss.item() - float(g["gnn_loss"])) < 1e-5
    loss.backward()
    for name, row, samp in zip(g["gnn_grad_names"], g["gnn_grad_rows"], g["gnn_grad_samples"]):
        gr = q[str(name)].grad.double().flatten()
        np.testing.assert_allclose(gr[sample_index(gr.numel(), 64)].numpy(), samp, rtol=1e-3, atol=2e-6)
        assert abs(gr.norm().item() - row[2]) <= 1e-3 * max(1e-3, row[2])


@pytest.mark.parametrize("n,layers", [(4, 2), (4, 3), (8, 2), (8, 3)])
def test_frozenlake_net_matches_reference(n, layers):
    g = golden(f"nets_fl_{n}_L{layers}")
    net = seeded_fl_module(n, layers)
    names, rows = checksum_rows(net.state_dict())
    assert list(names) == list(g["nnet_names"]) and np.array_equal(rows, g["nnet_checksum"])
    p = dict(net.state_dict())
    with torch.no_grad():
        for i, cell in enumerate(g["cells"]):
            pi, v = onets.fl_predict_cell(p, int(cell), n, layers)
            np.testing.assert_allclose(pi.numpy(), g["pi"][i], rtol=0, atol=1e-6)
            np.testing.assert_allclose(v.numpy(), g["v"][i], rtol=0, atol=1e-6)


# ------------------------------------------------------------------------------------ mcts
def _episode_game(tag):
    kind, n = tag.split("_")[:2]
    return _game(f"{kind}_{n}"), int(n)


@pytest.mark.parametrize("tag", ["c4_7_gnn", "c4_7_std", "c4_7_wide", "c4_5_gnn", "ttt_3_gnn", "ttt_4_std"])
def test_mcts_episode_bit_exact(tag):
    """Replay the reference's self-play episode (Coach.py:27-79 loop) through the oracle MCTS under
    the same fake net and NumPy seed: visit distributions, chosen actions, expand_tree outputs and
    the full Qsa/Nsa/Ns/Ps/Es/Vs dicts must be identical, bit for bit, value types included."""
    g = golden("mcts_" + tag)
    game, n = _episode_game(tag)
    A = game.getActionSize()
    args = dotdict(dict(numMCTSSims=int(g["numMCTSSims"]), cpuct=float(g["cpuct"]), use_gnn=bool(g["use_gnn"]),
                        expand_by=int(g["expand_by"]), tempThreshold=int(g["tempThreshold"])))
    net = FakeNet(A, salt=int(g["salt"]), spread=float(g["spread"]))
    m = OracleMCTS(game, net, args)
    np.random.seed(int(g["seed"]))
    board, player = game.getInitBoard(), 1
    dump_steps = set(int(x) for x in g["dump_steps"])
    for step in range(1, int(g["n_moves"]) + 1):
        canon = game.getCanonicalForm(board, player)
        assert np.array_equal(np.asarray(canon, dtype=np.int8).reshape(-1), g["roots"][step - 1])
        pi = m.getActionProb(canon, temp=int(step < args.tempThreshold))
        assert np.array_equal(np.asarray(pi, dtype=np.float64), g["pis"][step - 1])
        if args.use_gnn:
            (ip, iv, ep, ev), = m.expand_tree(canon, expand_by=args.expand_by).values()
            rec = g["expand"][step - 1]
            assert np.array_equal(ip, rec[:A]) and float(iv) == rec[A]
            assert np.array_equal(ep, rec[A + 1:2 * A + 1]) and float(np.asarray(ev)) == rec[2 * A + 1]
        action = np.random.choice(len(pi), p=pi)
        assert action == g["actions"][step - 1]
        if step in dump_steps:
            assert_tables_equal(mcts_as_tables(m, n, A), golden_as_tables(g, f"m{step}_", A))
        board, player = game.getNextState(board, player, action)
    assert float(game.getGameEnded(board, player)) == float(g["result"])
    assert net.calls == int(g["leaf_calls"])


def test_mcts_known_answer_position():
    """TicTacToe, X to move with an immediate win at a=2, uniform priors, 400 sims.  Q(root,2) is
    -1: the reference stores the child mover's perspective un-negated (MCTS.py:228-238)."""
    g = golden("mcts_ttt3_known_answer")
    game = orules.TicTacToeRules(3)

    class Uniform:
        def predict(self, b):
            return np.full(10, 0.1, dtype=np.float32), np.float32(0.0)
        predict_with_gnn = predict
    m = OracleMCTS(game, Uniform(), dotdict(dict(numMCTSSims=400, cpuct=1.0, use_gnn=False)))
    root = g["root"].astype(np.int64).reshape(3, 3)
    pi = m.getActionProb(root, temp=1)
    assert np.array_equal(np.asarray(pi), g["pi"])
    assert_tables_equal(mcts_as_tables(m, 3, 10), golden_as_tables(g, "m1_", 10))
    s = game.stringRepresentation(root)
    assert m.Qsa[(s, 2)] == -1


@pytest.mark.parametrize("n", [4, 8])
def test_mcts_frozenlake_until_first_cycle(n):
    """The reference FrozenLake search never terminates once a simulation cycles;
    parity is per-simulation dict equality for every simulation the reference completed."""
    g = golden(f"mcts_fl_{n}")
    game = orules.FrozenLakeRules(n)
    net = FakeNet(4, salt=int(g["salt"]), v_as_array=True)
    m = OracleMCTS(game, net, dotdict(dict(numMCTSSims=50, cpuct=float(g["cpuct"]), use_gnn=False)))
    b = game.getInitBoard()
    for i in range(1, int(g["n_completed"]) + 1):
        m.search(b)
        assert_tables_equal(mcts_as_tables(m, n, 4), golden_as_tables(g, f"m{i}_", 4))
    # with the documented depth cap the search terminates and every simulation returns
    m2 = OracleMCTS(game, FakeNet(4, salt=5, v_as_array=True), dotdict(dict(numMCTSSims=50, cpuct=2.0, use_gnn=False)),
                    max_depth=4 * n * n)
    pi = m2.getActionProb(b, temp=1)
    assert abs(sum(pi) - 1) < 1e-9
