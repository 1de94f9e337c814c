"""Batched self-play driver (azgnn_b200/selfplay.py) on the host check arena with a fake net."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck"))
from host_arena import HostArena  # noqa: E402

from azgnn_b200 import games  # noqa: E402
from azgnn_b200.selfplay import BatchedSelfPlay, probs_from_counts, sample_actions, unpack_boards  # noqa: E402
from azgnn_b200.mcts import pack_states  # noqa: E402
from oracle.mcts import FakeNet  # noqa: E402
from helpers import dotdict  # noqa: E402


def test_probs_match_reference_arithmetic():
    """MCTS.py:46-57: [(x + EPS) ** (1./temp)], float(sum(.)), x / sum -- bit for bit."""
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 40, size=(500, 8))
    counts[counts.sum(1) == 0, 0] = 1
    got = probs_from_counts(counts, np.ones(500, dtype=np.int64), rng)
    for g in range(500):
        cs = [(int(x) + 1e-8) ** (1. / 1) for x in counts[g]]
        tot = float(sum(cs))
        assert [x / tot for x in cs] == list(got[g])
    greedy = probs_from_counts(counts, np.zeros(500, dtype=np.int64), rng)
    assert (greedy.sum(1) == 1).all()
    assert (counts[np.arange(500), greedy.argmax(1)] == counts.max(1)).all()


def test_sampling_is_numpy_choice():
    """Coach.py:63 np.random.choice(len(pi), p=pi): same index for the same uniform draw."""
    rng = np.random.default_rng(1)
    for seed in range(200):
        p = rng.random(8)
        p[rng.integers(0, 8)] = 0
        p /= p.sum()
        want = np.random.RandomState(seed).choice(8, p=p)
        u = np.random.RandomState(seed).random_sample()
        assert sample_actions(p[None], None, u=[u])[0] == want


def test_unpack_round_trip():
    rng = np.random.default_rng(2)
    b = rng.integers(-1, 2, size=(50, 7, 7)).astype(np.int64)
    assert np.array_equal(unpack_boards("connect4", 7, pack_states("connect4", b)), b)


@pytest.mark.parametrize("kind,n,use_gnn", [("connect4", 5, True), ("tictactoe", 3, False)])
def test_episodes_are_well_formed(kind, n, use_gnn):
    game = games.Connect4Game(n) if kind == "connect4" else games.TicTacToeGame(n)
    A = game.getActionSize()
    args = dotdict(dict(numMCTSSims=8, cpuct=1.0, use_gnn=use_gnn, expand_by=3, tempThreshold=4))
    G = 5
    arena = HostArena(kind, n, G, 11, 1.0, capacity=11 * (n * n + 2) + 16)
    sp = BatchedSelfPlay(game, FakeNet(A, salt=3), args, G, seed=7, arena=arena)
    eps = sp.play(8)
    assert len(eps) == 8 and sp.episodes_done >= 8
    n_sym = len(game.getSymmetries(game.getInitBoard(), [0] * A))
    for std, gnn in eps:
        assert len(std) % n_sym == 0 and len(std) > 0
        plies = len(std) // n_sym
        assert plies <= n * n
        for b, p, r in std:
            assert abs(sum(p) - 1) < 1e-6 and abs(r) in (1, 1e-4)
            assert np.asarray(b).shape == (n, n)
        # first stored position is the empty board seen by player 1; the last mover's sign is +-r
        assert not np.asarray(std[0][0]).any()
        if use_gnn:
            assert len(gnn) == plies
            for b, pl, ip, iv, ep, ev, r in gnn:
                assert abs(ip.sum() - 1) < 1e-9 and abs(ep.sum() - 1) < 1e-9 and isinstance(iv, np.float32)
        # results alternate sign with the player to move (Coach.py:70)
        rs = [std[i * n_sym][2] for i in range(plies)]
        if abs(rs[0]) == 1:
            assert all(rs[i] == -rs[i + 1] for i in range(plies - 1))


def test_vectorised_expanded_value_has_the_scalar_loops_values_and_types():
    """expand_tree's value (MCTS.py:132-143) for all games at once vs the reference's scalar loop, including roots
    whose edges mix Python-number Q values (pure-terminal subtrees) with np.float32 ones (NEP 50 promotion order)."""
    from azgnn_b200 import _lib
    from azgnn_b200.mcts import expanded_value_scalar, expanded_values, typed_value
    rng = np.random.default_rng(0)
    G, A = 4000, 8
    N = rng.integers(0, 6, size=(G, A))
    T = rng.choice([_lib.TAG_NONE, _lib.TAG_F32, _lib.TAG_F32, _lib.TAG_PYFLOAT, _lib.TAG_PYINT], size=(G, A)).astype(np.int8)
    Q = np.where(T == _lib.TAG_F32, rng.uniform(-1, 1, (G, A)).astype(np.float32).astype(np.float64),
                 np.where(T == _lib.TAG_PYINT, rng.choice([-1.0, 1.0, 0.0], size=(G, A)), rng.choice([1e-4, -1e-4, 0.5], size=(G, A))))
    N[:50] = 0  # no visited edge: the network's root value is returned
    v0 = rng.uniform(-1, 1, G).astype(np.float32)
    val, tag = expanded_values(N, Q, T, v0)
    for g in range(G):
        want = expanded_value_scalar(N[g], Q[g], T[g], v0[g])
        got = typed_value(val[g], int(tag[g]))
        assert type(got) is type(want), (g, type(got), type(want))
        assert got == want, (g, got, want)


class _FakeDeviceNet:
    """FakeNet behind the batched `forward_states` interface of the CUDA wrappers (CPU tensors, with the device-count
    contract of the compacted leaf batches), so that the device-evaluation code paths of BatchedMCTS / BatchedSelfPlay
    -- searches queued without a status check, expand_tree launched before the host turns counts into policies, the
    next move's searches queued before this move's records -- run on the host check arena."""
    def __init__(self, kind, n, A, salt, dynamic):
        self.kind, self.n, self.A = kind, n, A
        self.net = FakeNet(A, salt=salt)
        self.supports_dynamic_count = dynamic
        self.calls = 0

    def predict(self, board):
        return self.net.predict(board)

    def predict_with_gnn(self, board):
        return self.net.predict_with_gnn(board)

    def forward_states(self, states, eval_mask=None, precision=None, count=None, out=None):
        import torch
        from azgnn_b200 import _lib
        self.calls += 1
        B = int(states.shape[0])
        live = B if count is None else int(count.reshape(-1)[0])
        boards = unpack_boards(self.kind, self.n, states.numpy()[:live])
        o = {}
        for bit, names, fn in ((_lib.EVAL_STD, ("pi", "v"), self.net.predict), (_lib.EVAL_GNN, ("pi_gnn", "v_gnn"), self.net.predict_with_gnn)):
            if eval_mask & bit:
                pi = np.full((B, self.A), np.nan, dtype=np.float32)  # rows beyond the count must never be read
                v = np.full(B, np.nan, dtype=np.float32)
                for i, b in enumerate(boards):
                    pi[i], v[i] = fn(b)
                o[names[0]], o[names[1]] = torch.from_numpy(pi), torch.from_numpy(v)
        return o


def _play_fixed_moves(kind, n, net, moves, collect, G=6, seed=11):
    game = games.Connect4Game(n) if kind == "connect4" else games.TicTacToeGame(n)
    args = dotdict(dict(numMCTSSims=7, cpuct=1.0, use_gnn=True, expand_by=3, tempThreshold=3))
    arena = HostArena(kind, n, G, 10, 1.0, capacity=10 * (n * n + 2) + 16)
    sp = BatchedSelfPlay(game, net, args, G, seed=seed, collect_examples=collect, arena=arena)
    out = []
    for _ in range(moves):
        out.append(sp.step_all())
    return sp, out


def _same_examples(a, b):
    assert len(a) == len(b)
    for (s1, g1), (s2, g2) in zip(a, b):
        assert len(s1) == len(s2) and len(g1) == len(g2)
        for (b1, p1, r1), (b2, p2, r2) in zip(s1, s2):
            assert np.array_equal(b1, b2) and list(p1) == list(p2) and r1 == r2 and type(r1) is type(r2)
        for x, y in zip(g1, g2):
            assert np.array_equal(x[0], y[0]) and x[1] == y[1] and np.array_equal(x[2], y[2]) and x[3] == y[3]
            assert np.array_equal(x[4], y[4]) and x[5] == y[5] and type(x[5]) is type(y[5]) and x[6] == y[6]


@pytest.mark.parametrize("kind,n", [("connect4", 5), ("tictactoe", 3)])
@pytest.mark.parametrize("dynamic", [False, True])
def test_pipelined_device_loop_equals_the_sequential_host_loop(kind, n, dynamic):
    """The move loop with batched ("device") evaluation -- pipelined against the host -- produces, move by move, the
    episodes of the sequential per-leaf loop: same examples, same types, same random stream."""
    A = n + 1 if kind == "connect4" else n * n + 1
    sp_h, host = _play_fixed_moves(kind, n, FakeNet(A, salt=5), 30, True)
    sp_d, dev = _play_fixed_moves(kind, n, _FakeDeviceNet(kind, n, A, 5, dynamic), 30, True)
    assert sp_d.mcts.device_eval and sp_d.mcts.compact == dynamic and not sp_h.mcts.device_eval
    assert sum(len(x) for x in host) >= 2  # episodes did finish (and restart) inside the window
    for h, d in zip(host, dev):
        _same_examples(h, d)
    assert sp_h.moves_played == sp_d.moves_played and sp_h.episodes_done == sp_d.episodes_done
    assert np.array_equal(sp_h.step, sp_d.step) and np.array_equal(sp_h.player, sp_d.player)
    # throughput mode (no examples): the same games are played
    sp_t, _ = _play_fixed_moves(kind, n, _FakeDeviceNet(kind, n, A, 5, dynamic), 30, False)
    assert sp_t.episodes_done == sp_h.episodes_done and np.array_equal(sp_t.step, sp_h.step)
    roots_t = sp_t.mcts.arena.to_host(sp_t.mcts.arena.get_roots())
    roots_h = sp_h.mcts.arena.to_host(sp_h.mcts.arena.get_roots())
    assert np.array_equal(roots_t, roots_h)


def test_expand_tree_in_two_halves_equals_expand_tree_arrays():
    """expand_tree_launch + expand_tree_readback + expand_tree_records (the split used by the pipelined loop) ==
    expand_tree_arrays, outputs and arena statistics alike."""
    from azgnn_b200.mcts import BatchedMCTS
    kind, n, G = "connect4", 5, 4
    game = games.Connect4Game(n)
    A = game.getActionSize()
    args = dotdict(dict(numMCTSSims=9, cpuct=1.0, use_gnn=True, expand_by=4))
    res = []
    for split in (False, True):
        arena = HostArena(kind, n, G, 13, 1.0, capacity=13 * (n * n + 2) + 16)
        bm = BatchedMCTS(game, _FakeDeviceNet(kind, n, A, 9, True), args, n_games=G, arena=arena)
        bm.set_root_boards([game.getInitBoard()] * G)
        bm.search(9)
        N0, _, _ = bm.root_stats()
        if split:
            pending = bm.expand_tree_launch(4, N0)
            assert pending is not None
            recs = bm.expand_tree_records(bm.expand_tree_readback(pending, 4))
        else:
            recs = bm.expand_tree_arrays(4)
        res.append((recs, bm.root_stats()))
    for x, y in zip(res[0][0], res[1][0]):
        assert np.array_equal(x, y) and x.dtype == y.dtype
    for x, y in zip(res[0][1], res[1][1]):
        assert np.array_equal(x, y)
