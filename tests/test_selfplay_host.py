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
