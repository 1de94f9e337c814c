"""Coach.executeEpisode (azgnn_b200/coach.py) on the host check arena reproduces the reference's
self-play episode (golden dump): same moves under the same NumPy seed and fake net, same example
tuples (Coach.py:27-79)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck"))
from host_arena import HostArena  # noqa: E402

from azgnn_b200 import games  # noqa: E402
from azgnn_b200.coach import Coach  # noqa: E402
from oracle.mcts import FakeNet  # noqa: E402
from helpers import dotdict, golden  # noqa: E402


@pytest.mark.parametrize("tag", ["c4_7_gnn", "c4_7_std", "c4_5_gnn", "ttt_3_gnn"])
def test_execute_episode_matches_reference(tag):
    g = golden("mcts_" + tag)
    kind, n = tag.split("_")[:2]
    n = int(n)
    game = games.Connect4Game(n) if kind == "c4" else games.TicTacToeGame(n)
    name = "connect4" if kind == "c4" else "tictactoe"
    A = game.getActionSize()
    args = dotdict(dict(numMCTSSims=int(g["numMCTSSims"]), cpuct=float(g["cpuct"]), use_gnn=bool(g["use_gnn"]),
                        expand_by=int(g["expand_by"]), tempThreshold=int(g["tempThreshold"])))
    sims = args.numMCTSSims + args.expand_by
    coach = Coach(game, FakeNet(A, salt=int(g["salt"]), spread=float(g["spread"])), args,
                  arena_factory=lambda: HostArena(name, n, 1, sims, args.cpuct, capacity=sims * (n * n + 1) + 64))
    np.random.seed(int(g["seed"]))
    std, gnn = coach.executeEpisode()
    moves = int(g["n_moves"])
    n_sym = len(game.getSymmetries(game.getInitBoard(), [0.0] * A))
    assert len(std) == moves * n_sym
    result = float(g["result"])  # getGameEnded for the player to move after the last move
    for i in range(moves):
        group = std[i * n_sym:(i + 1) * n_sym]
        assert any(np.array_equal(np.asarray(p, dtype=np.float64), g["pis"][i]) for _, p, _ in group)
        assert any(np.array_equal(np.asarray(b).reshape(-1), g["roots"][i]) for b, _, _ in group)
        # mover of position i vs the player to move at the end: players alternate
        same = ((moves - i) % 2 == 0)
        assert group[0][2] == result * (1 if same else -1)
    if args.use_gnn:
        assert len(gnn) == moves
        for i, (b, pl, ip, iv, ep, ev, r) in enumerate(gnn):
            rec = g["expand"][i]
            assert np.array_equal(ip, rec[:A]) and float(iv) == rec[A]
            assert np.array_equal(ep, rec[A + 1:2 * A + 1]) and float(np.asarray(ev)) == rec[2 * A + 1]
            assert pl == (1 if i % 2 == 0 else -1)
    else:
        assert gnn == []
