"""BatchedArena on the host check arena (CPU): all games advance through the arena's own rules (packed states,
`advance`), and the results equal the sequential Arena loop with fresh trees per game, move for move (the tie-break
among equally visited moves is pinned to the lowest index in both runs)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck"))
from host_arena import HostArena  # noqa: E402

from azgnn_b200 import games  # noqa: E402
from azgnn_b200.mcts import MCTS  # noqa: E402
from azgnn_b200.pit import BatchedArena  # noqa: E402
from oracle.mcts import FakeNet  # noqa: E402
from helpers import dotdict  # noqa: E402


@pytest.mark.parametrize("kind,n", [("connect4", 5), ("tictactoe", 3), ("tictactoe", 4)])
def test_batched_arena_equals_sequential_games(kind, n, monkeypatch):
    monkeypatch.setattr(np.random, "choice", lambda a, *args, **kw: np.asarray(a).reshape(-1)[0])
    game = games.Connect4Game(n) if kind == "connect4" else games.TicTacToeGame(n)
    A = game.getActionSize()
    args = dotdict(dict(numMCTSSims=8, cpuct=1.0, use_gnn=False))
    na, nb = FakeNet(A, salt=1), FakeNet(A, salt=2)

    def arena(G):
        return HostArena(kind, n, G, 8, 1.0, capacity=4096)
    got = BatchedArena(game, na, nb, args, arena_factory=arena).playGames(6)
    one = two = draws = 0
    for grp in range(2):
        for _ in range(3):
            first, second = (na, nb) if grp == 0 else (nb, na)
            players = {1: MCTS(game, first, args, arena=arena(1)), -1: MCTS(game, second, args, arena=arena(1))}
            board, cur = game.getInitBoard(), 1
            while game.getGameEnded(board, cur) == 0:
                canon = game.getCanonicalForm(board, cur)
                action = int(np.argmax(players[cur].getActionProb(canon, temp=0)))
                assert game.getValidMoves(canon, 1)[action] > 0
                board, cur = game.getNextState(board, cur, action)
            r = cur * game.getGameEnded(board, cur)  # Arena.py:152
            if r == 1:
                one, two = (one + 1, two) if grp == 0 else (one, two + 1)
            elif r == -1:
                one, two = (one, two + 1) if grp == 0 else (one + 1, two)
            else:
                draws += 1
    assert got == (one, two, draws)
    assert sum(got) == 6
