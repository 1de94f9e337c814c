"""The package's host-side Game mirrors (azgnn_b200/games.py) against the reference's outputs."""
import numpy as np
import pytest

from azgnn_b200 import games
from helpers import golden


def _game(name):
    kind, n = name.split("_")
    return {"c4": games.Connect4Game, "ttt": games.TicTacToeGame, "fl": games.FrozenLakeGame}[kind](int(n))


@pytest.mark.parametrize("name", ["c4_7", "c4_5", "c4_4", "ttt_3", "ttt_4", "fl_4", "fl_8"])
def test_game_api_matches_reference(name):
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
        for a in range(game.getActionSize()):
            if valids[a] and e == 0:
                nb, npl = game.getNextState(np.array(b, copy=True), player, a)
                assert np.array_equal(np.asarray(nb, dtype=np.float64), g["next"][i, a]) and npl == g["next_player"][i, a]


def test_value_types():
    c4 = games.Connect4Game(7)
    assert isinstance(c4.getGameEnded(c4.getInitBoard(), 1), int)
    assert len(c4.stringRepresentation(c4.getInitBoard())) == 392
    fl = games.FrozenLakeGame(4)
    assert fl.getValidMoves(fl.getInitBoard(), 1).dtype == np.int8
    assert fl.stringRepresentation(fl.getInitBoard()) == "0,0"
    b = np.zeros((4, 4)); b[3, 3] = 1
    assert fl.getGameEnded(b, 1) == 1.0 and isinstance(fl.getGameEnded(b, 1), float)
