"""Host-side regressions for three ADVICE findings (round 1), on the host check arena:
  * the drop-in MCTS object outgrows any fixed node capacity when Coach / Arena reuse it across all arenaCompare games
    (Coach.py:128-142): the table is re-homed in a larger arena and the search continues as in the reference's dicts;
  * BatchedSelfPlay.play(n) keeps the n episodes that START first, not the first n to finish (no bias to short games);
  * TicTacToe boards with more than 32 actions are refused loudly (32-bit valid-move masks)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck"))
from host_arena import HostArena  # noqa: E402

from azgnn_b200 import games  # noqa: E402
from azgnn_b200.mcts import MCTS  # noqa: E402
from azgnn_b200.selfplay import BatchedSelfPlay  # noqa: E402
from oracle import rules as orules  # noqa: E402
from oracle.mcts import FakeNet, OracleMCTS  # noqa: E402
from helpers import dotdict  # noqa: E402


def test_persistent_mcts_grows_like_the_reference_dicts():
    """Twelve sequential games on ONE MCTS object whose arena starts with room for 24 nodes: visit distributions equal
    the oracle's unbounded-dict MCTS move for move, and the table has been re-homed several times on the way."""
    n = 4
    game, ogame = games.Connect4Game(n), orules.Connect4Rules(n)
    args = dotdict(dict(numMCTSSims=6, cpuct=1.0, use_gnn=False, expand_by=0, tempThreshold=15))
    net = FakeNet(n + 1, salt=8)
    arena = HostArena("connect4", n, 1, 6, 1.0, capacity=24)
    m = MCTS(game, net, args, arena=arena)
    om = OracleMCTS(ogame, net, args)
    grown = 0
    rng = np.random.default_rng(0)
    for _game_no in range(12):
        board, player = game.getInitBoard(), 1
        while game.getGameEnded(board, player) == 0:
            canon = game.getCanonicalForm(board, player)
            cap0 = arena.capacity
            got = m.getActionProb(canon, temp=1)
            grown += arena.capacity > cap0
            want = om.getActionProb(np.asarray(canon), temp=1)
            assert list(got) == list(want)
            action = int(rng.choice(len(got), p=np.asarray(got)))
            board, player = game.getNextState(board, player, action)
    assert grown >= 3 and arena.node_count(0) > 24 * 4
    # the dict views survive the moves: same node set as the oracle
    assert len(m.Ns) == len(om.Ns) and len(m.Nsa) == len(om.Nsa)


def test_play_keeps_the_episodes_that_started_first():
    """G = 3 slots, n = 7 episodes: slots restart as their games end and episodes are numbered in START order; play(7)
    returns episodes 0..6 (in the order they finished), each with its own length, and ignores episodes 7+ that finish while the last
    kept ones are still running (the old rule, first n to finish, let those short late games displace long early ones)."""
    game = games.TicTacToeGame(3)
    args = dotdict(dict(numMCTSSims=4, cpuct=1.0, use_gnn=False, expand_by=0, tempThreshold=3))
    G, n_eps = 3, 7
    sp = BatchedSelfPlay(game, FakeNet(10, salt=2, spread=3.0), args, G, seed=5, arena=HostArena("tictactoe", 3, G, 4, 1.0, capacity=4 * 11 + 16))
    lengths, finish_order = {}, []
    real_step = sp.step_all

    def spy():
        before_idx, before_len = sp.ep_index.copy(), sp.step.copy()
        out = real_step()
        for idx in sp.last_done_index:
            g = int(np.flatnonzero(before_idx == idx)[0])
            lengths[int(idx)] = int(before_len[g]) + 1
            finish_order.append(int(idx))
        return out
    sp.step_all = spy
    eps = sp.play(n_eps)
    assert len(eps) == n_eps
    assert set(range(n_eps)) <= set(lengths)
    kept_order = [i for i in finish_order if i < n_eps]
    for i, (std, _gnn) in zip(kept_order, eps):  # returned in finish order; 8 symmetric forms per stored position
        assert len(std) == lengths[i] * 8, (i, len(std), lengths[i])
    assert sp._next_ep >= n_eps + 1  # slots kept restarting while the last kept episodes ran
    assert len(set(lengths[i] for i in range(n_eps))) > 1  # the kept games really differ in length
    assert sorted(i for i in finish_order if i < n_eps) == list(range(n_eps))  # each kept episode finished exactly once


def test_tictactoe_with_more_than_32_actions_is_refused():
    with pytest.raises(RuntimeError, match="32"):
        HostArena("tictactoe", 6, 1, 4, 1.0, capacity=64)
    HostArena("tictactoe", 5, 1, 4, 1.0, capacity=64)  # 26 actions: fine
