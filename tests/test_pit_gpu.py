"""Batched Arena: all evaluation games in flight at once.  With fresh trees per game in the
sequential loop (the flagged difference from the reference's persistent trees) it reproduces the sequential games
move for move; the tie-break among equally visited moves is pinned to the lowest index in both runs."""
import numpy as np
import pytest
import torch

from azgnn_b200 import games
from azgnn_b200.mcts import MCTS
from azgnn_b200.nets import B200Connect4GNNWrapper, B200TicTacToeGNNWrapper
from azgnn_b200.pit import BatchedArena
from helpers import dotdict

pytestmark = pytest.mark.gpu


def _sequential(game, nets, args, num):
    """Arena.playGamesForTwoPlayer (Arena.py:249-283) with a new MCTS pair per game"""
    half = num // 2
    one = two = draws = 0
    for grp in range(2):
        for _ in range(half):
            first, second = (nets[0], nets[1]) if grp == 0 else (nets[1], nets[0])
            players = {1: MCTS(game, first, args), -1: MCTS(game, second, args)}
            board, cur = game.getInitBoard(), 1
            while game.getGameEnded(board, cur) == 0:
                canon = game.getCanonicalForm(board, cur)
                action = int(np.argmax(players[cur].getActionProb(canon, temp=0)))
                assert game.getValidMoves(canon, 1)[action] > 0
                board, cur = game.getNextState(board, cur, action)
            r = cur * game.getGameEnded(board, cur)
            if r == 1:
                one, two = (one + 1, two) if grp == 0 else (one, two + 1)
            elif r == -1:
                one, two = (one, two + 1) if grp == 0 else (one + 1, two)
            else:
                draws += 1
    return one, two, draws


@pytest.mark.parametrize("kind,n", [("c4", 5), ("ttt", 3)])
def test_batched_arena_equals_sequential_games_with_fresh_trees(kind, n, monkeypatch):
    monkeypatch.setattr(np.random, "choice", lambda a, *args, **kw: np.asarray(a).reshape(-1)[0])
    game = games.Connect4Game(n) if kind == "c4" else games.TicTacToeGame(n)
    args = dotdict(dict(lr=1e-3, dropout=0.3, gnn_layers=2, use_gnn=True, numMCTSSims=8, cpuct=1.0, expand_by=3,
                        b200_precision="fp32"))
    W = B200Connect4GNNWrapper if kind == "c4" else B200TicTacToeGNNWrapper
    torch.manual_seed(0)
    net_a = W(game, args)
    torch.manual_seed(1)
    net_b = W(game, args)
    got = BatchedArena(game, net_a, net_b, args).playGames(6)
    want = _sequential(game, (net_a, net_b), args, 6)
    assert got == want and sum(got) == 6
    # the two halves mirror each other when a network meets itself
    one, two, draws = BatchedArena(game, net_a, net_a, args).playGames(4)
    assert one == two and one + two + draws == 4


def test_pit_gnn_vs_regular():
    """--pit_gnn (main.py:60-138): GNN wrapper searching with predict_with_gnn vs regular wrapper searching with predict"""
    from azgnn_b200.nets import B200Connect4NNetWrapper
    from azgnn_b200.pit import pit_gnn_vs_regular
    game = games.Connect4Game(5)
    args = dotdict(dict(lr=1e-3, dropout=0.3, gnn_layers=2, use_gnn=True, numMCTSSims=6, cpuct=1.0, expand_by=3, arenaCompare=6,
                        b200_precision="fp32"))
    torch.manual_seed(0)
    gnn_net = B200Connect4GNNWrapper(game, args)
    torch.manual_seed(1)
    reg_net = B200Connect4NNetWrapper(game, args)
    g, r, d = pit_gnn_vs_regular(game, gnn_net, reg_net, args)
    assert g + r + d == 6


def test_single_player_arena_frozenlake():
    """Arena.playGamesForSinglePlayer semantics on FrozenLake 4x4: results partition the games; a network that was
    trained to walk to the goal beats an untrained one."""
    from azgnn_b200.nets import B200FrozenLakeNet
    from azgnn_b200.pit import BatchedSinglePlayerArena
    game = games.FrozenLakeGame(4)
    args = dotdict(dict(lr=1e-2, epochs=30, batch_size=8, embedding_dim=128, gnn_layers=2, numMCTSSims=25, cpuct=2.0, use_gnn=False))
    torch.manual_seed(0)
    fresh = B200FrozenLakeNet(game, args)
    torch.manual_seed(1)
    taught = B200FrozenLakeNet(game, args)
    # teach the safe path of the standard map: down, down, right, down, right, right (cells 0,4,8,9,13,14 -> 15);
    # actions: 0 up, 1 right, 2 down, 3 left (FrozenLakeGame.py:88-110)
    path = [(0, 2), (4, 2), (8, 1), (9, 2), (13, 1), (14, 1)]
    examples = []
    for cell, act in path * 4:
        b = np.zeros((4, 4)); b[cell // 4, cell % 4] = 1
        pi = np.zeros(4); pi[act] = 1.0
        examples.append((b, list(pi), 1.0))
    np.random.seed(0)
    taught.train(examples)
    one, two, draws = BatchedSinglePlayerArena(game, fresh, taught, args).playGames(8)
    assert one + two + draws == 8
    assert two >= one


@pytest.mark.parametrize("kind,n,prec", [("c4", 7, "f16f8"), ("c4", 5, "fp32"), ("c4", 6, "bf16x3")])
def test_graph_replayed_search_equals_eager_search(kind, n, prec):
    """BatchedMCTS.search(graph=True): the whole lock-step search replayed as one CUDA graph (the arena's launch-bound
    50-game searches) leaves bit-identical trees: same root statistics ply after ply, same leaf count."""
    from azgnn_b200.mcts import BatchedMCTS
    game = games.Connect4Game(n) if kind == "c4" else games.TicTacToeGame(n)
    args = dotdict(dict(lr=1e-3, dropout=0.3, gnn_layers=2, use_gnn=True, numMCTSSims=9, cpuct=1.0, expand_by=3,
                        b200_precision=prec))
    W = B200Connect4GNNWrapper if kind == "c4" else B200TicTacToeGNNWrapper
    torch.manual_seed(3)
    net = W(game, args)
    G = 50
    eager, graphed = (BatchedMCTS(game, net, args, n_games=G) for _ in range(2))
    assert eager.compact and graphed.compact
    for m in (eager, graphed):
        m.set_root_boards([game.getInitBoard()] * G)
    rng = np.random.default_rng(0)
    for ply in range(5):
        eager.search(9)
        graphed.search(9, graph=True)
        (N1, Q1, t1), (N2, Q2, t2) = eager.root_stats(), graphed.root_stats()
        assert np.array_equal(N1, N2) and np.array_equal(Q1, Q2) and np.array_equal(t1, t2), ply
        # play a visited move (not always the best one, so that the games diverge from each other)
        actions = np.array([rng.choice(np.flatnonzero(N1[g] > 0)) for g in range(G)], dtype=np.int32)
        e1, _ = eager.advance_arrays(actions.copy())
        e2, _ = graphed.advance_arrays(actions.copy())
        assert np.array_equal(e1, e2)
        if (e1 != 0).any():
            break
    assert ply >= 2 and isinstance(graphed._graphs[9], tuple), "the third search must have been a graph replay"
    assert eager.leaf_evaluations() == graphed.leaf_evaluations() > 0
    # a weight update invalidates the captured graph (it holds the address of the packed weight images)
    net.weights_changed()
    graphed.search(9, graph=True)
    assert graphed._graphs[9] == "warm"
