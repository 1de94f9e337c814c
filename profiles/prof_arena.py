"""cProfile of BatchedArena.playGames(100) on Connect4 7x7 (the Coach's arena phase): where the host time goes.
usage: python profiles/prof_arena.py [graph 0|1]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from azgnn_b200.games import Connect4Game
from azgnn_b200.nets import B200Connect4GNNWrapper
from azgnn_b200.pit import BatchedArena

graph = int(sys.argv[1]) if len(sys.argv) > 1 else 1
a = dict(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5,
         b200_precision="f16f8", b200_graph_search=bool(graph))
game = Connect4Game(7)
torch.manual_seed(0)
n1 = B200Connect4GNNWrapper(game, a)
torch.manual_seed(1)
n2 = B200Connect4GNNWrapper(game, a)
np.random.seed(0)
BatchedArena(game, n1, n2, a).playGames(100)  # warm-up: kernels, packed weights
torch.cuda.synchronize()
t0 = time.perf_counter()
res = BatchedArena(game, n1, n2, a).playGames(100)
torch.cuda.synchronize()
print(f"graph={graph}: 100 games in {time.perf_counter() - t0:.3f} s, result {res}")
pr = cProfile.Profile()
pr.enable()
BatchedArena(game, n1, n2, a).playGames(100)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
