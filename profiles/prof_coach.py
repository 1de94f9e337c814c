"""cProfile of one warm Coach.learn() iteration (Connect4 7x7, 4,096 episodes, 20 epochs, 100 arena games): where the host
time of the phases goes.  usage: python profiles/prof_coach.py [episodes]"""
import cProfile
import os
import pstats
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from azgnn_b200.coach import Coach
from azgnn_b200.games import Connect4Game
from azgnn_b200.nets import B200Connect4GNNWrapper


class Args(dict):
    __getattr__ = dict.__getitem__


eps = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
a = Args(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5,
         tempThreshold=15, numIters=1, numEps=eps, n_parallel_games=eps, maxlenOfQueue=200000, numItersForTrainExamplesHistory=5,
         arenaCompare=100, updateThreshold=0.6, checkpoint=tempfile.mkdtemp(prefix="azg_coach_"), b200_precision="f16f8",
         save_examples=False)
game = Connect4Game(7)
torch.manual_seed(0)
np.random.seed(0)
coach = Coach(game, B200Connect4GNNWrapper(game, a), a)
coach.learn()  # pays the one-time costs
coach.learn()
print("warm iteration:", {k: round(v, 4) for k, v in coach.timings.items()})
pr = cProfile.Profile()
pr.enable()
coach.learn()
pr.disable()
print("profiled iteration:", {k: round(v, 4) for k, v in coach.timings.items()})
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(25)
