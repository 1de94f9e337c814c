"""ncu driver: a few lock-step self-play moves (Connect4 games, default precision) -- arena kernels in their real mix.
usage: run_selfplay.py [moves] [games]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from azgnn_b200.games import Connect4Game
from azgnn_b200.nets import B200Connect4GNNWrapper
from azgnn_b200.selfplay import BatchedSelfPlay

a = dict(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5,
         tempThreshold=15)
game = Connect4Game(7)
torch.manual_seed(0)
net = B200Connect4GNNWrapper(game, a)
sp = BatchedSelfPlay(game, net, a, int(sys.argv[2]) if len(sys.argv) > 2 else 4096, seed=0, collect_examples=False)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    sp.step_all()
torch.cuda.synchronize()
print("ok", sp.moves_played)
