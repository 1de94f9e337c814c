"""ncu driver: TicTacToe 4x4 GNN leaf evaluation, 65,536 positions (default precision: auto -> bf16x3 on tcgen05)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from azgnn_b200 import games
from azgnn_b200.nets import B200TicTacToeGNNWrapper

a = dict(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5)
game = games.TicTacToeGame(4)
torch.manual_seed(0)
w = B200TicTacToeGNNWrapper(game, a)
boards = np.random.default_rng(0).integers(-1, 2, size=(65536, 4, 4)).astype(np.int8)
states = w.states_from_boards(boards)
for _ in range(2):
    o = w.forward_states(states)
torch.cuda.synchronize()
print("ok", float(o["v"].sum()))
