"""ncu driver: FrozenLake 8x8 network (`fl_forward_kernel`, graph build fused) on 65,536 agent cells, and the 64-cell table
the wrapper actually evaluates per weight version."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from azgnn_b200 import games
from azgnn_b200.nets import B200FrozenLakeNet

torch.manual_seed(0)
w = B200FrozenLakeNet(games.FrozenLakeGame(8), dict(lr=1e-3, epochs=1, batch_size=32, embedding_dim=128, gnn_layers=3))
cells = torch.zeros(65536, 2, dtype=torch.int64, device="cuda")
cells[:, 0] = torch.randint(0, 64, (65536,), device="cuda")
for _ in range(2):
    o = w._forward_cells(cells)
torch.cuda.synchronize()
print("ok", float(o["v"].sum()))
