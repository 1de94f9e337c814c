"""ncu driver: three Connect4 leaf-evaluation steps (65,536 positions; AZG_RUN_PRECISION, default f16f8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from azgnn_b200 import _lib
from azgnn_b200.games import Connect4Game
from azgnn_b200.nets import B200Connect4GNNWrapper

a = dict(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5,
         b200_precision=os.environ.get("AZG_RUN_PRECISION", "f16f8"))
torch.manual_seed(0)
net = B200Connect4GNNWrapper(Connect4Game(7), a)
states = net.states_from_boards(np.random.default_rng(0).integers(-1, 2, size=(65536, 7, 7)).astype(np.int8))
for _ in range(3):
    o = net.forward_states(states, _lib.EVAL_STD | _lib.EVAL_GNN)
torch.cuda.synchronize()
print("ok", float(o["v_gnn"].sum()))
