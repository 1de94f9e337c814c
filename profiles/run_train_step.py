"""ncu driver: two eager training epochs of Connect4GNNWrapper.train's step pair (B = 64) -- the K3 kernels."""
import os
import sys

os.environ["AZG_TRAIN_EAGER"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from azgnn_b200 import training
from azgnn_b200.games import Connect4Game
from azgnn_b200.nets import B200Connect4GNNWrapper

a = dict(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5)
torch.manual_seed(0)
w = B200Connect4GNNWrapper(Connect4Game(7), a)
rng = np.random.default_rng(0)
boards = torch.FloatTensor(rng.integers(-1, 2, size=(64, 7, 7)).astype(np.float64)).cuda()
pi = torch.FloatTensor(rng.dirichlet(np.ones(8), size=64)).cuda()
v = torch.FloatTensor(rng.uniform(-1, 1, 64)).cuda()
for _ in range(2):
    for step in (training.std_step, training.gnn_step):
        for p in list(w.nnet.parameters()) + list(w.gnn.parameters()):
            p.grad = None
        step(training.CudaOps, w, boards, pi, v).backward()
torch.cuda.synchronize()
print("ok")
