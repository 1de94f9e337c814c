"""A/B driver: phase times (library CUDA-event hooks) of N Connect4 leaf-evaluation steps, 65,536 positions, bf16x3.
usage: [AZG_LIBRARY=path/to/variant.so] python profiles/bench_trunk.py [steps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from azgnn_b200 import _lib
from azgnn_b200.games import Connect4Game
from azgnn_b200.nets import B200Connect4GNNWrapper

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
a = dict(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5,
         b200_precision="bf16x3")
torch.manual_seed(0)
net = B200Connect4GNNWrapper(Connect4Game(7), a)
states = [net.states_from_boards(np.random.default_rng(i).integers(-1, 2, size=(65536, 7, 7)).astype(np.int8)) for i in range(4)]
mask = _lib.EVAL_STD | _lib.EVAL_GNN
for i in range(5):
    net.forward_states(states[i % 4], mask)
torch.cuda.synchronize()
lib = _lib.lib()
lib.azg_timing_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    net.forward_states(states[i % 4], mask)
e1.record()
torch.cuda.synchronize()
out = {"lib": os.path.basename(_lib.LIB_PATH), "ms_per_step": e0.elapsed_time(e1) / steps}
for pid, name in ((0, "trunk"), (1, "gemm"), (2, "heads")):
    tot, cnt = C.c_double(), C.c_int()
    _lib.check(lib.azg_timing_read(pid, C.byref(tot), C.byref(cnt)))
    out[name] = tot.value / steps
print(out)
