"""compute-sanitizer driver: one small call of every kernel family added in round 1 (grid layer fwd/bwd, weight
gradients, replay emit/gather, TicTacToe tcgen05 forward, Connect4 tcgen05 forward (bf16x3 and f16f8) incl. folded heads, training
step pieces).  Sizes are tiny: the tool slows kernels down by 10-100x."""
import os
import sys

os.environ["AZG_TRAIN_EAGER"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from azgnn_b200 import _lib, games, training
from azgnn_b200.gridgnn import GridGNNStack
from azgnn_b200.nets import B200Connect4GNNWrapper, B200TicTacToeGNNWrapper
from azgnn_b200.replay import DeviceExamples

a = dict(lr=1e-3, dropout=0.0, epochs=1, batch_size=8, gnn_layers=2, use_gnn=True, numMCTSSims=4, cpuct=1.0, expand_by=2,
         b200_precision="bf16x3")
torch.manual_seed(0)
for gh, gw, H in ((3, 3, 64), (7, 7, 128), (8, 8, 256)):
    net = GridGNNStack(gh, gw, H, layers=2).cuda()
    x = torch.randn(301, gh * gw, H, device="cuda", requires_grad=True)
    net(x).sum().backward()
c4 = B200Connect4GNNWrapper(games.Connect4Game(7), a)
boards = np.random.default_rng(0).integers(-1, 2, size=(300, 7, 7)).astype(np.int8)
for prec in (_lib.PREC_BF16X3, _lib.PREC_F16F8):
    for fold in (False, True):
        c4.fold_heads = fold
        c4.forward_states(c4.states_from_boards(boards), _lib.EVAL_STD | _lib.EVAL_GNN, precision=prec)
ttt = B200TicTacToeGNNWrapper(games.TicTacToeGame(4), a)
ttt.forward_states(ttt.states_from_boards(np.random.default_rng(1).integers(-1, 2, size=(300, 4, 4)).astype(np.int8)))
bt = torch.FloatTensor(boards[:8].astype(np.float64)).cuda()
pi = torch.full((8, 8), 0.125, device="cuda")
v = torch.zeros(8, device="cuda")
for step in (training.std_step, training.gnn_step):
    step(training.CudaOps, c4, bt, pi, v).backward()
ex = DeviceExamples.from_examples(games.Connect4Game(7), [(boards[i].astype(np.int64), list(np.full(8, 0.125)), 1) for i in range(20)])
ex.sample(8)
torch.cuda.synchronize()
print("sanitize_small ok")
