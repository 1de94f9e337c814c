"""ncu driver: a few launches of the fused grid-graph layer (7x7, H=128, 16384 graphs, bf16x3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from azgnn_b200.gridgnn import GridGNNStack

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
torch.manual_seed(0)
net = GridGNNStack(7, 7, H, layers=2, precision=prec).cuda()
x = torch.randn(16384, 49, H, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = net(x)
torch.cuda.synchronize()
print("ok", float(y.sum()))
