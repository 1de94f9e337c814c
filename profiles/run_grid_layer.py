"""ncu driver: a few launches of the fused grid-graph layer (default 7x7, H=128, 16384 graphs, bf16x3).
usage: python profiles/run_grid_layer.py [H] [precision] [gh] [gw] [graphs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from azgnn_b200.gridgnn import GridGNNStack

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
torch.manual_seed(0)
gh = int(sys.argv[3]) if len(sys.argv) > 3 else 7
gw = int(sys.argv[4]) if len(sys.argv) > 4 else 7
B = int(sys.argv[5]) if len(sys.argv) > 5 else 16384
net = GridGNNStack(gh, gw, H, layers=2, precision=prec).cuda()
x = torch.randn(B, gh * gw, H, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = net(x)
torch.cuda.synchronize()
print("ok", float(y.sum()))
