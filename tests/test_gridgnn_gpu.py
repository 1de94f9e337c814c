"""Grid-graph sweep operator (BASELINE configs[4]): CUDA forward/backward of the GNNLayer stack on
grid graphs vs the oracle (torch bmm with the dense normalised adjacency)."""
import numpy as np
import pytest
import torch

from oracle import nets as onets
from azgnn_b200.gridgnn import GridGNNStack

pytestmark = pytest.mark.gpu
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("gh,gw", [(3, 3), (4, 4), (6, 7), (7, 7), (8, 8), (16, 16)])
@pytest.mark.parametrize("hidden", [64, 128, 256])
def test_grid_gnn_forward_backward(gh, gw, hidden):
    torch.manual_seed(gh * 100 + hidden)
    B = 37
    net = GridGNNStack(gh, gw, hidden, layers=2).cuda()
    x = torch.randn(B, gh * gw, hidden, device="cuda", requires_grad=True)
    y = net(x)
    g = torch.randn_like(y)
    y.backward(g)
    got = [x.grad.clone()] + [p.grad.clone() for p in net.parameters()]
    x2 = x.detach().clone().requires_grad_(True)
    ws = [l.weight.detach().clone().requires_grad_(True) for l in net.gnn_layers]
    bs = [l.bias.detach().clone().requires_grad_(True) for l in net.gnn_layers]
    y2 = onets.grid_gnn_forward(ws, bs, x2, gh, gw)
    y2.backward(g)
    ref = [x2.grad] + [t.grad for pair in zip(ws, bs) for t in pair]
    assert (y - y2).abs().max().item() <= 1e-5 * max(1.0, y2.abs().max().item())
    for a, b in zip(got, ref):
        bad = (a - b).abs() > 1e-3 * b.abs().max() + 2e-6
        assert bad.float().mean().item() < 1e-3, (a - b).abs().max().item()  # isolated ReLU knife-edge flips only
