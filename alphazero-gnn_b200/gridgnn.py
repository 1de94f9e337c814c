"""Grid-graph GNN stack of the roofline sweep (BASELINE.json configs[4]).

Operator = the reference's `FrozenLakeNet.GNNLayer` (frozenlake/FrozenLakeNet.py:8-33),
`relu(bmm(adj, W x + b))`, applied to gh x gw 4-neighbour grid graphs with the symmetric
normalisation of `create_adjacency` (:68-72) and self loops.  The dense Linear runs on the
library's SGEMM, the aggregation on `azg_grid_aggregate_relu_*` (the adjacency is never
materialised); `torch.autograd.Function`s route tensors, parameters live in `nn.Linear`s.
"""
import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr, stream
from .training import _Linear


class _GridAggRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sup, gh, gw):
        sup = sup.contiguous()
        B, n, H = sup.shape
        out = torch.empty_like(sup)
        _lib.check(_lib.lib().azg_grid_aggregate_relu_forward(ptr(sup), B, gh, gw, H, ptr(out), stream()))
        ctx.save_for_backward(out)
        ctx.g = (gh, gw)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        dout = dout.contiguous()
        B, n, H = out.shape
        dsup = torch.empty_like(out)
        _lib.check(_lib.lib().azg_grid_aggregate_relu_backward(ptr(dout), ptr(out), B, ctx.g[0], ctx.g[1], H, ptr(dsup), stream()))
        return dsup, None, None


class GridGNNStack(nn.Module):
    """`layers` x GNNLayer(hidden, hidden) over [B, gh*gw, hidden] node features."""

    def __init__(self, gh, gw, hidden, layers=2):
        super().__init__()
        _lib.require_device()
        self.gh, self.gw, self.hidden = gh, gw, hidden
        self.gnn_layers = nn.ModuleList([nn.Linear(hidden, hidden) for _ in range(layers)])

    def forward(self, x):
        B, n, H = x.shape
        assert n == self.gh * self.gw and H == self.hidden
        for lin in self.gnn_layers:
            sup = _Linear.apply(x.reshape(B * n, H), lin.weight, lin.bias, False)
            x = _GridAggRelu.apply(sup.reshape(B, n, H), self.gh, self.gw)
        return x
