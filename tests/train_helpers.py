"""Oracle-backed compute vocabulary with the same interface as azgnn_b200.training.CudaOps
(TEST ONLY): lets the host-side training logic (graphs, row sharding, gradient all-reduce) run on
CPU under gloo, and provides the autograd reference for the GPU parity tests."""
import torch
import torch.nn.functional as F

from oracle import nets as onets


class OracleOps:
    @staticmethod
    def conv_relu(x, conv):
        return F.relu(F.conv2d(x, conv.weight, conv.bias, padding=int(conv.padding[0])))

    @staticmethod
    def linear(x, lin, relu=False):
        y = F.linear(x, lin.weight, lin.bias)
        return F.relu(y) if relu else y

    @staticmethod
    def dropout(x, p):
        assert p == 0, "parity tests disable dropout (RNG dependent)"
        return x

    @staticmethod
    def pv_loss(logits, vraw, target_pi, target_v, norm):
        logp = F.log_softmax(logits, dim=1)
        v = torch.tanh(vraw)
        loss = (-torch.sum(target_pi * logp) + torch.sum((target_v - v.view(-1)) ** 2)) / norm
        return loss.reshape(1), logp, v.view(-1)

    @staticmethod
    def gnn_layer(f0, path, layer):
        p = {"L." + k: v for k, v in layer.named_parameters()}
        out = onets.gnn_layer(p, "L.", torch.cat([f0.unsqueeze(0), path], dim=0))
        return out[0]

    @staticmethod
    def graph_mean_relu(sup, counts):
        out = torch.zeros_like(sup)
        for b in range(sup.shape[0]):
            k = int(counts[b])
            adj = onets.fl_adjacency(k).to(sup.device)
            out[b, :k] = F.relu(torch.mm(adj, sup[b, :k]))
        return out

    @staticmethod
    def fl_graph(states, n):
        B = states.shape[0]
        nodes = torch.zeros(B, 5, n * n, device=states.device)
        counts = torch.zeros(B, dtype=torch.int32, device=states.device)
        for b in range(B):
            cells = onets.fl_node_cells(int(states[b, 0]), n)
            counts[b] = len(cells)
            for j, c in enumerate(cells):
                nodes[b, j, c] = 1.0
        return nodes, counts
