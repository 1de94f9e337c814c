"""K5 host logic without GPUs: two gloo ranks run the package's training graph (row sharding,
feature all-gather, gradient all-reduce) with the oracle compute vocabulary and must reproduce the
single-process gradients of the full minibatch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _W:
    """minimal wrapper stand-in on CPU: the training graph only needs these attributes"""

    def __init__(self, kind, n):
        from azgnn_b200 import modules
        self.kind = "connect4" if kind == "c4" else "tictactoe"
        self.board_x = self.board_y = n
        torch.manual_seed(0)
        if kind == "c4":
            self.nnet = modules.Connect4Trunk(n, n + 1, dropout=0.0)
            self.feature_dim = 64 * n * n
        else:
            self.nnet = modules.TicTacToeTrunk(n, n * n + 1)
            self.feature_dim = 128 * (n - 2) ** 2
        self.gnn = modules.PolicyValueGNN(self.feature_dim, 2)


def _batch(kind, n, B):
    rng = np.random.default_rng(3)
    A = n + 1 if kind == "c4" else n * n + 1
    boards = torch.FloatTensor(rng.integers(-1, 2, size=(B, n, n)).astype(np.float64))
    tpi = torch.FloatTensor(rng.dirichlet(np.ones(A), size=B))
    tv = torch.FloatTensor(rng.uniform(-1, 1, B))
    return boards, tpi, tv


def _grads(kind, n, B):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from azgnn_b200 import training
    from train_helpers import OracleOps
    w = _W(kind, n)
    boards, tpi, tv = _batch(kind, n, B)
    out = {}
    for name, step, mod in (("std", training.std_step, w.nnet), ("gnn", training.gnn_step, w.gnn)):
        for p in list(w.nnet.parameters()) + list(w.gnn.parameters()):
            p.grad = None
        loss = step(OracleOps, w, boards, tpi, tv)
        if loss is not None:
            loss.backward()
        params = list(mod.parameters())
        training.exchange_grads(w, name)
        out[name] = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in params]
    return out


def _worker(rank, world, port, kind, n, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = _grads(kind, n, B)
    q.put((rank, {k: [t.numpy() for t in v] for k, v in g.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,n,B", [("c4", 4, 10), ("ttt", 3, 7)])
def test_two_ranks_reproduce_single_process_gradients(kind, n, B):
    torch.set_num_threads(2)
    single = _grads(kind, n, B)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, n, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):  # every rank ends the exchange with the full-batch gradients (replicated optimizer step)
        for name in ("std", "gnn"):
            for a, b in zip(got[rank][name], single[name]):
                np.testing.assert_allclose(a, b.numpy(), rtol=1e-4, atol=1e-6)


def test_shard_rows_cover_the_batch_once():
    from azgnn_b200.training import shard_rows
    for B in (1, 2, 7, 64, 65):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                lo, hi = shard_rows(B, r, world)
                rows += list(range(lo, hi))
            assert rows == list(range(B))
            assert shard_rows(B, 0, world)[0] == 0  # row 0 (the GNN target) stays on rank 0
