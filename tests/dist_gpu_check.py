"""Multi-GPU check (run with torchrun, one process per GPU, NCCL):
  1. data-parallel training step (row sharding + feature all-gather + NCCL gradient all-reduce,
     azgnn_b200/training.py) reproduces the single-GPU gradients of the full minibatch;
  2. sharded leaf evaluation: every rank evaluates its slice, the concatenation equals one GPU's result.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_gpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from azgnn_b200 import games, training  # noqa: E402
from azgnn_b200.nets import B200Connect4GNNWrapper  # noqa: E402
from helpers import dotdict  # noqa: E402


def grads(w, boards, tpi, tv):
    out = {}
    for name, step, mod in (("std", training.std_step, w.nnet), ("gnn", training.gnn_step, w.gnn)):
        for p in list(w.nnet.parameters()) + list(w.gnn.parameters()):
            p.grad = None
        loss = step(training.CudaOps, w, boards, tpi, tv)
        if loss is not None:
            loss.backward()
        params = list(mod.parameters())
        training.exchange_grads(w, name)
        out[name] = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in params]
    return out


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    n, B = 7, 64
    args = dotdict(dict(lr=1e-3, dropout=0.0, epochs=1, batch_size=64, gnn_layers=2, use_gnn=True))
    torch.manual_seed(0)
    w = B200Connect4GNNWrapper(games.Connect4Game(n), args)
    rng = np.random.default_rng(0)
    dev = w.device
    boards = torch.FloatTensor(rng.integers(-1, 2, size=(B, n, n)).astype(np.float64)).to(dev)
    tpi = torch.FloatTensor(rng.dirichlet(np.ones(n + 1), size=B)).to(dev)
    tv = torch.FloatTensor(rng.uniform(-1, 1, B)).to(dev)
    single = grads(w, boards, tpi, tv)  # process group not initialised yet: full batch on this GPU
    ref_eval = w.predict_batch(boards.to(torch.int8).cpu().numpy())
    dist.init_process_group("nccl", device_id=dev)
    multi = grads(w, boards, tpi, tv)
    ok = True
    for name in ("std", "gnn"):
        for a, b in zip(multi[name], single[name]):
            err = (a - b).abs().max().item()
            tol = 1e-4 * b.abs().max().item() + 1e-7
            if err > tol:
                ok = False
                print(f"[rank {rank}] {name} grad mismatch: {err:.3e} > {tol:.3e}")
    # sharded leaf evaluation, no collective on the data path
    lo, hi = training.shard_rows(B, rank, world)
    mine = w.predict_batch(boards[lo:hi].to(torch.int8).cpu().numpy())
    parts = [None] * world
    dist.all_gather_object(parts, mine["pi_gnn"])
    if rank == 0:
        ok = ok and np.array_equal(np.concatenate(parts, 0), ref_eval["pi_gnn"])
        print("DIST_GPU_CHECK", "OK" if ok else "FAILED", f"world={world}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
