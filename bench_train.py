#!/usr/bin/env python
"""bench_train.py -- the training step of Connect4GNNWrapper.train (connect4/Connect4GNN.py:122-197)
on the CUDA library: one "epoch" = one standard-network step + one GNN step on B=64 rows, Adam as in
the reference.  Reports ms per phase (forward+backward of each step, optimizer, gradient exchange)
against the HBM bytes each phase must move ("K3/K5 ... HBM bound").

    python bench_train.py [--iters 20] [--cpu]           one GPU
    torchrun --nproc-per-node N ... bench_train.py       rows sharded, NCCL all-reduce of the gradients
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

N_BOARD = 7


class Args(dict):
    __getattr__ = dict.__getitem__


def reference_args():
    return Args(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10,
                cpuct=1.0, expand_by=5, tempThreshold=15)


def batch(B, A, seed, dev):
    rng = np.random.default_rng(seed)
    boards = torch.FloatTensor(rng.integers(-1, 2, size=(B, N_BOARD, N_BOARD)).astype(np.float64)).to(dev)
    pi = torch.FloatTensor(rng.dirichlet(np.ones(A), size=B)).to(dev)
    v = torch.FloatTensor(rng.uniform(-1, 1, B)).to(dev)
    return boards, pi, v


def cpu_epoch(iters):
    """Oracle restatement of one reference "epoch" (std step + GNN step, Adam) on the host cores."""
    from azgnn_b200 import modules, training
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from train_helpers import OracleOps
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)

    class W:
        kind, board_x, board_y, feature_dim = "connect4", N_BOARD, N_BOARD, 64 * N_BOARD * N_BOARD
    w = W()
    w.nnet = modules.Connect4Trunk(N_BOARD, N_BOARD + 1, 0.0)
    w.gnn = modules.PolicyValueGNN(w.feature_dim, 2)
    o1 = torch.optim.Adam(w.nnet.parameters(), lr=1e-3)
    o2 = torch.optim.Adam(w.gnn.parameters(), lr=1e-3)
    boards, pi, v = batch(64, N_BOARD + 1, 0, "cpu")

    def epoch():
        o1.zero_grad()
        training.std_step(OracleOps, w, boards, pi, v).backward()
        o1.step()
        o2.zero_grad()
        training.gnn_step(OracleOps, w, boards, pi, v).backward()
        o2.step()
    epoch()
    t0 = time.perf_counter()
    for _ in range(iters):
        epoch()
    return (time.perf_counter() - t0) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cpu", action="store_true", help="also time the oracle's CPU restatement (3 epochs)")
    ap.add_argument("--profile", action="store_true", help="print torch.profiler's kernel table for 3 epochs (stderr)")
    ap.add_argument("--fused-adam", action="store_true", help="torch.optim.Adam(fused=True) instead of the reference default")
    a = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from azgnn_b200 import training
    from azgnn_b200.games import Connect4Game
    from azgnn_b200.nets import B200Connect4GNNWrapper
    torch.manual_seed(0)
    w = B200Connect4GNNWrapper(Connect4Game(N_BOARD), reference_args())
    ops = training.CudaOps
    boards, pi, v = batch(a.batch, w.action_size, 0, dev)
    kw = {"fused": True} if a.fused_adam else {}
    nnet_params, gnn_params = list(w.nnet.parameters()), list(w.gnn.parameters())
    o1 = torch.optim.Adam(nnet_params, lr=1e-3, **kw)
    o2 = torch.optim.Adam(gnn_params, lr=1e-3, **kw)

    phases = ["std_fwd_bwd", "std_allreduce", "std_adam", "gnn_fwd", "gnn_bwd", "gnn_allreduce", "gnn_adam"]
    acc = {p: 0.0 for p in phases}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def epoch(timed):
        marks = [ev()]
        o1.zero_grad()
        loss = training.std_step(ops, w, boards, pi, v)
        if loss is not None:
            loss.backward()
        marks.append(ev())
        training.exchange_grads(w, "std")
        marks.append(ev())
        o1.step()
        marks.append(ev())
        o2.zero_grad()
        loss = training.gnn_step(ops, w, boards, pi, v)
        marks.append(ev())
        if loss is not None:
            loss.backward()
        marks.append(ev())
        training.exchange_grads(w, "gnn")
        marks.append(ev())
        o2.step()
        marks.append(ev())
        if timed:
            torch.cuda.synchronize()
            for i, p in enumerate(phases):
                acc[p] += marks[i].elapsed_time(marks[i + 1])

    for _ in range(3):
        epoch(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = ev()
    for _ in range(a.iters):
        epoch(False)
    t1 = ev()
    torch.cuda.synchronize()
    total_ms = t0.elapsed_time(t1) / a.iters
    for _ in range(a.iters):
        epoch(True)
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                epoch(False)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), file=sys.stderr)
    # ---- the reference-facing call: NeuralNet.train(examples, gnn_examples) with host example lists (sampling,
    # H2D copies of each minibatch, captured steps) -- 20 "epochs" per call as connect4/config.yaml
    api_ms = None
    if True:  # every rank calls train(): multi-rank steps are captured with their NCCL collectives
        rng = np.random.default_rng(1)
        A = w.action_size
        ex = [(rng.integers(-1, 2, size=(N_BOARD, N_BOARD)).astype(np.int64), rng.dirichlet(np.ones(A)), float(rng.uniform(-1, 1)))
              for _ in range(512)]
        gex = [(b, None, None, None, p_, v_) for b, p_, v_ in ex]
        w.train(ex, gex)  # first call: Adam state + graph capture
        torch.cuda.synchronize()
        t_0 = time.perf_counter()
        for _ in range(3):
            w.train(ex, gex)
        torch.cuda.synchronize()
        api_ms = (time.perf_counter() - t_0) / 3 / int(w.args["epochs"]) * 1e3
    t = torch.tensor([total_ms, api_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    api_ms = float(t[1])
    if rank == 0:
        n_gnn = sum(p.numel() for p in gnn_params)
        n_std = sum(p.numel() for p in nnet_params)
        F = w.feature_dim
        # bytes that must cross HBM (fp32): every GNN weight read once forward and once for the input gradient,
        # every GNN gradient written once; Adam reads p, g, m, v and writes p, m, v
        grad_bytes = 4 * n_gnn
        fwd_bytes = 4 * n_gnn
        line = {"metric": "connect4_gnn_train_epoch_ms", "value": float(t[0]), "unit": "ms per (std step + GNN step), eager phases",
                "n_gpus": world, "batch": a.batch, "iters": a.iters, "higher_is_better": False,
                "phase_ms": {p: acc[p] / a.iters for p in phases},
                "train_api_ms_per_epoch": api_ms,
                "params": {"nnet": n_std, "gnn": n_gnn, "feature_dim": F},
                "hbm_bytes_algorithmic": {"gnn_forward_weights": fwd_bytes, "gnn_backward": 2 * grad_bytes,
                                          "adam": 7 * grad_bytes, "allreduce_payload": 4 * sum(p.numel() for p in w.gnn.output_transform.parameters()) if world > 1 else 0},
                "adam": "torch fused" if a.fused_adam else "torch default (as the reference)"}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm = peaks["hbm_gbs"] * 1e9
            line["hbm_floor_ms"] = {"gnn_fwd": fwd_bytes / hbm * 1e3, "gnn_bwd": 2 * grad_bytes / hbm * 1e3,
                                    "gnn_adam": 7 * grad_bytes / hbm * 1e3}
        except Exception:
            pass
        if a.cpu and world == 1:
            line["cpu_baseline"] = {"value": cpu_epoch(3), "unit": "ms per epoch", "cores": os.cpu_count(), "kind": "port"}
        print(json.dumps(line))
    if world > 1:
        from bench import shutdown_ranks
        shutdown_ranks(w)


if __name__ == "__main__":
    main()
