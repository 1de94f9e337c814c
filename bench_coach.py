#!/usr/bin/env python
"""bench_coach.py -- BASELINE configs[3]: one full Coach iteration of Connect4 (connect4/config.yaml search and
training settings) on N GPUs: self-play episodes and arena games sharded over the ranks with no communication,
examples all-gathered once, the training step data-parallel with an NCCL gradient all-reduce (Coach.py:87-176).

    python bench_coach.py [--eps-per-gpu 2048] [--games 2048]                 one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_coach.py ...

The reference's config plays numEps = 20 episodes per iteration; a GPU needs thousands of concurrent games to be
busy, so the episode count is scaled up (stated in the output) -- per-episode and per-move work is unchanged.
Two iterations run; the timings are those of the SECOND (the first pays one-time costs: torch's lazy imports behind
the fused/capturable Adam (3.7 s), CUDA-graph capture, weight re-tiling).
`--cpu-episodes k` also times k sequential episodes of the oracle's CPU restatement (rank 0).
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


class Args(dict):
    __getattr__ = dict.__getitem__


def coach_iteration(eps_per_gpu, games, arena, precision, rank, world, dev, iters=2):
    """Runs `iters` Coach iterations (process group already initialised by the caller when world > 1) and returns, on
    rank 0, the record of the LAST one: phase times (max over ranks), self-play moves/s (sum over ranks), example counts."""
    from azgnn_b200.coach import Coach
    from azgnn_b200.games import Connect4Game
    from azgnn_b200.nets import B200Connect4GNNWrapper
    folder = [tempfile.mkdtemp(prefix="azg_coach_")] if rank == 0 else [None]
    if world > 1:
        dist.broadcast_object_list(folder, src=0)
    a = Args(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5,
             tempThreshold=15, numIters=iters, numEps=eps_per_gpu * world, n_parallel_games=games, maxlenOfQueue=200000,
             numItersForTrainExamplesHistory=5, arenaCompare=arena, updateThreshold=0.6, checkpoint=folder[0],
             b200_precision=precision, save_examples=False)
    game = Connect4Game(7)
    torch.manual_seed(0)
    np.random.seed(rank)
    coach = Coach(game, B200Connect4GNNWrapper(game, a), a)
    coach.learn()
    from azgnn_b200.training import release_captured_steps
    release_captured_steps(coach.nnet)  # graphs with NCCL kernels must not outlive the process group
    t = coach.timings
    tt = torch.tensor([t[k] for k in sorted(t)] + [float(coach.selfplay_moves)], dtype=torch.float64, device=dev)
    mx, sm = tt.clone(), tt.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    import shutil
    shutil.rmtree(folder[0], ignore_errors=True)
    keys = sorted(t)
    phases = {k: mx[i].item() for i, k in enumerate(keys)}
    moves = sm[-1].item()
    n_ot = sum(p.numel() for p in coach.nnet.gnn.output_transform.parameters())
    return {"metric": "connect4_coach_iteration", "n_gpus": world, "episodes": a.numEps, "concurrent_games_per_gpu": games,
            "selfplay_moves": moves, "selfplay_moves_per_s": moves / phases["selfplay_s"], "phases_s": phases,
            "examples": {"standard": len(coach.trainExamplesHistory[-1][0]), "gnn": len(coach.trainExamplesHistory[-1][1])},
            "iteration": iters, "train_epochs": a.epochs, "arena_games": arena, "arena_prev_new_draws": coach.arena_result,
            "precision": precision, "chosen_precision": coach.nnet.precision_report,
            "allreduce_payload_bytes_per_gnn_step": 4 * n_ot if world > 1 else 0,
            "config": "connect4/config.yaml: 10 sims + 5 expand_by, cpuct 1.0, tempThreshold 15, 20 epochs x batch 64, "
                      "arenaCompare 100; numEps scaled from 20 to keep the GPUs busy"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--eps-per-gpu", type=int, default=2048)
    ap.add_argument("--games", type=int, default=2048, help="concurrent games per GPU")
    ap.add_argument("--arena", type=int, default=100)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--cpu-episodes", type=int, default=0)
    o = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = coach_iteration(o.eps_per_gpu, o.games, o.arena, o.precision, rank, world, dev)
    if rank == 0:
        if o.cpu_episodes:
            from bench import cpu_selfplay
            mps, mv, dt = cpu_selfplay(o.cpu_episodes)
            line["cpu_baseline"] = {"selfplay_moves_per_s": mps, "episodes": o.cpu_episodes, "cores": os.cpu_count(), "kind": "port"}
        print(json.dumps(line))
    if world > 1:
        from bench import shutdown_ranks
        shutdown_ranks()


if __name__ == "__main__":
    main()
