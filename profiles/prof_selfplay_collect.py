"""cProfile of lock-step self-play WITH example collection (the Coach path): 16,384 Connect4 games, default precision,
steady state (AZG_PROF_WARM move-steps first, default 50: every slot has turned over)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from azgnn_b200.games import Connect4Game
from azgnn_b200.nets import B200Connect4GNNWrapper
from azgnn_b200.selfplay import BatchedSelfPlay

a = dict(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=True, numMCTSSims=10, cpuct=1.0, expand_by=5,
         tempThreshold=15)
game = Connect4Game(7)
torch.manual_seed(0)
net = B200Connect4GNNWrapper(game, a)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for collect in (False, "device"):
    sp = BatchedSelfPlay(game, net, a, G, seed=0, collect_examples=collect)
    for _ in range(int(os.environ.get("AZG_PROF_WARM", "50"))):
        sp.step_all()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for _ in range(30):
        sp.step_all()
    torch.cuda.synchronize()
    pr.disable()
    dt = time.perf_counter() - t0
    print(f"collect={collect}: {30 * G / dt:.0f} moves/s, {dt / 30 * 1e3:.1f} ms per move-step")
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
