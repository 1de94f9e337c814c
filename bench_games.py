#!/usr/bin/env python
"""bench_games.py -- BASELINE configs[0] and configs[2]: TicTacToe (4x4 as shipped
in tictactoe/config.yaml, and 3x3) and FrozenLake (4x4, 8x8) through the same arena + batched leaf evaluation,
next to the reference's CPU path (oracle restatement: sequential MCTS with one B=1 network call per leaf).

Per configuration, one JSON line:
  leaf_evals_per_s     batched network forward on device-resident states (CUDA events)
  moves_per_s, sims_per_s   lock-step self-play of G concurrent games/trees with the config's search settings
  cpu_baseline         moves/s and leaf evals/s of the sequential CPU path on a bounded sample

FrozenLake cycle policy: the reference's search recurses forever once a simulation cycles;
here a search call entered at depth >= 4 n^2 returns 0, and episodes are capped at
4 n^2 moves (result 0) so that random-init policies terminate.
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


class Args(dict):
    __getattr__ = dict.__getitem__


def ttt_args(use_gnn, precision="bf16x3"):
    return Args(lr=1e-3, dropout=0.3, epochs=20, batch_size=64, gnn_layers=2, use_gnn=use_gnn, numMCTSSims=10, cpuct=1.0,
                expand_by=5, tempThreshold=15, b200_precision=precision)


def fl_args():
    return Args(lr=1e-3, dropout=0.3, epochs=20, batch_size=32, embedding_dim=128, gnn_layers=3, numMCTSSims=50, cpuct=2.0,
                tempThreshold=15, use_gnn=False)


class OracleTTTNet:
    def __init__(self, w, n):
        self.p, self.g, self.n = dict(w.nnet.state_dict()), dict(w.gnn.state_dict()) if w.gnn is not None else None, n
        self.p = {k: v.cpu() for k, v in self.p.items()}
        self.g = {k: v.cpu() for k, v in self.g.items()} if self.g else None

    def predict(self, board):
        from oracle import nets as onets
        with torch.no_grad():
            pi, v = onets.ttt_predict(self.p, onets.boards_to_tensor(np.asarray(board)[None]), self.n)
        return pi.numpy()[0], v.numpy()[0]

    def predict_with_gnn(self, board):
        from oracle import nets as onets
        with torch.no_grad():
            pi, v = onets.ttt_predict_with_gnn(self.p, self.g, onets.boards_to_tensor(np.asarray(board)[None]), self.n)
        return pi.numpy()[0], v.numpy()[0]


class OracleFLNet:
    def __init__(self, w, n, layers):
        self.p, self.n, self.layers = {k: v.cpu() for k, v in w.nnet.state_dict().items()}, n, layers

    def predict(self, board):
        from oracle import nets as onets
        with torch.no_grad():
            pi, v = onets.fl_predict_cell(self.p, int(np.argmax(np.asarray(board).reshape(-1))), self.n, self.layers)
        return pi.numpy(), v.numpy()


def cpu_selfplay(game, net, a, episodes, max_depth, max_moves):
    from oracle.mcts import OracleMCTS
    from oracle.selfplay import execute_episode
    torch.set_num_threads(os.cpu_count() or 1)
    np.random.seed(0)
    moves, evals, t0 = 0, 0, time.perf_counter()
    for _ in range(episodes):
        m = OracleMCTS(game, net, a, max_depth=max_depth)
        k, _r = execute_episode(game, m, a, max_moves=max_moves)
        moves += k
        evals += m.n_leaf_evals
    dt = time.perf_counter() - t0
    return moves / dt, evals / dt, moves, dt


def run(name, game, w, a, G, leaf_batch, moves, cpu_episodes, max_steps, oracle_net, max_depth):
    from azgnn_b200 import _lib
    from azgnn_b200.selfplay import BatchedSelfPlay
    dev = w.device
    n = game.getBoardSize()[0]
    rng = np.random.default_rng(0)
    if w.kind == "frozenlake":
        boards = np.zeros((leaf_batch, n, n))
        boards.reshape(leaf_batch, -1)[np.arange(leaf_batch), rng.integers(0, n * n, leaf_batch)] = 1
    else:
        boards = rng.integers(-1, 2, size=(leaf_batch, n, n)).astype(np.int8)
    states = w.states_from_boards(boards)
    for _ in range(3):
        w.forward_states(states)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        w.forward_states(states)
    e1.record()
    torch.cuda.synchronize()
    leaf_ms = e0.elapsed_time(e1) / reps

    sp = BatchedSelfPlay(game, w, a, G, seed=0, collect_examples=False, max_episode_steps=max_steps)
    for _ in range(2):
        sp.step_all()
    torch.cuda.synchronize()
    m0, l0 = sp.moves_played, sp.mcts.leaf_evaluations()
    e0.record()
    for _ in range(moves):
        sp.step_all()
    e1.record()
    torch.cuda.synchronize()
    sp_ms = e0.elapsed_time(e1)
    played = sp.moves_played - m0
    sims = int(a.numMCTSSims) + (int(a.expand_by) if a.use_gnn else 0)
    cpu_mps, cpu_eps, cpu_moves, cpu_dt = cpu_selfplay(game, oracle_net, a, cpu_episodes, max_depth, max_steps or 10_000)
    line = {"config": name, "board": f"{n}x{n}", "leaf_batch": leaf_batch, "leaf_evals_per_s": leaf_batch / (leaf_ms / 1e3),
            "leaf_ms": leaf_ms, "concurrent_games": G, "sims_per_move": sims,
            "moves_per_s": played / (sp_ms / 1e3), "sims_per_s": played * sims / (sp_ms / 1e3),
            "leaf_evals_per_s_in_search": (sp.mcts.leaf_evaluations() - l0) / (sp_ms / 1e3),
            "ms_per_move_step": sp_ms / moves, "episode_step_cap": max_steps,
            "cpu_baseline": {"moves_per_s": cpu_mps, "leaf_evals_per_s": cpu_eps, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{cpu_episodes} sequential episodes ({cpu_moves} moves, {cpu_dt:.1f} s)"}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=16384)
    ap.add_argument("--leaf-batch", type=int, default=65536)
    ap.add_argument("--moves", type=int, default=6)
    ap.add_argument("--cpu-episodes", type=int, default=2)
    a = ap.parse_args()
    from azgnn_b200 import games
    from azgnn_b200.nets import B200FrozenLakeNet, B200TicTacToeGNNWrapper, B200TicTacToeNNetWrapper
    torch.cuda.set_device(0)
    for n, use_gnn, prec in ((4, False, "bf16x3"), (4, True, "bf16x3"), (3, True, "bf16x3"), (4, True, "fp32")):
        game, ar = games.TicTacToeGame(n), ttt_args(use_gnn, prec)
        torch.manual_seed(0)
        w = (B200TicTacToeGNNWrapper if use_gnn else B200TicTacToeNNetWrapper)(game, ar)
        run(f"tictactoe_{n}x{n}_{'gnn' if use_gnn else 'std'}_{prec}", game, w, ar, a.games, a.leaf_batch, a.moves, a.cpu_episodes,
            None, OracleTTTNet(w, n), None)
    for n in (4, 8):
        game, ar = games.FrozenLakeGame(n), fl_args()
        torch.manual_seed(0)
        w = B200FrozenLakeNet(game, ar)
        run(f"frozenlake_{n}x{n}", game, w, ar, a.games, a.leaf_batch, a.moves, a.cpu_episodes, 4 * n * n,
            OracleFLNet(w, n, 3), 4 * n * n)


if __name__ == "__main__":
    main()
