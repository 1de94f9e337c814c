"""Batched self-play: thousands of concurrent `Coach.executeEpisode` loops (Coach.py:27-79) on
one GPU, advanced in lock step over the arena.

Per move of every live game:  getActionProb (numMCTSSims searches)  ->  training example(s)
->  [use_gnn] expand_tree (expand_by more searches)  ->  sample a ~ pi  ->  play it  ->
finished games are scored, turned into the reference's example tuples and restarted.

Host work is vectorised over games with NumPy but keeps the reference's arithmetic:
visit-count normalisation is the left-to-right Python sum of `x + 1e-8` (MCTS.py:46-57) and
move sampling is `numpy.random.choice`'s inverse-CDF rule (Coach.py:63), one uniform per game
from this object's own Generator (the reference's global-RNG stream is only defined for a
single sequential game; the compat class `azgnn_b200.mcts.MCTS` keeps that behaviour).
"""
import numpy as np
import torch

from . import _lib
from .mcts import BatchedMCTS, arg, pack_states

EPS = 1e-8


def unpack_boards(kind, n, states):
    """[G,2] packed states -> [G,n,n] boards (int64 cells; FrozenLake float64 one-hot), vectorised."""
    s = np.ascontiguousarray(states).view(np.uint64).reshape(-1, 2)
    G = s.shape[0]
    if kind == "frozenlake":
        b = np.zeros((G, n * n))
        b[np.arange(G), s[:, 0].astype(np.int64)] = 1
        return b.reshape(G, n, n)
    bits = np.arange(n * n, dtype=np.uint64)
    m = ((s[:, 0:1] >> bits) & np.uint64(1)).astype(np.int64)
    t = ((s[:, 1:2] >> bits) & np.uint64(1)).astype(np.int64)
    return (m - t).reshape(G, n, n)


def python_float_sum(x):
    """Row-wise `float(sum(row))` as CPython >= 3.12 computes it for floats: left to right with
    Neumaier compensation (bltinmodule.c builtin_sum), which is what MCTS.py:47 runs."""
    f = x[:, 0].copy()
    c = np.zeros_like(f)
    for a in range(1, x.shape[1]):
        v = x[:, a]
        t = f + v
        c = c + np.where(np.abs(f) >= np.abs(v), (f - t) + v, (v - t) + f)
        f = t
    return np.where((c != 0) & np.isfinite(c), f + c, f)


def probs_from_counts(counts, temps, rng):
    """Vectorised MCTS.getActionProb tail (MCTS.py:36-58) for counts [G,A] and temps [G] in {0,1}."""
    G, A = counts.shape
    x = counts.astype(np.float64) + EPS           # (x + EPS) ** 1.0 == x + EPS
    tot = python_float_sum(x)
    probs = x / tot[:, None]
    greedy = np.flatnonzero(np.asarray(temps) == 0)
    if greedy.size:
        c = counts[greedy]
        best = c == c.max(axis=1, keepdims=True)
        # uniform choice among the arg-max actions (np.random.choice(bestAs), MCTS.py:40-41)
        r = rng.random(greedy.size)
        k = (r * best.sum(axis=1)).astype(np.int64)
        pick = (np.cumsum(best, axis=1) > k[:, None]).argmax(axis=1)
        probs[greedy] = 0.0
        probs[greedy, pick] = 1.0
    return probs


def sample_actions(probs, rng, u=None):
    """numpy.random.choice(len(p), p=p) per row: cdf = cumsum(p); cdf /= cdf[-1];
    searchsorted(cdf, u, side='right').  `u` (one uniform per row) may be supplied to replay a
    reference run's RandomState.random_sample() draws."""
    cdf = np.cumsum(probs, axis=1)
    cdf = cdf / cdf[:, -1:]
    u = rng.random(probs.shape[0]) if u is None else np.asarray(u, dtype=np.float64)
    return np.minimum((cdf <= u[:, None]).sum(axis=1), probs.shape[1] - 1).astype(np.int32)


class BatchedSelfPlay:
    def __init__(self, game, nnet, args, n_games, seed=0, collect_examples=True, arena=None, max_episode_steps=None):
        self.game, self.nnet, self.args = game, nnet, args
        self.G = int(n_games)
        self.use_gnn = bool(arg(args, "use_gnn", False))
        self.mcts = BatchedMCTS(game, nnet, args, n_games=self.G, arena=arena)
        self.kind, self.n, self.A = self.mcts.kind, self.mcts.n, self.mcts.A
        self.rng = np.random.default_rng(seed)
        self.collect = collect_examples
        self.temp_threshold = int(arg(args, "tempThreshold", 15))
        self.expand_by = int(arg(args, "expand_by", 5))
        self.two_player = bool(getattr(game, "is_two_player", True))
        self.init_state = pack_states(self.kind, np.asarray(game.getInitBoard())[None])[0]
        # The reference has no step cap (Coach.py:34-79); single-player episodes can wander forever, so an
        # optional cap ends an episode with result 0 (off by default = reference behaviour).
        self.max_episode_steps = max_episode_steps
        self.moves_played = 0
        self.episodes_done = 0
        self._start_all()

    # ------------------------------------------------------------------ episode bookkeeping
    def _start_all(self):
        G = self.G
        self.mcts.reset()
        self.mcts.arena.set_roots(np.tile(self.init_state, (G, 1)))
        self.step = np.zeros(G, dtype=np.int64)      # episodeStep (Coach.py:32-35)
        self.player = np.ones(G, dtype=np.int64)     # curPlayer
        self.history = [[] for _ in range(G)]        # per game: (canonical board, player, pi, gnn record)

    def _restart(self, ids):
        ids = np.asarray(ids, dtype=np.int32)
        self.mcts.reset(ids)                         # new MCTS per episode, Coach.py:96
        roots = self.mcts.arena.to_host(self.mcts.arena.get_roots())
        roots[ids] = self.init_state
        self.mcts.arena.set_roots(roots)
        self.step[ids] = 0
        self.player[ids] = 1
        for g in ids:
            self.history[g] = []

    def _finish(self, g, r):
        """Coach.py:68-79: sign the result for every stored position; symmetries as the reference."""
        cur = self.player[g]
        std, gnn = [], []
        for board, pl, pi, rec in self.history[g]:
            sign = r * ((-1) ** (pl != cur))
            sym = self.game.getSymmetries(board, pi)
            for b, p in sym:
                std.append((b, p, sign))
            if rec is not None:
                ip, iv, ep, ev = rec
                gnn.append((board, pl, ip, iv, ep, ev, sign))
        return std, gnn

    # ------------------------------------------------------------------ one lock-step move
    def step_all(self):
        """Every live game plays one move.  Returns the list of (std_examples, gnn_examples) of the
        episodes that finished on this move."""
        m, G = self.mcts, self.G
        self.step += 1
        temps = (self.step < self.temp_threshold).astype(np.int64)   # Coach.py:37
        m.search(int(arg(self.args, "numMCTSSims")))
        N, _, _ = m.root_stats()
        probs = probs_from_counts(N, temps, self.rng)
        recs = None
        if self.use_gnn:
            recs = m.expand_tree(self.expand_by) if self.collect else self._expand_only()
        if self.collect:
            boards = unpack_boards(self.kind, self.n, m.arena.to_host(m.arena.get_roots()))
            for g in range(G):
                self.history[g].append((boards[g], int(self.player[g]), list(probs[g]),
                                        recs[g] if recs is not None else None))
        actions = sample_actions(probs, self.rng)
        ended = m.advance(actions)
        self.moves_played += G
        if self.two_player:
            self.player = -self.player
        if self.max_episode_steps is not None:
            ended = [e if (e != 0 or self.step[g] < self.max_episode_steps) else 0.0 for g, e in enumerate(ended)]
            done = [g for g in range(G) if ended[g] != 0 or self.step[g] >= self.max_episode_steps]
        else:
            done = [g for g in range(G) if ended[g] != 0]
        out = []
        for g in done:
            out.append(self._finish(g, ended[g]) if self.collect else ([], []))
        if done:
            self.episodes_done += len(done)
            self._restart(done)
        return out

    def _expand_only(self):
        """expand_tree's searches without building the example records (throughput runs)."""
        if self.mcts.device_eval:
            self.mcts.nnet.forward_states(self.mcts.arena.get_roots(), _lib.EVAL_STD)  # MCTS.py:108-111
        self.mcts.search(self.expand_by)
        return None

    def play(self, n_episodes):
        """Run until n_episodes episodes have finished; returns their example lists."""
        finished = []
        while len(finished) < n_episodes:
            finished.extend(self.step_all())
        return finished[:n_episodes]
