"""oracle.mcts -- TEST INFRASTRUCTURE: restatement of the reference MCTS (MCTS.py:10-240).

Same public surface (``getActionProb``, ``expand_tree``, ``search`` and the dicts
``Qsa Nsa Ns Ps Es Vs``), same NumPy >= 2 (NEP 50) value types, but the recursion
of MCTS.search (MCTS.py:151-240) is unrolled into descend + backup loops so the
structure matches what the GPU arena does (select -> leaf -> expand/backup).

Deviation from the reference, explicit and off by default: ``max_depth``.  The
reference has no depth cap and FrozenLake searches recurse forever (SURVEY.md
section 0 item 7).  With ``max_depth=d`` a search call entered at depth >= d returns
the Python int 0 (the same value MCTS.py:218-219 returns for "no action").  With
``max_depth=None`` behaviour is the reference's.

Pinned by tests/test_oracle_golden.py against tests/golden/mcts_*.npz (dumps of
the reference's dicts under a deterministic fake net).
"""
import logging
import math

import numpy as np

EPS = 1e-8  # MCTS.py:6
log = logging.getLogger(__name__)


class OracleMCTS:
    def __init__(self, game, nnet, args, max_depth=None):
        self.game, self.nnet, self.args = game, nnet, args
        self.max_depth = max_depth
        self.Qsa, self.Nsa, self.Ns, self.Ps, self.Es, self.Vs = {}, {}, {}, {}, {}, {}
        self.standard_predictions, self.gnn_predictions = {}, {}
        self.expanded, self.expanded_nodes = False, {}
        self.n_leaf_evals = 0

    # ------------------------------------------------------------------ MCTS.py:29-58
    def getActionProb(self, canonicalBoard, temp=1):
        self.standard_predictions, self.gnn_predictions = {}, {}
        for _ in range(self.args.numMCTSSims):
            self.search(canonicalBoard)
        s = self.game.stringRepresentation(canonicalBoard)
        counts = [self.Nsa.get((s, a), 0) for a in range(self.game.getActionSize())]
        return self.probs_from_counts(counts, temp, canonicalBoard)

    def probs_from_counts(self, counts, temp, canonicalBoard):
        if temp == 0:  # :39-44
            bestAs = np.array(np.argwhere(counts == np.max(counts))).flatten()
            bestA = np.random.choice(bestAs)
            probs = [0] * len(counts)
            probs[bestA] = 1
            return probs
        counts = [(x + EPS) ** (1. / temp) for x in counts]  # :46
        counts_sum = float(sum(counts))
        if counts_sum <= 0:  # :49-55
            valids = self.game.getValidMoves(canonicalBoard, 1)
            vs = np.sum(valids)
            return valids / vs if vs > 0 else np.ones(len(counts)) / len(counts)
        return [x / counts_sum for x in counts]

    # ------------------------------------------------------------------ MCTS.py:60-149
    def _root_counts(self, s):
        # `for key in self.Nsa: if key[0] == s` -- dict order = first-visit order of the edges
        return {k[1]: n for k, n in self.Nsa.items() if k[0] == s}

    def expand_tree(self, canonicalBoard, expand_by=5):
        s = self.game.stringRepresentation(canonicalBoard)
        A = self.game.getActionSize()
        self.expanded, self.expanded_nodes = True, {}
        init = self._root_counts(s)
        if not init:  # :83-92
            for _ in range(self.args.numMCTSSims):
                self.search(canonicalBoard)
            init = self._root_counts(s)
        initial_policy = np.zeros(A)
        for a, c in init.items():
            initial_policy[a] = c
        isum = np.sum(initial_policy)
        if isum > 0:
            initial_policy = initial_policy / isum
        else:
            valids = self.game.getValidMoves(canonicalBoard, 1)
            initial_policy = valids / np.sum(valids)
        if s not in self.standard_predictions:  # :108-111 (cache cleared by getActionProb)
            self.standard_predictions[s] = self.nnet.predict(canonicalBoard)
        initial_value = self.standard_predictions[s][1]
        for _ in range(expand_by):  # :116-117
            self.search(canonicalBoard)
        expanded_policy = np.zeros(A)
        for a, c in self._root_counts(s).items():
            expanded_policy[a] = c
        esum = np.sum(expanded_policy)
        expanded_policy = expanded_policy / esum if esum > 0 else initial_policy
        expanded_value, valid_count = 0, 0  # :132-143
        for a in range(A):
            if (s, a) in self.Qsa and (s, a) in self.Nsa and self.Nsa[(s, a)] > 0:
                expanded_value += self.Qsa[(s, a)] * self.Nsa[(s, a)]
                valid_count += self.Nsa[(s, a)]
        expanded_value = expanded_value / valid_count if valid_count > 0 else initial_value
        self.expanded_nodes[s] = (initial_policy, initial_value, expanded_policy, expanded_value)
        self.expanded = False
        return self.expanded_nodes

    # ------------------------------------------------------------------ MCTS.py:151-240
    def _use_gnn(self):
        return hasattr(self.args, 'use_gnn') and self.args.use_gnn

    def _expand_leaf(self, s, board):
        """MCTS.py:162-200: valid mask, the two net calls, mask+renormalise in float64."""
        if s not in self.Vs:
            self.Vs[s] = self.game.getValidMoves(board, 1)
        valids = self.Vs[s]
        try:
            std = self.nnet.predict(board)
            self.standard_predictions[s] = std
            self.n_leaf_evals += 1
            if self._use_gnn():
                gnn = self.nnet.predict_with_gnn(board)
                self.gnn_predictions[s] = gnn
                prior, v = gnn
            else:
                prior, v = std
            ps = prior * valids  # float32 * int64 -> float64 (NEP 50)
            tot = np.sum(ps)
            if tot > 0:
                ps /= tot
            else:
                log.warning("All valid moves were masked, using uniform policy")
                ps = valids / np.sum(valids)
            self.Ps[s] = ps
            self.Ns[s] = 0
            return v
        except Exception as e:  # :195-200 silent fallback kept for fidelity
            log.error(f"Error in neural network prediction: {e}")
            self.Ps[s] = valids / np.sum(valids)
            self.Ns[s] = 0
            return 0

    def _select(self, s):
        """MCTS.py:202-216: strict '>' scan in ascending action order."""
        valids, best, best_a = self.Vs[s], -float('inf'), -1
        for a in range(self.game.getActionSize()):
            if valids[a]:
                if (s, a) in self.Qsa:
                    u = self.Qsa[(s, a)] + self.args.cpuct * self.Ps[s][a] * math.sqrt(self.Ns[s]) / (
                        1 + self.Nsa[(s, a)])
                else:
                    u = self.args.cpuct * self.Ps[s][a] * math.sqrt(self.Ns[s] + EPS)
                if u > best:
                    best, best_a = u, a
        return best_a

    def search(self, canonicalBoard, expansion=False):
        game = self.game
        path, board = [], canonicalBoard
        while True:
            if self.max_depth is not None and len(path) >= self.max_depth:
                v = 0
                break
            s = game.stringRepresentation(board)
            if s not in self.Es:
                self.Es[s] = game.getGameEnded(board, 1)
            if self.Es[s] != 0:
                v = self.Es[s]
                break
            if expansion and self.Ns.get(s, 0) >= self.args.numMCTSSims:  # dead in the reference
                v = 0
                break
            if s not in self.Ps:
                v = self._expand_leaf(s, board)
                break
            a = self._select(s)
            if a == -1:
                v = 0
                break
            path.append((s, a))
            nxt, nplayer = game.getNextState(board, 1, a)
            board = game.getCanonicalForm(nxt, nplayer)
        two = hasattr(game, 'is_two_player') and game.is_two_player
        for (s, a) in reversed(path):  # :228-240 on the way back up
            if (s, a) in self.Qsa:
                self.Qsa[(s, a)] = (self.Nsa[(s, a)] * self.Qsa[(s, a)] + v) / (self.Nsa[(s, a)] + 1)
                self.Nsa[(s, a)] += 1
            else:
                self.Qsa[(s, a)] = v
                self.Nsa[(s, a)] = 1
            self.Ns[s] += 1
            v = -v if two else v
        return v


class FakeNet:
    """Deterministic stand-in for a NeuralNet (SURVEY.md section 4 'fake-net fixture'): priors
    and values are a pure function of the board bytes, returned with the reference's types
    (np.float32 vector, np.float32 scalar; FrozenLake style returns v as float32[1])."""

    def __init__(self, action_size, salt=0, v_as_array=False, spread=1.0):
        self.action_size, self.salt, self.v_as_array, self.spread = action_size, salt, v_as_array, spread
        self.calls = 0

    def _rng(self, board, which):
        import zlib
        h = zlib.crc32(np.ascontiguousarray(board).astype(np.int8).tobytes()) ^ (self.salt * 2654435761 & 0xFFFFFFFF)
        return np.random.Generator(np.random.PCG64([h, which]))

    def _eval(self, board, which):
        self.calls += 1
        g = self._rng(board, which)
        logits = (g.standard_normal(self.action_size) * self.spread).astype(np.float32)
        e = np.exp(logits - logits.max()).astype(np.float32)
        pi = (e / e.sum(dtype=np.float32)).astype(np.float32)
        v = np.float32(np.tanh(g.standard_normal()))
        if self.v_as_array:
            v = np.array([v], dtype=np.float32)
        return pi, v

    def predict(self, board):
        return self._eval(board, 0)

    def predict_with_gnn(self, board):
        return self._eval(board, 1)
