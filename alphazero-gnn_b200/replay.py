"""Device-resident training examples: the pipeline between self-play and `NeuralNet.train`
(SURVEY.md section 8f.1).

    DeviceExamples      a list of (board, pi, v) examples living in HBM as columns
                        states int64[N,2] | pi float64[N,A] | v float64[N] | vtag int8[N] | sym int8[N]
      .emit(...)        Coach.executeEpisode's tail (Coach.py:45-49, 68-79) for finished episodes: symmetries
                        (Connect4Game.py:189-215, TicTacToeGame.py:187-200) + value signing, one kernel launch
      .sample(B)        the minibatch of `train` (Connect4GNN.py:141-148): np.random.randint with replacement
                        (the reference's RNG call) + one gather kernel -> float32 boards [B,n,n], pi [B,A], v [B]
      .to_examples()    the reference's tuples, with the reference's Python/NumPy value types, for the pickle
      .from_examples()  format of Coach.saveTrainExamples / loadTrainExamples (Coach.py:178-201)

Symmetries are cell permutations.  Their tables are built by applying the game's own `getSymmetries` (the host
mirror of the reference code, numpy calls included) to an index board and an index policy, so the kernel carries
no game knowledge and the C4 axis quirk comes along for free.
"""
import random

import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream
from .mcts import game_kind, pack_states, typed_value, unpack_state

_I8 = torch.int8


def symmetry_tables(game):
    """(board_perm [S, n*n] int32, pi_perm [S, A] int32): output cell d of symmetry s shows input cell board_perm[s, d]"""
    n = game.getBoardSize()[0]
    A = game.getActionSize()
    idx_board = np.arange(n * n, dtype=np.int64).reshape(n, n)
    forms = game.getSymmetries(idx_board, list(range(A)))
    bp = np.stack([np.asarray(b, dtype=np.int64).reshape(-1) for b, _ in forms]).astype(np.int32)
    pp = np.stack([np.asarray(p, dtype=np.int64).reshape(-1) for _, p in forms]).astype(np.int32)
    return bp, pp


class DeviceExamples:
    def __init__(self, game, device=None):
        self.game = game
        self.kind = game_kind(game)
        self.n = game.getBoardSize()[0]
        self.A = game.getActionSize()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        bp, pp = symmetry_tables(game)
        self.S = bp.shape[0]
        self._bp = torch.from_numpy(bp).to(self.device)
        self._pp = torch.from_numpy(pp).to(self.device)
        d = self.device
        self.states = torch.empty(0, 2, dtype=torch.int64, device=d)
        self.pi = torch.empty(0, self.A, dtype=torch.float64, device=d)
        self.v = torch.empty(0, dtype=torch.float64, device=d)
        self.vtag = torch.empty(0, dtype=_I8, device=d)
        self.sym = torch.empty(0, dtype=_I8, device=d)

    def __len__(self):
        return int(self.states.shape[0])

    def _append(self, states, pi, v, vtag, sym):
        self.states = torch.cat([self.states, states])
        self.pi = torch.cat([self.pi, pi])
        self.v = torch.cat([self.v, v])
        self.vtag = torch.cat([self.vtag, vtag])
        self.sym = torch.cat([self.sym, sym])

    # ------------------------------------------------------------------ Coach.py:45-49, 68-79
    def emit(self, states, pi, player, game, result, result_tag, cur, pi_int=None):
        """states int64 [E,2], pi float64 [E,A], player int32 [E], game int32 [E] (row into the per-finished-game
        tables result float64 / result_tag int8 / cur int32), all on the device.  Appends E*S examples in the
        reference's order (entry by entry, symmetry by symmetry).  pi_int [E] int8: the policy came from the temp = 0
        branch of getActionProb, whose list holds Python ints (MCTS.py:39-44) -- only the exported types differ."""
        E = int(states.shape[0])
        if E == 0:
            return
        d, S, A = self.device, self.S, self.A
        out_states = torch.empty(E * S, 2, dtype=torch.int64, device=d)
        out_pi = torch.empty(E * S, A, dtype=torch.float64, device=d)
        out_v = torch.empty(E * S, dtype=torch.float64, device=d)
        out_tag = torch.empty(E * S, dtype=_I8, device=d)
        fl = int(self.kind == "frozenlake")
        _lib.check(_lib.lib().azg_emit_examples(fl, ptr(states.contiguous()), ptr(pi.contiguous()), ptr(player.contiguous()),
                                                ptr(game.contiguous()), ptr(result.contiguous()), ptr(result_tag.contiguous()),
                                                ptr(cur.contiguous()), E, self.n * self.n, A, S, ptr(self._bp), ptr(self._pp),
                                                ptr(out_states), ptr(out_pi), ptr(out_v), ptr(out_tag), stream()))
        sym = torch.arange(S, dtype=_I8, device=d).repeat(E)
        if pi_int is not None:
            sym = sym + 16 * pi_int.to(_I8).repeat_interleave(S)
        self._append(out_states, out_pi, out_v, out_tag, sym)

    def extend(self, other):
        self._append(other.states, other.pi, other.v, other.vtag, other.sym)

    def shuffled(self):
        """`random.shuffle` of the example list (Coach.py:118): the swaps depend only on the length, so shuffling an
        index list with the same `random` state gives the permutation the reference's list would get."""
        idx = list(range(len(self)))
        random.shuffle(idx)
        t = torch.as_tensor(idx, dtype=torch.int64, device=self.device)
        out = DeviceExamples.__new__(DeviceExamples)
        out.__dict__.update(self.__dict__)
        out.states, out.pi, out.v = self.states[t], self.pi[t], self.v[t]
        out.vtag, out.sym = self.vtag[t], self.sym[t]
        return out

    # ------------------------------------------------------------------ Connect4GNN.py:141-148
    def sample(self, batch_size, idx=None):
        """float32 boards [B,n,n], pi [B,A], v [B] of examples drawn with replacement through the global NumPy RNG
        (`np.random.randint(len(examples), size=...)`), or of the given indices."""
        if idx is None:
            idx = np.random.randint(0, len(self), min(len(self), batch_size))
        B = len(idx)
        d = self.device
        t = torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(d)
        boards = torch.empty(B, self.n, self.n, dtype=torch.float32, device=d)
        pi = torch.empty(B, self.A, dtype=torch.float32, device=d)
        v = torch.empty(B, dtype=torch.float32, device=d)
        _lib.check(_lib.lib().azg_gather_examples(int(self.kind == "frozenlake"), ptr(self.states), ptr(self.pi), ptr(self.v),
                                                  ptr(t), B, self.n * self.n, self.A, ptr(boards), ptr(pi), ptr(v), stream()))
        return boards, pi, v

    # ------------------------------------------------------------------ pickle format (Coach.py:178-201)
    def to_examples(self):
        """list of (board, pi, v) with the reference's container and scalar types: Connect4's first form keeps the
        Python list from getActionProb and its mirror is the ndarray made by np.copy (Connect4Game.py:208-214);
        TicTacToe's forms are lists of numpy float64 with the pass entry a Python float (TicTacToeGame.py:199)."""
        states = self.states.cpu().numpy()
        pis = self.pi.cpu().numpy()
        vs, tags, syms = self.v.cpu().numpy(), self.vtag.cpu().numpy(), self.sym.cpu().numpy()
        out = []
        for i in range(len(states)):
            board = unpack_state(self.kind, self.n, states[i])
            row = pis[i]
            s, ints = int(syms[i]) & 15, bool(int(syms[i]) & 16)
            if self.kind == "connect4":
                if s == 0:
                    pi = [int(x) for x in row] if ints else [float(x) for x in row]
                else:
                    pi = row.astype(np.int64) if ints else row.copy()
            elif self.kind == "tictactoe":
                pi = ([np.int64(x) for x in row[:-1]] + [int(row[-1])]) if ints else ([np.float64(x) for x in row[:-1]] + [float(row[-1])])
            else:
                pi = [int(x) for x in row] if ints else [float(x) for x in row]
            out.append((board, pi, typed_value(vs[i], int(tags[i]))))
        return out

    @classmethod
    def from_examples(cls, game, examples, device=None):
        ex = cls(game, device)
        if not examples:
            return ex
        boards = np.stack([np.asarray(e[0]) for e in examples])
        d = ex.device
        tags, syms = [], []
        for e in examples:
            v = e[2]
            tags.append(_lib.TAG_F32 if isinstance(v, np.floating) and not isinstance(v, float) else
                        _lib.TAG_PYINT if isinstance(v, (int, np.integer)) else _lib.TAG_PYFLOAT)
            first = np.asarray(e[1]).reshape(-1)[0]
            ints = 16 if isinstance(first, (int, np.integer)) and not isinstance(first, bool) else 0
            syms.append((1 if (ex.kind == "connect4" and isinstance(e[1], np.ndarray)) else 0) + ints)
        ex._append(torch.as_tensor(pack_states(ex.kind, boards)).to(d),
                   torch.as_tensor(np.array([np.asarray(e[1], dtype=np.float64) for e in examples])).to(d),
                   torch.as_tensor(np.array([float(e[2]) for e in examples], dtype=np.float64)).to(d),
                   torch.as_tensor(np.array(tags, dtype=np.int8)).to(d), torch.as_tensor(np.array(syms, dtype=np.int8)).to(d))
        return ex
