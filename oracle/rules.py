"""oracle.rules -- TEST INFRASTRUCTURE: numpy restatement of the reference game rules.

Each class exposes the reference ``Game`` API (Game.py:14-113) so the oracle
MCTS and the parity tests read like the reference's own call sites.  The
bodies are written independently (shift-and-compare line detection instead of
nested Python scans) but must return *identical values and types*:

  Connect4   connect4/Connect4Game.py:116-219  (+ Board :38-110)
  TicTacToe  tictactoe/TicTacToeGame.py:122-204 (+ Board :37-116)
  FrozenLake frozenlake/FrozenLakeGame.py:60-202

Pinned by tests/test_oracle_golden.py against tests/golden/rules_*.npz
(generated from the reference by tests/golden/make_golden.py).
"""
import numpy as np

DRAW_VALUE = 1e-4  # Connect4Game.py:183, TicTacToeGame.py:181


def _line_of(mask, length, dx, dy):
    """True if boolean n x n `mask` holds `length` consecutive cells stepping (dx, dy)."""
    n = mask.shape[0]
    xs = range(0, n - (length - 1) * dx) if dx >= 0 else range(-(length - 1) * dx, n)
    ys = range(0, n - (length - 1) * dy) if dy >= 0 else range(-(length - 1) * dy, n)
    for x in xs:
        for y in ys:
            ok = True
            for i in range(length):
                if not mask[x + i * dx, y + i * dy]:
                    ok = False
                    break
            if ok:
                return True
    return False


class Connect4Rules:
    """connect4/Connect4Game.py:116-219.  Board is [column][row], row 0 = bottom."""
    is_two_player = True  # Connect4Game.py:121

    def __init__(self, board_size=7):
        self.board_size = board_size

    def getInitBoard(self):
        return np.zeros((self.board_size, self.board_size), dtype=np.int64)  # :129-132

    def getBoardSize(self):
        return (self.board_size, self.board_size)

    def getActionSize(self):
        return self.board_size + 1  # :139-141 columns + pass

    def getNextState(self, board, player, action):
        n = self.board_size
        if action == n:  # pass, :146-147 (returns the SAME array object)
            return (board, -player)
        nxt = np.copy(board)
        col = nxt[action]
        empties = np.flatnonzero(col == 0)
        assert empties.size > 0, "Column is full!"  # Board.execute_move :109
        col[empties[0]] = player
        return (nxt, -player)

    def getValidMoves(self, board, player):
        n = self.board_size
        valids = [0] * (n + 1)
        open_cols = [x for x in range(n) if board[x][n - 1] == 0]  # Board.get_legal_moves :46-48
        if not open_cols:
            valids[-1] = 1
        for x in open_cols:
            valids[x] = 1
        return np.array(valids)

    @staticmethod
    def _wins(board, color, n):
        k = min(4, n)  # Board.is_win :73
        m = (np.asarray(board) == color)
        return (_line_of(m, k, 1, 0) or _line_of(m, k, 0, 1)
                or _line_of(m, k, 1, -1) or _line_of(m, k, 1, 1))

    def getGameEnded(self, board, player):
        n = self.board_size
        if self._wins(board, player, n):
            return 1
        if self._wins(board, -player, n):
            return -1
        if any(board[x][n - 1] == 0 for x in range(n)):
            return 0
        return DRAW_VALUE

    def getCanonicalForm(self, board, player):
        return player * board

    def stringRepresentation(self, board):
        return board.tobytes()


class TicTacToeRules:
    """tictactoe/TicTacToeGame.py:122-204.  action a -> cell (a // n, a % n)."""
    is_two_player = True

    def __init__(self, n=3):
        self.n = n

    def getInitBoard(self):
        return np.zeros((self.n, self.n), dtype=np.int64)

    def getBoardSize(self):
        return (self.n, self.n)

    def getActionSize(self):
        return self.n * self.n + 1

    def getNextState(self, board, player, action):
        n = self.n
        if action == n * n:
            return (board, -player)
        nxt = np.copy(board)
        x, y = int(action / n), action % n  # :153
        assert nxt[x][y] == 0
        nxt[x][y] = player
        return (nxt, -player)

    def getValidMoves(self, board, player):
        n = self.n
        valids = [0] * (n * n + 1)
        empt = np.argwhere(np.asarray(board) == 0)
        if len(empt) == 0:
            valids[-1] = 1
            return np.array(valids)
        for x, y in empt:
            valids[n * x + y] = 1
        return np.array(valids)

    @staticmethod
    def _wins(board, color, n):
        m = (np.asarray(board) == color)
        return bool(m.all(axis=0).any() or m.all(axis=1).any()
                    or np.diag(m).all() or np.diag(m[:, ::-1]).all())

    def getGameEnded(self, board, player):
        n = self.n
        if self._wins(board, player, n):
            return 1
        if self._wins(board, -player, n):
            return -1
        if (np.asarray(board) == 0).any():
            return 0
        return DRAW_VALUE

    def getCanonicalForm(self, board, player):
        return player * board

    def stringRepresentation(self, board):
        return board.tobytes()


# The two standard gymnasium FrozenLake maps (public constants); the reference
# reads them through gym.make(...).unwrapped.desc (FrozenLakeGame.py:26-43).
FROZENLAKE_MAPS = {
    4: ["SFFF", "FHFH", "FFFH", "HFFG"],
    8: ["SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF",
        "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG"],
}
_FL_DIRS = [(-1, 0), (0, 1), (1, 0), (0, -1)]  # up, right, down, left (:104)


class FrozenLakeRules:
    """frozenlake/FrozenLakeGame.py:60-202 without gymnasium (map constants above)."""

    def __init__(self, map_size=4, custom_map=None):
        self.is_two_player = False  # :18
        rows = custom_map if custom_map is not None else FROZENLAKE_MAPS[8 if map_size == 8 else 4]
        self.desc = np.array([[c.encode() for c in r] for r in rows], dtype="|S1")
        self.map_size = len(self.desc)
        self.action_size = 4
        self.board = None

    def _start(self):
        for i in range(self.map_size):
            for j in range(self.map_size):
                if self.desc[i][j] == b"S":
                    return (i, j)
        return (0, 0)

    def getInitBoard(self):
        b = np.zeros((self.map_size, self.map_size))
        b[self._start()] = 1
        return b

    def getBoardSize(self):
        return (self.map_size, self.map_size)

    def getActionSize(self):
        return self.action_size

    @staticmethod
    def _pos(board):
        return np.unravel_index(np.argmax(board), board.shape)

    def getNextState(self, board, player, action):
        if np.sum(board) == 0:
            return self.getInitBoard(), player
        r, c = self._pos(board)
        dr, dc = _FL_DIRS[action]
        nr, nc = r + dr, c + dc
        if not (0 <= nr < self.map_size and 0 <= nc < self.map_size):
            nr, nc = r, c
        nxt = np.zeros_like(board)
        nxt[nr, nc] = 1
        self.board = nxt  # side effect kept (:117-118)
        return nxt, player

    def getValidMoves(self, board, player):
        v = np.ones(self.action_size, dtype=np.int8)
        if self.getGameEnded(board, player) != 0:
            return np.zeros(self.action_size, dtype=np.int8)
        if np.sum(board) == 0:
            return v
        r, c = self._pos(board)
        if r == 0:
            v[0] = 0
        if r == self.map_size - 1:
            v[2] = 0
        if c == 0:
            v[3] = 0
        if c == self.map_size - 1:
            v[1] = 0
        return v

    def getGameEnded(self, board, player):
        if np.sum(board) == 0:
            return 0
        r, c = self._pos(board)
        if self.desc[r][c] == b"G":
            return 1.0
        if self.desc[r][c] == b"H":
            return -1.0
        return 0

    def getCanonicalForm(self, board, player):
        return board

    def stringRepresentation(self, board):
        if np.sum(board) == 0:
            return "empty"
        r, c = self._pos(board)
        return f"{r},{c}"
