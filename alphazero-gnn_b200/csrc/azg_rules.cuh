// azg_rules.cuh -- game rules on compact states, shared by the arena kernels and the host
// check build (tests/hostcheck).  Bit x*n+y of {mine, theirs} is cell board[x][y] of the
// reference's canonical n x n array; FrozenLake keeps the agent cell index in `mine`.
//
//   Connect4   connect4/Connect4Game.py:143-187 (+ Board :38-110)
//   TicTacToe  tictactoe/TicTacToeGame.py:145-183 (+ Board :37-116)
//   FrozenLake frozenlake/FrozenLakeGame.py:88-187
#pragma once
#include "azg_common.cuh"

struct AzgState {
  uint64_t mine, theirs;
};

// Value with the reference's NumPy/Python type tag (AZG_TAG_*), SURVEY.md section 0.3.
struct AzgVal {
  double d;
  int tag;
};

struct AzgRules {
  int game, n, A, win_len;
  uint64_t board_mask;   // n*n low bits
  uint64_t y_low_mask;   // cells with y <= n - win_len      (Connect4)
  uint64_t y_high_mask;  // cells with y >= win_len - 1      (Connect4)
  uint64_t lines[18];    // TicTacToe: n rows, n columns, 2 diagonals
  int n_lines;
  uint8_t fl_map[64];    // FrozenLake map characters
};

static inline int azg_rules_init(AzgRules* r, int game, int n, const uint8_t* fl_map) {
  memset(r, 0, sizeof(*r));
  r->game = game;
  r->n = n;
  if (n < 2 || n > 8) return 1;
  r->board_mask = (n * n == 64) ? ~0ull : ((1ull << (n * n)) - 1);
  if (game == AZG_GAME_CONNECT4) {
    r->A = n + 1;                 // Connect4Game.py:139-141
    r->win_len = n < 4 ? n : 4;   // Board.is_win :73
    for (int x = 0; x < n; ++x)
      for (int y = 0; y < n; ++y) {
        if (y <= n - r->win_len) r->y_low_mask |= 1ull << (x * n + y);
        if (y >= r->win_len - 1) r->y_high_mask |= 1ull << (x * n + y);
      }
  } else if (game == AZG_GAME_TICTACTOE) {
    r->A = n * n + 1;  // TicTacToeGame.py:141-143
    // valid-move masks (azg_valids, AzgArenaView::valids, the exported Vs) are 32-bit: boards whose action count does
    // not fit (n >= 6: 37+ actions) are refused here, loudly, instead of silently dropping the high cells
    if (r->A > 32) return 1;
    int k = 0;
    for (int y = 0; y < n; ++y) {  // Board.is_win :66-75: fixed y, all x
      uint64_t m = 0;
      for (int x = 0; x < n; ++x) m |= 1ull << (x * n + y);
      r->lines[k++] = m;
    }
    for (int x = 0; x < n; ++x) {  // :78-87
      uint64_t m = 0;
      for (int y = 0; y < n; ++y) m |= 1ull << (x * n + y);
      r->lines[k++] = m;
    }
    uint64_t d0 = 0, d1 = 0;
    for (int i = 0; i < n; ++i) {
      d0 |= 1ull << (i * n + i);
      d1 |= 1ull << (i * n + (n - i - 1));
    }
    r->lines[k++] = d0;
    r->lines[k++] = d1;
    r->n_lines = k;
  } else if (game == AZG_GAME_FROZENLAKE) {
    r->A = 4;
    if (!fl_map) return 1;
    for (int i = 0; i < n * n; ++i) r->fl_map[i] = fl_map[i];
  } else {
    return 1;
  }
  return 0;
}

AZG_HD int azg_popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}

// ---- Connect4 ---------------------------------------------------------------------------
AZG_HD bool azg_c4_wins(const AzgRules& r, uint64_t b) {
  const int n = r.n, k = r.win_len;
  uint64_t h = b, v = b, d = b, e = b;
  for (int i = 1; i < k; ++i) {
    h &= b >> (i * n);        // (x+i, y)    horizontal   Board.is_win :76-79
    v &= b >> i;              // (x, y+i)    vertical     :82-85
    e &= b >> (i * (n - 1));  // (x+i, y-i)  diagonal /   :88-91
    d &= b >> (i * (n + 1));  // (x+i, y+i)  diagonal \   :94-97
  }
  return (h | (v & r.y_low_mask) | (d & r.y_low_mask) | (e & r.y_high_mask)) != 0;
}

AZG_HD uint32_t azg_c4_valids(const AzgRules& r, AzgState s) {
  const int n = r.n;
  const uint64_t occ = s.mine | s.theirs;
  uint32_t m = 0;
  for (int x = 0; x < n; ++x)
    if (!((occ >> (x * n + n - 1)) & 1ull)) m |= 1u << x;  // top cell empty, Board.get_legal_moves :46-48
  if (m == 0) m = 1u << n;                                 // pass, Connect4Game.py:162-164
  return m;
}

// ---- shared two-player pieces -----------------------------------------------------------
AZG_HD uint32_t azg_ttt_valids(const AzgRules& r, AzgState s) {
  const uint64_t empty = ~(s.mine | s.theirs) & r.board_mask;
  if (empty == 0) return 1u << (r.n * r.n);  // TicTacToeGame.py:161-163
  return (uint32_t)empty;                    // action a = n*x + y = bit index
}

AZG_HD bool azg_ttt_wins(const AzgRules& r, uint64_t b) {
  for (int i = 0; i < r.n_lines; ++i)
    if ((b & r.lines[i]) == r.lines[i]) return true;
  return false;
}

// getValidMoves(board, 1)
AZG_HD uint32_t azg_valids(const AzgRules& r, AzgState s) {
  if (r.game == AZG_GAME_CONNECT4) return azg_c4_valids(r, s);
  if (r.game == AZG_GAME_TICTACTOE) return azg_ttt_valids(r, s);
  // FrozenLakeGame.py:122-161 (terminal cells have no valid move, :128-129)
  const int n = r.n, cell = (int)s.mine, row = cell / n, col = cell % n;
  const uint8_t c = r.fl_map[cell];
  if (c == 'G' || c == 'H') return 0;
  uint32_t m = 0xF;
  if (row == 0) m &= ~1u;      // up
  if (col == n - 1) m &= ~2u;  // right
  if (row == n - 1) m &= ~4u;  // down
  if (col == 0) m &= ~8u;      // left
  return m;
}

// getGameEnded(board, 1): tag AZG_TAG_NONE = not ended (the reference's 0)
AZG_HD AzgVal azg_ended(const AzgRules& r, AzgState s) {
  AzgVal out;
  out.d = 0.0;
  out.tag = AZG_TAG_NONE;
  if (r.game == AZG_GAME_FROZENLAKE) {  // FrozenLakeGame.py:163-187: Python floats +-1.0
    const uint8_t c = r.fl_map[(int)s.mine];
    if (c == 'G') { out.d = 1.0; out.tag = AZG_TAG_PYFLOAT; }
    else if (c == 'H') { out.d = -1.0; out.tag = AZG_TAG_PYFLOAT; }
    return out;
  }
  const bool c4 = (r.game == AZG_GAME_CONNECT4);
  const bool win = c4 ? azg_c4_wins(r, s.mine) : azg_ttt_wins(r, s.mine);
  if (win) { out.d = 1.0; out.tag = AZG_TAG_PYINT; return out; }            // Connect4Game.py:175-176
  const bool loss = c4 ? azg_c4_wins(r, s.theirs) : azg_ttt_wins(r, s.theirs);
  if (loss) { out.d = -1.0; out.tag = AZG_TAG_PYINT; return out; }          // :177-178
  bool moves;
  if (c4) {
    const uint64_t occ = s.mine | s.theirs;
    moves = false;
    for (int x = 0; x < r.n; ++x) moves |= !((occ >> (x * r.n + r.n - 1)) & 1ull);
  } else {
    moves = (~(s.mine | s.theirs) & r.board_mask) != 0;
  }
  if (moves) return out;                                                     // :179-180
  out.d = 1e-4;                                                              // :183 draw
  out.tag = AZG_TAG_PYFLOAT;
  return out;
}

// canonical(getNextState(board, 1, a)): the mover's stones become `theirs`
AZG_HD AzgState azg_next(const AzgRules& r, AzgState s, int a) {
  AzgState o;
  if (r.game == AZG_GAME_FROZENLAKE) {  // FrozenLakeGame.py:88-120; single player, no flip
    const int n = r.n, cell = (int)s.mine;
    int row = cell / n, col = cell % n;
    const int dr = (a == 0) ? -1 : (a == 2) ? 1 : 0;
    const int dc = (a == 1) ? 1 : (a == 3) ? -1 : 0;
    int nr = row + dr, nc = col + dc;
    if (nr < 0 || nr >= n || nc < 0 || nc >= n) { nr = row; nc = col; }
    o.mine = (uint64_t)(nr * n + nc);
    o.theirs = 0;
    return o;
  }
  uint64_t placed = s.mine;
  if (r.game == AZG_GAME_CONNECT4) {
    if (a < r.n) {  // Connect4Game.py:146-152; drop to the first empty row, Board.get_drop_position :60-65
      const uint64_t col = ((s.mine | s.theirs) >> (a * r.n)) & ((1ull << r.n) - 1);
      int y = 0;
      while (y < r.n && ((col >> y) & 1ull)) ++y;
      placed |= 1ull << (a * r.n + y);
    }
  } else {
    if (a < r.n * r.n) placed |= 1ull << a;  // TicTacToeGame.py:149-155
  }
  o.mine = s.theirs;  // getCanonicalForm(next, -1) = -next
  o.theirs = placed;
  return o;
}
