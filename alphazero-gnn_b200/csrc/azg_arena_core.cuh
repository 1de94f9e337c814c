// azg_arena_core.cuh -- per-game search logic of the arena, host/device.
//
// One *group* of W lanes owns one game (W = 8 on the GPU: four games per warp; W = 1 in the
// host check build of tests/hostcheck, where the same source runs sequentially).  All lanes of
// a group hold identical scalar state, so control flow is uniform inside a group; lane 0 does
// the table writes, every lane scores its own actions for the PUCT arg-max.
//
// What is reproduced, bit for bit:
//   MCTS.search      MCTS.py:151-240   descend / leaf / select / backup
//   value types      NEP 50: float64 priors and PUCT score, float32 running-mean Q once a
//                    network value has joined, Python int/float before that
//   tie-breaking     strict '>' in ascending action order -> lowest index wins (MCTS.py:206-216)
// No FMA contraction anywhere in this file: every rounding is spelled out.
#pragma once
#include "azg_rules.cuh"

// ---- exactly-rounded arithmetic ---------------------------------------------------------
#if defined(__CUDA_ARCH__)
AZG_HD double azg_dmul(double a, double b) { return __dmul_rn(a, b); }
AZG_HD double azg_dadd(double a, double b) { return __dadd_rn(a, b); }
AZG_HD double azg_ddiv(double a, double b) { return __ddiv_rn(a, b); }
AZG_HD double azg_dsqrt(double a) { return __dsqrt_rn(a); }
AZG_HD float azg_fmul(float a, float b) { return __fmul_rn(a, b); }
AZG_HD float azg_fadd(float a, float b) { return __fadd_rn(a, b); }
AZG_HD float azg_fdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
#include <math.h>
// host check build: compiled with -ffp-contract=off -O1 (see tests/hostcheck/build.py)
static inline double azg_dmul(double a, double b) { volatile double r = a * b; return r; }
static inline double azg_dadd(double a, double b) { volatile double r = a + b; return r; }
static inline double azg_ddiv(double a, double b) { volatile double r = a / b; return r; }
static inline double azg_dsqrt(double a) { volatile double r = sqrt(a); return r; }
static inline float azg_fmul(float a, float b) { volatile float r = a * b; return r; }
static inline float azg_fadd(float a, float b) { volatile float r = a + b; return r; }
static inline float azg_fdiv(float a, float b) { volatile float r = a / b; return r; }
#endif

#define AZG_MAX_A 65  // TicTacToe 8x8 + pass

struct AzgArenaView {
  AzgRules rules;
  int G, cap, hcap, max_depth, A, two_player;
  int p_f32;  // valid masks are int8 (FrozenLakeGame.py:125): Ps and the PUCT score stay float32 (NEP 50)
  double cpuct;
  // per game
  AzgState* root;       // [G]
  int32_t* sims_left;   // [G]
  int32_t* node_count;  // [G]
  int32_t* pending;     // [G] node index awaiting a prediction, -1 = none
  int32_t* path_len;    // [G]
  int32_t* path_node;   // [G, max_depth]
  int8_t* path_act;     // [G, max_depth]
  int32_t* status;      // [G] sticky error code
  int32_t* hslot;       // [G, hcap] node index + 1, 0 = empty
  // per node [G, cap]
  AzgState* key;
  double* es;       // Es value (MCTS.py:154-157); es_tag == AZG_TAG_NONE: not ended
  int8_t* es_tag;
  int32_t* ns;      // Ns; -1 = not expanded yet (s not in Ps)
  uint32_t* valids; // Vs bit mask
  int8_t* ptag;     // dtype of Ps[s]: 0 = float64, 1 = float32 (stored widened)
  // per edge [G, cap, A]
  double* P;    // Ps, float64
  double* Q;    // Qsa (float32 values stored widened, exact)
  int8_t* qtag; // AZG_TAG_NONE = (s,a) not in Qsa
  int32_t* N;   // Nsa
};

// ---- group helpers ----------------------------------------------------------------------
template <int W>
AZG_HD void azg_group_sync(unsigned mask) {
#if defined(__CUDA_ARCH__)
  if (W > 1) __syncwarp(mask);
#else
  (void)mask;
#endif
}

AZG_HD uint64_t azg_hash(AzgState s) {
  uint64_t h = s.mine * 0x9E3779B97F4A7C15ull ^ (s.theirs + 0x7F4A7C15F39CC060ull) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

// node index of state s in game g's table, or -1
AZG_HD int azg_lookup(const AzgArenaView& a, int g, AzgState s) {
  const int32_t* hs = a.hslot + (size_t)g * a.hcap;
  const AzgState* keys = a.key + (size_t)g * a.cap;
  uint32_t slot = (uint32_t)azg_hash(s) & (uint32_t)(a.hcap - 1);
  for (int probe = 0; probe < a.hcap; ++probe) {
    const int32_t v = hs[slot];
    if (v == 0) return -1;
    const AzgState k = keys[v - 1];
    if (k.mine == s.mine && k.theirs == s.theirs) return v - 1;
    slot = (slot + 1) & (uint32_t)(a.hcap - 1);
  }
  return -1;
}

// lane 0 only: claim a new node for s (caller checked capacity); returns its index
AZG_HD int azg_insert(const AzgArenaView& a, int g, AzgState s, int idx) {
  int32_t* hs = a.hslot + (size_t)g * a.hcap;
  uint32_t slot = (uint32_t)azg_hash(s) & (uint32_t)(a.hcap - 1);
  while (hs[slot] != 0) slot = (slot + 1) & (uint32_t)(a.hcap - 1);
  hs[slot] = idx + 1;
  const size_t node = (size_t)g * a.cap + idx;
  a.key[node] = s;
  a.ns[node] = -1;
  a.valids[node] = 0;
  a.es[node] = 0.0;
  a.es_tag[node] = AZG_TAG_NONE;
  a.ptag[node] = 0;
  const size_t e0 = node * a.A;
  for (int i = 0; i < a.A; ++i) {
    a.qtag[e0 + i] = AZG_TAG_NONE;
    a.N[e0 + i] = 0;
    a.Q[e0 + i] = 0.0;
    a.P[e0 + i] = 0.0;
  }
  return idx;
}

AZG_HD AzgVal azg_neg(AzgVal v) {
  // -v: Python int 0 stays +0 (there is no negative int zero); floats flip the sign bit
  v.d = (v.tag == AZG_TAG_PYINT) ? (0.0 - v.d) : -v.d;
  return v;
}

// Qsa <- (Nsa*Qsa + v)/(Nsa+1), MCTS.py:229, with NEP 50 promotion
AZG_HD AzgVal azg_q_update(int n, AzgVal q, AzgVal v) {
  AzgVal o;
  if (q.tag != AZG_TAG_F32 && v.tag != AZG_TAG_F32) {
    // all-Python arithmetic: int*int, int+int exact; true division -> float (binary64)
    const double nq = azg_dmul((double)n, q.d);
    const double s = azg_dadd(nq, v.d);
    o.d = azg_ddiv(s, (double)(n + 1));
    o.tag = AZG_TAG_PYFLOAT;
    return o;
  }
  float nq;
  if (q.tag == AZG_TAG_F32) nq = azg_fmul((float)n, (float)q.d);  // int * float32 -> float32
  else nq = (float)azg_dmul((double)n, q.d);                       // Python product, then weak-cast to float32
  const float vv = (float)v.d;                                     // float32 as is; Python scalar weak-cast
  const float s = azg_fadd(nq, vv);
  o.d = (double)azg_fdiv(s, (float)(n + 1));
  o.tag = AZG_TAG_F32;
  return o;
}

// numpy.sum of a contiguous float64 vector (pairwise_sum in loops_utils.h.src: <8 sequential,
// <=128: eight running lanes combined as a tree, then the tail)
AZG_HD double azg_np_sum(const double* x, int n) {
  if (n < 8) {
    double r = -0.0;
    for (int i = 0; i < n; ++i) r = azg_dadd(r, x[i]);
    return azg_dadd(0.0, r);
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = x[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = azg_dadd(r[j], x[i + j]);
  double res = azg_dadd(azg_dadd(azg_dadd(r[0], r[1]), azg_dadd(r[2], r[3])),
                        azg_dadd(azg_dadd(r[4], r[5]), azg_dadd(r[6], r[7])));
  for (; i < n; ++i) res = azg_dadd(res, x[i]);
  return azg_dadd(0.0, res);
}

// numpy.sum of a contiguous float32 vector: same blocking, float32 accumulators
AZG_HD float azg_np_sum_f32(const float* x, int n) {
  if (n < 8) {
    float r = -0.0f;
    for (int i = 0; i < n; ++i) r = azg_fadd(r, x[i]);
    return azg_fadd(0.0f, r);
  }
  float r[8];
  for (int j = 0; j < 8; ++j) r[j] = x[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = azg_fadd(r[j], x[i + j]);
  float res = azg_fadd(azg_fadd(azg_fadd(r[0], r[1]), azg_fadd(r[2], r[3])),
                       azg_fadd(azg_fadd(r[4], r[5]), azg_fadd(r[6], r[7])));
  for (; i < n; ++i) res = azg_fadd(res, x[i]);
  return azg_fadd(0.0f, res);
}

// lane 0: walk the recorded path back up, MCTS.py:228-240
AZG_HD void azg_backup(const AzgArenaView& a, int g, int depth, AzgVal v) {
  const int32_t* pn = a.path_node + (size_t)g * a.max_depth;
  const int8_t* pa = a.path_act + (size_t)g * a.max_depth;
  for (int i = depth - 1; i >= 0; --i) {
    const size_t node = (size_t)g * a.cap + pn[i];
    const size_t e = node * a.A + pa[i];
    if (a.qtag[e] != AZG_TAG_NONE) {
      AzgVal q;
      q.d = a.Q[e];
      q.tag = a.qtag[e];
      const AzgVal nq = azg_q_update(a.N[e], q, v);
      a.Q[e] = nq.d;
      a.qtag[e] = (int8_t)nq.tag;
      a.N[e] += 1;
    } else {
      a.Q[e] = v.d;
      a.qtag[e] = (int8_t)v.tag;
      a.N[e] = 1;
    }
    a.ns[node] += 1;
    if (a.two_player) v = azg_neg(v);
  }
}

// PUCT arg-max over the valid actions of an expanded node, MCTS.py:202-216
template <int W>
AZG_HD int azg_select_action(const AzgArenaView& a, size_t node, int lane, unsigned mask) {
  const uint32_t valids = a.valids[node];
  const double ns = (double)a.ns[node];
  const double sq_visited = azg_dsqrt(ns);                      // math.sqrt(Ns)
  const double sq_fresh = azg_dsqrt(azg_dadd(ns, 1e-8));        // math.sqrt(Ns + EPS)
  double best_u = -INFINITY;
  int best_a = -1;
  const bool pf32 = a.ptag[node] != 0;
  for (int act = lane; act < a.A; act += W) {
    if ((valids >> act) & 1u) {
      const size_t e = node * a.A + act;
      double u;
      if (!pf32) {  // float64 prior: every factor promotes to float64
        const double cp = azg_dmul(a.cpuct, a.P[e]);
        if (a.qtag[e] != AZG_TAG_NONE)
          u = azg_dadd(a.Q[e], azg_ddiv(azg_dmul(cp, sq_visited), (double)(1 + a.N[e])));
        else
          u = azg_dmul(cp, sq_fresh);
      } else {  // float32 prior: Python scalars are weak, the whole score stays float32
        const float cp = azg_fmul((float)a.cpuct, (float)a.P[e]);
        if (a.qtag[e] != AZG_TAG_NONE) {
          const float ex = azg_fdiv(azg_fmul(cp, (float)sq_visited), (float)(1 + a.N[e]));
          u = (double)azg_fadd((float)a.Q[e], ex);
        } else {
          u = (double)azg_fmul(cp, (float)sq_fresh);
        }
      }
      if (u > best_u) {
        best_u = u;
        best_a = act;
      }
    }
  }
#if defined(__CUDA_ARCH__)
  if (W > 1) {  // warp arg-max: larger score wins, equal scores -> lower action index
#pragma unroll
    for (int off = W / 2; off > 0; off >>= 1) {
      const double ou = __shfl_xor_sync(mask, best_u, off, W);
      const int oa = __shfl_xor_sync(mask, best_a, off, W);
      if (oa >= 0 && (best_a < 0 || ou > best_u || (ou == best_u && oa < best_a))) {
        best_u = ou;
        best_a = oa;
      }
    }
  }
#else
  (void)mask;
#endif
  return best_a;
}

// MCTS.search descend phase for game g; runs searches until one needs a prediction or the
// simulation budget is spent.
// leaf_game/leaf_count (optional): compacted output -- games that wait for a prediction claim consecutive
// slots; leaf_states then holds their positions densely and leaf_game the owning game of each slot.
template <int W>
AZG_HD void azg_select_game(const AzgArenaView& a, int g, int lane, unsigned mask, AzgState* leaf_states,
                            int32_t* leaf_mask, int32_t* leaf_game = nullptr, int32_t* leaf_count = nullptr) {
  int sims = a.sims_left[g];
  int count = a.node_count[g];
  const bool waiting = a.pending[g] >= 0;
  int emit = waiting ? 1 : 0;
  AzgState emit_state = a.root[g];
  if (waiting) emit_state = a.key[(size_t)g * a.cap + a.pending[g]];
  bool failed = a.status[g] != 0;
  while (!waiting && !failed && sims > 0) {
    AzgState s = a.root[g];
    int depth = 0;
    AzgVal v;
    v.d = 0.0;
    v.tag = AZG_TAG_PYINT;
    bool need_eval = false;
    while (true) {
      if (depth >= a.max_depth) break;  // cycle policy: return 0
      int idx = azg_lookup(a, g, s);
      bool fresh = false;
      if (idx < 0) {
        if (count >= a.cap) {
          failed = true;
          if (lane == 0) a.status[g] = AZG_ERR_CAPACITY;
          break;
        }
        idx = count++;
        fresh = true;
        const AzgVal e = azg_ended(a.rules, s);  // Es[s] = getGameEnded(board, 1), MCTS.py:154-155
        if (lane == 0) {
          azg_insert(a, g, s, idx);
          const size_t node = (size_t)g * a.cap + idx;
          a.es[node] = e.d;
          a.es_tag[node] = (int8_t)e.tag;
        }
        azg_group_sync<W>(mask);
      }
      const size_t node = (size_t)g * a.cap + idx;
      const int etag = a.es_tag[node];
      if (etag != AZG_TAG_NONE) {  // terminal: return Es[s], MCTS.py:156-157
        v.d = a.es[node];
        v.tag = etag;
        break;
      }
      if (fresh || a.ns[node] < 0) {  // s not in Ps: leaf, MCTS.py:162-193 continues in expand_backup
        if (lane == 0) {
          a.valids[node] = azg_valids(a.rules, s);
          a.pending[g] = idx;
          a.path_len[g] = depth;
        }
        emit_state = s;
        need_eval = true;
        break;
      }
      const int act = azg_select_action<W>(a, node, lane, mask);
      if (act < 0) break;  // no action: return 0, MCTS.py:218-219
      if (lane == 0) {
        a.path_node[(size_t)g * a.max_depth + depth] = idx;
        a.path_act[(size_t)g * a.max_depth + depth] = (int8_t)act;
      }
      ++depth;
      s = azg_next(a.rules, s, act);  // MCTS.py:221-224
    }
    if (failed) break;
    if (need_eval) {
      emit = 1;
      break;
    }
    if (lane == 0) azg_backup(a, g, depth, v);
    --sims;
    azg_group_sync<W>(mask);
  }
  if (lane == 0) {
    a.sims_left[g] = sims;
    a.node_count[g] = count;
    leaf_mask[g] = emit;
    if (leaf_count) {
      if (emit) {
#if defined(__CUDA_ARCH__)
        const int slot = atomicAdd(leaf_count, 1);
#else
        const int slot = (*leaf_count)++;
#endif
        leaf_states[slot] = emit_state;
        leaf_game[slot] = g;
      }
    } else if (emit) {
      leaf_states[g] = emit_state;
    }
  }
}

// MCTS.search leaf phase + backup for a game whose prediction arrived (lane 0 only)
// `row` = row of pi / vpred that holds game g's prediction (g itself, or its compact slot)
AZG_HD void azg_expand_backup_game(const AzgArenaView& a, int g, const float* pi, const float* vpred, int64_t row = -1) {
  if (row < 0) row = g;
  const int idx = a.pending[g];
  if (idx < 0) return;
  const size_t node = (size_t)g * a.cap + idx;
  const uint32_t valids = a.valids[node];
  double ps[AZG_MAX_A];
  const int A = a.A;
  int ptag = 0;
  if (!a.p_f32) {
    for (int i = 0; i < A; ++i)  // Ps = pi * valids: float32 * int64 -> float64, MCTS.py:180
      ps[i] = azg_dmul((double)pi[(size_t)row * A + i], (double)((valids >> i) & 1u));
    const double tot = azg_np_sum(ps, A);
    if (tot > 0) {
      for (int i = 0; i < A; ++i) ps[i] = azg_ddiv(ps[i], tot);  // MCTS.py:182-183
    } else {
      const double k = (double)azg_popc64((uint64_t)valids);      // valids / np.sum(valids), :186
      for (int i = 0; i < A; ++i) ps[i] = azg_ddiv((double)((valids >> i) & 1u), k);
    }
  } else {
    float pf[AZG_MAX_A];
    for (int i = 0; i < A; ++i)  // float32 * int8 -> float32
      pf[i] = azg_fmul(pi[(size_t)row * A + i], (float)((valids >> i) & 1u));
    const float tot = azg_np_sum_f32(pf, A);
    if (tot > 0) {
      for (int i = 0; i < A; ++i) ps[i] = (double)azg_fdiv(pf[i], tot);
      ptag = 1;
    } else {  // int8 array / int64 sum -> float64
      const double k = (double)azg_popc64((uint64_t)valids);
      for (int i = 0; i < A; ++i) ps[i] = azg_ddiv((double)((valids >> i) & 1u), k);
    }
  }
  a.ptag[node] = (int8_t)ptag;
  for (int i = 0; i < A; ++i) a.P[node * A + i] = ps[i];
  a.ns[node] = 0;  // MCTS.py:188
  AzgVal v;
  v.d = (double)vpred[row];  // numpy.float32 from the net, MCTS.py:190-193
  v.tag = AZG_TAG_F32;
  azg_backup(a, g, a.path_len[g], v);
  a.pending[g] = -1;
  a.sims_left[g] -= 1;
}

AZG_HD void azg_root_stats_game(const AzgArenaView& a, int g, int32_t* N, double* Q, int8_t* qtag) {
  const int idx = azg_lookup(a, g, a.root[g]);
  for (int i = 0; i < a.A; ++i) {
    const size_t o = (size_t)g * a.A + i;
    if (idx < 0) {
      N[o] = 0; Q[o] = 0.0; qtag[o] = AZG_TAG_NONE;
    } else {
      const size_t e = ((size_t)g * a.cap + idx) * a.A + i;
      N[o] = a.N[e]; Q[o] = a.Q[e]; qtag[o] = a.qtag[e];
    }
  }
}

AZG_HD void azg_advance_game(const AzgArenaView& a, int g, int action, double* ended, int8_t* ended_tag) {
  if (action >= 0) a.root[g] = azg_next(a.rules, a.root[g], action);
  const AzgVal e = azg_ended(a.rules, a.root[g]);
  ended[g] = e.d;
  ended_tag[g] = (int8_t)e.tag;
}

// ---- table growth: copy game g from `s` into the larger `d` (same game, A, max_depth; d.cap >= s.cap) ------------
// part 1 (any number of lanes): clear d's hash slots and copy the node and edge arrays verbatim (node indices are kept)
AZG_HD void azg_copy_game_nodes(const AzgArenaView& d, const AzgArenaView& s, int g, int lane, int lanes) {
  int32_t* hs = d.hslot + (size_t)g * d.hcap;
  for (int i = lane; i < d.hcap; i += lanes) hs[i] = 0;
  const int count = s.node_count[g];
  const size_t sn = (size_t)g * s.cap, dn = (size_t)g * d.cap, A = (size_t)s.A;
  for (int i = lane; i < count; i += lanes) {
    d.key[dn + i] = s.key[sn + i];
    d.es[dn + i] = s.es[sn + i];
    d.es_tag[dn + i] = s.es_tag[sn + i];
    d.ns[dn + i] = s.ns[sn + i];
    d.valids[dn + i] = s.valids[sn + i];
    d.ptag[dn + i] = s.ptag[sn + i];
    for (size_t e = 0; e < A; ++e) {
      d.P[(dn + i) * A + e] = s.P[(sn + i) * A + e];
      d.Q[(dn + i) * A + e] = s.Q[(sn + i) * A + e];
      d.qtag[(dn + i) * A + e] = s.qtag[(sn + i) * A + e];
      d.N[(dn + i) * A + e] = s.N[(sn + i) * A + e];
    }
  }
}
// part 2 (one lane, after part 1 is complete): per-game scalars, the search path, and the hash rebuilt for d.hcap
AZG_HD void azg_copy_game_finish(const AzgArenaView& d, const AzgArenaView& s, int g) {
  d.root[g] = s.root[g];
  d.sims_left[g] = s.sims_left[g];
  d.node_count[g] = s.node_count[g];
  d.pending[g] = s.pending[g];
  d.path_len[g] = s.path_len[g];
  d.status[g] = s.status[g];
  for (int i = 0; i < s.path_len[g] && i < d.max_depth; ++i) {
    d.path_node[(size_t)g * d.max_depth + i] = s.path_node[(size_t)g * s.max_depth + i];
    d.path_act[(size_t)g * d.max_depth + i] = s.path_act[(size_t)g * s.max_depth + i];
  }
  int32_t* hs = d.hslot + (size_t)g * d.hcap;
  const AzgState* keys = d.key + (size_t)g * d.cap;
  for (int i = 0; i < s.node_count[g]; ++i) {
    uint32_t slot = (uint32_t)azg_hash(keys[i]) & (uint32_t)(d.hcap - 1);
    while (hs[slot] != 0) slot = (slot + 1) & (uint32_t)(d.hcap - 1);
    hs[slot] = i + 1;
  }
}

// ---- memory carve-up shared by device and host builds -------------------------------------
static inline size_t azg_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline int azg_hash_capacity(int cap) {
  int h = 16;
  while (h < 2 * cap) h <<= 1;
  return h;
}

// lays the arrays out inside `base` (may be NULL to measure); returns total bytes
static inline size_t azg_arena_carve(AzgArenaView* v, char* base) {
  size_t off = 0;
  const size_t G = v->G, cap = v->cap, A = v->A, D = v->max_depth, H = v->hcap;
#define AZG_CARVE(field, type, count)                       \
  do {                                                      \
    off = azg_align_up(off, 256);                           \
    v->field = base ? (type*)(base + off) : (type*)nullptr; \
    off += sizeof(type) * (count);                          \
  } while (0)
  AZG_CARVE(root, AzgState, G);
  AZG_CARVE(sims_left, int32_t, G);
  AZG_CARVE(node_count, int32_t, G);
  AZG_CARVE(pending, int32_t, G);
  AZG_CARVE(path_len, int32_t, G);
  AZG_CARVE(path_node, int32_t, G * D);
  AZG_CARVE(path_act, int8_t, G * D);
  AZG_CARVE(status, int32_t, G);
  AZG_CARVE(hslot, int32_t, G * H);
  AZG_CARVE(key, AzgState, G * cap);
  AZG_CARVE(es, double, G * cap);
  AZG_CARVE(es_tag, int8_t, G * cap);
  AZG_CARVE(ns, int32_t, G * cap);
  AZG_CARVE(valids, uint32_t, G * cap);
  AZG_CARVE(ptag, int8_t, G * cap);
  AZG_CARVE(P, double, G * cap * A);
  AZG_CARVE(Q, double, G * cap * A);
  AZG_CARVE(qtag, int8_t, G * cap * A);
  AZG_CARVE(N, int32_t, G * cap * A);
#undef AZG_CARVE
  return azg_align_up(off, 256);
}
