// TEST INFRASTRUCTURE ONLY -- x86 build of the arena's host/device core
// (alphazero-gnn_b200/csrc/azg_arena_core.cuh, azg_rules.cuh) so the *same source* the GPU
// kernels run can be checked against the oracle on machines without a GPU.  Not part of the
// package, never loaded by it: the product path is libazgnn_b200.so (CUDA) only.
// Same entry points as include/azgnn_b200.h with an `h` infix (azgh_arena_*), all pointers are
// HOST pointers, `stream` is ignored.
#include "../../alphazero-gnn_b200/csrc/azg_arena_core.cuh"

#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";
void azg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct azgh_arena {
  AzgArenaView view;
};

extern "C" {

const char* azgh_last_error(void) { return g_err; }

size_t azgh_arena_bytes(int game, int n, int n_games, int cap, int max_depth) {
  AzgArenaView v;
  memset(&v, 0, sizeof(v));
  uint8_t blank[64];
  memset(blank, 'F', 64);
  if (azg_rules_init(&v.rules, game, n, blank)) return 0;
  v.G = n_games; v.cap = cap; v.hcap = azg_hash_capacity(cap); v.max_depth = max_depth; v.A = v.rules.A;
  return azg_arena_carve(&v, nullptr);
}

int azgh_arena_reset(azgh_arena* a, const int32_t* ids, int count, void*) {
  AzgArenaView& v = a->view;
  if (!ids) count = v.G;
  for (int i = 0; i < count; ++i) {
    const int g = ids ? ids[i] : i;
    memset(v.hslot + (size_t)g * v.hcap, 0, sizeof(int32_t) * v.hcap);
    v.node_count[g] = 0; v.sims_left[g] = 0; v.pending[g] = -1; v.path_len[g] = 0; v.status[g] = 0;
  }
  return 0;
}

int azgh_arena_create(azgh_arena** out, int game, int n, int n_games, int cap, int max_depth, double cpuct, void* mem,
                      size_t bytes, const uint8_t* fl_map, void*) {
  azgh_arena* a = new azgh_arena();
  memset(a, 0, sizeof(*a));
  if (azg_rules_init(&a->view.rules, game, n, fl_map)) { delete a; azg_set_error("bad game"); return AZG_ERR_INVALID; }
  AzgArenaView& v = a->view;
  v.G = n_games; v.cap = cap; v.hcap = azg_hash_capacity(cap); v.max_depth = max_depth; v.A = v.rules.A;
  v.two_player = (game != AZG_GAME_FROZENLAKE);
  v.cpuct = cpuct;
  v.p_f32 = (game == AZG_GAME_FROZENLAKE);
  if (azg_arena_carve(&v, (char*)mem) > bytes) { delete a; azg_set_error("arena memory too small"); return AZG_ERR_INVALID; }
  *out = a;
  memset(v.root, 0, (size_t)n_games * sizeof(AzgState));  // as azg_arena_create: roots start as the all-zero state
  return azgh_arena_reset(a, nullptr, n_games, nullptr);
}

int azgh_arena_destroy(azgh_arena* a) { delete a; return 0; }
int azgh_arena_action_size(const azgh_arena* a) { return a->view.A; }

int azgh_arena_copy_from(azgh_arena* dst, const azgh_arena* src, void*) {
  const AzgArenaView &d = dst->view, &s = src->view;
  if (d.G != s.G || d.A != s.A || d.max_depth < s.max_depth || d.cap < s.cap) { azg_set_error("copy_from: incompatible arenas"); return AZG_ERR_INVALID; }
  for (int g = 0; g < d.G; ++g) {
    azg_copy_game_nodes(d, s, g, 0, 1);
    azg_copy_game_finish(d, s, g);
  }
  return 0;
}

int azgh_arena_set_roots(azgh_arena* a, const uint64_t* s, void*) {
  memcpy(a->view.root, s, sizeof(AzgState) * a->view.G);
  return 0;
}
int azgh_arena_get_roots(azgh_arena* a, uint64_t* s, void*) {
  memcpy(s, a->view.root, sizeof(AzgState) * a->view.G);
  return 0;
}
int azgh_arena_begin(azgh_arena* a, int n_sims, void*) {
  for (int g = 0; g < a->view.G; ++g) a->view.sims_left[g] += n_sims;
  return 0;
}
int azgh_arena_select(azgh_arena* a, uint64_t* leaf_states, int32_t* leaf_mask, void*) {
  for (int g = 0; g < a->view.G; ++g) azg_select_game<1>(a->view, g, 0, 1u, (AzgState*)leaf_states, leaf_mask);
  return 0;
}
int azgh_arena_select_compact(azgh_arena* a, uint64_t* leaf_states, int32_t* leaf_mask, int32_t* leaf_game,
                              int32_t* leaf_count, void*) {
  *leaf_count = 0;
  for (int g = 0; g < a->view.G; ++g)
    azg_select_game<1>(a->view, g, 0, 1u, (AzgState*)leaf_states, leaf_mask, leaf_game, leaf_count);
  return 0;
}
int azgh_arena_expand_backup_compact(azgh_arena* a, const float* pi, const float* v, const int32_t* leaf_game,
                                     const int32_t* leaf_count, void*) {
  for (int i = 0; i < *leaf_count; ++i) azg_expand_backup_game(a->view, leaf_game[i], pi, v, i);
  return 0;
}
int azgh_arena_expand_backup(azgh_arena* a, const float* pi, const float* v, void*) {
  for (int g = 0; g < a->view.G; ++g) azg_expand_backup_game(a->view, g, pi, v);
  return 0;
}
int azgh_arena_root_stats(azgh_arena* a, int32_t* N, double* Q, int8_t* qtag, void*) {
  for (int g = 0; g < a->view.G; ++g) azg_root_stats_game(a->view, g, N, Q, qtag);
  return 0;
}
int azgh_arena_advance(azgh_arena* a, const int32_t* actions, double* ended, int8_t* ended_tag, void*) {
  for (int g = 0; g < a->view.G; ++g) azg_advance_game(a->view, g, actions[g], ended, ended_tag);
  return 0;
}
int azgh_arena_status(azgh_arena* a, int32_t* status, void*) {
  memcpy(status, a->view.status, sizeof(int32_t) * a->view.G);
  return 0;
}
int azgh_arena_export(azgh_arena* a, int g, int* n_nodes, uint64_t* keys, double* es, int8_t* es_tag, int32_t* ns,
                      uint32_t* valids, int8_t* ptag, double* P, double* Q, int8_t* qtag, int32_t* N, void*) {
  const AzgArenaView& v = a->view;
  const size_t c = v.node_count[g], n0 = (size_t)g * v.cap, A = v.A;
  *n_nodes = (int)c;
  if (keys) memcpy(keys, v.key + n0, c * sizeof(AzgState));
  if (es) memcpy(es, v.es + n0, c * sizeof(double));
  if (es_tag) memcpy(es_tag, v.es_tag + n0, c);
  if (ns) memcpy(ns, v.ns + n0, c * sizeof(int32_t));
  if (valids) memcpy(valids, v.valids + n0, c * sizeof(uint32_t));
  if (ptag) memcpy(ptag, v.ptag + n0, c);
  if (P) memcpy(P, v.P + n0 * A, c * A * sizeof(double));
  if (Q) memcpy(Q, v.Q + n0 * A, c * A * sizeof(double));
  if (qtag) memcpy(qtag, v.qtag + n0 * A, c * A);
  if (N) memcpy(N, v.N + n0 * A, c * A * sizeof(int32_t));
  return 0;
}
int azgh_rules_eval(int game, int n, const uint8_t* fl_map, const uint64_t* states, int64_t B, uint32_t* valids,
                    double* ended, int8_t* ended_tag, uint64_t* next, void*) {
  AzgRules r;
  if (azg_rules_init(&r, game, n, fl_map)) return AZG_ERR_INVALID;
  const AzgState* s = (const AzgState*)states;
  AzgState* nx = (AzgState*)next;
  for (int64_t i = 0; i < B; ++i) {
    const uint32_t m = azg_valids(r, s[i]);
    const AzgVal e = azg_ended(r, s[i]);
    valids[i] = m; ended[i] = e.d; ended_tag[i] = (int8_t)e.tag;
    for (int a = 0; a < r.A; ++a) {
      AzgState o = {0, 0};
      if ((m >> a) & 1u) o = azg_next(r, s[i], a);
      nx[i * r.A + a] = o;
    }
  }
  return 0;
}
double azgh_np_sum(const double* x, int n) { return azg_np_sum(x, n); }

}  // extern "C"
