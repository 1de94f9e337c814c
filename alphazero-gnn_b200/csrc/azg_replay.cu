// azg_replay.cu -- the example pipeline between self-play and training, on the device:
//
//   Coach.executeEpisode tail (Coach.py:45-49, 68-79): every stored position of a finished episode becomes
//     S symmetric examples (Connect4Game.getSymmetries :189-215 -- mirror, with its axis quirk;
//     TicTacToeGame.getSymmetries :187-200 -- 4 rotations x {flipped, not}; FrozenLake: identity) whose value is
//     the result signed for the player who was to move:  v = r * (-1)^(player != cur)
//   minibatch assembly of *.train (Connect4GNN.py:141-148): boards[idx] -> float32 [B,n,n], pi[idx], v[idx]
//
// Positions are the packed 16-byte states of the arena; policies stay float64 (the reference keeps Python floats)
// so that exported examples round-trip the reference's pickle format exactly.  A symmetry is a cell permutation
// given as a table (built on the host by applying the reference's own numpy calls to an index array, so the
// kernels hold no game knowledge): out cell d takes in cell perm[s][d], out policy entry a takes pi[pi_perm[s][a]].
#include "azg_common.cuh"

namespace {

struct State2 {
  uint64_t mine, theirs;
};

inline int grid_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// One thread per (history entry e, symmetry s): writes example e*S + s.
//   states/pi/player: [E] history entries;  game[e]: index into result/cur/out_base tables of finished games
//   v = result[g] * (player[e] != cur[g] ? -1 : 1);  vtag carries the result's Python type (int / float)
__global__ void emit_examples_kernel(const State2* __restrict__ states, const double* __restrict__ pi,
                                     const int32_t* __restrict__ player, const int32_t* __restrict__ game,
                                     const double* __restrict__ result, const int8_t* __restrict__ result_tag,
                                     const int32_t* __restrict__ cur, int64_t E, int ncells, int A, int S,
                                     const int32_t* __restrict__ board_perm, const int32_t* __restrict__ pi_perm,
                                     State2* __restrict__ out_states, double* __restrict__ out_pi, double* __restrict__ out_v,
                                     int8_t* __restrict__ out_vtag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * S) return;
  const int64_t e = i / S;
  const int s = (int)(i - e * S);
  const State2 in = states[e];
  State2 o{0ull, 0ull};
  const int32_t* bp = board_perm + (size_t)s * ncells;
  for (int d = 0; d < ncells; ++d) {
    const int src = bp[d];
    o.mine |= ((in.mine >> src) & 1ull) << d;
    o.theirs |= ((in.theirs >> src) & 1ull) << d;
  }
  out_states[i] = o;
  const int32_t* pp = pi_perm + (size_t)s * A;
  for (int a = 0; a < A; ++a) out_pi[i * A + a] = pi[e * A + pp[a]];
  if (out_v) {
    const int g = game[e];
    out_v[i] = result[g] * (player[e] != cur[g] ? -1.0 : 1.0);
    out_vtag[i] = result_tag[g];
  }
}

// FrozenLake states keep the agent cell index in `mine` (FrozenLakeGame.py:197-202): no bit permutation
__global__ void emit_examples_fl_kernel(const State2* __restrict__ states, const double* __restrict__ pi,
                                        const int32_t* __restrict__ player, const int32_t* __restrict__ game,
                                        const double* __restrict__ result, const int8_t* __restrict__ result_tag,
                                        const int32_t* __restrict__ cur, int64_t E, int A, State2* __restrict__ out_states,
                                        double* __restrict__ out_pi, double* __restrict__ out_v, int8_t* __restrict__ out_vtag) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  out_states[e] = states[e];
  for (int a = 0; a < A; ++a) out_pi[e * A + a] = pi[e * A + a];
  if (out_v) {
    const int g = game[e];
    out_v[e] = result[g] * (player[e] != cur[g] ? -1.0 : 1.0);
    out_vtag[e] = result_tag[g];
  }
}

// minibatch gather: one thread per (row, cell) for the boards, per (row, action) for pi, per row for v
__global__ void gather_examples_kernel(const State2* __restrict__ states, const double* __restrict__ pi, const double* __restrict__ v,
                                       const int64_t* __restrict__ idx, int B, int ncells, int A, int frozenlake,
                                       float* __restrict__ boards, float* __restrict__ out_pi, float* __restrict__ out_v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = ncells + A + 1;
  if (i >= (int64_t)B * per) return;
  const int b = (int)(i / per), k = (int)(i - (int64_t)b * per);
  const int64_t src = idx[b];
  if (k < ncells) {
    const State2 s = states[src];
    float cell;
    if (frozenlake) cell = (int64_t)s.mine == k ? 1.0f : 0.0f;
    else cell = (float)((int)((s.mine >> k) & 1ull) - (int)((s.theirs >> k) & 1ull));
    boards[(size_t)b * ncells + k] = cell;
  } else if (k < ncells + A) {
    out_pi[(size_t)b * A + (k - ncells)] = (float)pi[src * A + (k - ncells)];
  } else {
    out_v[b] = (float)v[src];
  }
}

}  // namespace

extern "C" {

int azg_emit_examples(int frozenlake, const uint64_t* states, const double* pi, const int32_t* player, const int32_t* game,
                      const double* result, const int8_t* result_tag, const int32_t* cur, int64_t E, int ncells, int A, int S,
                      const int32_t* board_perm, const int32_t* pi_perm, uint64_t* out_states, double* out_pi, double* out_v,
                      int8_t* out_vtag, azg_stream stream) {
  AZG_REQUIRE(states && pi && out_states && out_pi && A >= 1 && S >= 1, "azg_emit_examples: bad argument");
  AZG_REQUIRE(!out_v || (player && game && result && result_tag && cur && out_vtag), "azg_emit_examples: value inputs missing");
  AZG_REQUIRE(frozenlake || (board_perm && pi_perm && ncells >= 1 && ncells <= 64), "azg_emit_examples: permutation tables missing");
  if (E <= 0) return AZG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (frozenlake) {
    AZG_REQUIRE(S == 1, "azg_emit_examples: FrozenLake has no symmetries");
    emit_examples_fl_kernel<<<grid_for(E, 256), 256, 0, st>>>((const State2*)states, pi, player, game, result, result_tag, cur, E, A,
                                                              (State2*)out_states, out_pi, out_v, out_vtag);
  } else {
    emit_examples_kernel<<<grid_for(E * S, 256), 256, 0, st>>>((const State2*)states, pi, player, game, result, result_tag, cur, E,
                                                               ncells, A, S, board_perm, pi_perm, (State2*)out_states, out_pi,
                                                               out_v, out_vtag);
  }
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

int azg_gather_examples(int frozenlake, const uint64_t* states, const double* pi, const double* v, const int64_t* idx, int B,
                        int ncells, int A, float* boards, float* out_pi, float* out_v, azg_stream stream) {
  AZG_REQUIRE(states && pi && v && idx && boards && out_pi && out_v && ncells >= 1 && ncells <= 64 && A >= 1,
              "azg_gather_examples: bad argument");
  if (B <= 0) return AZG_OK;
  gather_examples_kernel<<<grid_for((int64_t)B * (ncells + A + 1), 256), 256, 0, (cudaStream_t)stream>>>(
      (const State2*)states, pi, v, idx, B, ncells, A, frozenlake, boards, out_pi, out_v);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

}  // extern "C"
