// azg_nets.cu -- K1 (encode) and the fp32 CUDA-core forward path (K2, AZG_PREC_FP32) of the
// three games' networks.  This is the 1e-5 parity path; the dense F x F contractions of the
// Connect4 GNN head have a tcgen05 implementation in azg_gemm_tc.cu.
//
//   encode            Connect4GNN.py:71-72 / FrozenLakeNet.py:197-213
//   conv3x3_relu      Connect4Net.py:45-46, TicTacToeNet.py:33-35
//   linear (sgemm)    F.linear everywhere; output_transform gnn_utils.py:99-103
//   heads             Connect4Net.py:55-60 / Connect4GNN.py:48-57 / TicTacToeGNN.py:36-45
//   fl_forward        FrozenLakeNet.py:297-334
#include "azg_common.cuh"
#include "azg_rules.cuh"

int azg_train_linear_fwd(const float* X, const float* W, const float* bias, float* Y, int64_t M, int N, int K, int relu,
                         cudaStream_t st);  // azg_train.cu

namespace {

inline int grid_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------ K1 encode
template <typename T>
__global__ void pack_boards_kernel(const T* __restrict__ cells, int nn, int64_t B, AzgState* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint64_t mine = 0, theirs = 0;
  for (int i = 0; i < nn; ++i) {
    const T c = cells[b * nn + i];
    mine |= (uint64_t)(c > (T)0) << i;
    theirs |= (uint64_t)(c < (T)0) << i;
  }
  out[b].mine = mine;
  out[b].theirs = theirs;
}

// one thread per cell: coalesced 4-byte stores, the 16-byte state is a warp-broadcast load
__global__ void encode_planes_kernel(const AzgState* __restrict__ states, int nn, int64_t B, float* __restrict__ planes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nn) return;
  const int64_t b = i / nn;
  const int c = (int)(i - b * nn);
  const AzgState s = states[b];
  planes[i] = (float)((int)((s.mine >> c) & 1ull) - (int)((s.theirs >> c) & 1ull));
}

__device__ __forceinline__ int fl_nodes(int cell, int n, int* cells) {
  // node 0 = current cell; then successors of the valid actions up,right,down,left
  const int r = cell / n, c = cell % n;
  int k = 0;
  cells[k++] = cell;
  if (r > 0) cells[k++] = cell - n;
  if (c < n - 1) cells[k++] = cell + 1;
  if (r < n - 1) cells[k++] = cell + n;
  if (c > 0) cells[k++] = cell - 1;
  return k;
}

__global__ void fl_encode_graph_kernel(const AzgState* __restrict__ states, int n, int64_t B, float* __restrict__ nodes,
                                       int32_t* __restrict__ counts) {
  const int nn = n * n;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 5 * nn) return;
  const int64_t b = i / (5 * nn);
  const int rem = (int)(i - b * 5 * nn), node = rem / nn, c = rem % nn;
  int cells[5];
  const int k = fl_nodes((int)states[b].mine, n, cells);
  nodes[i] = (node < k && cells[node] == c) ? 1.0f : 0.0f;
  if (rem == 0) counts[b] = k;
}

// ------------------------------------------------------------------------------ conv 3x3
// One CTA per board.  The padded input lives in shared memory; each thread produces one
// output row (fixed cout, ox; all oy), so an input row and three weights feed 3*Wo FMAs.
template <int MAXW>
__global__ void __launch_bounds__(256) conv3x3_relu_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           int64_t B, int Cin, int Cout, int H, int W, int pad) {
  extern __shared__ float sin_[];  // [Cin][H+2p][W+2p]
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int Ho = Hp - 2, Wo = Wp - 2;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < Cin * Hp * Wp; i += blockDim.x) {
      const int ci = i / (Hp * Wp), r = i % (Hp * Wp), x = r / Wp - pad, y = r % Wp - pad;
      sin_[i] = (x >= 0 && x < H && y >= 0 && y < W) ? in[((b * Cin + ci) * H + x) * W + y] : 0.0f;
    }
    __syncthreads();
    for (int row = threadIdx.x; row < Cout * Ho; row += blockDim.x) {
      const int co = row / Ho, ox = row % Ho;
      float acc[MAXW];
#pragma unroll
      for (int j = 0; j < MAXW; ++j) acc[j] = 0.0f;
      const float* wc = w + (size_t)co * Cin * 9;
      for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* irow = sin_ + (ci * Hp + ox + kx) * Wp;
          const float w0 = __ldg(wc + ci * 9 + kx * 3 + 0), w1 = __ldg(wc + ci * 9 + kx * 3 + 1),
                      w2 = __ldg(wc + ci * 9 + kx * 3 + 2);
          float v[MAXW + 2];
#pragma unroll
          for (int j = 0; j < MAXW + 2; ++j) v[j] = (j < Wp) ? irow[j] : 0.0f;
#pragma unroll
          for (int j = 0; j < MAXW; ++j) acc[j] = fmaf(v[j], w0, fmaf(v[j + 1], w1, fmaf(v[j + 2], w2, acc[j])));
        }
      }
      const float bb = __ldg(bias + co);
      float* o = out + ((b * Cout + co) * Ho + ox) * Wo;
#pragma unroll
      for (int j = 0; j < MAXW; ++j)
        if (j < Wo) o[j] = fmaxf(acc[j] + bb, 0.0f);
    }
  }
}

// ------------------------------------------------------------------------------ sgemm
// C[M,N] = act(A[M,K] . W[N,K]^T + bias).  128x128x16 CTA tile, 8x8 register tile, register
// prefetch of the next K slab.  fp32 FFMA: this is the parity path, not the fast path.
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_T = 256;

__global__ void __launch_bounds__(SG_T) sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ C,
                                                        int64_t M, int N, int K, int relu) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;  // M on grid.x (2^31 limit), N on grid.y
  const int n0 = blockIdx.y * SG_BN;
  // loader mapping: 128 rows x 16 k = 512 float4; thread loads rows (tid>>2) and (tid>>2)+64, k4 = (tid&3)*4
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  float4 ra[2], rw[2];
  auto load = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t m = m0 + lr + h * 64;
      const int n = n0 + lr + h * 64;
      const int k = k0 + lk;
      ra[h] = (m < M && k < K) ? *reinterpret_cast<const float4*>(A + m * K + k) : make_float4(0, 0, 0, 0);
      rw[h] = (n < N && k < K) ? *reinterpret_cast<const float4*>(W + (int64_t)n * K + k) : make_float4(0, 0, 0, 0);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      As[lk + 0][r] = ra[h].x; As[lk + 1][r] = ra[h].y; As[lk + 2][r] = ra[h].z; As[lk + 3][r] = ra[h].w;
      Ws[lk + 0][r] = rw[h].x; Ws[lk + 1][r] = rw[h].y; Ws[lk + 2][r] = rw[h].z; Ws[lk + 3][r] = rw[h].w;
    }
  };
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 (m) x 8 (n)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  load(0);
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    __syncthreads();
    stash();
    __syncthreads();
    if (k0 + SG_BK < K) load(k0 + SG_BK);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.0f);
      if (relu) v = fmaxf(v, 0.0f);
      C[m * N + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------ heads
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per position: policy logits (A x Kp), value (1 x Kv), then exp(log_softmax) and tanh
// exactly as predict() does (Connect4GNN.py:76-82).  Xp/Xv may alias (Connect4) or differ
// (TicTacToe: relu(fc1 f) and relu(fc2 f)).
__global__ void __launch_bounds__(256) heads_kernel(const float* __restrict__ Xp, int Kp, const float* __restrict__ Wp,
                                                    const float* __restrict__ bp, int A, const float* __restrict__ Xv,
                                                    int Kv, const float* __restrict__ Wv, const float* __restrict__ bv,
                                                    int64_t B, float* __restrict__ pi, float* __restrict__ v) {
  __shared__ float logits[8][72];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < B; row += (int64_t)gridDim.x * 8) {
    const float* xp = Xp + row * Kp;
    for (int a0 = 0; a0 < A; a0 += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
      for (int k = lane * 4; k < Kp; k += 128) {
        const float4 x = *reinterpret_cast<const float4*>(xp + k);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (a0 + j < A) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(Wp + (size_t)(a0 + j) * Kp + k));
            acc[j] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[j]))));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = warp_sum(acc[j]);
        if (lane == 0 && a0 + j < A) logits[warp][a0 + j] = s + __ldg(bp + a0 + j);
      }
    }
    float accv = 0.0f;
    const float* xv = Xv + row * Kv;
    for (int k = lane * 4; k < Kv; k += 128) {
      const float4 x = *reinterpret_cast<const float4*>(xv + k);
      const float4 w = __ldg(reinterpret_cast<const float4*>(Wv + k));
      accv = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, accv))));
    }
    accv = warp_sum(accv);
    __syncwarp();
    float m = -INFINITY;
    for (int a = 0; a < A; ++a) m = fmaxf(m, logits[warp][a]);
    float sum = 0.0f;
    for (int a = 0; a < A; ++a) sum += expf(logits[warp][a] - m);
    const float lse = logf(sum);
    for (int a = lane; a < A; a += 32) pi[row * A + a] = expf((logits[warp][a] - m) - lse);
    if (lane == 0) v[row] = tanhf(accv + __ldg(bv));
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------ FrozenLake
// One CTA (128 threads) per state.  The graph has k+1 <= 5 nodes and an all-ones adjacency
// normalised to c = d*d, d = k^-1/2 (FrozenLakeNet.py:55-74), so aggregation is c * sum over
// nodes.  First Linear on a one-hot input is a column gather.
constexpr int FL_MAXE = 256;
struct FlLayers {
  const float* w[8];
  const float* b[8];
};

__global__ void __launch_bounds__(128) fl_forward_kernel(const AzgState* __restrict__ states, int n, int E, int layers,
                                                         const float* __restrict__ fe0_w, const float* __restrict__ fe0_b,
                                                         const float* __restrict__ fe2_w, const float* __restrict__ fe2_b,
                                                         FlLayers gl,
                                                         const float* __restrict__ pw, const float* __restrict__ pb,
                                                         const float* __restrict__ vw, const float* __restrict__ vb,
                                                         int64_t B, float* __restrict__ pi, float* __restrict__ v) {
  __shared__ float h1[5][128];
  __shared__ float x[5][FL_MAXE];
  __shared__ float sup[5][FL_MAXE];
  __shared__ float head[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nn = n * n;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    int cells[5];
    const int k = fl_nodes((int)states[b].mine, n, cells);
    __syncthreads();
    for (int i = threadIdx.x; i < k * 128; i += blockDim.x) {
      const int node = i >> 7, j = i & 127;
      h1[node][j] = fmaxf(__ldg(fe0_w + (size_t)j * nn + cells[node]) + __ldg(fe0_b + j), 0.0f);
    }
    __syncthreads();
    for (int e = warp; e < E; e += 4) {  // feature_extractor.2: 128 -> E, warp per output
      float acc[5] = {0, 0, 0, 0, 0};
      for (int j = lane; j < 128; j += 32) {
        const float w = __ldg(fe2_w + (size_t)e * 128 + j);
#pragma unroll
        for (int node = 0; node < 5; ++node)
          if (node < k) acc[node] = fmaf(w, h1[node][j], acc[node]);
      }
#pragma unroll
      for (int node = 0; node < 5; ++node) {
        const float s = warp_sum(acc[node]);
        if (lane == 0 && node < k) x[node][e] = fmaxf(s + __ldg(fe2_b + e), 0.0f);
      }
    }
    __syncthreads();
    const float d = 1.0f / sqrtf((float)k);
    const float c = d * d;
    for (int l = 0; l < layers; ++l) {
      const float* wl = gl.w[l];
      const float* bl = gl.b[l];
      for (int e = warp; e < E; e += 4) {
        float acc[5] = {0, 0, 0, 0, 0};
        for (int j = lane; j < E; j += 32) {
          const float w = __ldg(wl + (size_t)e * E + j);
#pragma unroll
          for (int node = 0; node < 5; ++node)
            if (node < k) acc[node] = fmaf(w, x[node][j], acc[node]);
        }
#pragma unroll
        for (int node = 0; node < 5; ++node) {
          const float s = warp_sum(acc[node]);
          if (lane == 0 && node < k) sup[node][e] = s + __ldg(bl + e);
        }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float agg = 0.0f;
        for (int node = 0; node < k; ++node) agg = fmaf(c, sup[node][e], agg);  // bmm(adj, support) row
        agg = fmaxf(agg, 0.0f);
        for (int node = 0; node < k; ++node) x[node][e] = agg;
      }
      __syncthreads();
    }
    // heads on node 0: 4 policy logits + value, one warp each (warp 0 also takes the value)
    for (int o = warp; o < 5; o += 4) {
      const float* wrow = (o < 4) ? pw + (size_t)o * E : vw;
      float acc = 0.0f;
      for (int j = lane; j < E; j += 32) acc = fmaf(__ldg(wrow + j), x[0][j], acc);
      acc = warp_sum(acc);
      if (lane == 0) head[o] = acc + ((o < 4) ? __ldg(pb + o) : __ldg(vb));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float m = fmaxf(fmaxf(head[0], head[1]), fmaxf(head[2], head[3]));
      float ex[4], s = 0.0f;
      for (int a = 0; a < 4; ++a) { ex[a] = expf(head[a] - m); s += ex[a]; }
      for (int a = 0; a < 4; ++a) pi[b * 4 + a] = ex[a] / s;  // F.softmax, FrozenLakeNet.py:330
      v[b] = tanhf(head[4]);
    }
  }
}

// ------------------------------------------------------------------------------ launch helpers
int launch_conv(const float* in, const float* w, const float* b, float* out, int64_t B, int Cin, int Cout, int H, int W,
                int pad, cudaStream_t st) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const size_t smem = (size_t)Cin * Hp * Wp * sizeof(float);
  const int grid = (int)(B < 148 * 8 ? B : 148 * 8);
  AZG_REQUIRE(Wp - 2 <= 8 && smem <= 48 * 1024, "conv3x3: board too large (W=%d, Cin=%d)", W, Cin);
  conv3x3_relu_kernel<8><<<grid, 256, smem, st>>>(in, w, b, out, B, Cin, Cout, H, W, pad);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

int launch_linear(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K, int relu,
                  cudaStream_t st) {
  AZG_REQUIRE(K % 4 == 0, "linear: K=%d must be a multiple of 4", K);
  AZG_REQUIRE(N <= 65535 * SG_BN, "linear: N=%d too large for one launch", N);
  dim3 grid((unsigned)((M + SG_BM - 1) / SG_BM), (N + SG_BN - 1) / SG_BN);
  sgemm_tn_kernel<<<grid, SG_T, 0, st>>>(A, W, bias, C, M, N, K, relu);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

// ---- 3x3 convolution as im2col + SGEMM (fp32 path) -----------------------------------------------------------
// The direct kernel above re-reads the Cin*9 weights of an output channel in every thread of every board (conv3 of
// TicTacToe: 590 KB of L1/L2 weight traffic per board, 20 ms per 65,536 boards).  Here the patches of a chunk of
// boards are written once as rows of a [boards*Ho*Wo, Cin*9] matrix in the weight tensor's own (ci, kx, ky) order, so
// conv.weight viewed as [Cout, Cin*9] is the SGEMM's W operand in place and every output row keeps a fixed summation
// order (batch-invariant, like launch_linear).  Output rows are (board, cell): NHWC.
__global__ void im2col3x3_kernel(const float* __restrict__ in, int in_nhwc, int64_t b0, int64_t nb, int Cin, int H, int W,
                                 int pad, float* __restrict__ col) {
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2, K = Cin * 9;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * Ho * Wo * K) return;
  const int k = (int)(i % K), ci = k / 9, kx = (k % 9) / 3, ky = k % 3;
  const int64_t r = i / K;
  const int p = (int)(r % (Ho * Wo)), x = p / Wo + kx - pad, y = p % Wo + ky - pad;
  const int64_t b = b0 + r / (Ho * Wo);
  float v = 0.0f;
  if (x >= 0 && x < H && y >= 0 && y < W)
    v = in_nhwc ? in[((b * H + x) * W + y) * Cin + ci] : in[((b * Cin + ci) * H + x) * W + y];
  col[i] = v;
}

// [B*P, C] (board, cell, channel) -> [B, C*P] (the reference's NCHW flatten order, Connect4Net.py:48 / TicTacToeNet.py:38)
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int64_t B, int P, int C, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P * C) return;
  const int p = (int)(i % P), c = (int)((i / P) % C);
  const int64_t b = i / ((int64_t)P * C);
  dst[i] = src[(b * P + p) * C + c];
}

constexpr size_t CONV_COL_FLOATS = (size_t)32 << 20;  // 128 MB im2col chunk

// relu(conv3x3(in) + b) for all B boards; out_nhwc [B*Ho*Wo, Cout]; col: CONV_COL_FLOATS floats of scratch
int launch_conv_gemm(const float* in, int in_nhwc, const float* w, const float* b, float* out_nhwc, int64_t B, int Cin,
                     int Cout, int H, int W, int pad, float* col, cudaStream_t st) {
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2, K = Cin * 9, P = Ho * Wo;
  int64_t chunk = (int64_t)(CONV_COL_FLOATS / ((size_t)P * K));
  if (chunk < 1) chunk = 1;
  for (int64_t b0 = 0; b0 < B; b0 += chunk) {
    const int64_t nb = B - b0 < chunk ? B - b0 : chunk;
    const int64_t n = nb * P * K;
    im2col3x3_kernel<<<grid_for(n, 256), 256, 0, st>>>(in, in_nhwc, b0, nb, Cin, H, W, pad, col);
    AZG_LAUNCH_CHECK();
    int rc = launch_linear(col, w, b, out_nhwc + (size_t)b0 * P * Cout, nb * P, Cout, K, 1, st);
    if (rc) return rc;
  }
  return AZG_OK;
}

int launch_nhwc_to_nchw(const float* src, int64_t B, int P, int C, float* dst, cudaStream_t st) {
  nhwc_to_nchw_kernel<<<grid_for(B * P * C, 256), 256, 0, st>>>(src, B, P, C, dst);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

int launch_heads(const float* Xp, int Kp, const float* Wp, const float* bp, int A, const float* Xv, int Kv,
                 const float* Wv, const float* bv, int64_t B, float* pi, float* v, cudaStream_t st) {
  AZG_REQUIRE(A <= 72 && Kp % 4 == 0 && Kv % 4 == 0, "heads: unsupported sizes A=%d Kp=%d Kv=%d", A, Kp, Kv);
  const int grid = (int)(azg_ceil_div(B, 8) < 148 * 8 ? azg_ceil_div(B, 8) : 148 * 8);
  heads_kernel<<<grid, 256, 0, st>>>(Xp, Kp, Wp, bp, A, Xv, Kv, Wv, bv, B, pi, v);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

struct Carver {
  char* base;
  size_t off = 0;
  float* take(size_t floats) {
    off = (off + 255) / 256 * 256;
    float* p = base ? (float*)(base + off) : nullptr;
    off += floats * sizeof(float);
    return p;
  }
};

}  // namespace

// fp32 pieces used by the TicTacToe tensor-core path (azg_gemm_tc.cu)
int azg_ttt_fp32_front(const float* conv1_w, const float* conv1_b, int n, const uint64_t* states, int64_t B, float* planes, float* c1,
                       cudaStream_t st) {
  if (B <= 0) return AZG_OK;
  encode_planes_kernel<<<grid_for(B * n * n, 256), 256, 0, st>>>((const AzgState*)states, n * n, B, planes);
  AZG_LAUNCH_CHECK();
  return launch_conv(planes, conv1_w, conv1_b, c1, B, 1, 32, n, n, 1, st);
}
int azg_ttt_fp32_heads(const float* h1, const float* pw, const float* pb, int A, const float* h2, const float* vw, const float* vb,
                       int64_t B, float* pi, float* v, cudaStream_t st) {
  return launch_heads(h1, 512, pw, pb, A, h2, 512, vw, vb, B, pi, v, st);
}
int azg_nhwc_to_nchw(const float* src, int64_t B, int P, int C, float* dst, cudaStream_t st) {
  return launch_nhwc_to_nchw(src, B, P, C, dst, st);
}

// tcgen05 path (azg_gemm_tc.cu)
int azg_tc_c4_forward(const void* packed, const azg_c4_params* p, int n, int prec, const uint64_t* states, int64_t B,
                      int eval_mask, float* pi_std, float* v_std, float* pi_gnn, float* v_gnn, void* scratch,
                      size_t scratch_bytes, const int32_t* dyn_rows, cudaStream_t st);
size_t azg_tc_scratch_bytes(int n, int64_t B, int prec);

extern "C" {

int azg_pack_boards(const void* cells, int cell_dtype, int n, int64_t B, uint64_t* states, azg_stream stream) {
  AZG_REQUIRE(cells && states && n >= 2 && n <= 8, "azg_pack_boards: bad argument");
  if (B <= 0) return AZG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(B, 128), nn = n * n;
  AzgState* out = (AzgState*)states;
  switch (cell_dtype) {
    case AZG_CELL_I8: pack_boards_kernel<int8_t><<<g, 128, 0, st>>>((const int8_t*)cells, nn, B, out); break;
    case AZG_CELL_I64: pack_boards_kernel<int64_t><<<g, 128, 0, st>>>((const int64_t*)cells, nn, B, out); break;
    case AZG_CELL_F32: pack_boards_kernel<float><<<g, 128, 0, st>>>((const float*)cells, nn, B, out); break;
    case AZG_CELL_F64: pack_boards_kernel<double><<<g, 128, 0, st>>>((const double*)cells, nn, B, out); break;
    default: azg_set_error("azg_pack_boards: bad cell_dtype %d", cell_dtype); return AZG_ERR_INVALID;
  }
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

int azg_encode_planes(const uint64_t* states, int n, int64_t B, float* planes, azg_stream stream) {
  AZG_REQUIRE(states && planes && n >= 2 && n <= 8, "azg_encode_planes: bad argument");
  if (B <= 0) return AZG_OK;
  encode_planes_kernel<<<grid_for(B * n * n, 256), 256, 0, (cudaStream_t)stream>>>((const AzgState*)states, n * n, B, planes);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

int azg_fl_encode_graph(const uint64_t* states, int n, int64_t B, float* nodes, int32_t* counts, azg_stream stream) {
  AZG_REQUIRE(states && nodes && counts && n >= 2 && n <= 8, "azg_fl_encode_graph: bad argument");
  if (B <= 0) return AZG_OK;
  fl_encode_graph_kernel<<<grid_for(B * 5 * n * n, 256), 256, 0, (cudaStream_t)stream>>>((const AzgState*)states, n, B,
                                                                                         nodes, counts);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

int azg_conv3x3_relu_forward(const float* in, const float* w, const float* b, float* out, int64_t B, int Cin, int Cout,
                             int H, int W, int pad, azg_stream stream) {
  AZG_REQUIRE(in && w && b && out, "azg_conv3x3_relu_forward: null pointer");
  if (B <= 0) return AZG_OK;
  return launch_conv(in, w, b, out, B, Cin, Cout, H, W, pad, (cudaStream_t)stream);
}

int azg_linear_f32(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K, int relu,
                   azg_stream stream) {
  AZG_REQUIRE(A && W && C, "azg_linear_f32: null pointer");
  if (M <= 0) return AZG_OK;
  // Training-step entry point: with fewer than ~1.5 waves of 128 x 128 tiles the split-K / matrix-vector kernels of
  // azg_train.cu fill the SMs instead.  The inference path (launch_linear) never takes this route: there a row's
  // result must not depend on how many other rows share the batch (bit-exact tree statistics, tests/test_arena_gpu.py).
  if (azg_ceil_div(M, SG_BM) * azg_ceil_div(N, SG_BN) < 222) return azg_train_linear_fwd(A, W, bias, C, M, N, K, relu, (cudaStream_t)stream);
  return launch_linear(A, W, bias, C, M, N, K, relu, (cudaStream_t)stream);
}

// ---- Connect4 ---------------------------------------------------------------------------
static size_t c4_carve(Carver& c, int n, int64_t B, int eval_mask, int prec, float** planes, float** c1, float** feat,
                       float** hid, float** enh, void** scratch, size_t* scratch_bytes, float** col = nullptr,
                       float** nhwc = nullptr) {
  const size_t nn = (size_t)n * n, F = 64 * nn;
  *planes = *c1 = *feat = *hid = *enh = nullptr;
  *scratch = nullptr;
  *scratch_bytes = 0;
  if (prec == AZG_PREC_FP32) {
    if (eval_mask & AZG_EVAL_GNN) *enh = c.take(B * F);
    *planes = c.take(B * nn);
    *c1 = c.take(B * 32 * nn);
    *feat = c.take(B * F);
    if (eval_mask & AZG_EVAL_GNN) *hid = c.take(B * F);
    float* colp = c.take(CONV_COL_FLOATS);  // im2col chunk of conv2
    float* nh = c.take(B * F);              // conv2 output as (board, cell, channel) before the NCHW flatten
    if (col) *col = colp;
    if (nhwc) *nhwc = nh;
  } else {  // tensor-core path: operand images live in the scratch area
    *scratch_bytes = azg_tc_scratch_bytes(n, B, prec);
    *scratch = c.take((*scratch_bytes + 3) / 4);
  }
  return (c.off + 255) / 256 * 256;
}

size_t azg_c4_workspace_bytes(int n, int64_t B, int eval_mask, int prec) {
  Carver c{nullptr};
  float *a, *b, *d, *e, *f;
  void* s;
  size_t sb;
  return c4_carve(c, n, B, eval_mask, prec, &a, &b, &d, &e, &f, &s, &sb);
}

int azg_c4_forward(const azg_c4_params* p, int n, const uint64_t* states, int64_t B, int eval_mask, int prec,
                   float* pi_std, float* v_std, float* pi_gnn, float* v_gnn, void* workspace, size_t workspace_bytes,
                   azg_stream stream) {
  return azg_c4_forward_dyn(p, n, states, B, nullptr, eval_mask, prec, pi_std, v_std, pi_gnn, v_gnn, workspace,
                            workspace_bytes, stream);
}

int azg_c4_forward_dyn(const azg_c4_params* p, int n, const uint64_t* states, int64_t B, const int32_t* dyn_rows,
                       int eval_mask, int prec, float* pi_std, float* v_std, float* pi_gnn, float* v_gnn, void* workspace,
                       size_t workspace_bytes, azg_stream stream) {
  AZG_REQUIRE(p && states && workspace, "azg_c4_forward: null pointer");
  AZG_REQUIRE(n >= 4 && n <= 8, "azg_c4_forward: board size %d unsupported (4..8)", n);
  AZG_REQUIRE((eval_mask & ~7) == 0 && (eval_mask & 3) != 0, "azg_c4_forward: bad eval_mask %d", eval_mask);
  AZG_REQUIRE(prec == AZG_PREC_FP32 || prec == AZG_PREC_BF16X3 || prec == AZG_PREC_BF16 || prec == AZG_PREC_F16F8 || prec == AZG_PREC_F16F8_KS ||
                  prec == AZG_PREC_BF16X3_KS,
              "azg_c4_forward: bad prec %d", prec);
  if (B <= 0) return AZG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Carver c{(char*)workspace};
  float *planes, *c1, *feat, *hid, *enh, *col = nullptr, *nhwc = nullptr;
  void* scratch;
  size_t scratch_bytes;
  const size_t need = c4_carve(c, n, B, eval_mask, prec, &planes, &c1, &feat, &hid, &enh, &scratch, &scratch_bytes, &col, &nhwc);
  AZG_REQUIRE(need <= workspace_bytes, "azg_c4_forward: workspace %zu < %zu bytes", workspace_bytes, need);
  const int nn = n * n, F = 64 * nn, A = n + 1;
  int rc;
  if (eval_mask & AZG_EVAL_STD) AZG_REQUIRE(pi_std && v_std, "azg_c4_forward: null std outputs");
  if (eval_mask & AZG_EVAL_GNN) AZG_REQUIRE(pi_gnn && v_gnn && p->ot0_b && p->ot2_b, "azg_c4_forward: null gnn outputs/params");
  if (prec == AZG_PREC_FP32) {
    azg_phase_begin(AZG_PHASE_TRUNK, st);
    if ((rc = azg_encode_planes(states, n, B, planes, stream))) return rc;
    if ((rc = launch_conv(planes, p->conv1_w, p->conv1_b, c1, B, 1, 32, n, n, 1, st))) return rc;
    if (nn <= 25) {  // small boards: the direct kernel wastes most of its 8-wide rows; measured slower than the GEMM form
      if ((rc = launch_conv_gemm(c1, 0, p->conv2_w, p->conv2_b, nhwc, B, 32, 64, n, n, 1, col, st))) return rc;
      if ((rc = launch_nhwc_to_nchw(nhwc, B, nn, 64, feat, st))) return rc;
    } else {  // 6x6 and larger: direct kernel (7x7: 8.9 ms vs ~12 ms through im2col, which writes 3.7 GB of patches)
      if ((rc = launch_conv(c1, p->conv2_w, p->conv2_b, feat, B, 32, 64, n, n, 1, st))) return rc;
    }
    azg_phase_end(AZG_PHASE_TRUNK, st);
    if (eval_mask & AZG_EVAL_STD) {
      azg_phase_begin(AZG_PHASE_HEADS, st);
      if ((rc = launch_heads(feat, F, p->fc_policy_w, p->fc_policy_b, A, feat, F, p->fc_value_w, p->fc_value_b, B, pi_std,
                             v_std, st)))
        return rc;
      azg_phase_end(AZG_PHASE_HEADS, st);
    }
    if (eval_mask & AZG_EVAL_GNN) {
      AZG_REQUIRE(p->ot0_w && p->ot2_w, "azg_c4_forward: null output_transform weights");
      azg_phase_begin(AZG_PHASE_GEMM, st);
      if ((rc = launch_linear(feat, p->ot0_w, p->ot0_b, hid, B, F, F, 1, st))) return rc;
      if ((rc = launch_linear(hid, p->ot2_w, p->ot2_b, enh, B, F, F, 0, st))) return rc;
      azg_phase_end(AZG_PHASE_GEMM, st);
    }
    if (eval_mask & AZG_EVAL_GNN) {
      azg_phase_begin(AZG_PHASE_HEADS, st);
      if ((rc = launch_heads(enh, F, p->fc_policy_w, p->fc_policy_b, A, enh, F, p->fc_value_w, p->fc_value_b, B, pi_gnn,
                             v_gnn, st)))
        return rc;
      azg_phase_end(AZG_PHASE_HEADS, st);
    }
  } else {
    AZG_REQUIRE(p->ot_packed, "azg_c4_forward: prec %d needs ot_packed (azg_c4_pack)", prec);
    if ((rc = azg_tc_c4_forward(p->ot_packed, p, n, prec, states, B, eval_mask, pi_std, v_std, pi_gnn, v_gnn, scratch,
                                scratch_bytes, dyn_rows, st)))
      return rc;
  }
  return AZG_OK;
}

// ---- TicTacToe --------------------------------------------------------------------------
static size_t ttt_carve(Carver& c, int n, int64_t B, int eval_mask, float** planes, float** c1, float** c2, float** feat,
                        float** h1, float** h2, float** hid, float** enh, float** col = nullptr, float** nhwc = nullptr) {
  const size_t nn = (size_t)n * n, F = 128 * (size_t)(n - 2) * (n - 2);
  *planes = c.take(B * nn);
  *c1 = c.take(B * 32 * nn);
  *c2 = c.take(B * 64 * nn);
  *feat = c.take(B * F);
  *h1 = c.take(B * 512);
  *h2 = c.take(B * 512);
  float* colp = c.take(CONV_COL_FLOATS);  // im2col chunk of conv2 / conv3
  float* nh = c.take(B * F);              // conv3 output as (board, cell, channel) before the NCHW flatten
  if (col) *col = colp;
  if (nhwc) *nhwc = nh;
  *hid = *enh = nullptr;
  if (eval_mask & AZG_EVAL_GNN) {
    *hid = c.take(B * F);
    *enh = c.take(B * F);
  }
  return (c.off + 255) / 256 * 256;
}

size_t azg_ttt_workspace_bytes(int n, int64_t B, int eval_mask) {
  Carver c{nullptr};
  float* q[8];
  return ttt_carve(c, n, B, eval_mask, q, q + 1, q + 2, q + 3, q + 4, q + 5, q + 6, q + 7);
}

int azg_ttt_forward(const azg_ttt_params* p, int n, const uint64_t* states, int64_t B, int eval_mask, float* pi_std,
                    float* v_std, float* pi_gnn, float* v_gnn, void* workspace, size_t workspace_bytes,
                    azg_stream stream) {
  AZG_REQUIRE(p && states && workspace, "azg_ttt_forward: null pointer");
  AZG_REQUIRE(n >= 3 && n <= 8, "azg_ttt_forward: board size %d unsupported (3..8)", n);
  AZG_REQUIRE((eval_mask & ~3) == 0 && eval_mask != 0, "azg_ttt_forward: bad eval_mask %d", eval_mask);
  if (B <= 0) return AZG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Carver c{(char*)workspace};
  float *planes, *c1, *c2, *feat, *h1, *h2, *hid, *enh, *col, *nhwc;
  const size_t need = ttt_carve(c, n, B, eval_mask, &planes, &c1, &c2, &feat, &h1, &h2, &hid, &enh, &col, &nhwc);
  AZG_REQUIRE(need <= workspace_bytes, "azg_ttt_forward: workspace %zu < %zu bytes", workspace_bytes, need);
  const int F = 128 * (n - 2) * (n - 2), A = n * n + 1;
  int rc;
  if ((rc = azg_encode_planes(states, n, B, planes, stream))) return rc;
  if ((rc = launch_conv(planes, p->conv1_w, p->conv1_b, c1, B, 1, 32, n, n, 1, st))) return rc;
  // conv2 / conv3 as im2col + SGEMM; c2 holds (board, cell, channel) rows, conv3's patches are read from them
  int c2_nhwc = 1;
  if (n * n <= 25) {
    if ((rc = launch_conv_gemm(c1, 0, p->conv2_w, p->conv2_b, c2, B, 32, 64, n, n, 1, col, st))) return rc;
  } else {
    c2_nhwc = 0;
    if ((rc = launch_conv(c1, p->conv2_w, p->conv2_b, c2, B, 32, 64, n, n, 1, st))) return rc;
  }
  if ((rc = launch_conv_gemm(c2, c2_nhwc, p->conv3_w, p->conv3_b, nhwc, B, 64, 128, n, n, 0, col, st))) return rc;
  if ((rc = launch_nhwc_to_nchw(nhwc, B, (n - 2) * (n - 2), 128, feat, st))) return rc;
  for (int pass = 0; pass < 2; ++pass) {
    const int bit = pass == 0 ? AZG_EVAL_STD : AZG_EVAL_GNN;
    if (!(eval_mask & bit)) continue;
    const float* f = feat;
    float *pi = pi_std, *v = v_std;
    if (pass == 1) {
      AZG_REQUIRE(p->ot0_w && p->ot2_w, "azg_ttt_forward: null output_transform weights");
      if ((rc = launch_linear(feat, p->ot0_w, p->ot0_b, hid, B, F, F, 1, st))) return rc;
      if ((rc = launch_linear(hid, p->ot2_w, p->ot2_b, enh, B, F, F, 0, st))) return rc;
      f = enh;
      pi = pi_gnn;
      v = v_gnn;
    }
    AZG_REQUIRE(pi && v, "azg_ttt_forward: null outputs");
    if ((rc = launch_linear(f, p->fc1_w, p->fc1_b, h1, B, 512, F, 1, st))) return rc;
    if ((rc = launch_linear(f, p->fc2_w, p->fc2_b, h2, B, 512, F, 1, st))) return rc;
    if ((rc = launch_heads(h1, 512, p->fc_policy_w, p->fc_policy_b, A, h2, 512, p->fc_value_w, p->fc_value_b, B, pi, v, st)))
      return rc;
  }
  return AZG_OK;
}

// ---- FrozenLake -------------------------------------------------------------------------
int azg_fl_forward(const azg_fl_params* p, int n, int embedding_dim, int layers, const uint64_t* states, int64_t B,
                   float* pi, float* v, azg_stream stream) {
  AZG_REQUIRE(p && states && pi && v, "azg_fl_forward: null pointer");
  AZG_REQUIRE(n >= 2 && n <= 8 && embedding_dim >= 1 && embedding_dim <= FL_MAXE && layers >= 0 && layers <= 8,
              "azg_fl_forward: unsupported sizes n=%d E=%d L=%d", n, embedding_dim, layers);
  if (B <= 0) return AZG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FlLayers gl;
  for (int l = 0; l < 8; ++l) {
    gl.w[l] = l < layers ? p->gnn_w[l] : nullptr;
    gl.b[l] = l < layers ? p->gnn_b[l] : nullptr;
  }
  const int grid = (int)(B < 148 * 8 ? B : 148 * 8);
  fl_forward_kernel<<<grid, 128, 0, st>>>((const AzgState*)states, n, embedding_dim, layers, p->fe0_w, p->fe0_b, p->fe2_w,
                                          p->fe2_b, gl, p->policy_w, p->policy_b, p->value_w, p->value_b, B, pi, v);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

}  // extern "C"
