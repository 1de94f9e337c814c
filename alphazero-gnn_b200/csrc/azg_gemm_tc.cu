// azg_gemm_tc.cu -- K2 main loop on the 5th-gen tensor cores: the two F x F contractions of
// PolicyValueGNN.output_transform (gnn_utils.py:99-103, F = 64 n^2 = 3136 for 7x7) for a batch
// of leaf positions:   H = relu(X W0^T + b0),   E = H W2^T + b2.
//
// Design (sm_100a only):
//   * Operands live in HBM as *tile images*: every [rows x 64] bf16 block is stored exactly as
//     the tensor core wants to see it in shared memory (8-row groups of 128-byte rows, 16-byte
//     chunks XOR-swizzled by the row index = the SWIZZLE_128B K-major canonical layout), so a
//     pipeline stage is filled by two plain bulk-async copies (cp.async.bulk -> UBLKCP) that
//     complete on an mbarrier; no tensor maps, no in-kernel shuffling.  Weights are re-tiled
//     once per optimizer step (azg_c4_pack_gnn); activations are written as images by the
//     producer (the f32->image kernel for X, the GEMM epilogue for H).
//   * One persistent CTA per SM, warp-specialised: warp 0 = bulk-copy producer, warp 1 = MMA
//     issuer (a single thread issues tcgen05.mma, M=128 x N=BN x K=16 per instruction),
//     warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> bias/ReLU -> store).
//     Accumulators live in TMEM, double-buffered (2 x BN fp32 columns) so the epilogue of tile
//     i overlaps the main loop of tile i+1.  smem ring: 4 stages x (16 KB A + BN*128 B W).
//   * AZG_PREC_BF16X3 (the 1e-5 parity mode): x = hi + lo with hi = bf16(x), lo = bf16(x - hi);
//     X W^T ~= Xhi Whi^T + Xhi Wlo^T + Xlo Whi^T, fp32 accumulate.  Implemented as ONE GEMM with a
//     3x longer K: the producer walks the k-blocks of [Xhi|Xhi|Xlo] against [Whi|Wlo|Whi]; the
//     MMA and epilogue code is the same as for plain bf16.
#include "azg_common.cuh"

#include <cuda_bf16.h>

namespace tc {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // k-block = one 128-byte swizzle row of bf16
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_THREADS = 256;

// ---- tile image addressing ---------------------------------------------------------------
// byte offset of element (row r, k) inside one [R x 64] bf16 tile image
__host__ __device__ __forceinline__ uint32_t image_offset(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// global -> shared bulk async copy, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM lane quarter
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                   // layout type SWIZZLE_128B
  return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- epilogue output modes -----------------------------------------------------------------
enum { OUT_F32 = 0, OUT_IMG = 1, OUT_IMG_HILO = 2 };

struct GemmArgs {
  const uint8_t* a_hi;  // activation images, tiles [128 x 64], (mt * KB + kb) * 16384
  const uint8_t* a_lo;  // (x3 only)
  const uint8_t* w_hi;  // weight images, tiles [BN x 64], (nt * KB + kb) * BN * 128
  const uint8_t* w_lo;  // (x3 only)
  const float* bias;    // [N]
  float* out_f32;       // OUT_F32: row-major [M, N]
  uint8_t* out_hi;      // OUT_IMG*: image with tiles [128 x 64] over (mt, N/64)
  uint8_t* out_lo;
  int64_t M;            // valid rows
  int m_tiles, n_tiles, KB;  // KB = K / 64 (per operand, before the x3 expansion)
  int x3;               // 0: bf16, 1: 3-term split
  int relu;
  int out_mode;
};

template <int BN>
struct Smem {
  static constexpr int W_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + W_STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_bf16_tc_kernel(GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B wants 1024-B alignment
  using S = Smem<BN>;
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFFSET);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;  // 2 accumulator stages of BN <= 256 fp32 columns

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = g.m_tiles * g.n_tiles;
  const int kb_total = g.x3 ? 3 * g.KB : g.KB;

  if (warp == 0) {
    // ================= producer: bulk copies of operand tile images =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t / g.n_tiles, nt = t % g.n_tiles;
        for (int kb = 0; kb < kb_total; ++kb) {
          // x3: [Xhi|Xhi|Xlo] against [Whi|Wlo|Whi]
          const int seg = g.x3 ? kb / g.KB : 0, kk = g.x3 ? kb % g.KB : kb;
          const uint8_t* a_src = (seg == 2 ? g.a_lo : g.a_hi) + ((size_t)mt * g.KB + kk) * A_STAGE_BYTES;
          const uint8_t* w_src = (seg == 1 ? g.w_lo : g.w_hi) + ((size_t)nt * g.KB + kk) * S::W_STAGE_BYTES;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], S::STAGE_BYTES);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          bulk_g2s(sa, a_src, A_STAGE_BYTES, &full[stage]);
          bulk_g2s(sa + A_STAGE_BYTES, w_src, S::W_STAGE_BYTES, &full[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: one thread =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full[stage], phase);  // operands have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // advance 32 bytes (2 x 16-byte units) per K=16 step
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> bias/ReLU -> HBM =================
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int r_local = q * 32 + lane;   // row of the tile == TMEM lane
    int acc = 0;
    uint32_t acc_phase = 0;
    const int NKB = (g.n_tiles * BN) / BK;  // k-blocks of the output image
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mt = t / g.n_tiles, nt = t % g.n_tiles;
      const int64_t row = (int64_t)mt * BM + r_local;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t rr[32];
        tmem_ld32(taddr + (uint32_t)c0, rr);
        tmem_ld_wait();
        const int n0 = nt * BN + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(rr[j]) + __ldg(g.bias + n0 + j);
          v[j] = g.relu ? fmaxf(x, 0.0f) : x;
        }
        if (g.out_mode == OUT_F32) {
          if (row < g.M) {
            float4* dst = reinterpret_cast<float4*>(g.out_f32 + row * (int64_t)(g.n_tiles * BN) + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          // next GEMM's A operand: columns are its K index; 32 columns = 4 x 16-byte chunks of one image row
          const int kb = n0 / BK, koff = n0 % BK;
          const size_t tile = ((size_t)mt * NKB + kb) * A_STAGE_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = v[c * 8 + 2 * e], x1 = v[c * 8 + 2 * e + 1];
              const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
              hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              if (g.out_mode == OUT_IMG_HILO) {
                const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              }
            }
            const size_t off = tile + image_offset(r_local, koff + c * 8);
            *reinterpret_cast<uint4*>(g.out_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (g.out_mode == OUT_IMG_HILO) *reinterpret_cast<uint4*>(g.out_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);  // 128 arrivals release the accumulator
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- image builders --------------------------------------------------------------------------
// fp32 row-major [rows, K] -> bf16 tile images (hi, optionally lo) with R-row tiles.
// One thread per 16-byte output chunk (8 elements): 32-byte coalesced reads, 16-byte writes.
__global__ void __launch_bounds__(256) f32_to_image_kernel(const float* __restrict__ src, int64_t rows, int64_t rows_padded,
                                                           int K, int R, uint8_t* __restrict__ hi, uint8_t* __restrict__ lo) {
  const int chunks_per_row = K / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_padded * chunks_per_row) return;
  const int64_t row = idx / chunks_per_row;
  const int c = (int)(idx % chunks_per_row);
  float x[8];
  if (row < rows) {
    const float4 a = *reinterpret_cast<const float4*>(src + row * K + c * 8);
    const float4 b = *reinterpret_cast<const float4*>(src + row * K + c * 8 + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.0f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * e]), h1 = __float2bfloat16_rn(x[2 * e + 1]);
    h[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * e] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * e + 1] - __bfloat162float(h1));
    l[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  const int KB = K / BK;
  const int64_t tile_row = row / R;
  const int r = (int)(row % R), kb = (c * 8) / BK, k = (c * 8) % BK;
  const size_t off = ((size_t)tile_row * KB + kb) * ((size_t)R * 128) + image_offset(r, k);
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

inline int pick_bn(int F) {
  if (F % 224 == 0) return 224;
  if (F % 256 == 0) return 256;
  if (F % 160 == 0) return 160;
  return 0;
}

template <int BN>
int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  static bool configured = false;
  int dev = 0, sms = 0;
  AZG_CUDA_CHECK(cudaGetDevice(&dev));
  AZG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!configured) {
    AZG_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL));
    configured = true;
  }
  const int tiles = g.m_tiles * g.n_tiles;
  const int grid = tiles < sms ? tiles : sms;
  gemm_bf16_tc_kernel<BN><<<grid, NUM_THREADS, Smem<BN>::TOTAL, st>>>(g);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

int run_gemm(int BN, const GemmArgs& g, cudaStream_t st) {
  switch (BN) {
    case 224: return launch_gemm<224>(g, st);
    case 256: return launch_gemm<256>(g, st);
    case 160: return launch_gemm<160>(g, st);
  }
  azg_set_error("tcgen05 GEMM: no tile width for this feature size");
  return AZG_ERR_INVALID;
}

int to_image(const float* src, int64_t rows, int64_t rows_padded, int K, int R, uint8_t* hi, uint8_t* lo, cudaStream_t st) {
  const int64_t n = rows_padded * (K / 8);
  f32_to_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, rows, rows_padded, K, R, hi, lo);
  AZG_LAUNCH_CHECK();
  return AZG_OK;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// packed weight blob: [W0_hi | W0_lo | W2_hi | W2_lo] images (lo parts only for BF16X3)
static size_t image_bytes(int F) { return (size_t)F * F * 2; }

size_t azg_tc_scratch_bytes(int n, int64_t B, int prec) {
  const size_t F = 64 * (size_t)n * n;
  const size_t Mp = (size_t)azg_ceil_div(B, tc::BM) * tc::BM;
  const size_t per = Mp * F * 2;  // one bf16 image of [Mp, F]
  const int parts = (prec == AZG_PREC_BF16X3) ? 2 : 1;
  return 2 * parts * per + 1024;  // X image(s) + H image(s)
}

int azg_tc_output_transform(const void* packed, int n, int prec, const float* feat, const float* b0, const float* b2,
                            float* enh, int64_t B, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  const int F = 64 * n * n, BN = tc::pick_bn(F);
  AZG_REQUIRE(BN != 0, "tcgen05 path: unsupported board size %d", n);
  AZG_REQUIRE(scratch && scratch_bytes >= azg_tc_scratch_bytes(n, B, prec), "tcgen05 path: scratch too small");
  const bool x3 = prec == AZG_PREC_BF16X3;
  const int64_t m_tiles = azg_ceil_div(B, tc::BM), Mp = m_tiles * tc::BM;
  const size_t per = (size_t)Mp * F * 2;
  uint8_t* base = (uint8_t*)(((uintptr_t)scratch + 1023) & ~(uintptr_t)1023);
  uint8_t* x_hi = base;
  uint8_t* x_lo = x3 ? x_hi + per : nullptr;
  uint8_t* h_hi = base + (x3 ? 2 : 1) * per;
  uint8_t* h_lo = x3 ? h_hi + per : nullptr;
  const uint8_t* w = (const uint8_t*)packed;
  const size_t wb = image_bytes(F);
  const uint8_t *w0_hi = w, *w0_lo = x3 ? w + wb : nullptr;
  const uint8_t *w2_hi = w + (x3 ? 2 : 1) * wb, *w2_lo = x3 ? w2_hi + wb : nullptr;
  int rc;
  if ((rc = tc::to_image(feat, B, Mp, F, tc::BM, x_hi, x_lo, st))) return rc;
  tc::GemmArgs g{};
  g.M = B;
  g.m_tiles = (int)m_tiles;
  g.n_tiles = F / BN;
  g.KB = F / tc::BK;
  g.x3 = x3;
  // H = relu(X W0^T + b0), written as the next GEMM's operand image
  g.a_hi = x_hi; g.a_lo = x_lo; g.w_hi = w0_hi; g.w_lo = w0_lo; g.bias = b0; g.relu = 1;
  g.out_mode = x3 ? tc::OUT_IMG_HILO : tc::OUT_IMG; g.out_hi = h_hi; g.out_lo = h_lo; g.out_f32 = nullptr;
  if ((rc = tc::run_gemm(BN, g, st))) return rc;
  // E = H W2^T + b2, fp32 row-major for the heads
  g.a_hi = h_hi; g.a_lo = h_lo; g.w_hi = w2_hi; g.w_lo = w2_lo; g.bias = b2; g.relu = 0;
  g.out_mode = tc::OUT_F32; g.out_f32 = enh; g.out_hi = g.out_lo = nullptr;
  return tc::run_gemm(BN, g, st);
}

extern "C" {

size_t azg_c4_packed_bytes(int n, int prec) {
  const int F = 64 * n * n;
  if (tc::pick_bn(F) == 0 || prec == AZG_PREC_FP32) return 0;
  return (prec == AZG_PREC_BF16X3 ? 4 : 2) * image_bytes(F) + 1024;
}

int azg_c4_pack_gnn(const float* ot0_w, const float* ot2_w, int n, int prec, void* packed, size_t packed_bytes,
                    azg_stream stream) {
  const int F = 64 * n * n, BN = tc::pick_bn(F);
  AZG_REQUIRE(ot0_w && ot2_w && packed, "azg_c4_pack_gnn: null pointer");
  AZG_REQUIRE(BN != 0 && (prec == AZG_PREC_BF16X3 || prec == AZG_PREC_BF16), "azg_c4_pack_gnn: unsupported n=%d prec=%d", n, prec);
  AZG_REQUIRE(packed_bytes >= azg_c4_packed_bytes(n, prec), "azg_c4_pack_gnn: buffer too small");
  AZG_REQUIRE(((uintptr_t)packed & 15) == 0, "azg_c4_pack_gnn: buffer must be 16-byte aligned (bulk-copy source)");
  const bool x3 = prec == AZG_PREC_BF16X3;
  uint8_t* w = (uint8_t*)packed;
  const size_t wb = image_bytes(F);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = tc::to_image(ot0_w, F, F, F, BN, w, x3 ? w + wb : nullptr, st))) return rc;
  uint8_t* w2 = w + (x3 ? 2 : 1) * wb;
  return tc::to_image(ot2_w, F, F, F, BN, w2, x3 ? w2 + wb : nullptr, st);
}

// Stand-alone dense layer on the tensor-core path, for parity tests of the GEMM itself:
// C[M,F] = act(A[M,F] . W[F,F]^T + bias), fp32 in/out, operands converted on the fly.
int azg_tc_linear(const float* A, const float* W, const float* bias, float* C, int64_t M, int F, int prec, int relu,
                  void* scratch, size_t scratch_bytes, azg_stream stream) {
  const int BN = tc::pick_bn(F);
  AZG_REQUIRE(A && W && bias && C && scratch, "azg_tc_linear: null pointer");
  AZG_REQUIRE(BN != 0 && F % 64 == 0 && (prec == AZG_PREC_BF16X3 || prec == AZG_PREC_BF16), "azg_tc_linear: unsupported F=%d prec=%d", F, prec);
  const bool x3 = prec == AZG_PREC_BF16X3;
  const int64_t m_tiles = azg_ceil_div(M, tc::BM), Mp = m_tiles * tc::BM;
  const size_t a_img = (size_t)Mp * F * 2, w_img = image_bytes(F);
  const size_t need = (x3 ? 2 : 1) * (a_img + w_img) + 1024;
  AZG_REQUIRE(scratch_bytes >= need, "azg_tc_linear: scratch %zu < %zu", scratch_bytes, need);
  uint8_t* base = (uint8_t*)(((uintptr_t)scratch + 1023) & ~(uintptr_t)1023);
  uint8_t *a_hi = base, *a_lo = x3 ? base + a_img : nullptr;
  uint8_t* wbase = base + (x3 ? 2 : 1) * a_img;
  uint8_t *w_hi = wbase, *w_lo = x3 ? wbase + w_img : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = tc::to_image(A, M, Mp, F, tc::BM, a_hi, a_lo, st))) return rc;
  if ((rc = tc::to_image(W, F, F, F, BN, w_hi, w_lo, st))) return rc;
  tc::GemmArgs g{};
  g.M = M; g.m_tiles = (int)m_tiles; g.n_tiles = F / BN; g.KB = F / tc::BK; g.x3 = x3;
  g.a_hi = a_hi; g.a_lo = a_lo; g.w_hi = w_hi; g.w_lo = w_lo; g.bias = bias; g.relu = relu;
  g.out_mode = tc::OUT_F32; g.out_f32 = C;
  return tc::run_gemm(BN, g, st);
}

}  // extern "C"
