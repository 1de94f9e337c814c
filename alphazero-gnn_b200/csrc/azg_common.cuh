// azg_common.cuh -- shared helpers: error plumbing, host/device portability macros.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define AZG_HD __host__ __device__ __forceinline__
#else
#define AZG_HD inline
#endif

#include "../../include/azgnn_b200.h"

// ---- error plumbing (thread-local message, C ABI returns a code) -------------------------
void azg_set_error(const char* fmt, ...);

#if defined(__CUDACC__)
#define AZG_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      azg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                    __LINE__);                                                            \
      return AZG_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)
// every kernel launch of the library goes through this: counts launches for bench.py's gpu_launches
extern unsigned long long g_azg_launches;
#define AZG_LAUNCH_CHECK()                  \
  do {                                      \
    ++g_azg_launches;                       \
    AZG_CUDA_CHECK(cudaGetLastError());     \
  } while (0)

// optional per-phase device timing (bench.py roofline): CUDA events recorded on the launching stream
enum { AZG_PHASE_TRUNK = 0, AZG_PHASE_GEMM = 1, AZG_PHASE_HEADS = 2, AZG_PHASE_ARENA = 3, AZG_NUM_PHASES = 4 };
void azg_phase_begin(int phase, cudaStream_t st);
void azg_phase_end(int phase, cudaStream_t st);
#endif

#define AZG_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      azg_set_error(__VA_ARGS__);   \
      return AZG_ERR_INVALID;       \
    }                               \
  } while (0)

static inline int64_t azg_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
