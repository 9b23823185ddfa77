// Shared helpers for the mpnn_b200 CUDA library (sm_100a).  Not a public header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MPNN_OK 0
#define MPNN_ERR_ARG -1
#define MPNN_ERR_UNSUPPORTED -2
#define MPNN_ERR_CUDA -3
#define MPNN_ERR_WORKSPACE -4

void mpnn_set_error(const char* fmt, ...);

#define MPNN_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      mpnn_set_error(__VA_ARGS__);      \
      return (code);                    \
    }                                   \
  } while (0)

// Launch check: no synchronisation, only the launch status (sticky errors surface on the next call).
#define MPNN_CHECK_LAUNCH(what)                                                        \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      mpnn_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e__));    \
      return MPNN_ERR_CUDA;                                                            \
    }                                                                                  \
  } while (0)

#define MPNN_CUDA(call)                                                                \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      mpnn_set_error("%s failed: %s", #call, cudaGetErrorString(e__));                \
      return MPNN_ERR_CUDA;                                                            \
    }                                                                                  \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
static inline int pad4(int a) { return (a + 3) & ~3; }
static inline int pow2_at_least(int a, int lo) {
  int p = lo;
  while (p < a) p <<= 1;
  return p;
}

int mpnn_num_sms();

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
#endif
