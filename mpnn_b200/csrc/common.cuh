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

// Exclusive scan of one or two int arrays of length n <= 1024 * SMALL_SCAN_PER_T in ONE single-block launch (n + 1
// outputs each, the last one the total).  At the reference's batch sizes (thousands of atoms / edges) the three-phase
// grid scan is three dependent microsecond launches; this is one.  Launch with <<<1, 1024>>>.
// The array is walked in chunks of 1024 consecutive elements (thread t owns element chunk*1024 + t): every load and
// store is coalesced and all loads are issued up front; per chunk there is one warp scan, ONE block barrier (warp
// totals double-buffered by chunk parity) and a redundant per-warp scan of the 32 warp totals.  (A first version gave
// each thread 16 CONSECUTIVE elements: 32 uncoalesced wavefronts per load instruction through one SM's LSU, 13.7 us
// for 7 424 elements.)
constexpr int SMALL_SCAN_PER_T = 16;
constexpr int SMALL_SCAN_MAX = 1024 * SMALL_SCAN_PER_T;
__device__ __forceinline__ void small_scan_block(const int* __restrict__ a, const int* __restrict__ b, int n,
                                                 int* __restrict__ outa, int* __restrict__ outb) {
  __shared__ int wsum[2][2][32];   // [chunk parity][array][warp]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = (n + 1023) >> 10;
  int va[SMALL_SCAN_PER_T], vb[SMALL_SCAN_PER_T];
#pragma unroll
  for (int c = 0; c < SMALL_SCAN_PER_T; ++c) {
    const int idx = (c << 10) + tid;
    va[c] = (c < nchunks && idx < n) ? a[idx] : 0;
    vb[c] = (b && c < nchunks && idx < n) ? b[idx] : 0;
  }
  int carry_a = 0, carry_b = 0;
#pragma unroll
  for (int c = 0; c < SMALL_SCAN_PER_T; ++c) {
    if (c < nchunks) {   // block-uniform
      int ia = va[c], ib = vb[c];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int xa = __shfl_up_sync(0xffffffffu, ia, o);
        const int xb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) {
          ia += xa;
          ib += xb;
        }
      }
      if (lane == 31) {
        wsum[c & 1][0][warp] = ia;
        wsum[c & 1][1][warp] = ib;
      }
      __syncthreads();
      int wa = wsum[c & 1][0][lane], wb = wsum[c & 1][1][lane];   // every warp scans the 32 warp totals itself
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int xa = __shfl_up_sync(0xffffffffu, wa, o);
        const int xb = __shfl_up_sync(0xffffffffu, wb, o);
        if (lane >= o) {
          wa += xa;
          wb += xb;
        }
      }
      const int tot_a = __shfl_sync(0xffffffffu, wa, 31), tot_b = __shfl_sync(0xffffffffu, wb, 31);
      int off_a = __shfl_sync(0xffffffffu, wa, (warp + 31) & 31), off_b = __shfl_sync(0xffffffffu, wb, (warp + 31) & 31);
      if (warp == 0) off_a = off_b = 0;
      const int idx = (c << 10) + tid;
      if (idx < n) {
        outa[idx] = carry_a + off_a + ia - va[c];
        if (b) outb[idx] = carry_b + off_b + ib - vb[c];
      }
      carry_a += tot_a;
      carry_b += tot_b;
    }
  }
  if (tid == 0) {
    outa[n] = carry_a;
    if (b) outb[n] = carry_b;
  }
}
#endif
