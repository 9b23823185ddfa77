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
constexpr int SMALL_SCAN_PER_T = 16;
constexpr int SMALL_SCAN_MAX = 1024 * SMALL_SCAN_PER_T;
__device__ __forceinline__ void small_scan_block(const int* __restrict__ a, const int* __restrict__ b, int n,
                                                 int* __restrict__ outa, int* __restrict__ outb) {
  __shared__ int wsum[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = tid * SMALL_SCAN_PER_T;
  int la[SMALL_SCAN_PER_T], lb[SMALL_SCAN_PER_T];
  int ta = 0, tb = 0;
#pragma unroll
  for (int i = 0; i < SMALL_SCAN_PER_T; ++i) {
    const int idx = base + i;
    la[i] = idx < n ? a[idx] : 0;
    lb[i] = (b && idx < n) ? b[idx] : 0;
    ta += la[i];
    tb += lb[i];
  }
  int ia = ta, ib = tb;
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
    wsum[0][warp] = ia;
    wsum[1][warp] = ib;
  }
  __syncthreads();
  if (warp == 0) {
    int va = wsum[0][lane], vb = wsum[1][lane];
    const int sa = va, sb = vb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int xa = __shfl_up_sync(0xffffffffu, va, o);
      const int xb = __shfl_up_sync(0xffffffffu, vb, o);
      if (lane >= o) {
        va += xa;
        vb += xb;
      }
    }
    wsum[0][lane] = va - sa;   // exclusive offsets of the warps
    wsum[1][lane] = vb - sb;
  }
  __syncthreads();
  int oa = wsum[0][warp] + ia - ta, ob = wsum[1][warp] + ib - tb;
  if (n == 0 && tid == 0) {
    outa[0] = 0;
    if (b) outb[0] = 0;
  }
#pragma unroll
  for (int i = 0; i < SMALL_SCAN_PER_T; ++i) {
    const int idx = base + i;
    if (idx < n) {
      outa[idx] = oa;
      if (b) outb[idx] = ob;
    }
    oa += la[i];
    ob += lb[i];
    if (idx == n - 1) {
      outa[n] = oa;
      if (b) outb[n] = ob;
    }
  }
}
#endif
