// Library-wide state: thread-local error string, version, device properties, small utility kernels.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void mpnn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int mpnn_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  }
  return sms;
}

namespace {

// out[r, :] = sum over k in [ptr[r], ptr[r+1]) of scale * src[idx[k] (or k), :]   (deterministic order)
__global__ void k_segment_sum(const float* __restrict__ src, const int* __restrict__ ptr, const int* __restrict__ idx,
                              int rows, int width, long long lds, float* __restrict__ out, long long ldo,
                              int accumulate, float scale) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * width) return;
  int r = (int)(t / width), c = (int)(t - (long long)r * width);
  float acc = 0.f;
  int kb = ptr[r], ke = ptr[r + 1];
  for (int k = kb; k < ke; ++k) {
    int s = idx ? idx[k] : k;
    acc += src[(long long)s * lds + c];
  }
  acc *= scale;
  float* o = out + (long long)r * ldo + c;
  *o = accumulate ? *o + acc : acc;
}

// two-stage deterministic column sums of X[rows, width] (optionally of X*Y elementwise)
constexpr int CS_T = 256;
__global__ void k_colsum_partial(const float* __restrict__ X, const float* __restrict__ Y, long long rows, int width,
                                 long long ldx, long long ldy, int rows_per_block, float* __restrict__ partial) {
  // thread -> column (tid % width), row lane (tid / width); blockDim is a multiple of width (or width > CS_T)
  extern __shared__ float sm[];
  int lanes = blockDim.x / width;
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  int c = threadIdx.x % width, rl = threadIdx.x / width;
  float acc = 0.f;
  if (rl < lanes) {
    // four independent loads in flight per thread (fixed combination order: bit-reproducible)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    long long r = r0 + rl;
    const long long st = lanes;
    for (; r + 3 * st < r1; r += 4 * st) {
      float v0 = X[r * ldx + c], v1 = X[(r + st) * ldx + c], v2 = X[(r + 2 * st) * ldx + c], v3 = X[(r + 3 * st) * ldx + c];
      if (Y) {
        v0 *= Y[r * ldy + c];
        v1 *= Y[(r + st) * ldy + c];
        v2 *= Y[(r + 2 * st) * ldy + c];
        v3 *= Y[(r + 3 * st) * ldy + c];
      }
      a0 += v0;
      a1 += v1;
      a2 += v2;
      a3 += v3;
    }
    for (; r < r1; r += st) {
      float v = X[r * ldx + c];
      if (Y) v *= Y[r * ldy + c];
      a0 += v;
    }
    acc = (a0 + a1) + (a2 + a3);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sm[l * width + c];
    partial[(size_t)blockIdx.x * width + c] = s;
  }
}
// A block is 32 columns x 8 slices of the partials: slice sl sums blocks sl, sl + 8, ... with four loads in flight, the
// eight slice sums are added in slice order (fixed order: bit-reproducible).  (One thread per column walking all the
// partials took 41 us for the 16 columns of config 4's per-atom tensors: 16 threads, thousands of dependent round trips.)
__global__ void __launch_bounds__(256) k_colsum_final(const float* __restrict__ partial, int nblk, int width,
                                                      float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < width) {
    int b = sl;
    for (; b + 24 < nblk; b += 32) {
      const float v0 = partial[(size_t)b * width + c], v1 = partial[(size_t)(b + 8) * width + c],
                  v2 = partial[(size_t)(b + 16) * width + c], v3 = partial[(size_t)(b + 24) * width + c];
      s += v0;
      s += v1;
      s += v2;
      s += v3;
    }
    for (; b < nblk; b += 8) s += partial[(size_t)b * width + c];
  }
  red[sl][cx] = s;
  __syncthreads();
  if (sl == 0 && c < width) {
    float t = red[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][cx];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// float4 form (width, lds, ldo multiples of 4, 16-byte aligned bases): one thread per (row, 4 columns); four
// independent 16-byte loads in flight per thread, summed in list order (same order as the scalar kernel)
__global__ void __launch_bounds__(256) k_segment_sum_v4(const float* __restrict__ src, const int* __restrict__ ptr,
                                                        const int* __restrict__ idx, int rows, int w4, long long lds,
                                                        float* __restrict__ out, long long ldo, int accumulate,
                                                        float scale) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)rows * w4) return;
  const int r = (int)(t / w4), c = (int)(t - (long long)r * w4) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int kb = __ldg(ptr + r), ke = __ldg(ptr + r + 1);
  int k = kb;
  for (; k + 4 <= ke; k += 4) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = idx ? __ldg(idx + k + j) : k + j;
      v[j] = __ldg(reinterpret_cast<const float4*>(src + (long long)s * lds + c));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc.x += v[j].x;
      acc.y += v[j].y;
      acc.z += v[j].z;
      acc.w += v[j].w;
    }
  }
  {
    float4 v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int kk = k + j < ke ? k + j : (ke > kb ? ke - 1 : 0);
      const int s = idx ? (ke > kb ? __ldg(idx + kk) : 0) : kk;
      v[j] = (k + j < ke) ? __ldg(reinterpret_cast<const float4*>(src + (long long)s * lds + c))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (k + j < ke) {
        acc.x += v[j].x;
        acc.y += v[j].y;
        acc.z += v[j].z;
        acc.w += v[j].w;
      }
    }
  }
  acc.x *= scale;
  acc.y *= scale;
  acc.z *= scale;
  acc.w *= scale;
  float4* o = reinterpret_cast<float4*>(out + (long long)r * ldo + c);
  if (accumulate) {
    const float4 p = *o;
    acc.x += p.x;
    acc.y += p.y;
    acc.z += p.z;
    acc.w += p.w;
  }
  *o = acc;
}

// wide matrices (width > 1024: the flat [rows, mf*nf] output of edge_map's last Linear on the distinct bond rows):
// one thread per column, coalesced across the warp, rows walked in order
__global__ void k_colsum_wide(const float* __restrict__ X, const float* __restrict__ Y, long long rows, int width,
                              long long ldx, long long ldy, float* __restrict__ out, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= width) return;
  float s0 = 0.f, s1 = 0.f;
  long long r = 0;
  for (; r + 1 < rows; r += 2) {
    float a = X[r * ldx + c], b = X[(r + 1) * ldx + c];
    if (Y) {
      a *= Y[r * ldy + c];
      b *= Y[(r + 1) * ldy + c];
    }
    s0 += a;
    s1 += b;
  }
  if (r < rows) s0 += X[r * ldx + c] * (Y ? Y[r * ldy + c] : 1.f);
  const float s = s0 + s1;
  out[c] = accumulate ? out[c] + s : s;
}

}  // namespace

static int colsum_blocks(long long rows, int* rows_per_block) {
  int target = 8 * mpnn_num_sms();
  long long rpb = (rows + target - 1) / target;
  if (rpb < 64) rpb = 64;
  *rows_per_block = (int)rpb;
  return (int)((rows + rpb - 1) / rpb);
}

extern "C" {

int mpnn_version(void) { return 100; }  // 0.1.0

// bytes of device memory <- 0 on `stream` (a memset node under CUDA-graph capture): the one clear of a captured step's
// arena of zero-initialised buffers (mpnn_b200/functional.py zeros)
int mpnn_zero_bytes(void* p, size_t bytes, cudaStream_t stream) {
  MPNN_REQUIRE(p || bytes == 0, MPNN_ERR_ARG, "zero_bytes: null pointer");
  if (bytes) MPNN_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
  return MPNN_OK;
}

const char* mpnn_last_error(void) { return g_err; }

int mpnn_segment_sum(const float* src, const int* ptr, const int* idx, int rows, int width, long long lds, float* out,
                     long long ldo, int accumulate, float scale, cudaStream_t stream) {
  MPNN_REQUIRE(rows >= 0 && width > 0, MPNN_ERR_ARG, "segment_sum: bad dims");
  if (rows == 0) return MPNN_OK;
  if ((width & 3) == 0 && (lds & 3) == 0 && (ldo & 3) == 0 && (((uintptr_t)src | (uintptr_t)out) & 15) == 0) {
    const int w4 = width / 4;
    k_segment_sum_v4<<<ceil_div((long long)rows * w4, 256), 256, 0, stream>>>(src, ptr, idx, rows, w4, lds, out, ldo,
                                                                            accumulate, scale);
    MPNN_CHECK_LAUNCH("k_segment_sum_v4");
    return MPNN_OK;
  }
  k_segment_sum<<<ceil_div((long long)rows * width, 256), 256, 0, stream>>>(src, ptr, idx, rows, width, lds, out, ldo,
                                                                           accumulate, scale);
  MPNN_CHECK_LAUNCH("k_segment_sum");
  return MPNN_OK;
}

size_t mpnn_colsum_workspace_bytes(long long rows, int width) {
  int rpb;
  int nblk = colsum_blocks(rows, &rpb);
  return (size_t)nblk * width * sizeof(float);
}

// out[c] (+)= sum_r X[r,c] * (Y ? Y[r,c] : 1)
int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy, float* out,
                int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows >= 0 && width > 0, MPNN_ERR_ARG, "colsum: bad dims");
  if (width > 1024) {
    k_colsum_wide<<<ceil_div(width, 256), 256, 0, stream>>>(X, Y, rows, width, ldx, ldy, out, accumulate);
    MPNN_CHECK_LAUNCH("k_colsum_wide");
    return MPNN_OK;
  }
  int rpb;
  int nblk = colsum_blocks(rows, &rpb);
  if (rows == 0) nblk = 0;
  MPNN_REQUIRE(workspace_bytes >= (size_t)nblk * width * sizeof(float), MPNN_ERR_WORKSPACE, "colsum: workspace");
  int threads = width >= CS_T ? width : (CS_T / width) * width;
  if (nblk > 0) {
    k_colsum_partial<<<nblk, threads, threads * sizeof(float), stream>>>(X, Y, rows, width, ldx, ldy, rpb,
                                                                        (float*)workspace);
    MPNN_CHECK_LAUNCH("k_colsum_partial");
  }
  k_colsum_final<<<ceil_div(width, 32), 256, 0, stream>>>((const float*)workspace, nblk, width, out, accumulate);
  MPNN_CHECK_LAUNCH("k_colsum_final");
  return MPNN_OK;
}

}  // extern "C"
