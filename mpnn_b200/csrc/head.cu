// Prediction head + loss of the reference drivers as ONE launch each way:
//   test_graph_norm.py:86-90   nn.BatchNorm1d(out) -> nn.Linear(out, targets)  + nn.MSELoss()
// At the reference's batch sizes (256 graphs x 64 features x 12 targets) the stock modules are 14 kernels of 2-5 us
// each on the critical path of a ~0.45 ms step (BN fwd, GEMM + bias, MSE, mean, fills, MSE bwd, three GEMMs with
// split-K reduce, bias column sum, BN bwd).  The whole problem is 64 KB, but one CTA is too slow for it (measured:
// 23 + 40 us, no better than the stock kernels), and the batch statistics couple all rows, so independent CTAs would
// need several launches.  A thread-block CLUSTER is the fit: up to 8 CTAs split the rows, every batch-wide sum
// (column moments, loss, weight gradients) is a per-CTA partial in shared memory that the other CTAs read over
// distributed shared memory after a cluster barrier, always in rank order (bit-reproducible).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int HT = 256;         // threads per CTA
constexpr int MAX_CLUSTER = 8;  // portable cluster size
constexpr size_t HEAD_SMEM_MAX = 200 * 1024;

__host__ __device__ inline int head_cluster(int B) { return B >= 8 * MAX_CLUSTER ? MAX_CLUSTER : 1; }
__host__ __device__ inline int head_rows_per(int B) {
  const int nc = head_cluster(B);
  return (B + nc - 1) / nc;
}
__host__ __device__ inline size_t head_fwd_floats(int B, int C, int T) {
  const size_t RP = head_rows_per(B);
  return RP * (C + 1) + (size_t)T * (C + 1) + 4 * (size_t)C + T + 8 + HT;
}
__host__ __device__ inline size_t head_bwd_floats(int B, int C, int T) {
  const size_t RP = head_rows_per(B);
  return 2 * RP * (C + 1) + RP * (T + 1) + (size_t)T * (C + 1) + ((size_t)T * C + T) + 8 * (size_t)C + HT;
}

// global rows of a contiguous [nr, cols] matrix -> shared [nr][ld] through f(value, col).  All loads of a thread are
// issued before its first store (one memory latency per batch of 8, not one per loop trip).
template <typename F>
__device__ __forceinline__ void stage_rows(const float* __restrict__ g, int nr, int cols, float* s, int ld, F f) {
  constexpr int U = 8;
  const int n = nr * cols;
  for (int base = threadIdx.x; base < n; base += HT * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * HT;
      v[u] = i < n ? __ldg(g + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * HT;
      if (i < n) {
        const int r = i / cols, c = i - r * cols;
        s[r * ld + c] = f(v[u], c);
      }
    }
  }
}

// out[c] = sum over this CTA's rows of f(r, c)   (slices of rows combined in a fixed order)
template <typename F>
__device__ __forceinline__ void column_partial(int nr, int C, float* scratch, float* out, F f) {
  if (C <= HT) {
    const int nsl = HT / C;
    const int c = threadIdx.x % C, s = threadIdx.x / C;
    const bool on = threadIdx.x < nsl * C;
    float a = 0.f;
    if (on)
      for (int r = s; r < nr; r += nsl) a += f(r, c);
    __syncthreads();
    if (on) scratch[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < C) {
      float t = 0.f;
      for (int k = 0; k < nsl; ++k) t += scratch[k * C + threadIdx.x];
      out[threadIdx.x] = t;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += HT) {
      float a = 0.f;
      for (int r = 0; r < nr; ++r) a += f(r, c);
      out[c] = a;
    }
  }
  __syncthreads();
}

// total[c] = sum over the cluster's CTAs (rank order) of their `part[c]`; a cluster barrier must separate the last
// write of `part` from this call
__device__ __forceinline__ void cluster_gather(cg::cluster_group& cluster, float* part, int n, float* total) {
  const unsigned nc = cluster.num_blocks();
  for (int c = threadIdx.x; c < n; c += HT) {
    float t = 0.f;
    for (unsigned k = 0; k < nc; ++k) t += cluster.map_shared_rank(part, k)[c];
    total[c] = t;
  }
  __syncthreads();
}

struct HeadArgs {
  const float *x, *target, *gamma, *beta, *W, *b;
  float *running_mean, *running_var;
  long long* num_batches_tracked;
  int B, C, T, training;
  float momentum, eps;
};

// y [B,T], loss [1], stats [2C] (mean, inverse std actually used)
__global__ void __launch_bounds__(HT) k_head_fwd(HeadArgs a, float* __restrict__ y, float* __restrict__ loss,
                                                 float* __restrict__ stats) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  const int B = a.B, C = a.C, T = a.T, LD = C + 1;
  const int RP = head_rows_per(B);
  const int rank = (int)cluster.block_rank();
  const int r0 = rank * RP;
  const int nr = max(0, min(RP, B - r0));
  float* xs = sm;                       // [RP][LD]  x, then x_hat
  float* Ws = xs + (size_t)RP * LD;     // [T][LD]   gamma_c W[o][c]
  float* p0 = Ws + (size_t)T * LD;      // [C]       partial column sums      (read by the other CTAs)
  float* p1 = p0 + C;                   // [C]       partial centred squares  (read by the other CTAs)
  float* mean = p1 + C;                 // [C]
  float* istd = mean + C;               // [C]
  float* bs = istd + C;                 // [T]       b_o + sum_c beta_c W[o][c]
  float* lp = bs + T;                   // [8]       partial loss             (read by rank 0)
  float* scratch = lp + 8;              // [HT]
  const int tid = threadIdx.x;
  // every global operand is staged first (independent loads, one latency); the folding of the affine parameters into
  // the Linear then runs out of shared memory: W' = W gamma, b' = b + W beta
  stage_rows(a.x + (size_t)r0 * C, nr, C, xs, LD, [](float v, int) { return v; });
  stage_rows(a.W, T, C, Ws, LD, [](float v, int) { return v; });
  for (int c = tid; c < C; c += HT) {
    p0[c] = a.gamma ? __ldg(a.gamma + c) : 1.f;
    p1[c] = a.beta ? __ldg(a.beta + c) : 0.f;
  }
  for (int o = tid; o < T; o += HT) bs[o] = a.b ? __ldg(a.b + o) : 0.f;
  __syncthreads();
  for (int o = tid; o < T; o += HT) {
    float s = bs[o];
    for (int c = 0; c < C; ++c) s = fmaf(p1[c], Ws[o * LD + c], s);
    bs[o] = s;
  }
  __syncthreads();
  for (int i = tid; i < T * C; i += HT) {
    const int o = i / C, c = i - o * C;
    Ws[o * LD + c] *= p0[c];
  }
  __syncthreads();
  if (a.training) {
    column_partial(nr, C, scratch, p0, [&](int r, int c) { return xs[r * LD + c]; });
    cluster.sync();
    cluster_gather(cluster, p0, C, mean);
    for (int c = tid; c < C; c += HT) mean[c] *= 1.f / B;
    __syncthreads();
    column_partial(nr, C, scratch, p1, [&](int r, int c) {
      const float d = xs[r * LD + c] - mean[c];
      return d * d;
    });
    cluster.sync();
    cluster_gather(cluster, p1, C, istd);
    for (int c = tid; c < C; c += HT) {
      const float ss = istd[c];
      const float var = ss / B;                                 // biased: what normalises (torch BatchNorm1d)
      if (rank == 0 && a.running_mean) {
        const float unb = B > 1 ? ss / (B - 1) : var;           // unbiased: what is tracked
        a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean[c];
        a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * unb;
      }
      istd[c] = rsqrtf(var + a.eps);
    }
    if (rank == 0 && tid == 0 && a.num_batches_tracked) *a.num_batches_tracked += 1;
  } else {
    for (int c = tid; c < C; c += HT) {
      mean[c] = a.running_mean[c];
      istd[c] = rsqrtf(a.running_var[c] + a.eps);
    }
  }
  __syncthreads();
  if (rank == 0)
    for (int c = tid; c < C; c += HT) {
      stats[c] = mean[c];
      stats[C + c] = istd[c];
    }
  for (int i = tid; i < nr * C; i += HT) {
    const int r = i / C, c = i - r * C;
    xs[r * LD + c] = (xs[r * LD + c] - mean[c]) * istd[c];
  }
  __syncthreads();
  float lsum = 0.f;
  for (int i = tid; i < nr * T; i += HT) {
    const float tg = __ldg(a.target + (size_t)r0 * T + i);
    const int r = i / T, o = i - r * T;
    const float* xr = xs + (size_t)r * LD;
    const float* wo = Ws + (size_t)o * LD;
    float s0 = bs[o], s1 = 0.f;
    int c = 0;
    for (; c + 1 < C; c += 2) {
      s0 = fmaf(xr[c], wo[c], s0);
      s1 = fmaf(xr[c + 1], wo[c + 1], s1);
    }
    if (c < C) s0 = fmaf(xr[c], wo[c], s0);
    const float s = s0 + s1;
    y[(size_t)r0 * T + i] = s;
    const float d = s - tg;
    lsum = fmaf(d, d, lsum);
  }
  // loss: warp -> CTA -> cluster, fixed order
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) scratch[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < HT / 32; ++w) s += scratch[w];
    lp[0] = s;
  }
  cluster.sync();
  if (rank == 0 && tid == 0) {
    float s = 0.f;
    for (unsigned k = 0; k < cluster.num_blocks(); ++k) s += cluster.map_shared_rank(lp, k)[0];
    *loss = s / ((float)B * T);
  }
  cluster.sync();   // nobody leaves while its shared memory may still be read
}

struct HeadBwd {
  const float *x, *target, *gamma, *beta, *W, *y, *stats, *gloss;
  int B, C, T, training;
  float *dx, *dgamma, *dbeta, *dW, *db;
};

__global__ void __launch_bounds__(HT) k_head_bwd(HeadBwd a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  const int B = a.B, C = a.C, T = a.T, LD = C + 1, LT = T + 1;
  const int RP = head_rows_per(B);
  const int rank = (int)cluster.block_rank();
  const int nc = (int)cluster.num_blocks();
  const int r0 = rank * RP;
  const int nr = max(0, min(RP, B - r0));
  float* xh = sm;                          // [RP][LD]  x_hat
  float* dz = xh + (size_t)RP * LD;        // [RP][LD]  gradient at the Linear's input
  float* dy = dz + (size_t)RP * LD;        // [RP][LT]
  float* Ws = dy + (size_t)RP * LT;        // [T][LD]   raw W
  float* pW = Ws + (size_t)T * LD;         // [T*C + T] partial dW (x_hat form), db   (read by the other CTAs)
  float* ps = pW + (size_t)T * C + T;      // [2C]      partial sum dz, sum dz x_hat  (read by the other CTAs)
  float* mean = ps + 2 * C;                // [C]
  float* istd = mean + C;
  float* s12 = istd + C;                   // [2C]      cluster totals of ps
  float* gb = s12 + 2 * C;                 // [2C]      gamma, beta
  float* scratch = gb + 2 * C;             // [HT]
  const int tid = threadIdx.x;
  for (int c = tid; c < C; c += HT) {
    mean[c] = __ldg(a.stats + c);
    istd[c] = __ldg(a.stats + C + c);
    gb[c] = a.gamma ? __ldg(a.gamma + c) : 1.f;
    gb[C + c] = a.beta ? __ldg(a.beta + c) : 0.f;
  }
  const float scale = 2.f * __ldg(a.gloss) / ((float)B * T);
  stage_rows(a.W, T, C, Ws, LD, [](float v, int) { return v; });
  stage_rows(a.y + (size_t)r0 * T, nr, T, dy, LT, [](float v, int) { return v; });
  stage_rows(a.x + (size_t)r0 * C, nr, C, xh, LD, [](float v, int) { return v; });
  __syncthreads();
  {
    constexpr int U = 8;
    const float* tg = a.target + (size_t)r0 * T;
    for (int base = tid; base < nr * T; base += HT * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * HT;
        v[u] = i < nr * T ? __ldg(tg + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * HT;
        if (i < nr * T) {
          const int r = i / T, o = i - r * T;
          dy[r * LT + o] = scale * (dy[r * LT + o] - v[u]);
        }
      }
    }
  }
  for (int i = tid; i < nr * C; i += HT) {
    const int r = i / C, c = i - r * C;
    xh[r * LD + c] = (xh[r * LD + c] - mean[c]) * istd[c];
  }
  __syncthreads();
  // partial db[o] and partial sum_r dy[r][o] x_hat[r][c] over this CTA's rows
  for (int i = tid; i < T * C + T; i += HT) {
    float s = 0.f;
    if (i < T * C) {
      const int o = i / C, c = i - o * C;
      for (int r = 0; r < nr; ++r) s = fmaf(dy[r * LT + o], xh[r * LD + c], s);
    } else {
      const int o = i - T * C;
      for (int r = 0; r < nr; ++r) s += dy[r * LT + o];
    }
    pW[i] = s;
  }
  // dz[r][c] = sum_o dy[r][o] W[o][c]
  for (int i = tid; i < nr * C; i += HT) {
    const int r = i / C, c = i - r * C;
    float s = 0.f;
    for (int o = 0; o < T; ++o) s = fmaf(dy[r * LT + o], Ws[o * LD + c], s);
    dz[r * LD + c] = s;
  }
  __syncthreads();
  column_partial(nr, C, scratch, ps, [&](int r, int c) { return dz[r * LD + c]; });                     // -> dbeta
  column_partial(nr, C, scratch, ps + C, [&](int r, int c) { return dz[r * LD + c] * xh[r * LD + c]; });  // -> dgamma
  cluster.sync();
  cluster_gather(cluster, ps, 2 * C, s12);
  if (rank == 0)
    for (int c = tid; c < C; c += HT) {
      if (a.dbeta) a.dbeta[c] = s12[c];
      if (a.dgamma) a.dgamma[c] = s12[C + c];
    }
  // parameter gradients: element e is finished by CTA e % nc (partials of all CTAs in rank order)
  for (int e = rank + nc * tid; e < T * C + T; e += nc * HT) {
    float s = 0.f;
    for (int k = 0; k < nc; ++k) s += cluster.map_shared_rank(pW, k)[e];
    if (e < T * C) {
      const int o = e / C, c = e - o * C;
      // the Linear's input is x_hat gamma + beta:  dW = gamma_c sum dy x_hat + beta_c db[o]
      float dbo = 0.f;
      if (a.beta) {
        for (int k = 0; k < nc; ++k) dbo += cluster.map_shared_rank(pW, k)[T * C + o];
        dbo *= gb[C + c];
      }
      a.dW[e] = fmaf(gb[c], s, dbo);
    } else {
      a.db[e - T * C] = s;
    }
  }
  for (int i = tid; i < nr * C; i += HT) {
    const int r = i / C, c = i - r * C;
    const float gm = gb[c];
    float g = dz[r * LD + c] * gm;
    if (a.training) g -= gm * (s12[c] + xh[r * LD + c] * s12[C + c]) * (1.f / B);
    a.dx[(size_t)(r0 + r) * C + c] = g * istd[c];
  }
  cluster.sync();   // nobody leaves while its shared memory may still be read
}

template <typename K, typename... Args>
cudaError_t launch_cluster(K kernel, int nc, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nc, 1, 1);
  cfg.blockDim = dim3(HT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace

extern "C" {

int mpnn_head_supported(int B, int C, int T) {
  if (B <= 0 || C <= 0 || T <= 0 || C > 4096 || T > 4096) return 0;
  return head_bwd_floats(B, C, T) * sizeof(float) <= HEAD_SMEM_MAX &&
                 head_fwd_floats(B, C, T) * sizeof(float) <= HEAD_SMEM_MAX
             ? 1
             : 0;
}

// BatchNorm1d(C) -> Linear(C, T) -> mean squared error against target [B,T].
// gamma/beta NULL = no affine; running_* NULL = statistics not tracked (training only); num_batches_tracked may be NULL.
// Outputs: y [B,T] (predictions), loss [1], stats [2C] (saved for backward).
int mpnn_head_bn_linear_mse_fwd(const float* x, const float* target, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, long long* num_batches_tracked,
                                const float* W, const float* b, int B, int C, int T, int training, float momentum,
                                float eps, float* y, float* loss, float* stats, cudaStream_t stream) {
  MPNN_REQUIRE(mpnn_head_supported(B, C, T), MPNN_ERR_UNSUPPORTED, "head_fwd: B=%d C=%d T=%d exceeds one cluster", B, C,
               T);
  MPNN_REQUIRE(training || (running_mean && running_var), MPNN_ERR_ARG, "head_fwd: eval mode needs running statistics");
  HeadArgs a = {x, target, gamma, beta, W, b, running_mean, running_var, num_batches_tracked, B, C, T, training,
                momentum, eps};
  const size_t smem = head_fwd_floats(B, C, T) * sizeof(float);
  MPNN_CUDA(cudaFuncSetAttribute(k_head_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM_MAX));
  MPNN_CUDA(launch_cluster(k_head_fwd, head_cluster(B), smem, stream, a, y, loss, stats));
  return MPNN_OK;
}

// gloss: DEVICE scalar (gradient of the loss).  Writes dx [B,C], dgamma/dbeta [C] (may be NULL), dW [T,C], db [T].
int mpnn_head_bn_linear_mse_bwd(const float* x, const float* target, const float* gamma, const float* beta,
                                const float* W, const float* y, const float* stats, const float* gloss, int B, int C,
                                int T, int training, float* dx, float* dgamma, float* dbeta, float* dW, float* db,
                                cudaStream_t stream) {
  MPNN_REQUIRE(mpnn_head_supported(B, C, T), MPNN_ERR_UNSUPPORTED, "head_bwd: B=%d C=%d T=%d exceeds one cluster", B, C,
               T);
  HeadBwd a = {x, target, gamma, beta, W, y, stats, gloss, B, C, T, training, dx, dgamma, dbeta, dW, db};
  const size_t smem = head_bwd_floats(B, C, T) * sizeof(float);
  MPNN_CUDA(cudaFuncSetAttribute(k_head_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM_MAX));
  MPNN_CUDA(launch_cluster(k_head_bwd, head_cluster(B), smem, stream, a));
  return MPNN_OK;
}

}  // extern "C"
