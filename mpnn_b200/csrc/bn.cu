// Masked batch norms that sit inside every message-passing step of the "normed" models
// (reference models/mask_batch_norm.py).  Both are batch-wide column statistics over the [rows, C] view
// (rows = B*N, or B*N*N with adj as the mask: batch_norm_graph_wrapper.py:14), i.e. HBM-bound reductions.
// All reductions are two-stage with a fixed summation order (bit-reproducible, no float atomics).
//
//  MaskBatchNorm   (:9-15):  mean = sum_rows(x)/M  (UNMASKED sum), c = (x-mean)*mu, var = sum(c^2)/M,
//                            y = c / sqrt(var + eps)
//  MaskBatchNorm1d (:20-38): mean = sum(x*mu)/M, c = (x-mean)*mu, var = sum(c^2)/M (biased),
//                            train: y = ((x-mean)/(sqrt(var)+eps) * w + b) * mu, running <- .9 old + .1 new
//                            eval : y = ((x-rm)/(sqrt(rv)+eps) * w + b) * mu
#include "common.cuh"

namespace {

constexpr int MAXQ = 5;

enum Mode { STATS1 = 0, STATS2 = 1, BWD_PLAIN = 2, BWD_1D_TRAIN = 3, BWD_1D_EVAL = 4 };
enum Fin { FIN_NONE = 0, FIN_MEAN_PLAIN = 1, FIN_MEAN_1D = 2, FIN_SD_PLAIN = 3, FIN_SD_1D = 4 };

struct RedArgs {
  const float* x;
  const float* mask;
  const float* dy;
  const float* p0;  // per-column parameter vectors (meaning depends on mode)
  const float* p1;
  const float* p2;
  long long rows;
  int C;
  int rows_per_block;
  int mode;
  int nq;
  // finalisation, done by the LAST block to finish (fixed summation order over the per-block partials)
  int fin;
  float eps, momentum;
  float* red;           // [nq*C] reduced sums
  float* stats;         // [2C+1] mean | sd | M
  float* running_mean;  // optional (1d, training)
  float* running_var;
  unsigned int* counter;  // zeroed by the caller
};

// thread -> (column tid % C, row lane tid / C); partial[block][q][C]
template <int MODE>
__global__ void k_bn_stage(RedArgs a, float* __restrict__ partial) {
  extern __shared__ float sm[];
  __shared__ int is_last;
  const int C = a.C;
  const int lanes = blockDim.x / C;
  const int c = threadIdx.x % C, rl = threadIdx.x / C;
  long long r0 = (long long)blockIdx.x * a.rows_per_block;
  long long r1 = r0 + a.rows_per_block;
  if (r1 > a.rows) r1 = a.rows;
  float q[MAXQ] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < lanes) {
    const float m0 = a.p0 ? a.p0[c] : 0.f;
    float m1 = a.p1 ? a.p1[c] : 0.f;
    const float m2 = a.p2 ? a.p2[c] : 1.f;  // weight (affine off -> 1)
    if (MODE == BWD_1D_TRAIN) m1 = 1.f / (m1 + a.eps);           // p1 = sd
    if (MODE == BWD_1D_EVAL) m1 = 1.f / (sqrtf(m1) + a.eps);     // p1 = running_var
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const float xv = a.x[r * C + c];
      const float mu = a.mask[r];
      switch (MODE) {
        case STATS1:
          q[0] += xv;
          q[1] += xv * mu;
          q[2] += mu;   // M = sum(mask): sums of 0/1 values are exact, any order
          break;
        case STATS2: {  // p0 = mean
          float cc = (xv - m0) * mu;
          q[0] += cc * cc;
          break;
        }
        case BWD_PLAIN: {  // p0 = mean
          float cc = (xv - m0) * mu;
          float d = a.dy[r * C + c];
          q[0] += d * cc;
          q[1] += d * mu;
          q[2] += cc * mu;
          break;
        }
        case BWD_1D_TRAIN: {  // p0 = mean, m1 = 1/(s+eps), p2 = weight
          float d = a.dy[r * C + c] * mu;
          float xc = xv - m0;
          float dyh = d * m2;
          q[0] += dyh;
          q[1] += dyh * xc;
          q[2] += xc * mu * mu;
          q[3] += d * xc * m1;
          q[4] += d;
          break;
        }
        default: {  // BWD_1D_EVAL: p0 = running_mean, m1 = 1/(sqrt(rv)+eps)
          float d = a.dy[r * C + c] * mu;
          q[0] += d * (xv - m0) * m1;
          q[1] += d;
          break;
        }
      }
    }
  }
  for (int k = 0; k < a.nq; ++k) {
    __syncthreads();
    sm[threadIdx.x] = q[k];
    __syncthreads();
    if (rl == 0) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += sm[l * C + c];
      partial[((size_t)blockIdx.x * a.nq + k) * C + c] = s;
    }
  }
  // ---- last block: reduce the partials (fixed order) and derive the statistics ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int nqC = a.nq * C;
  const int nblk = gridDim.x;
  const int nsl = blockDim.x >= nqC ? blockDim.x / nqC : 1;
  for (int e0 = 0; e0 < nqC; e0 += blockDim.x) {
    const int e = e0 + (int)threadIdx.x % (nsl > 1 ? nqC : blockDim.x);
    const int sl = nsl > 1 ? threadIdx.x / nqC : 0;
    float s = 0.f;
    if (e < nqC && sl < nsl) {
#pragma unroll 8
      for (int b = sl; b < nblk; b += nsl) s += __ldcg(partial + (size_t)b * nqC + e);
    }
    __syncthreads();
    sm[threadIdx.x] = s;
    __syncthreads();
    if (sl == 0 && e < nqC) {
      float t = 0.f;
      for (int j = 0; j < nsl; ++j) t += sm[j * nqC + (e - e0)];
      a.red[e] = t;
    }
  }
  __syncthreads();
  if (a.fin == FIN_NONE) {
    if (threadIdx.x == 0) *a.counter = 0u;
    return;
  }
  for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
    if (a.fin == FIN_MEAN_PLAIN || a.fin == FIN_MEAN_1D) {
      const float M = a.red[2 * C];
      a.stats[cc] = a.red[(a.fin == FIN_MEAN_1D ? C : 0) + cc] / M;
      if (cc == 0) a.stats[2 * C] = M;
    } else {
      const float var = a.red[cc] / a.stats[2 * C];
      if (a.fin == FIN_SD_PLAIN) {
        a.stats[C + cc] = sqrtf(var + a.eps);
      } else {
        a.stats[C + cc] = sqrtf(var);
        if (a.running_mean) {
          a.running_mean[cc] = (1.f - a.momentum) * a.running_mean[cc] + a.momentum * a.stats[cc];
          a.running_var[cc] = (1.f - a.momentum) * a.running_var[cc] + a.momentum * var;
        }
      }
    }
  }
  if (threadIdx.x == 0) *a.counter = 0u;
}

// stats layout written by the forward and consumed by the backward: [mean | scale | M (1 float, at 2C)]
__global__ void k_plain_apply(const float* __restrict__ x, const float* __restrict__ mask,
                              const float* __restrict__ stats, long long rows, int C, float* __restrict__ y) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C) return;
  long long r = t / C;
  int c = (int)(t - r * C);
  y[t] = ((x[t] - stats[c]) * mask[r]) / stats[C + c];
}
// dx = dc*mu - S2/M,  dc = dy/s - c*S1/(s^3 M),  S2 = sum(dy*mu)/s - S1*sum(c*mu)/(s^3 M)
__global__ void k_plain_bwd_apply(const float* __restrict__ x, const float* __restrict__ mask,
                                  const float* __restrict__ dy, const float* __restrict__ stats,
                                  const float* __restrict__ red, long long rows, int C, float* __restrict__ dx) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C) return;
  long long r = t / C;
  int c = (int)(t - r * C);
  const float M = stats[2 * C], mean = stats[c], s = stats[C + c];
  const float S1 = red[c], Sdm = red[C + c], Scm = red[2 * C + c];
  const float k = S1 / (s * s * s * M);
  const float S2 = Sdm / s - k * Scm;
  const float mu = mask[r];
  const float cc = (x[t] - mean) * mu;
  const float dc = dy[t] / s - cc * k;
  dx[t] = dc * mu - S2 / M;
}

// --- MaskBatchNorm1d ---
__global__ void k_1d_apply(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ mean,
                           const float* __restrict__ sd, int sd_is_var, const float* __restrict__ w,
                           const float* __restrict__ b, float eps, long long rows, int C, float* __restrict__ y) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C) return;
  long long r = t / C;
  int c = (int)(t - r * C);
  float s = sd_is_var ? sqrtf(sd[c]) : sd[c];
  float v = (x[t] - mean[c]) / (s + eps);
  if (w) v = w[c] * v + b[c];
  y[t] = v * mask[r];
}
// train: red = [A1 | A2 | A3 | dgamma | dbeta]
__global__ void k_1d_bwd_apply_train(const float* __restrict__ x, const float* __restrict__ mask,
                                     const float* __restrict__ dy, const float* __restrict__ w,
                                     const float* __restrict__ stats, float eps, const float* __restrict__ red,
                                     long long rows, int C, float* __restrict__ dx, float* __restrict__ dweight,
                                     float* __restrict__ dbias) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < C) {
    if (dweight) dweight[t] = red[3 * C + t];
    if (dbias) dbias[t] = red[4 * C + t];
  }
  if (t >= rows * C) return;
  long long r = t / C;
  int c = (int)(t - r * C);
  const float M = stats[2 * C], mean = stats[c], s = stats[C + c];
  const float iv = 1.f / (s + eps);
  const float A1 = red[c], A2 = red[C + c], A3 = red[2 * C + c];
  const float mu = mask[r];
  const float gam = w ? w[c] : 1.f;
  const float dyh = dy[t] * mu * gam;
  // ds = -A2 * iv^2 ; dvar = ds / (2 s) ; dc = dvar * 2 c / M
  const float dvar = s > 0.f ? (-A2 * iv * iv) / (2.f * s) : 0.f;
  const float cc = (x[t] - mean) * mu;
  // dmean = -A1*iv - sum(dc*mu) = -A1*iv - dvar*2*A3/M
  const float dmean = -A1 * iv - dvar * 2.f * A3 / M;
  dx[t] = dyh * iv + dvar * 2.f * cc / M * mu + dmean * mu / M;
}
__global__ void k_1d_bwd_apply_eval(const float* __restrict__ mask, const float* __restrict__ dy,
                                    const float* __restrict__ w, const float* __restrict__ running_var, float eps,
                                    const float* __restrict__ red, long long rows, int C, float* __restrict__ dx,
                                    float* __restrict__ dweight, float* __restrict__ dbias) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < C) {
    if (dweight) dweight[t] = red[t];
    if (dbias) dbias[t] = red[C + t];
  }
  if (t >= rows * C) return;
  long long r = t / C;
  int c = (int)(t - r * C);
  dx[t] = dy[t] * mask[r] * (w ? w[c] : 1.f) / (sqrtf(running_var[c]) + eps);
}

int red_blocks(long long rows, int* rpb) {
  int target = 2 * mpnn_num_sms();
  long long per = (rows + target - 1) / target;
  if (per < 128) per = 128;   // small inputs: fewer partials for the last block to sum
  *rpb = (int)per;
  return (int)((rows + per - 1) / per);
}

int run_stage(RedArgs a, float* partial, cudaStream_t stream) {
  int rpb;
  int nblk = red_blocks(a.rows, &rpb);
  a.rows_per_block = rpb;
  int threads = a.C >= 256 ? a.C : (256 / a.C) * a.C;
  switch (a.mode) {
    case STATS1: k_bn_stage<STATS1><<<nblk, threads, threads * sizeof(float), stream>>>(a, partial); break;
    case STATS2: k_bn_stage<STATS2><<<nblk, threads, threads * sizeof(float), stream>>>(a, partial); break;
    case BWD_PLAIN: k_bn_stage<BWD_PLAIN><<<nblk, threads, threads * sizeof(float), stream>>>(a, partial); break;
    case BWD_1D_TRAIN: k_bn_stage<BWD_1D_TRAIN><<<nblk, threads, threads * sizeof(float), stream>>>(a, partial); break;
    default: k_bn_stage<BWD_1D_EVAL><<<nblk, threads, threads * sizeof(float), stream>>>(a, partial); break;
  }
  MPNN_CHECK_LAUNCH("k_bn_stage");
  return MPNN_OK;
}

}  // namespace

extern "C" {

// workspace: partial sums + reduced vectors + the completion counter
size_t mpnn_bn_workspace_bytes(long long rows, int C) {
  int rpb;
  int nblk = red_blocks(rows, &rpb);
  return align_up((size_t)nblk * MAXQ * C * sizeof(float), 256) + align_up((size_t)(MAXQ + 2) * C * sizeof(float), 256) +
         256;
}

static void carve(void* workspace, long long rows, int C, float** partial, float** red, unsigned int** counter) {
  int rpb;
  int nblk = red_blocks(rows, &rpb);
  char* wp = (char*)workspace;
  *partial = (float*)wp;
  wp += align_up((size_t)nblk * MAXQ * C * sizeof(float), 256);
  *red = (float*)wp;
  wp += align_up((size_t)(MAXQ + 2) * C * sizeof(float), 256);
  *counter = (unsigned int*)wp;
}

// stats: [2*C + 1] floats (mean, sqrt(var+eps), M) saved for backward
int mpnn_mask_bn_fwd(const float* x, const float* mask, long long rows, int C, float eps, float* y, float* stats,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && C > 0 && C <= 1024, MPNN_ERR_ARG, "mask_bn_fwd: bad dims rows=%lld C=%d", rows, C);
  MPNN_REQUIRE(workspace_bytes >= mpnn_bn_workspace_bytes(rows, C), MPNN_ERR_WORKSPACE, "mask_bn_fwd: workspace");
  float *partial, *red;
  unsigned int* counter;
  carve(workspace, rows, C, &partial, &red, &counter);
  MPNN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  RedArgs a = {x, mask, nullptr, nullptr, nullptr, nullptr, rows, C, 0, STATS1, 3,
               FIN_MEAN_PLAIN, eps, 0.f, red, stats, nullptr, nullptr, counter};
  int rc = run_stage(a, partial, stream);
  if (rc) return rc;
  RedArgs b = {x, mask, nullptr, stats, nullptr, nullptr, rows, C, 0, STATS2, 1,
               FIN_SD_PLAIN, eps, 0.f, red, stats, nullptr, nullptr, counter};
  if ((rc = run_stage(b, partial, stream))) return rc;
  k_plain_apply<<<ceil_div(rows * C, 256), 256, 0, stream>>>(x, mask, stats, rows, C, y);
  MPNN_CHECK_LAUNCH("mask_bn_fwd");
  return MPNN_OK;
}

int mpnn_mask_bn_bwd(const float* x, const float* mask, const float* dy, const float* stats, long long rows, int C,
                     float* dx, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && C > 0 && C <= 1024, MPNN_ERR_ARG, "mask_bn_bwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_bn_workspace_bytes(rows, C), MPNN_ERR_WORKSPACE, "mask_bn_bwd: workspace");
  float *partial, *red;
  unsigned int* counter;
  carve(workspace, rows, C, &partial, &red, &counter);
  MPNN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  RedArgs a = {x, mask, dy, stats, nullptr, nullptr, rows, C, 0, BWD_PLAIN, 3,
               FIN_NONE, 0.f, 0.f, red, nullptr, nullptr, nullptr, counter};
  int rc = run_stage(a, partial, stream);
  if (rc) return rc;
  k_plain_bwd_apply<<<ceil_div(rows * C, 256), 256, 0, stream>>>(x, mask, dy, stats, red, rows, C, dx);
  MPNN_CHECK_LAUNCH("mask_bn_bwd");
  return MPNN_OK;
}

// training != 0: batch statistics, running buffers updated in place (pass NULL to skip tracking).
// stats: [2*C + 1] floats (mean, sqrt(var), M) saved for backward (training only).
int mpnn_mask_bn1d_fwd(const float* x, const float* mask, const float* weight, const float* bias, float* running_mean,
                       float* running_var, long long rows, int C, int training, float momentum, float eps, float* y,
                       float* stats, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && C > 0 && C <= 1024, MPNN_ERR_ARG, "mask_bn1d_fwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_bn_workspace_bytes(rows, C), MPNN_ERR_WORKSPACE, "mask_bn1d_fwd: workspace");
  if (!training) {
    MPNN_REQUIRE(running_mean && running_var, MPNN_ERR_ARG, "mask_bn1d_fwd: eval mode needs running statistics");
    k_1d_apply<<<ceil_div(rows * C, 256), 256, 0, stream>>>(x, mask, running_mean, running_var, 1, weight, bias, eps,
                                                            rows, C, y);
    MPNN_CHECK_LAUNCH("k_1d_apply");
    return MPNN_OK;
  }
  float *partial, *red;
  unsigned int* counter;
  carve(workspace, rows, C, &partial, &red, &counter);
  MPNN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  RedArgs a = {x, mask, nullptr, nullptr, nullptr, nullptr, rows, C, 0, STATS1, 3,
               FIN_MEAN_1D, eps, momentum, red, stats, nullptr, nullptr, counter};
  int rc = run_stage(a, partial, stream);
  if (rc) return rc;
  RedArgs b = {x, mask, nullptr, stats, nullptr, nullptr, rows, C, 0, STATS2, 1,
               FIN_SD_1D, eps, momentum, red, stats, running_mean, running_var, counter};
  if ((rc = run_stage(b, partial, stream))) return rc;
  k_1d_apply<<<ceil_div(rows * C, 256), 256, 0, stream>>>(x, mask, stats, stats + C, 0, weight, bias, eps, rows, C, y);
  MPNN_CHECK_LAUNCH("mask_bn1d_fwd");
  return MPNN_OK;
}

int mpnn_mask_bn1d_bwd(const float* x, const float* mask, const float* dy, const float* weight, const float* stats,
                       const float* running_mean, const float* running_var, long long rows, int C, int training,
                       float eps, float* dx, float* dweight, float* dbias, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && C > 0 && C <= 1024, MPNN_ERR_ARG, "mask_bn1d_bwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_bn_workspace_bytes(rows, C), MPNN_ERR_WORKSPACE, "mask_bn1d_bwd: workspace");
  float *partial, *red;
  unsigned int* counter;
  carve(workspace, rows, C, &partial, &red, &counter);
  MPNN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  int rc;
  if (training) {
    RedArgs a = {x, mask, dy, stats, stats + C, weight, rows, C, 0, BWD_1D_TRAIN, 5,
                 FIN_NONE, eps, 0.f, red, nullptr, nullptr, nullptr, counter};
    if ((rc = run_stage(a, partial, stream))) return rc;
    k_1d_bwd_apply_train<<<ceil_div(rows * C, 256), 256, 0, stream>>>(x, mask, dy, weight, stats, eps, red, rows, C, dx,
                                                                      dweight, dbias);
  } else {
    RedArgs a = {x, mask, dy, running_mean, running_var, nullptr, rows, C, 0, BWD_1D_EVAL, 2,
                 FIN_NONE, eps, 0.f, red, nullptr, nullptr, nullptr, counter};
    if ((rc = run_stage(a, partial, stream))) return rc;
    k_1d_bwd_apply_eval<<<ceil_div(rows * C, 256), 256, 0, stream>>>(mask, dy, weight, running_var, eps, red, rows, C,
                                                                     dx, dweight, dbias);
  }
  MPNN_CHECK_LAUNCH("mask_bn1d_bwd");
  return MPNN_OK;
}

}  // extern "C"
