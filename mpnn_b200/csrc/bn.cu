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

constexpr int MAXQ = 7;

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
  unsigned int* counter;  // zero on entry (workspace contract); the last block resets it to zero
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
  // all nq sums of the block behind ONE barrier pair (sm is [MAXQ][blockDim])
#pragma unroll
  for (int k = 0; k < MAXQ; ++k)
    if (k < a.nq) sm[k * blockDim.x + threadIdx.x] = q[k];
  __syncthreads();
  if (rl == 0) {
    for (int k = 0; k < a.nq; ++k) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += sm[k * blockDim.x + l * C + c];
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

// ---------------------------------------------------------------------------------------------------
// Forward statistics in ONE launch.  The reference's two passes (mean, then the variance of the centred
// values) become: every block centres its rows on its OWN mean (second look at rows that are still in L1),
// and the last block to finish combines the per-block moments exactly:
//   sum_r mu_r^2 (x_r - m)^2 = sum_b [ A_b + 2 (s_b - m) B_b + (s_b - m)^2 C_b ],
//   A_b = sum mu^2 (x - s_b)^2, B_b = sum mu^2 (x - s_b), C_b = sum mu^2, s_b = the block's shift.
// Valid for any mask values; no cancellation (the shift is within the data).  Fixed combination order.
// partial[block][7][C] = {sum x, sum x mu, sum mu, shift, A, B, C}
// ---------------------------------------------------------------------------------------------------
struct StatArgs {
  const float* x;
  const float* mask;
  long long rows;
  int C;
  int rows_per_block;
  int one_d;            // 0: MaskBatchNorm (mean = unmasked sum / M, sd = sqrt(var + eps)); 1: MaskBatchNorm1d
  float eps, momentum;
  float* stats;         // [2C+1] mean | sd | M
  float* running_mean;
  float* running_var;
  unsigned int* counter;
};

__global__ void k_bn_stats(StatArgs a, float* __restrict__ partial) {
  extern __shared__ float sm[];      // [blockDim] reduction scratch, then [C] shifts
  __shared__ int is_last;
  const int C = a.C;
  const int lanes = blockDim.x / C;
  const int c = threadIdx.x % C, rl = threadIdx.x / C;
  float* shift = sm + blockDim.x;
  long long r0 = (long long)blockIdx.x * a.rows_per_block;
  long long r1 = r0 + a.rows_per_block;
  if (r1 > a.rows) r1 = a.rows;
  float q[3] = {0.f, 0.f, 0.f};
  if (rl < lanes) {
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const float xv = a.x[r * C + c];
      const float mu = a.mask[r];
      q[0] += xv;
      q[1] += xv * mu;
      q[2] += mu;
    }
  }
  float* prt = partial + (size_t)blockIdx.x * 7 * C;
  float* r3 = sm + blockDim.x + C;   // [3][blockDim]: the three sums of a pass are reduced behind ONE barrier pair
#pragma unroll
  for (int k = 0; k < 3; ++k) r3[k * blockDim.x + threadIdx.x] = q[k];
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += r3[k * blockDim.x + l * C + c];
      q[k] = s;
      prt[k * C + c] = s;
    }
  }
  if (rl == 0) {
    const float sh = q[2] > 0.f ? q[1] / q[2] : 0.f;
    shift[c] = sh;
    prt[3 * C + c] = sh;
  }
  __syncthreads();
  float w[3] = {0.f, 0.f, 0.f};
  if (rl < lanes) {
    const float sh = shift[c];
#pragma unroll 4
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const float xv = a.x[r * C + c];
      const float mu = a.mask[r];
      const float m2 = mu * mu, dv = xv - sh;
      w[0] = fmaf(m2 * dv, dv, w[0]);
      w[1] = fmaf(m2, dv, w[1]);
      w[2] += m2;
    }
  }
  __syncthreads();   // the first pass's reads of r3 are done
#pragma unroll
  for (int k = 0; k < 3; ++k) r3[k * blockDim.x + threadIdx.x] = w[k];
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += r3[k * blockDim.x + l * C + c];
      prt[(4 + k) * C + c] = s;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- last block: column c, slice rl takes blocks rl, rl + lanes, ...; slices are combined in a fixed order ----
  const int nblk = gridDim.x;
  float* s3 = sm + blockDim.x + C;   // [3][blockDim]
  float S0 = 0.f, S1 = 0.f, M = 0.f;
  if (rl < lanes) {
#pragma unroll 4
    for (int b = rl; b < nblk; b += lanes) {
      const float* pb = partial + (size_t)b * 7 * C + c;
      S0 += __ldcg(pb);
      S1 += __ldcg(pb + C);
      M += __ldcg(pb + 2 * C);
    }
  }
  __syncthreads();
  s3[threadIdx.x] = S0;
  s3[blockDim.x + threadIdx.x] = S1;
  s3[2 * blockDim.x + threadIdx.x] = M;
  __syncthreads();
  if (rl == 0) {
    S0 = S1 = M = 0.f;
    for (int l = 0; l < lanes; ++l) {
      S0 += s3[l * C + c];
      S1 += s3[blockDim.x + l * C + c];
      M += s3[2 * blockDim.x + l * C + c];
    }
    shift[c] = (a.one_d ? S1 : S0) / M;   // the mean
    if (c == 0) a.stats[2 * C] = M;
    sm[c] = M;
  }
  __syncthreads();
  const float mean = shift[c];
  float V = 0.f;
  if (rl < lanes) {
#pragma unroll 4
    for (int b = rl; b < nblk; b += lanes) {
      const float* pb = partial + (size_t)b * 7 * C + c;
      const float dlt = __ldcg(pb + 3 * C) - mean;
      V += __ldcg(pb + 4 * C) + 2.f * dlt * __ldcg(pb + 5 * C) + dlt * dlt * __ldcg(pb + 6 * C);
    }
  }
  const float Mtot = sm[c];
  __syncthreads();
  s3[threadIdx.x] = V;
  __syncthreads();
  if (rl == 0) {
    V = 0.f;
    for (int l = 0; l < lanes; ++l) V += s3[l * C + c];
    const float var = fmaxf(V, 0.f) / Mtot;
    a.stats[c] = mean;
    if (!a.one_d) {
      a.stats[C + c] = sqrtf(var + a.eps);
    } else {
      a.stats[C + c] = sqrtf(var);
      if (a.running_mean) {
        a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
        a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * var;
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
    case STATS1: k_bn_stage<STATS1><<<nblk, threads, MAXQ * threads * sizeof(float), stream>>>(a, partial); break;
    case STATS2: k_bn_stage<STATS2><<<nblk, threads, MAXQ * threads * sizeof(float), stream>>>(a, partial); break;
    case BWD_PLAIN: k_bn_stage<BWD_PLAIN><<<nblk, threads, MAXQ * threads * sizeof(float), stream>>>(a, partial); break;
    case BWD_1D_TRAIN: k_bn_stage<BWD_1D_TRAIN><<<nblk, threads, MAXQ * threads * sizeof(float), stream>>>(a, partial); break;
    default: k_bn_stage<BWD_1D_EVAL><<<nblk, threads, MAXQ * threads * sizeof(float), stream>>>(a, partial); break;
  }
  MPNN_CHECK_LAUNCH("k_bn_stage");
  return MPNN_OK;
}

int run_stats(StatArgs a, float* partial, cudaStream_t stream) {
  int rpb;
  int nblk = red_blocks(a.rows, &rpb);
  a.rows_per_block = rpb;
  int threads = a.C >= 256 ? a.C : (256 / a.C) * a.C;
  k_bn_stats<<<nblk, threads, (4 * threads + a.C) * sizeof(float), stream>>>(a, partial);
  MPNN_CHECK_LAUNCH("k_bn_stats");
  return MPNN_OK;
}

// ---------------------------------------------------------------------------------------------------
// Masked batch norms of a categorical bond tensor in ROW SPACE (graph.TypedBonds): the dense tensor
// [B, N, N, F] has R distinct rows x_u with multiplicity cnt_u and mask (adjacency) value a_u, so every sum over its
// B N^2 rows is a count-weighted sum over R rows (R ~ tens).  One launch each way, one warp per column:
//   mean = sum_u wm_u x_u / M,  wm = cnt a (MaskBatchNorm1d, mask_batch_norm.py:24) or cnt (MaskBatchNorm, :13: unmasked sum)
//   var  = sum_u cnt_u ((x_u - mean) a_u)^2 / M,   M = sum_u cnt_u a_u
//   y_u  = (gamma (x_u - mean) / s + beta) a_u,    s = sqrt(var + eps) (MaskBatchNorm) or sqrt(var) + eps (MaskBatchNorm1d)
// (replaces ~15 aten kernels forward and as many backward on [R, F] tensors)
// ---------------------------------------------------------------------------------------------------
struct RowBN {
  const float *x, *a, *cnt;     // [R, F], [R], [R]
  const float *gamma, *beta;    // [F] or NULL
  float *running_mean, *running_var;   // [F] or NULL (updated in place when training)
  int R, F, masked_mean, eps_inside, training;
  float momentum, eps;
};

// stats: mean [F] | s [F] | sqrt(var) [F] | M
__global__ void __launch_bounds__(256) k_row_bn_fwd(RowBN p, float* __restrict__ y, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float M = 0.f;
  for (int u = lane; u < p.R; u += 32) M = fmaf(p.cnt[u], p.a[u], M);
  M = warp_sum(M);
  if (threadIdx.x == 0) stats[3 * p.F] = M;
  for (int j = warp; j < p.F; j += 8) {
    float mean, s, sv;
    if (p.training) {
      float acc = 0.f;
      for (int u = lane; u < p.R; u += 32) acc = fmaf(p.cnt[u] * (p.masked_mean ? p.a[u] : 1.f), p.x[(size_t)u * p.F + j], acc);
      mean = warp_sum(acc) / M;
      acc = 0.f;
      for (int u = lane; u < p.R; u += 32) {
        const float c = (p.x[(size_t)u * p.F + j] - mean) * p.a[u];
        acc = fmaf(p.cnt[u] * c, c, acc);
      }
      const float var = warp_sum(acc) / M;
      sv = sqrtf(var);
      s = p.eps_inside ? sqrtf(var + p.eps) : sv + p.eps;
      if (lane == 0 && p.running_mean) {
        p.running_mean[j] = (1.f - p.momentum) * p.running_mean[j] + p.momentum * mean;
        p.running_var[j] = (1.f - p.momentum) * p.running_var[j] + p.momentum * var;
      }
    } else {
      mean = p.running_mean[j];
      sv = sqrtf(p.running_var[j]);
      s = sv + p.eps;
    }
    if (lane == 0) {
      stats[j] = mean;
      stats[p.F + j] = s;
      stats[2 * p.F + j] = sv;
    }
    const float ga = p.gamma ? p.gamma[j] : 1.f, be = p.beta ? p.beta[j] : 0.f;
    for (int u = lane; u < p.R; u += 32)
      y[(size_t)u * p.F + j] = fmaf(ga, (p.x[(size_t)u * p.F + j] - mean) / s, be) * p.a[u];
  }
}

__global__ void __launch_bounds__(256) k_row_bn_bwd(RowBN p, const float* __restrict__ stats, const float* __restrict__ dy,
                                                    float* __restrict__ dx, float* __restrict__ dgamma,
                                                    float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float M = stats[3 * p.F];
  for (int j = warp; j < p.F; j += 8) {
    const float mean = stats[j], s = stats[p.F + j], sv = stats[2 * p.F + j];
    const float ga = p.gamma ? p.gamma[j] : 1.f;
    float sg = 0.f, sgx = 0.f, a1 = 0.f;
    for (int u = lane; u < p.R; u += 32) {
      const float au = p.a[u], xc = p.x[(size_t)u * p.F + j] - mean;
      const float g = dy[(size_t)u * p.F + j] * au;
      sg += g;
      sgx = fmaf(g, xc, sgx);
      a1 = fmaf(p.cnt[u] * au * au, xc, a1);
    }
    sg = warp_sum(sg);
    sgx = warp_sum(sgx);
    a1 = warp_sum(a1) / M;
    if (lane == 0) {
      if (dbeta) dbeta[j] = sg;
      if (dgamma) dgamma[j] = sgx / s;
    }
    if (!dx) continue;
    if (!p.training) {
      for (int u = lane; u < p.R; u += 32) dx[(size_t)u * p.F + j] = dy[(size_t)u * p.F + j] * p.a[u] * ga / s;
      continue;
    }
    // S1 = sum dxh, S2 = sum dxh (x - mean) with dxh = g gamma;  ds/dvar
    const float S1 = ga * sg, S2 = ga * sgx;
    const float sp = p.eps_inside ? 0.5f / s : 0.5f / fmaxf(sv, 1e-30f);
    const float k2 = S2 / (s * s) * sp;
    for (int u = lane; u < p.R; u += 32) {
      const float au = p.a[u], cu = p.cnt[u], xc = p.x[(size_t)u * p.F + j] - mean;
      const float wm = cu * (p.masked_mean ? au : 1.f) / M;
      const float dvar = 2.f * cu * au * au * xc / M - 2.f * a1 * wm;
      dx[(size_t)u * p.F + j] = dy[(size_t)u * p.F + j] * au * ga / s - S1 / s * wm - k2 * dvar;
    }
  }
}

}  // namespace

extern "C" {

// ---- masked batch norms in row space (graph.TypedBonds): x [R, F] distinct rows, a [R] mask values, cnt [R] counts ----
// masked_mean: 1 = MaskBatchNorm1d's mean (mask_batch_norm.py:24), 0 = MaskBatchNorm's unmasked sum (:13);
// eps_inside: 1 = sqrt(var + eps) (MaskBatchNorm), 0 = sqrt(var) + eps (MaskBatchNorm1d); gamma / beta / running_* may be
// NULL; training = 0 normalises with the running statistics.  stats [3F + 1] is saved for the backward.
int mpnn_row_bn_fwd(const float* x, const float* a, const float* cnt, int R, int F, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, int masked_mean, int eps_inside, int training,
                    float momentum, float eps, float* y, float* stats, cudaStream_t stream) {
  MPNN_REQUIRE(x && a && cnt && y && stats && R > 0 && F > 0, MPNN_ERR_ARG, "row_bn_fwd: bad argument");
  MPNN_REQUIRE(training || (running_mean && running_var), MPNN_ERR_ARG, "row_bn_fwd: eval mode needs running statistics");
  RowBN p = {x, a, cnt, gamma, beta, running_mean, running_var, R, F, masked_mean, eps_inside, training, momentum, eps};
  k_row_bn_fwd<<<1, 256, 0, stream>>>(p, y, stats);
  MPNN_CHECK_LAUNCH("k_row_bn_fwd");
  return MPNN_OK;
}

// dx [R, F] (may be NULL), dgamma / dbeta [F] (may be NULL)
int mpnn_row_bn_bwd(const float* x, const float* a, const float* cnt, int R, int F, const float* gamma, const float* stats,
                    const float* dy, int masked_mean, int eps_inside, int training, float* dx, float* dgamma,
                    float* dbeta, cudaStream_t stream) {
  MPNN_REQUIRE(x && a && cnt && stats && dy && R > 0 && F > 0, MPNN_ERR_ARG, "row_bn_bwd: bad argument");
  RowBN p = {x, a, cnt, gamma, nullptr, nullptr, nullptr, R, F, masked_mean, eps_inside, training, 0.f, 0.f};
  k_row_bn_bwd<<<1, 256, 0, stream>>>(p, stats, dy, dx, dgamma, dbeta);
  MPNN_CHECK_LAUNCH("k_row_bn_bwd");
  return MPNN_OK;
}


// workspace: partial sums + reduced vectors + the completion counter.  CONTRACT: the workspace is zero-filled ONCE by
// the caller; every call leaves the counter word zero again, so a workspace can be reused call after call on one
// stream without a memset node per launch (each one costs ~4 us of a ~500 us graph-replayed step).
size_t mpnn_bn_workspace_bytes(long long rows, int C) {
  int rpb;
  int nblk = red_blocks(rows, &rpb);
  return align_up((size_t)nblk * MAXQ * C * sizeof(float), 256) + align_up((size_t)(MAXQ + 2) * C * sizeof(float), 256) +
         256;
}

// The counter sits at offset 0 for EVERY (rows, C): a workspace that is reused across shapes of the same byte size must
// never present one shape's partial sums as another shape's counter (round 2: a batch norm on [rows, C] after one on a
// different [rows', C'] of equal workspace size started with a non-zero counter and normalised with garbage).
static void carve(void* workspace, long long rows, int C, float** partial, float** red, unsigned int** counter) {
  int rpb;
  int nblk = red_blocks(rows, &rpb);
  char* wp = (char*)workspace;
  *counter = (unsigned int*)wp;
  wp += 256;
  *partial = (float*)wp;
  wp += align_up((size_t)nblk * MAXQ * C * sizeof(float), 256);
  *red = (float*)wp;
}

// stats: [2*C + 1] floats (mean, sqrt(var+eps), M) saved for backward
int mpnn_mask_bn_fwd(const float* x, const float* mask, long long rows, int C, float eps, float* y, float* stats,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && C > 0 && C <= 1024, MPNN_ERR_ARG, "mask_bn_fwd: bad dims rows=%lld C=%d", rows, C);
  MPNN_REQUIRE(workspace_bytes >= mpnn_bn_workspace_bytes(rows, C), MPNN_ERR_WORKSPACE, "mask_bn_fwd: workspace");
  float *partial, *red;
  unsigned int* counter;
  carve(workspace, rows, C, &partial, &red, &counter);
  StatArgs a = {x, mask, rows, C, 0, 0, eps, 0.f, stats, nullptr, nullptr, counter};
  int rc = run_stats(a, partial, stream);
  if (rc) return rc;
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
  StatArgs a = {x, mask, rows, C, 0, 1, eps, momentum, stats, running_mean, running_var, counter};
  int rc = run_stats(a, partial, stream);
  if (rc) return rc;
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
