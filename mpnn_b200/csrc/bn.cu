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
};

// thread -> (column tid % C, row lane tid / C); partial[block][q][C]
__global__ void k_bn_reduce(RedArgs a, float* __restrict__ partial) {
  extern __shared__ float sm[];
  const int C = a.C;
  const int lanes = blockDim.x / C;
  const int c = threadIdx.x % C, rl = threadIdx.x / C;
  long long r0 = (long long)blockIdx.x * a.rows_per_block;
  long long r1 = r0 + a.rows_per_block;
  if (r1 > a.rows) r1 = a.rows;
  float q[MAXQ] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < lanes) {
    const float m0 = a.p0 ? a.p0[c] : 0.f;
    const float m1 = a.p1 ? a.p1[c] : 0.f;
    const float m2 = a.p2 ? a.p2[c] : 1.f;  // weight (affine off -> 1)
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const float xv = a.x[r * C + c];
      const float mu = a.mask[r];
      switch (a.mode) {
        case STATS1:
          q[0] += xv;
          q[1] += xv * mu;
          break;
        case STATS2: {  // p0 = mean
          float cc = (xv - m0) * mu;
          q[0] += cc * cc;
          q[1] += cc * mu;
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
        case BWD_1D_TRAIN: {  // p0 = mean, p1 = 1/(s+eps), p2 = weight
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
        default: {  // BWD_1D_EVAL: p0 = running_mean, p1 = 1/(sqrt(rv)+eps)
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
}

__global__ void k_bn_final(const float* __restrict__ partial, int nblk, int nq, int C, float* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * C) return;
  int k = t / C, c = t - k * C;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[((size_t)b * nq + k) * C + c];
  out[t] = s;
}

__global__ void k_mask_sum(const float* __restrict__ mask, long long rows, float* __restrict__ out) {
  // single block, fixed order
  __shared__ float sm[256];
  float s = 0.f;
  for (long long r = threadIdx.x; r < rows; r += 256) s += mask[r];
  sm[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += sm[i];
    *out = t;
  }
}

// stats layout written by the forward and consumed by the backward: [mean | scale | M (1 float, at 2C)]
__global__ void k_plain_derive1(const float* __restrict__ sums, const float* __restrict__ M, int C,
                                float* __restrict__ stats) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) stats[c] = sums[c] / *M;  // unmasked sum / M
  if (c == 0) stats[2 * C] = *M;
}
__global__ void k_plain_derive2(const float* __restrict__ sums, int C, float eps, float* __restrict__ stats) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) stats[C + c] = sqrtf(sums[c] / stats[2 * C] + eps);  // s = sqrt(var + eps)
}
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
__global__ void k_1d_derive1(const float* __restrict__ sums, const float* __restrict__ M, int C,
                             float* __restrict__ stats) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) stats[c] = sums[C + c] / *M;  // masked sum / M
  if (c == 0) stats[2 * C] = *M;
}
__global__ void k_1d_derive2(const float* __restrict__ sums, int C, float momentum, float* __restrict__ stats,
                             float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float var = sums[c] / stats[2 * C];
  stats[C + c] = sqrtf(var);
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * stats[c];
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var;
  }
}
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
__global__ void k_1d_inv(const float* __restrict__ sd, int sd_is_var, float eps, int C, float* __restrict__ inv) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) inv[c] = 1.f / ((sd_is_var ? sqrtf(sd[c]) : sd[c]) + eps);
}
// train: red = [A1 | A2 | A3 | dgamma | dbeta]
__global__ void k_1d_bwd_apply_train(const float* __restrict__ x, const float* __restrict__ mask,
                                     const float* __restrict__ dy, const float* __restrict__ w,
                                     const float* __restrict__ stats, const float* __restrict__ inv,
                                     const float* __restrict__ red, long long rows, int C, float* __restrict__ dx) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C) return;
  long long r = t / C;
  int c = (int)(t - r * C);
  const float M = stats[2 * C], mean = stats[c], s = stats[C + c], iv = inv[c];
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
                                    const float* __restrict__ w, const float* __restrict__ inv, long long rows, int C,
                                    float* __restrict__ dx) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * C) return;
  long long r = t / C;
  int c = (int)(t - r * C);
  dx[t] = dy[t] * mask[r] * (w ? w[c] : 1.f) * inv[c];
}

int red_blocks(long long rows, int* rpb) {
  int target = 2 * mpnn_num_sms();
  long long per = (rows + target - 1) / target;
  if (per < 32) per = 32;
  *rpb = (int)per;
  return (int)((rows + per - 1) / per);
}

int run_reduce(RedArgs a, float* partial, float* out, cudaStream_t stream) {
  int rpb;
  int nblk = red_blocks(a.rows, &rpb);
  a.rows_per_block = rpb;
  int threads = a.C >= 256 ? a.C : (256 / a.C) * a.C;
  k_bn_reduce<<<nblk, threads, threads * sizeof(float), stream>>>(a, partial);
  MPNN_CHECK_LAUNCH("k_bn_reduce");
  k_bn_final<<<ceil_div(a.nq * a.C, 128), 128, 0, stream>>>(partial, nblk, a.nq, a.C, out);
  MPNN_CHECK_LAUNCH("k_bn_final");
  return MPNN_OK;
}

}  // namespace

extern "C" {

// workspace: partial sums + reduced vectors
size_t mpnn_bn_workspace_bytes(long long rows, int C) {
  int rpb;
  int nblk = red_blocks(rows, &rpb);
  return align_up((size_t)nblk * MAXQ * C * sizeof(float), 256) + align_up((size_t)(MAXQ + 2) * C * sizeof(float), 256) +
         256;
}

static void carve(void* workspace, long long rows, int C, float** partial, float** red, float** scal) {
  int rpb;
  int nblk = red_blocks(rows, &rpb);
  char* wp = (char*)workspace;
  *partial = (float*)wp;
  wp += align_up((size_t)nblk * MAXQ * C * sizeof(float), 256);
  *red = (float*)wp;
  wp += align_up((size_t)(MAXQ + 2) * C * sizeof(float), 256);
  *scal = (float*)wp;
}

// stats: [2*C + 1] floats (mean, sqrt(var+eps), M) saved for backward
int mpnn_mask_bn_fwd(const float* x, const float* mask, long long rows, int C, float eps, float* y, float* stats,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && C > 0 && C <= 1024, MPNN_ERR_ARG, "mask_bn_fwd: bad dims rows=%lld C=%d", rows, C);
  MPNN_REQUIRE(workspace_bytes >= mpnn_bn_workspace_bytes(rows, C), MPNN_ERR_WORKSPACE, "mask_bn_fwd: workspace");
  float *partial, *red, *scal;
  carve(workspace, rows, C, &partial, &red, &scal);
  k_mask_sum<<<1, 256, 0, stream>>>(mask, rows, scal);
  RedArgs a = {x, mask, nullptr, nullptr, nullptr, nullptr, rows, C, 0, STATS1, 2};
  int rc = run_reduce(a, partial, red, stream);
  if (rc) return rc;
  k_plain_derive1<<<ceil_div(C, 128), 128, 0, stream>>>(red, scal, C, stats);
  RedArgs b = {x, mask, nullptr, stats, nullptr, nullptr, rows, C, 0, STATS2, 2};
  if ((rc = run_reduce(b, partial, red, stream))) return rc;
  k_plain_derive2<<<ceil_div(C, 128), 128, 0, stream>>>(red, C, eps, stats);
  k_plain_apply<<<ceil_div(rows * C, 256), 256, 0, stream>>>(x, mask, stats, rows, C, y);
  MPNN_CHECK_LAUNCH("mask_bn_fwd");
  return MPNN_OK;
}

int mpnn_mask_bn_bwd(const float* x, const float* mask, const float* dy, const float* stats, long long rows, int C,
                     float* dx, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && C > 0 && C <= 1024, MPNN_ERR_ARG, "mask_bn_bwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_bn_workspace_bytes(rows, C), MPNN_ERR_WORKSPACE, "mask_bn_bwd: workspace");
  float *partial, *red, *scal;
  carve(workspace, rows, C, &partial, &red, &scal);
  RedArgs a = {x, mask, dy, stats, nullptr, nullptr, rows, C, 0, BWD_PLAIN, 3};
  int rc = run_reduce(a, partial, red, stream);
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
  float *partial, *red, *scal;
  carve(workspace, rows, C, &partial, &red, &scal);
  k_mask_sum<<<1, 256, 0, stream>>>(mask, rows, scal);
  RedArgs a = {x, mask, nullptr, nullptr, nullptr, nullptr, rows, C, 0, STATS1, 2};
  int rc = run_reduce(a, partial, red, stream);
  if (rc) return rc;
  k_1d_derive1<<<ceil_div(C, 128), 128, 0, stream>>>(red, scal, C, stats);
  RedArgs b = {x, mask, nullptr, stats, nullptr, nullptr, rows, C, 0, STATS2, 2};
  if ((rc = run_reduce(b, partial, red, stream))) return rc;
  k_1d_derive2<<<ceil_div(C, 128), 128, 0, stream>>>(red, C, momentum, stats, running_mean, running_var);
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
  float *partial, *red, *scal;
  carve(workspace, rows, C, &partial, &red, &scal);
  float* inv = red + (size_t)MAXQ * C;
  int rc;
  if (training) {
    k_1d_inv<<<ceil_div(C, 128), 128, 0, stream>>>(stats + C, 0, eps, C, inv);
    RedArgs a = {x, mask, dy, stats, inv, weight, rows, C, 0, BWD_1D_TRAIN, 5};
    if ((rc = run_reduce(a, partial, red, stream))) return rc;
    k_1d_bwd_apply_train<<<ceil_div(rows * C, 256), 256, 0, stream>>>(x, mask, dy, weight, stats, inv, red, rows, C,
                                                                      dx);
    if (dweight) MPNN_CUDA(cudaMemcpyAsync(dweight, red + 3 * C, C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if (dbias) MPNN_CUDA(cudaMemcpyAsync(dbias, red + 4 * C, C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  } else {
    k_1d_inv<<<ceil_div(C, 128), 128, 0, stream>>>(running_var, 1, eps, C, inv);
    RedArgs a = {x, mask, dy, running_mean, inv, nullptr, rows, C, 0, BWD_1D_EVAL, 2};
    if ((rc = run_reduce(a, partial, red, stream))) return rc;
    k_1d_bwd_apply_eval<<<ceil_div(rows * C, 256), 256, 0, stream>>>(mask, dy, weight, inv, rows, C, dx);
    if (dweight) MPNN_CUDA(cudaMemcpyAsync(dweight, red, C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if (dbias) MPNN_CUDA(cudaMemcpyAsync(dbias, red + C, C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  MPNN_CHECK_LAUNCH("mask_bn1d_bwd");
  return MPNN_OK;
}

}  // extern "C"
