// Graph readout: GraphLevelOutput (reference mpnn_functions/readout/graph_level_output.py:30-47).
//   masked   : out[b,:] = sum_i softmax_o(i(x*mu))[b,i,:] * j(x*mu)[b,i,:] * mu[b,i]
//   unmasked : out[b,:] = softmax_o(sum_i i(x)[b,i,:]) * sum_i j(x)[b,i,:]
// i, j are Linear(2*nf -> O).  The two projections run as plain GEMMs (x is read from HBM once per
// projection); the feature softmax, gating, masking and the per-graph segmented sum are one kernel with a
// fixed reduction order.  Saved for backward: u = i(.), v = j(.) [rows, O].
#include "common.cuh"

extern "C" int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
                         long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_gemm_workspace_bytes(int M, int N, int K);
extern "C" int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy,
                           float* out, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_colsum_workspace_bytes(long long rows, int width);

namespace {

constexpr int KMAX = 32;  // O <= 32*KMAX = 1024

// u' = mu*u + bi, v' = mu*v + bj in place; out[b] = sum_i softmax(u') * v' * mu.  One block per graph.
__global__ void __launch_bounds__(256) k_glo_fwd_masked(float* __restrict__ u, float* __restrict__ v,
                                                        const float* __restrict__ mask, const float* __restrict__ bi,
                                                        const float* __restrict__ bj, int N, int O,
                                                        float* __restrict__ out) {
  extern __shared__ float sm[];  // [8][O]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (O + 31) / 32;
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  for (int i = warp; i < N; i += 8) {
    const size_t row = (size_t)b * N + i;
    const float mu = mask[row];
    float* ur = u + row * O;
    float* vr = v + row * O;
    float uv[KMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      int o = lane + 32 * k;
      if (k < nk && o < O) {
        uv[k] = mu * ur[o] + bi[o];
        ur[o] = uv[k];
        mx = fmaxf(mx, uv[k]);
      }
    }
    mx = warp_max(mx);
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      int o = lane + 32 * k;
      if (k < nk && o < O) {
        uv[k] = expf(uv[k] - mx);
        den += uv[k];
      }
    }
    den = warp_sum(den);
    const float inv = 1.f / den;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      int o = lane + 32 * k;
      if (k < nk && o < O) {
        float vv = mu * vr[o] + bj[o];
        vr[o] = vv;
        acc[k] += uv[k] * inv * vv * mu;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < O) sm[warp * O + o] = acc[k];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < O; o += 256) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sm[w * O + o];
    out[(size_t)b * O + o] = s;
  }
}

// du', dv' from dout (masked form).  One warp per row.
__global__ void __launch_bounds__(256) k_glo_bwd_masked(const float* __restrict__ u, const float* __restrict__ v,
                                                        const float* __restrict__ mask,
                                                        const float* __restrict__ dout, long long rows, int N, int O,
                                                        float* __restrict__ du, float* __restrict__ dv) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nk = (O + 31) / 32;
  const long long b = row / N;
  const float mu = mask[row];
  const float* ur = u + row * O;
  const float* vr = v + row * O;
  const float* dr = dout + b * O;
  float s[KMAX];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < O) {
      s[k] = ur[o];
      mx = fmaxf(mx, s[k]);
    }
  }
  mx = warp_max(mx);
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < O) {
      s[k] = expf(s[k] - mx);
      den += s[k];
    }
  }
  den = warp_sum(den);
  const float inv = 1.f / den;
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < O) {
      s[k] *= inv;
      dot += dr[o] * mu * vr[o] * s[k];
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < O) {
      float dg = dr[o] * mu;
      dv[row * O + o] = dg * s[k];
      du[row * O + o] = s[k] * (dg * vr[o] - dot);
    }
  }
}

// unmasked: U = sum_i (u + bi), V = sum_i (v + bj); out = softmax(U) * V.  One block per graph; saves U,V in UV[b][2][O]
__global__ void __launch_bounds__(256) k_glo_fwd_nomask(const float* __restrict__ u, const float* __restrict__ v,
                                                        const float* __restrict__ bi, const float* __restrict__ bj,
                                                        int N, int O, float* __restrict__ UV, float* __restrict__ out) {
  extern __shared__ float sm[];  // U[O] V[O] red[256]
  float* U = sm;
  float* V = sm + O;
  float* red = sm + 2 * O;
  const int b = blockIdx.x;
  for (int o = threadIdx.x; o < O; o += 256) {
    float su = 0.f, sv = 0.f;
    for (int i = 0; i < N; ++i) {
      su += u[((size_t)b * N + i) * O + o] + bi[o];
      sv += v[((size_t)b * N + i) * O + o] + bj[o];
    }
    U[o] = su;
    V[o] = sv;
    UV[((size_t)b * 2) * O + o] = su;
    UV[((size_t)b * 2 + 1) * O + o] = sv;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int o = threadIdx.x; o < O; o += 256) mx = fmaxf(mx, U[o]);
  red[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int i = 0; i < 256; ++i) m = fmaxf(m, red[i]);
    red[0] = m;
  }
  __syncthreads();
  mx = red[0];
  __syncthreads();
  float den = 0.f;
  for (int o = threadIdx.x; o < O; o += 256) den += expf(U[o] - mx);
  red[threadIdx.x] = den;
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 0.f;
    for (int i = 0; i < 256; ++i) d += red[i];
    red[0] = d;
  }
  __syncthreads();
  den = red[0];
  for (int o = threadIdx.x; o < O; o += 256) out[(size_t)b * O + o] = expf(U[o] - mx) / den * V[o];
}

// unmasked backward: dU = S*(dS - sum dS*S), dS = dout*V; dV = dout*S; broadcast to every row of the graph
__global__ void __launch_bounds__(256) k_glo_bwd_nomask(const float* __restrict__ UV, const float* __restrict__ dout,
                                                        int N, int O, float* __restrict__ du, float* __restrict__ dv) {
  extern __shared__ float sm[];  // S[O] red[256]
  float* S = sm;
  float* red = sm + O;
  const int b = blockIdx.x;
  const float* U = UV + ((size_t)b * 2) * O;
  const float* V = U + O;
  float mx = -INFINITY;
  for (int o = threadIdx.x; o < O; o += 256) mx = fmaxf(mx, U[o]);
  red[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int i = 0; i < 256; ++i) m = fmaxf(m, red[i]);
    red[0] = m;
  }
  __syncthreads();
  mx = red[0];
  __syncthreads();
  float den = 0.f;
  for (int o = threadIdx.x; o < O; o += 256) {
    float e = expf(U[o] - mx);
    S[o] = e;
    den += e;
  }
  red[threadIdx.x] = den;
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 0.f;
    for (int i = 0; i < 256; ++i) d += red[i];
    red[0] = d;
  }
  __syncthreads();
  den = red[0];
  __syncthreads();
  float dot = 0.f;
  for (int o = threadIdx.x; o < O; o += 256) {
    S[o] /= den;
    dot += dout[(size_t)b * O + o] * V[o] * S[o];
  }
  red[threadIdx.x] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 0.f;
    for (int i = 0; i < 256; ++i) d += red[i];
    red[0] = d;
  }
  __syncthreads();
  dot = red[0];
  for (int o = threadIdx.x; o < O; o += 256) {
    float d = dout[(size_t)b * O + o];
    float duo = S[o] * (d * V[o] - dot);
    float dvo = d * S[o];
    for (int i = 0; i < N; ++i) {
      du[((size_t)b * N + i) * O + o] = duo;
      dv[((size_t)b * N + i) * O + o] = dvo;
    }
  }
}

__global__ void k_row_scale(float* __restrict__ a, float* __restrict__ b, const float* __restrict__ mask,
                            long long rows, int O) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * O) return;
  float mu = mask[t / O];
  a[t] *= mu;
  b[t] *= mu;
}

}  // namespace

extern "C" {

size_t mpnn_glo_workspace_bytes(int B, int N, int F2, int O) {
  long long rows = (long long)B * N;
  size_t g = mpnn_gemm_workspace_bytes(O, F2, (int)rows);
  size_t c = mpnn_colsum_workspace_bytes(rows, O);
  return 2 * align_up((size_t)rows * O * sizeof(float), 256) + align_up(g > c ? g : c, 256);
}

// u, v: [B*N, O] saved for backward; UV: [B, 2, O] (unmasked form only)
int mpnn_glo_fwd(const float* x, const float* mask, const float* Wi, const float* bi, const float* Wj, const float* bj,
                 int B, int N, int F2, int O, float* out, float* u, float* v, float* UV, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && F2 > 0 && O > 0, MPNN_ERR_ARG, "glo_fwd: bad dims");
  MPNN_REQUIRE(O <= 32 * KMAX, MPNN_ERR_UNSUPPORTED, "glo_fwd: output_dim %d > %d", O, 32 * KMAX);
  int rows = B * N;
  int rc;
  if ((rc = mpnn_gemm(x, Wi, u, rows, O, F2, F2, 1, 1, F2, O, nullptr, 0, nullptr, 0, stream))) return rc;
  if ((rc = mpnn_gemm(x, Wj, v, rows, O, F2, F2, 1, 1, F2, O, nullptr, 0, nullptr, 0, stream))) return rc;
  if (mask) {
    k_glo_fwd_masked<<<B, 256, 8 * O * sizeof(float), stream>>>(u, v, mask, bi, bj, N, O, out);
  } else {
    MPNN_REQUIRE(UV != nullptr, MPNN_ERR_ARG, "glo_fwd: UV buffer required without mask");
    k_glo_fwd_nomask<<<B, 256, (2 * O + 256) * sizeof(float), stream>>>(u, v, bi, bj, N, O, UV, out);
  }
  MPNN_CHECK_LAUNCH("k_glo_fwd");
  return MPNN_OK;
}

int mpnn_glo_bwd(const float* x, const float* mask, const float* Wi, const float* Wj, const float* u, const float* v,
                 const float* UV, const float* dout, int B, int N, int F2, int O, float* dx, float* dWi, float* dbi,
                 float* dWj, float* dbj, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && F2 > 0 && O > 0 && O <= 32 * KMAX, MPNN_ERR_ARG, "glo_bwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_glo_workspace_bytes(B, N, F2, O), MPNN_ERR_WORKSPACE, "glo_bwd: workspace");
  long long rows = (long long)B * N;
  char* wp = (char*)workspace;
  float* du = (float*)wp;
  wp += align_up((size_t)rows * O * sizeof(float), 256);
  float* dv = (float*)wp;
  wp += align_up((size_t)rows * O * sizeof(float), 256);
  void* sub = wp;
  size_t sub_bytes = workspace_bytes - (size_t)(wp - (char*)workspace);
  if (mask) {
    k_glo_bwd_masked<<<ceil_div(rows * 32, 256), 256, 0, stream>>>(u, v, mask, dout, rows, N, O, du, dv);
  } else {
    k_glo_bwd_nomask<<<B, 256, (O + 256) * sizeof(float), stream>>>(UV, dout, N, O, du, dv);
  }
  MPNN_CHECK_LAUNCH("k_glo_bwd");
  int rc;
  if ((rc = mpnn_colsum(du, nullptr, rows, O, O, 0, dbi, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_colsum(dv, nullptr, rows, O, O, 0, dbj, 0, sub, sub_bytes, stream))) return rc;
  if (mask) {  // u' = mu*u_raw + b  ->  d u_raw = mu * du'
    k_row_scale<<<ceil_div(rows * O, 256), 256, 0, stream>>>(du, dv, mask, rows, O);
    MPNN_CHECK_LAUNCH("k_row_scale");
  }
  int R = (int)rows;
  // dx = du Wi + dv Wj  (W [O, F2] row-major)
  if ((rc = mpnn_gemm(du, Wi, dx, R, F2, O, O, 1, F2, 1, F2, nullptr, 0, nullptr, 0, stream))) return rc;
  if ((rc = mpnn_gemm(dv, Wj, dx, R, F2, O, O, 1, F2, 1, F2, nullptr, 2, nullptr, 0, stream))) return rc;
  // dW = du^T x
  if ((rc = mpnn_gemm(du, x, dWi, O, F2, R, 1, O, F2, 1, F2, nullptr, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_gemm(dv, x, dWj, O, F2, R, 1, O, F2, 1, F2, nullptr, 0, sub, sub_bytes, stream))) return rc;
  return MPNN_OK;
}

}  // extern "C"
