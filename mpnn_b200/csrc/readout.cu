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

extern "C" int mpnn_tc_linear_supported(int K, int N);
extern "C" size_t mpnn_tc_linear_workspace_bytes(int K, int N);
extern "C" int mpnn_tc_linear_fwd(const float* X, long long rows, int ldx, int K, const float* W, int N, const float* bias,
                                  float* Y, int ldy, int accumulate, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream);
extern "C" int mpnn_tc_linear_bwd_data(const float* dY, long long rows, int ldd, int N, const float* W, int K, float* dX,
                                       int ldx, int accumulate, void* workspace, size_t workspace_bytes,
                                       cudaStream_t stream);
extern "C" int mpnn_tc_linear_bwd_weight(const float* dY, long long rows, int ldd, int N, const float* X, int ldx, int K,
                                         float* dW, void* workspace, size_t workspace_bytes, cudaStream_t stream);

namespace {

constexpr int KMAX = 32;  // O <= 32*KMAX = 1024

// u' = mu*u + bi, v' = mu*v + bj in place; out[b] = sum_i softmax(u') * v' * mu.  One block per graph.
template <int NK>  // NK = ceil(O / 32) rounded up to a power of two: register arrays sized to the real width
__global__ void __launch_bounds__(256) k_glo_fwd_masked(float* __restrict__ u, float* __restrict__ v,
                                                        const float* __restrict__ mask, const float* __restrict__ bi,
                                                        const float* __restrict__ bj, int N, int O,
                                                        float* __restrict__ out) {
  extern __shared__ float sm[];  // [8][O]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (O + 31) / 32;
  float acc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = 0.f;
  for (int i = warp; i < N; i += 8) {
    const size_t row = (size_t)b * N + i;
    const float mu = mask[row];
    float* ur = u + row * O;
    float* vr = v + row * O;
    float uv[NK];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
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
    for (int k = 0; k < NK; ++k) {
      int o = lane + 32 * k;
      if (k < nk && o < O) {
        uv[k] = expf(uv[k] - mx);
        den += uv[k];
      }
    }
    den = warp_sum(den);
    const float inv = 1.f / den;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      int o = lane + 32 * k;
      if (k < nk && o < O) {
        float vv = mu * vr[o] + bj[o];
        vr[o] = vv;
        acc[k] += uv[k] * inv * vv * mu;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) {
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
template <int NK>
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
  float s[NK];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < O) {
      s[k] = ur[o];
      mx = fmaxf(mx, s[k]);
    }
  }
  mx = warp_max(mx);
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
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
  for (int k = 0; k < NK; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < O) {
      s[k] *= inv;
      dot += dr[o] * mu * vr[o] * s[k];
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int k = 0; k < NK; ++k) {
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


// ---------------------------------------------------------------------------------------------------
// Fused masked readout for O <= 64, F2 <= 64 (the reference's configurations): projections + feature softmax +
// gate + per-graph sum in ONE kernel forward; backward = one kernel (du', dv', dx, per-CTA weight-gradient
// partials) + a fixed-order reduction.  A warp owns a row; lane o (+32) owns output feature o.
// ---------------------------------------------------------------------------------------------------
constexpr int FO = 64;   // max O and F2 of the fused kernels
constexpr int FK = 2;    // 32-lane slices per row

__global__ void __launch_bounds__(256) k_glo_fwd_fused(const float* __restrict__ x, const float* __restrict__ mask,
                                                       const float* __restrict__ Wi, const float* __restrict__ bi,
                                                       const float* __restrict__ Wj, const float* __restrict__ bj,
                                                       int B, int N, int F2, int O, float* __restrict__ out,
                                                       float* __restrict__ u, float* __restrict__ v) {
  __shared__ float WiT[FO * FO];   // WiT[l][o]
  __shared__ float WjT[FO * FO];
  __shared__ float red[8][FO];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    constexpr int NW = FO * FO / 256;
    const int nw = O * F2;
    float a[NW], b[NW];
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = min(tid + t * 256, nw - 1);
      a[t] = __ldg(Wi + i);
      b[t] = __ldg(Wj + i);
    }
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = tid + t * 256;
      if (i < nw) {
        const int o = i / F2, l = i - o * F2;
        WiT[l * FO + o] = a[t];
        WjT[l * FO + o] = b[t];
      }
    }
  }
  float bio[FK], bjo[FK];
#pragma unroll
  for (int k = 0; k < FK; ++k) {
    const int o = min(lane + 32 * k, O - 1);
    bio[k] = bi[o];
    bjo[k] = bj[o];
  }
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float acc[FK] = {0.f, 0.f};
    // a warp's rows are loaded one iteration ahead (N / 8 ~ 4 dependent load latencies per graph otherwise)
    float n_mu = 0.f, n_x[FK];
    auto prefetch = [&](int i) {
      const size_t row = (size_t)b * N + min(i, N - 1);
      n_mu = __ldg(mask + row);
#pragma unroll
      for (int k = 0; k < FK; ++k) n_x[k] = __ldg(x + row * F2 + min(lane + 32 * k, F2 - 1));
    };
    prefetch(warp);
    for (int i = warp; i < N; i += 8) {
      const size_t row = (size_t)b * N + i;
      const float mu = n_mu;
      float xin[FK];
#pragma unroll
      for (int k = 0; k < FK; ++k) xin[k] = n_x[k];
      if (i + 8 < N) prefetch(i + 8);
      if (mu == 0.f) {   // padded atom (about half of a config-4 batch): u' = b_i, v' = b_j, no contribution
#pragma unroll
        for (int k = 0; k < FK; ++k) {
          const int o = lane + 32 * k;
          if (o < O) {
            u[row * O + o] = bio[k];
            v[row * O + o] = bjo[k];
          }
        }
        continue;
      }
      float xl[FK];
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const int l = lane + 32 * k;
        xl[k] = xin[k] * mu;
        if (l >= F2) xl[k] = 0.f;
      }
      float ua[FK] = {0.f, 0.f}, va[FK] = {0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < FK; ++kk) {
        const int lend = min(32, F2 - 32 * kk);
        for (int ll = 0; ll < lend; ++ll) {
          const int l = ll + 32 * kk;
          const float xv = __shfl_sync(0xffffffffu, xl[kk], ll);
#pragma unroll
          for (int k = 0; k < FK; ++k) {
            ua[k] = fmaf(xv, WiT[l * FO + lane + 32 * k], ua[k]);
            va[k] = fmaf(xv, WjT[l * FO + lane + 32 * k], va[k]);
          }
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        ua[k] += bio[k];
        va[k] += bjo[k];
        if (lane + 32 * k < O) mx = fmaxf(mx, ua[k]);
      }
      mx = warp_max(mx);
      float ex[FK], den = 0.f;
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        ex[k] = lane + 32 * k < O ? expf(ua[k] - mx) : 0.f;
        den += ex[k];
      }
      den = warp_sum(den);
      const float inv = 1.f / den;
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const int o = lane + 32 * k;
        if (o < O) {
          u[row * O + o] = ua[k];
          v[row * O + o] = va[k];
          acc[k] += ex[k] * inv * va[k] * mu;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FK; ++k) red[warp][lane + 32 * k] = acc[k];
    __syncthreads();
    if (tid < O) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][tid];
      out[(size_t)b * O + tid] = s;
    }
  }
}

// partial layout per CTA: [dWi O*F2 | dWj O*F2 | dbi O | dbj O]
__global__ void __launch_bounds__(256) k_glo_bwd_fused(const float* __restrict__ x, const float* __restrict__ mask,
                                                       const float* __restrict__ Wi, const float* __restrict__ Wj,
                                                       const float* __restrict__ u, const float* __restrict__ v,
                                                       const float* __restrict__ dout, int B, int N, int F2, int O,
                                                       float* __restrict__ dx, float* __restrict__ partial) {
  constexpr int LD = FO + 1;
  __shared__ float Wis[FO * LD];      // Wis[o][l]
  __shared__ float Wjs[FO * LD];
  __shared__ float du_s[8][FO], dv_s[8][FO], dub_s[8][FO], dvb_s[8][FO], x_s[8][FO];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nw = O * F2;
  {
    constexpr int NW = FO * FO / 256;
    float a[NW], b[NW];
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = min(tid + t * 256, nw - 1);
      a[t] = __ldg(Wi + i);
      b[t] = __ldg(Wj + i);
    }
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = tid + t * 256;
      if (i < nw) {
        const int o = i / F2, l = i - o * F2;
        Wis[o * LD + l] = a[t];
        Wjs[o * LD + l] = b[t];
      }
    }
  }
  // weight-gradient elements of this thread: (which, o = warp + 8 j, l = lane + 32 kf) -- no index divisions
  float acc[2][FK][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int k = 0; k < FK; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][k][j] = 0.f;
  float accb = 0.f;
  const int nkf = (F2 + 31) / 32;
  const long long rows = (long long)B * N;
  const long long ntiles = (rows + 7) / 8;
  // The inputs of a row (mask, u, v, the graph's dout, x) are loaded ONE TILE AHEAD: with 8 warps per SM nothing else
  // hides the ~1 us of dependent global-load latency a tile would otherwise start with (6 tiles per CTA at B = 256).
  float n_mu = 0.f, n_s[FK], n_vv[FK], n_dr[FK], n_x[FK];
  auto prefetch = [&](long long tile) {
    const long long row = tile * 8 + warp;
    const long long rr = row < rows ? row : rows - 1;
    const long long b = rr / N;
    n_mu = __ldg(mask + rr);
#pragma unroll
    for (int k = 0; k < FK; ++k) {
      const int o = min(lane + 32 * k, O - 1);
      n_s[k] = __ldg(u + rr * O + o);
      n_vv[k] = __ldg(v + rr * O + o);
      n_dr[k] = __ldg(dout + b * O + o);
      n_x[k] = __ldg(x + rr * F2 + min(lane + 32 * k, F2 - 1));
    }
  };
  if (blockIdx.x < ntiles) prefetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row = tile * 8 + warp;
    const bool live = row < rows;
    const float mu = n_mu;
    float s[FK], vv[FK], dr[FK], xin[FK];
#pragma unroll
    for (int k = 0; k < FK; ++k) {
      s[k] = n_s[k];
      vv[k] = n_vv[k];
      dr[k] = n_dr[k];
      xin[k] = n_x[k];
    }
    if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x);
    // (barrier: weights staged / previous tile's accumulation finished)  A tile of padded atoms only -- they are the
    // tail of every graph's rows -- contributes nothing: dx = 0 and on to the next tile.
    if (!__syncthreads_or(live && mu != 0.f)) {
      if (live) {
#pragma unroll
        for (int k = 0; k < FK; ++k)
          if (lane + 32 * k < F2) dx[row * F2 + lane + 32 * k] = 0.f;
      }
      continue;
    }
    {
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < FK; ++k)
        if (lane + 32 * k < O) mx = fmaxf(mx, s[k]);
      mx = warp_max(mx);
      float den = 0.f;
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        s[k] = lane + 32 * k < O ? expf(s[k] - mx) : 0.f;
        den += s[k];
      }
      den = warp_sum(den);
      const float inv = 1.f / den;
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        s[k] *= inv;
        dot += dr[k] * mu * vv[k] * s[k];
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const int o = lane + 32 * k;
        const float dg = dr[k] * mu;
        float dvp = dg * s[k];
        float dup = s[k] * (dg * vv[k] - dot);
        if (!live || o >= O) dvp = dup = 0.f;
        dub_s[warp][o] = dup;           // d u' (bias gradient)
        dvb_s[warp][o] = dvp;
        du_s[warp][o] = dup * mu;       // d u_raw (u' = mu * u_raw + b)
        dv_s[warp][o] = dvp * mu;
        const int l = lane + 32 * k;
        float xv = xin[k];
        if (!live || l >= F2) xv = 0.f;
        x_s[warp][l] = xv;
      }
    }
    __syncwarp();
    if (live) {
      // dx[row, l] = sum_o du_raw[o] Wi[o][l] + dv_raw[o] Wj[o][l]
      float dxa[FK] = {0.f, 0.f};
#pragma unroll 4
      for (int o = 0; o < O; ++o) {
        const float a0 = du_s[warp][o], b0 = dv_s[warp][o];
#pragma unroll
        for (int k = 0; k < FK; ++k) {
          if (k < nkf) {
            dxa[k] = fmaf(a0, Wis[o * LD + lane + 32 * k], dxa[k]);
            dxa[k] = fmaf(b0, Wjs[o * LD + lane + 32 * k], dxa[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < FK; ++k)
        if (lane + 32 * k < F2) dx[row * F2 + lane + 32 * k] = dxa[k];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float xv[FK];
#pragma unroll
      for (int k = 0; k < FK; ++k) xv[k] = x_s[r][lane + 32 * k];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gu = du_s[r][warp + 8 * j], gv = dv_s[r][warp + 8 * j];
#pragma unroll
        for (int k = 0; k < FK; ++k) {
          acc[0][k][j] = fmaf(gu, xv[k], acc[0][k][j]);
          acc[1][k][j] = fmaf(gv, xv[k], acc[1][k][j]);
        }
      }
    }
    if (tid < 2 * O) {
      const int which = tid >= O;
      const int o = tid - which * O;
      const float(*gs)[FO] = which ? dvb_s : dub_s;
      float a = accb;
#pragma unroll
      for (int r = 0; r < 8; ++r) a += gs[r][o];
      accb = a;
    }
  }
  float* part = partial + (size_t)blockIdx.x * (2 * nw + 2 * O);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int k = 0; k < FK; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = warp + 8 * j, l = lane + 32 * k;
        if (o < O && l < F2) part[(size_t)a * nw + o * F2 + l] = acc[a][k][j];
      }
  if (tid < 2 * O) part[2 * nw + tid] = accb;
}

__global__ void __launch_bounds__(256) k_glo_bwd_reduce(const float* __restrict__ partial, int nparts, int F2, int O,
                                                        float* __restrict__ dWi, float* __restrict__ dWj,
                                                        float* __restrict__ dbi, float* __restrict__ dbj) {
  __shared__ float sm[8][33];
  const int nw = O * F2;
  const int total = 2 * nw + 2 * O;
  const int el = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + el;
  float s = 0.f;
  if (e < total) {
#pragma unroll 4
    for (int p = sl; p < nparts; p += 8) s += partial[(size_t)p * total + e];
  }
  sm[sl][el] = s;
  __syncthreads();
  if (sl != 0 || e >= total) return;
  s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += sm[j][el];
  if (e < nw)
    dWi[e] = s;
  else if (e < 2 * nw)
    dWj[e - nw] = s;
  else if (e < 2 * nw + O)
    dbi[e - 2 * nw] = s;
  else
    dbj[e - 2 * nw - O] = s;
}

int glo_bwd_grid(long long rows) {
  long long tiles = (rows + 7) / 8;
  int cap = mpnn_num_sms();
  // 80 registers, 44.5 KB shared: three CTAs per SM once there is work for them (below that the 3x larger set of
  // per-CTA partials costs the reduction what the main kernel gains: measured at 928 tiles)
  if (tiles >= 16LL * cap) cap *= 3;
  else if (tiles >= 4LL * cap) cap *= 2;   // two CTAs per SM (122 registers): twice the warps to hide the LDS chains
  return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

}  // namespace

extern "C" {

size_t mpnn_glo_workspace_bytes(int B, int N, int F2, int O) {
  long long rows = (long long)B * N;
  size_t g = mpnn_gemm_workspace_bytes(O, F2, (int)rows);
  size_t c = mpnn_colsum_workspace_bytes(rows, O);
  size_t fusedb = 3 * (size_t)mpnn_num_sms() * (2 * (size_t)O * F2 + 2 * O) * sizeof(float);
  size_t sub = g > c ? g : c;
  const size_t t = mpnn_tc_linear_workspace_bytes(F2, O);   // projections on the tensor cores when the widths allow
  if (t > sub) sub = t;
  size_t need = 2 * align_up((size_t)rows * O * sizeof(float), 256) + align_up(sub, 256);
  return need > fusedb ? need : fusedb;
}

// u, v: [B*N, O] saved for backward; UV: [B, 2, O] (unmasked form only)
int mpnn_glo_fwd(const float* x, const float* mask, const float* Wi, const float* bi, const float* Wj, const float* bj,
                 int B, int N, int F2, int O, float* out, float* u, float* v, float* UV, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && F2 > 0 && O > 0, MPNN_ERR_ARG, "glo_fwd: bad dims");
  MPNN_REQUIRE(O <= 32 * KMAX, MPNN_ERR_UNSUPPORTED, "glo_fwd: output_dim %d > %d", O, 32 * KMAX);
  int rows = B * N;
  int rc;
  if (mask && O <= FO && F2 <= FO) {
    int grid = B < 4 * mpnn_num_sms() ? B : 4 * mpnn_num_sms();
    k_glo_fwd_fused<<<grid, 256, 0, stream>>>(x, mask, Wi, bi, Wj, bj, B, N, F2, O, out, u, v);
    MPNN_CHECK_LAUNCH("k_glo_fwd_fused");
    return MPNN_OK;
  }
  if (mpnn_tc_linear_supported(F2, O) && workspace && workspace_bytes >= mpnn_tc_linear_workspace_bytes(F2, O)) {
    // u = x Wi^T, v = x Wj^T on the tcgen05 dense-GEMM mode (TF32 operands, fp32 accumulate)
    if ((rc = mpnn_tc_linear_fwd(x, rows, F2, F2, Wi, O, nullptr, u, O, 0, workspace, workspace_bytes, stream))) return rc;
    if ((rc = mpnn_tc_linear_fwd(x, rows, F2, F2, Wj, O, nullptr, v, O, 0, workspace, workspace_bytes, stream))) return rc;
  } else {
    if ((rc = mpnn_gemm(x, Wi, u, rows, O, F2, F2, 1, 1, F2, O, nullptr, 0, nullptr, 0, stream))) return rc;
    if ((rc = mpnn_gemm(x, Wj, v, rows, O, F2, F2, 1, 1, F2, O, nullptr, 0, nullptr, 0, stream))) return rc;
  }
  if (mask) {
    const int nk = (O + 31) / 32;
    const size_t sm = 8 * (size_t)O * sizeof(float);
    if (nk <= 1) k_glo_fwd_masked<1><<<B, 256, sm, stream>>>(u, v, mask, bi, bj, N, O, out);
    else if (nk <= 2) k_glo_fwd_masked<2><<<B, 256, sm, stream>>>(u, v, mask, bi, bj, N, O, out);
    else if (nk <= 4) k_glo_fwd_masked<4><<<B, 256, sm, stream>>>(u, v, mask, bi, bj, N, O, out);
    else if (nk <= 8) k_glo_fwd_masked<8><<<B, 256, sm, stream>>>(u, v, mask, bi, bj, N, O, out);
    else if (nk <= 16) k_glo_fwd_masked<16><<<B, 256, sm, stream>>>(u, v, mask, bi, bj, N, O, out);
    else k_glo_fwd_masked<KMAX><<<B, 256, sm, stream>>>(u, v, mask, bi, bj, N, O, out);
  } else {
    MPNN_REQUIRE(UV != nullptr, MPNN_ERR_ARG, "glo_fwd: UV buffer required without mask");
    k_glo_fwd_nomask<<<B, 256, (2 * O + 256) * sizeof(float), stream>>>(u, v, bi, bj, N, O, UV, out);
  }
  MPNN_CHECK_LAUNCH("k_glo_fwd");
  return MPNN_OK;
}

// The fused masked backward (O <= 64, F2 <= 64) in two calls, so that the caller can put the parameter half on another
// stream: `mpnn_glo_bwd_data` writes dx and the per-CTA partials into the workspace, `mpnn_glo_bwd_params` reduces them in
// fixed order into dWi / dbi / dWj / dbj (only parameter gradients: nothing downstream of the readout waits for it).
int mpnn_glo_bwd_split_supported(int has_mask, int F2, int O) { return (has_mask && O <= FO && F2 <= FO) ? 1 : 0; }

int mpnn_glo_bwd_data(const float* x, const float* mask, const float* Wi, const float* Wj, const float* u, const float* v,
                      const float* dout, int B, int N, int F2, int O, float* dx, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && mpnn_glo_bwd_split_supported(mask != nullptr, F2, O), MPNN_ERR_UNSUPPORTED,
               "glo_bwd_data: shape not served by the fused kernel");
  MPNN_REQUIRE(workspace_bytes >= mpnn_glo_workspace_bytes(B, N, F2, O), MPNN_ERR_WORKSPACE, "glo_bwd_data: workspace");
  const long long rows = (long long)B * N;
  const int grid = glo_bwd_grid(rows);
  k_glo_bwd_fused<<<grid, 256, 0, stream>>>(x, mask, Wi, Wj, u, v, dout, B, N, F2, O, dx, (float*)workspace);
  MPNN_CHECK_LAUNCH("k_glo_bwd_fused");
  return MPNN_OK;
}

int mpnn_glo_bwd_params(const void* workspace, int B, int N, int F2, int O, float* dWi, float* dbi, float* dWj, float* dbj,
                        cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && mpnn_glo_bwd_split_supported(1, F2, O), MPNN_ERR_UNSUPPORTED, "glo_bwd_params: shape");
  const int grid = glo_bwd_grid((long long)B * N);
  k_glo_bwd_reduce<<<ceil_div(2 * O * F2 + 2 * O, 32), 256, 0, stream>>>((const float*)workspace, grid, F2, O, dWi, dWj, dbi,
                                                                        dbj);
  MPNN_CHECK_LAUNCH("k_glo_bwd_reduce");
  return MPNN_OK;
}

int mpnn_glo_bwd(const float* x, const float* mask, const float* Wi, const float* Wj, const float* u, const float* v,
                 const float* UV, const float* dout, int B, int N, int F2, int O, float* dx, float* dWi, float* dbi,
                 float* dWj, float* dbj, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && F2 > 0 && O > 0 && O <= 32 * KMAX, MPNN_ERR_ARG, "glo_bwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_glo_workspace_bytes(B, N, F2, O), MPNN_ERR_WORKSPACE, "glo_bwd: workspace");
  long long rows = (long long)B * N;
  if (mask && O <= FO && F2 <= FO) {
    const int grid = glo_bwd_grid(rows);
    float* partial = (float*)workspace;
    k_glo_bwd_fused<<<grid, 256, 0, stream>>>(x, mask, Wi, Wj, u, v, dout, B, N, F2, O, dx, partial);
    MPNN_CHECK_LAUNCH("k_glo_bwd_fused");
    k_glo_bwd_reduce<<<ceil_div(2 * O * F2 + 2 * O, 32), 256, 0, stream>>>(partial, grid, F2, O, dWi, dWj, dbi, dbj);
    MPNN_CHECK_LAUNCH("k_glo_bwd_reduce");
    return MPNN_OK;
  }
  char* wp = (char*)workspace;
  float* du = (float*)wp;
  wp += align_up((size_t)rows * O * sizeof(float), 256);
  float* dv = (float*)wp;
  wp += align_up((size_t)rows * O * sizeof(float), 256);
  void* sub = wp;
  size_t sub_bytes = workspace_bytes - (size_t)(wp - (char*)workspace);
  if (mask) {
    const int nk = (O + 31) / 32;
    const int gb = ceil_div(rows * 32, 256);
    if (nk <= 1) k_glo_bwd_masked<1><<<gb, 256, 0, stream>>>(u, v, mask, dout, rows, N, O, du, dv);
    else if (nk <= 2) k_glo_bwd_masked<2><<<gb, 256, 0, stream>>>(u, v, mask, dout, rows, N, O, du, dv);
    else if (nk <= 4) k_glo_bwd_masked<4><<<gb, 256, 0, stream>>>(u, v, mask, dout, rows, N, O, du, dv);
    else if (nk <= 8) k_glo_bwd_masked<8><<<gb, 256, 0, stream>>>(u, v, mask, dout, rows, N, O, du, dv);
    else if (nk <= 16) k_glo_bwd_masked<16><<<gb, 256, 0, stream>>>(u, v, mask, dout, rows, N, O, du, dv);
    else k_glo_bwd_masked<KMAX><<<gb, 256, 0, stream>>>(u, v, mask, dout, rows, N, O, du, dv);
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
  if (mpnn_tc_linear_supported(F2, O)) {
    if ((rc = mpnn_tc_linear_bwd_data(du, rows, O, O, Wi, F2, dx, F2, 0, sub, sub_bytes, stream))) return rc;
    if ((rc = mpnn_tc_linear_bwd_data(dv, rows, O, O, Wj, F2, dx, F2, 1, sub, sub_bytes, stream))) return rc;
    if ((rc = mpnn_tc_linear_bwd_weight(du, rows, O, O, x, F2, F2, dWi, sub, sub_bytes, stream))) return rc;
    if ((rc = mpnn_tc_linear_bwd_weight(dv, rows, O, O, x, F2, F2, dWj, sub, sub_bytes, stream))) return rc;
    return MPNN_OK;
  }
  // dx = du Wi + dv Wj  (W [O, F2] row-major)
  if ((rc = mpnn_gemm(du, Wi, dx, R, F2, O, O, 1, F2, 1, F2, nullptr, 0, nullptr, 0, stream))) return rc;
  if ((rc = mpnn_gemm(dv, Wj, dx, R, F2, O, O, 1, F2, 1, F2, nullptr, 2, nullptr, 0, stream))) return rc;
  // dW = du^T x
  if ((rc = mpnn_gemm(du, x, dWi, O, F2, R, 1, O, F2, 1, F2, nullptr, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_gemm(dv, x, dWj, O, F2, R, 1, O, F2, 1, F2, nullptr, 0, sub, sub_bytes, stream))) return rc;
  return MPNN_OK;
}

}  // extern "C"
