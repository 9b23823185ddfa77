// Masked GRU node update (reference mpnn_functions/update/gru_update.py:26-35, 66-68).
//   [ri|zi|ni] = m W_ih + b_ih ; [rh|zh|nh] = h W_hh + b_hh        (weights stored [d, 3d], gate order r,z,n)
//   r = sigmoid(ri+rh)*mu ; z = sigmoid(zi+zh)*mu ; n = tanh(ni + r*nh)*mu ; h' = ((1-z)*n + z*h)*mu
// Saved for backward: gates[rows, 4d] = (sigmoid_r, sigmoid_z, tanh_n, nh) -- un-masked activations, so any
// float mask value is differentiated exactly like the reference's autograd graph.
#include "common.cuh"

extern "C" int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
                         long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_gemm_workspace_bytes(int M, int N, int K);
extern "C" int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy,
                           float* out, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_colsum_workspace_bytes(long long rows, int width);

extern "C" int mpnn_tc_dp(int nf, int mf);
extern "C" int mpnn_tc_gru_supported(int d);
extern "C" size_t mpnn_tc_gru_workspace_bytes(int d);
extern "C" int mpnn_tc_gru_fwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                               const float* b_ih, const float* b_hh, long long rows, int d, float* h_out, float* gates,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_tc_dense_workspace_bytes(int n_blocks, int DP);
extern "C" int mpnn_tc_dense_gemm(const float* A, long long rows, int lda, int K, int kseg, int acol, const float* W,
                                  long long w_sn, long long w_sk, long long w_sg, long long w_ss, int G, int N,
                                  const float* bias, float* Y, int ldy, int ycol, int accumulate, int DP,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_tc_dense_grad_workspace_bytes(int G, int DP);
extern "C" int mpnn_tc_dense_gemm_tn(const float* X, long long rows, int ldx, int M, const float* D, int ldd, int dcol,
                                     int G, int N, int DP, float* out, long long o_sg, long long o_sl, void* workspace,
                                     size_t workspace_bytes, cudaStream_t stream);

extern "C" int mpnn_tc_dense_gemm_ll(const float* A, long long rows, int lda, int K, int kseg, int acol, const float* W,
                                     long long w_sn, long long w_sk, long long w_sg, long long w_ss, int G, int N,
                                     const float* bias, float* Y, int ldy, long long ycol, int nsplit, int accumulate,
                                     int DP, void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" int mpnn_tc_gru_fwd_agg(const float* Y, const int* row_ptr, const float* m, const float* h, const float* mask,
                                   const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh,
                                   long long rows, int d, float* m_out, float* h_out, float* gates, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_tc_gru_param_workspace_bytes(void);
extern "C" int mpnn_tc_gru_param_bias_parts(void);
extern "C" size_t mpnn_tc_gru_data_workspace_bytes(void);
extern "C" int mpnn_tc_gru_data_grad(const float* gates, const float* h, const float* dh_out, const float* mask,
                                     const float* Wc, long long rows, int d, float* dm, float* dh, void* workspace,
                                     size_t workspace_bytes, cudaStream_t stream);
extern "C" int mpnn_tc_gru_param_point(const float* m, const float* h, const float* mask, const float* gates,
                                       const float* dh_out, long long rows, int d, float* dg, float* bias_part,
                                       float* dW_ih, float* dW_hh, void* workspace, size_t workspace_bytes,
                                       cudaStream_t stream);
extern "C" int mpnn_tc_gru_param_grad(const float* m, const float* h, const float* dg, int ldg, long long rows, int d,
                                      float* dW_ih, float* dW_hh, void* workspace, size_t workspace_bytes,
                                      cudaStream_t stream);

namespace {
int g_point_in_param = 1;   // mpnn_gru_bwd_one_pass: pointwise pass inside the weight-gradient kernel (widths <= 64)


__global__ void k_gru_point_fwd(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h,
                                const float* __restrict__ mask, long long rows, int d, float* __restrict__ hout,
                                float* __restrict__ gates) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * d) return;
  long long row = t / d;
  int c = (int)(t - row * d);
  const float mu = mask[row];
  const float* gir = gi + row * 3 * d;
  const float* ghr = gh + row * 3 * d;
  float sr = 1.f / (1.f + expf(-(gir[c] + ghr[c])));
  float sz = 1.f / (1.f + expf(-(gir[d + c] + ghr[d + c])));
  float nh = ghr[2 * d + c];
  float r = sr * mu, z = sz * mu;
  float tn = tanhf(gir[2 * d + c] + r * nh);
  float n = tn * mu;
  float hv = h[t];
  hout[t] = ((1.f - z) * n + z * hv) * mu;
  float* g = gates + row * 4 * d;
  g[c] = sr;
  g[d + c] = sz;
  g[2 * d + c] = tn;
  g[3 * d + c] = nh;
}

__global__ void k_gru_point_bwd(const float* __restrict__ gates, const float* __restrict__ h,
                                const float* __restrict__ mask, const float* __restrict__ dhout, long long rows, int d,
                                float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * d) return;
  long long row = t / d;
  int c = (int)(t - row * d);
  const float mu = mask[row];
  const float* g = gates + row * 4 * d;
  float sr = g[c], sz = g[d + c], tn = g[2 * d + c], nh = g[3 * d + c];
  float r = sr * mu, z = sz * mu, n = tn * mu;
  float go = dhout[t] * mu;
  float dn = go * (1.f - z);
  float dz = go * (h[t] - n);
  float dan = dn * mu * (1.f - tn * tn);
  float dr = dan * nh;
  float dnh = dan * r;
  float dar = dr * mu * sr * (1.f - sr);
  float daz = dz * mu * sz * (1.f - sz);
  float* a = dgi + row * 3 * d;
  float* b = dgh + row * 3 * d;
  a[c] = dar;
  a[d + c] = daz;
  a[2 * d + c] = dan;
  b[c] = dar;
  b[d + c] = daz;
  b[2 * d + c] = dnh;
  dh[t] = go * z;
}


// Gate gradients of the wide path (d > 32), ONE array dg [rows, 6d] = dar | daz | dan | dnh | hi(go z) | lo(go z).  The
// last two blocks are the direct term of dh, which the data product adds through identity blocks instead of a
// read-modify-write epilogue; it is split into a part that is exact in TF32 (13 low mantissa bits cleared) and the
// remainder, so that the tensor core's operand truncation costs the dominant term of dh 2^-22 instead of 2^-11,
// and their column sums (the bias gradients) as per-CTA partials: a thread owns 4 columns and walks the rows, so the
// sums stay in registers; no second pass over the gate gradients.
__global__ void __launch_bounds__(256) k_gru_point_bwd5(const float* __restrict__ gates, const float* __restrict__ h,
                                                        const float* __restrict__ mask,
                                                        const float* __restrict__ dhout, long long rows, int d,
                                                        float* __restrict__ dg, float* __restrict__ bias_part) {
  __shared__ float4 red[256];
  const int cg = d >> 2;                      // column groups per row
  const int rpb = 256 / cg;                   // rows per block and iteration
  const int ty = threadIdx.x / cg, c = (threadIdx.x - ty * cg) * 4;
  const bool active = ty < rpb;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
  if (active) {
    for (long long row = (long long)blockIdx.x * rpb + ty; row < rows; row += (long long)gridDim.x * rpb) {
      const float mu = __ldg(mask + row);
      const float* g = gates + row * 4 * d + c;
      const float4 sr = __ldg(reinterpret_cast<const float4*>(g));
      const float4 sz = __ldg(reinterpret_cast<const float4*>(g + d));
      const float4 tn = __ldg(reinterpret_cast<const float4*>(g + 2 * d));
      const float4 nh = __ldg(reinterpret_cast<const float4*>(g + 3 * d));
      const float4 hv = __ldg(reinterpret_cast<const float4*>(h + row * d + c));
      const float4 dv = __ldg(reinterpret_cast<const float4*>(dhout + row * d + c));
      float4 o0, o1, o2, o3, o4, o5;
#define MPNN_GRU_POINT(X)                                  \
  {                                                        \
    const float r_ = sr.X * mu, z_ = sz.X * mu, n_ = tn.X * mu; \
    const float go = dv.X * mu;                            \
    const float dn = go * (1.f - z_);                      \
    const float dz = go * (hv.X - n_);                     \
    const float dan = dn * mu * (1.f - tn.X * tn.X);       \
    const float dr = dan * nh.X;                           \
    o0.X = dr * mu * sr.X * (1.f - sr.X);                  \
    o1.X = dz * mu * sz.X * (1.f - sz.X);                  \
    o2.X = dan;                                            \
    o3.X = dan * r_;                                       \
    const float gz = go * z_;                              \
    o4.X = __uint_as_float(__float_as_uint(gz) & 0xffffe000u); \
    o5.X = gz - o4.X;                                      \
  }
      MPNN_GRU_POINT(x) MPNN_GRU_POINT(y) MPNN_GRU_POINT(z) MPNN_GRU_POINT(w)
#undef MPNN_GRU_POINT
      float* o = dg + row * 6 * d + c;
      *reinterpret_cast<float4*>(o) = o0;
      *reinterpret_cast<float4*>(o + d) = o1;
      *reinterpret_cast<float4*>(o + 2 * d) = o2;
      *reinterpret_cast<float4*>(o + 3 * d) = o3;
      *reinterpret_cast<float4*>(o + 4 * d) = o4;
      *reinterpret_cast<float4*>(o + 5 * d) = o5;
      s0.x += o0.x; s0.y += o0.y; s0.z += o0.z; s0.w += o0.w;
      s1.x += o1.x; s1.y += o1.y; s1.z += o1.z; s1.w += o1.w;
      s2.x += o2.x; s2.y += o2.y; s2.z += o2.z; s2.w += o2.w;
      s3.x += o3.x; s3.y += o3.y; s3.z += o3.z; s3.w += o3.w;
    }
  }
  // fixed-order sum over the block's row slots, one gate block at a time
  float* out = bias_part + (size_t)blockIdx.x * 4 * d;
#pragma unroll 1
  for (int b = 0; b < 4; ++b) {
    red[threadIdx.x] = b == 0 ? s0 : b == 1 ? s1 : b == 2 ? s2 : s3;
    __syncthreads();
    if (threadIdx.x < cg) {
      float4 t = red[threadIdx.x];
      for (int y = 1; y < rpb; ++y) {
        const float4 v = red[y * cg + threadIdx.x];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      *reinterpret_cast<float4*>(out + b * d + threadIdx.x * 4) = t;
    }
    __syncthreads();
  }
}

// db_ih = (sum dar | sum daz | sum dan), db_hh = (sum dar | sum daz | sum dnh) over the per-CTA partials, fixed order:
// a block is 32 columns x 8 slices of the partials (slice sums in registers, then a fixed-order sum of the 8 slices)
__global__ void __launch_bounds__(256) k_gru_bias_final(const float* __restrict__ part, int n_part, int d,
                                                        float* __restrict__ db_ih, float* __restrict__ db_hh) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (j < 4 * d) {
    int i = sl;
    for (; i + 24 < n_part; i += 32) {
      const float v0 = part[(size_t)i * 4 * d + j], v1 = part[(size_t)(i + 8) * 4 * d + j],
                  v2 = part[(size_t)(i + 16) * 4 * d + j], v3 = part[(size_t)(i + 24) * 4 * d + j];
      s += v0;
      s += v1;
      s += v2;
      s += v3;
    }
    for (; i < n_part; i += 8) s += part[(size_t)i * 4 * d + j];
  }
  red[sl][cx] = s;
  __syncthreads();
  if (sl == 0 && j < 4 * d) {
    float t = red[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][cx];
    const int b = j / d, c = j - b * d;
    if (b < 2) {
      db_ih[j] = t;
      db_hh[j] = t;
    } else if (b == 2) {
      db_ih[j] = t;
    } else {
      db_hh[2 * d + c] = t;
    }
  }
}

// Wc[s][g*d + n][k] (g: 0 = dm, 1 = dh; s: the six blocks of dg): the B matrix of the data product, N = 2d
//   dm = dar Wr_ih^T + daz Wz_ih^T + dan Wn_ih^T,   dh = dar Wr_hh^T + daz Wz_hh^T + dnh Wn_hh^T + (hi + lo) I
__global__ void __launch_bounds__(256) k_gru_bwd_wcomb(const float* __restrict__ W_ih, const float* __restrict__ W_hh,
                                                       int d, float* __restrict__ Wc) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 12 * d * d) return;
  const int k = i % d, nn = (i / d) % (2 * d), s = i / (2 * d * d);   // [s][n of (dm | dh)][k]
  const int g = nn / d, n = nn - g * d;
  float v = 0.f;
  if (g == 0) {
    if (s < 3) v = W_ih[(size_t)n * 3 * d + s * d + k];
  } else {
    if (s < 2) v = W_hh[(size_t)n * 3 * d + s * d + k];
    else if (s == 3) v = W_hh[(size_t)n * 3 * d + 2 * d + k];
    else if (s >= 4) v = n == k ? 1.f : 0.f;
  }
  Wc[i] = v;
}

// ---------------------------------------------------------------------------------------------------
// Fused GRU for feature widths <= 32: one kernel forward, one kernel (+ a fixed-order reduction of the
// per-CTA weight-gradient partials) backward.  A group of DP lanes owns a row; lane c owns column c of
// every gate.  Both gate weight matrices live in shared memory for the whole kernel; nothing but the
// saved gates [rows, 4d] is written besides the outputs (no [rows, 3d] pre-activation round trip).
// ---------------------------------------------------------------------------------------------------
template <int DP>
__device__ __forceinline__ uint32_t grp_mask(int lane) {
  if constexpr (DP == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << DP) - 1u) << ((lane / DP) * DP);
  }
}

template <int DP>
__global__ void __launch_bounds__(256) k_gru_fwd_fused(const float* __restrict__ m, const float* __restrict__ h,
                                                       const float* __restrict__ mask, const float* __restrict__ W_ih,
                                                       const float* __restrict__ W_hh, const float* __restrict__ b_ih,
                                                       const float* __restrict__ b_hh, long long rows, int d,
                                                       float* __restrict__ hout, float* __restrict__ gates) {
  extern __shared__ __align__(16) float sm[];
  float* Wi = sm;               // [d][3d]
  float* Wh = sm + d * 3 * d;   // [d][3d]
  {
    constexpr int NW = (3 * DP * DP + 255) / 256;
    const int nw = d * 3 * d;
    float wi[NW], wh[NW];
#pragma unroll
    for (int t = 0; t < NW; ++t) {  // unconditional clamped loads: all in flight together
      const int i = min((int)threadIdx.x + t * 256, nw - 1);
      wi[t] = __ldg(W_ih + i);
      wh[t] = __ldg(W_hh + i);
    }
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = threadIdx.x + t * 256;
      if (i < nw) {
        Wi[i] = wi[t];
        Wh[i] = wh[t];
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int c = lane % DP;
  const uint32_t gm = grp_mask<DP>(lane);
  const int gpb = 256 / DP;
  const bool on = c < d;
  const int cc = on ? c : 0;
  const float bir = b_ih[cc], biz = b_ih[d + cc], bin = b_ih[2 * d + cc];
  const float bhr = b_hh[cc], bhz = b_hh[d + cc], bhn = b_hh[2 * d + cc];
  for (long long row = (long long)blockIdx.x * gpb + threadIdx.x / DP; row < rows; row += (long long)gridDim.x * gpb) {
    float mv = __ldg(m + row * d + cc);
    float hv = __ldg(h + row * d + cc);
    const float mu = __ldg(mask + row);
    if (!on) mv = hv = 0.f;
    float ir = bir, iz = biz, in_ = bin, hr = bhr, hz = bhz, hn = bhn;
#pragma unroll 4
    for (int l = 0; l < d; ++l) {
      const float ml = __shfl_sync(gm, mv, l, DP);
      const float hl = __shfl_sync(gm, hv, l, DP);
      const float* wi = Wi + l * 3 * d + cc;
      const float* wh = Wh + l * 3 * d + cc;
      ir = fmaf(ml, wi[0], ir);
      iz = fmaf(ml, wi[d], iz);
      in_ = fmaf(ml, wi[2 * d], in_);
      hr = fmaf(hl, wh[0], hr);
      hz = fmaf(hl, wh[d], hz);
      hn = fmaf(hl, wh[2 * d], hn);
    }
    if (on) {
      const float sr = 1.f / (1.f + expf(-(ir + hr)));
      const float sz = 1.f / (1.f + expf(-(iz + hz)));
      const float r = sr * mu, z = sz * mu;
      const float tn = tanhf(in_ + r * hn);
      const float n = tn * mu;
      hout[row * d + c] = ((1.f - z) * n + z * hv) * mu;
      float* g = gates + row * 4 * d;
      g[c] = sr;
      g[d + c] = sz;
      g[2 * d + c] = tn;
      g[3 * d + c] = hn;
    }
  }
}

// partial layout per CTA: [dW_ih d*3d | dW_hh d*3d | db_ih 3d | db_hh 3d]
template <int DP>
__global__ void __launch_bounds__(256) k_gru_bwd_fused(const float* __restrict__ m, const float* __restrict__ h,
                                                       const float* __restrict__ mask, const float* __restrict__ W_ih,
                                                       const float* __restrict__ W_hh, const float* __restrict__ gates,
                                                       const float* __restrict__ dhout, long long rows, int d,
                                                       float* __restrict__ dm, float* __restrict__ dh,
                                                       float* __restrict__ partial) {
  constexpr int TR = 256 / DP;                      // rows per tile (one per group)
  constexpr int NACC = (6 * DP * DP + 255) / 256;   // weight-gradient elements per thread
  extern __shared__ __align__(16) float sm[];
  const int d3 = 3 * d;
  const int ldt = d + 1;
  float* WiT = sm;                    // [3d][d+1]  WiT[g][l] = W_ih[l][g]
  float* WhT = WiT + d3 * ldt;        // [3d][d+1]
  float* Ms = WhT + d3 * ldt;         // [TR][d]
  float* Hs = Ms + TR * d;            // [TR][d]
  float* Gi = Hs + TR * d;            // [TR][3d]
  float* Gh = Gi + TR * d3;           // [TR][3d]
  {
    constexpr int NW = (3 * DP * DP + 255) / 256;
    const int nw = d * d3;
    float wi[NW], wh[NW];
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = min((int)threadIdx.x + t * 256, nw - 1);
      wi[t] = __ldg(W_ih + i);
      wh[t] = __ldg(W_hh + i);
    }
#pragma unroll
    for (int t = 0; t < NW; ++t) {
      const int i = threadIdx.x + t * 256;
      if (i < nw) {
        const int l = i / d3, g = i - l * d3;
        WiT[g * ldt + l] = wi[t];
        WhT[g * ldt + l] = wh[t];
      }
    }
  }
  const int lane = threadIdx.x & 31;
  const int c = lane % DP;
  const int grp = threadIdx.x / DP;
  const bool on = c < d;
  const int nW = d * d3;
  float acc[NACC];
  int pk[NACC];      // per element: offset into [Ms|Hs] (low 16 bits) and into [Gi|Gh] (high bits); -1 = none
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    acc[q] = 0.f;
    const int e = threadIdx.x + q * 256;
    pk[q] = -1;
    if (e < 2 * nW) {
      const int which = e >= nW;
      const int ee = e - which * nW;
      const int l = ee / d3, g = ee - l * d3;
      pk[q] = (which * TR * d + l) | ((which * TR * d3 + g) << 16);
    }
  }
  float accb = 0.f;  // threads [0, 6d): bias gradients
  const long long ntiles = (rows + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row = tile * TR + grp;
    const bool live = row < rows;
    __syncthreads();  // previous tile's accumulation is done with the staging buffers (and W*T are loaded)
    float dar = 0.f, daz = 0.f, dan = 0.f, dnh = 0.f, dhd = 0.f, mv = 0.f, hv = 0.f;
    {
      // unconditional clamped loads (8 requests in flight), masked afterwards
      const long long rr = live ? row : rows - 1;
      const int cq = on ? c : 0;
      const float* g = gates + rr * 4 * d;
      const float mu = __ldg(mask + rr);
      const float sr = __ldg(g + cq), sz = __ldg(g + d + cq), tn = __ldg(g + 2 * d + cq), nh = __ldg(g + 3 * d + cq);
      const float hl = __ldg(h + rr * d + cq), ml = __ldg(m + rr * d + cq), dl = __ldg(dhout + rr * d + cq);
      if (live && on) {
      const float r = sr * mu, z = sz * mu, n = tn * mu;
      hv = hl;
      mv = ml;
      const float go = dl * mu;
      const float dn = go * (1.f - z);
      const float dz = go * (hv - n);
      dan = dn * mu * (1.f - tn * tn);
      const float dr = dan * nh;
      dnh = dan * r;
      dar = dr * mu * sr * (1.f - sr);
      daz = dz * mu * sz * (1.f - sz);
      dhd = go * z;
      }
    }
    if (on) {
      Ms[grp * d + c] = mv;
      Hs[grp * d + c] = hv;
      Gi[grp * d3 + c] = dar;
      Gi[grp * d3 + d + c] = daz;
      Gi[grp * d3 + 2 * d + c] = dan;
      Gh[grp * d3 + c] = dar;
      Gh[grp * d3 + d + c] = daz;
      Gh[grp * d3 + 2 * d + c] = dnh;
    }
    __syncthreads();
    if (live && on) {
      float am = 0.f, ah = dhd;
      const float* gi = Gi + grp * d3;
      const float* gh = Gh + grp * d3;
#pragma unroll 4
      for (int g = 0; g < d3; ++g) {
        am = fmaf(gi[g], WiT[g * ldt + c], am);
        ah = fmaf(gh[g], WhT[g * ldt + c], ah);
      }
      dm[row * d + c] = am;
      dh[row * d + c] = ah;
    }
    // weight gradients: element q -> (which, l, g); rows of the tile that are out of range staged zeros
#pragma unroll
    for (int q = 0; q < NACC; ++q) {
      if (pk[q] >= 0) {
        const float* xs = Ms + (pk[q] & 0xffff);
        const float* gs = Gi + (pk[q] >> 16);
        float a = acc[q];
#pragma unroll 8
        for (int r = 0; r < TR; ++r) a = fmaf(xs[r * d], gs[r * d3], a);
        acc[q] = a;
      }
    }
    if (threadIdx.x < 2 * d3) {
      const int which = threadIdx.x >= d3;
      const int g = threadIdx.x - which * d3;
      const float* gs = which ? Gh : Gi;
      float a = accb;
      for (int r = 0; r < TR; ++r) a += gs[r * d3 + g];
      accb = a;
    }
  }
  float* part = partial + (size_t)blockIdx.x * (2 * nW + 2 * d3);
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    const int e = threadIdx.x + q * 256;
    if (e < 2 * nW) part[e] = acc[q];
  }
  if (threadIdx.x < 2 * d3) part[2 * nW + threadIdx.x] = accb;
}

// 32 elements x 8 partial slices per block; slices are combined in a fixed order
__global__ void __launch_bounds__(256) k_gru_bwd_reduce(const float* __restrict__ partial, int nparts, int d,
                                                        float* __restrict__ dW_ih, float* __restrict__ dW_hh,
                                                        float* __restrict__ db_ih, float* __restrict__ db_hh) {
  __shared__ float sm[8][33];
  const int nW = d * 3 * d, d3 = 3 * d;
  const int total = 2 * nW + 2 * d3;
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
  if (e < nW)
    dW_ih[e] = s;
  else if (e < 2 * nW)
    dW_hh[e - nW] = s;
  else if (e < 2 * nW + d3)
    db_ih[e - 2 * nW] = s;
  else
    db_hh[e - 2 * nW - d3] = s;
}

int fused_dp(int d) { return d <= 8 ? 8 : (d <= 16 ? 16 : 32); }

int gru_bwd_grid(long long rows, int DP) {
  long long tiles = (rows + (256 / DP) - 1) / (256 / DP);
  int cap = mpnn_num_sms();
  // 80 registers x 256 threads: three CTAs fit an SM at DP <= 16.  Small batches keep one CTA per SM (fewer partials
  // for the fixed-order reduction, the step is latency-bound anyway); 10^5-row batches fill the SMs.
  // (ncu at 7 424 rows, one CTA per SM: 12.6 % occupancy, issue slots 35 % busy, 17.4 us)
  if (DP <= 16 && tiles >= 2LL * cap) cap *= 3;
  return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

size_t gru_bwd_smem(int d, int DP) {
  int TR = 256 / DP;
  return (size_t)(2 * 3 * d * (d + 1) + 2 * TR * d + 2 * TR * 3 * d) * sizeof(float);
}

}  // namespace

extern "C" {

size_t mpnn_gru_workspace_bytes(long long rows, int d) {
  size_t pre = 2 * align_up((size_t)rows * 3 * d * sizeof(float), 256);
  size_t g = mpnn_gemm_workspace_bytes(d, 3 * d, (int)rows);
  size_t c = mpnn_colsum_workspace_bytes(rows, 3 * d);
  size_t fused = 3 * (size_t)mpnn_num_sms() * (6 * (size_t)d * d + 6 * d) * sizeof(float);
  size_t sub = g > c ? g : c;
  const int DP = d > 32 ? mpnn_tc_dp(d, d) : -1;   // widths 33..256: gate GEMMs on the tensor cores (tc_message.cu)
  if (DP > 0) {
    size_t t = align_up(mpnn_tc_dense_workspace_bytes(24, DP), 256) + mpnn_tc_dense_grad_workspace_bytes(3, DP);
    if (t > sub) sub = t;
    if (mpnn_tc_gru_param_workspace_bytes() > sub) sub = mpnn_tc_gru_param_workspace_bytes();
    if (mpnn_tc_gru_data_workspace_bytes() > sub) sub = mpnn_tc_gru_data_workspace_bytes();
    // dg [rows, 6d] (= the two [rows, 3d] arrays of the fp32 path) + bias partials + the combined weights
    pre += align_up((size_t)16 * mpnn_num_sms() * 4 * d * sizeof(float), 256) + align_up((size_t)12 * d * d * sizeof(float), 256);
  }
  size_t need = pre + align_up(sub, 256);
  const size_t tg = mpnn_tc_gru_workspace_bytes(d);
  if (tg > need) need = tg;
  return need > fused ? need : fused;
}

int mpnn_gru_fwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                 const float* b_ih, const float* b_hh, long long rows, int d, float* h_out, float* gates,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && d > 0 && rows < (1ll << 31), MPNN_ERR_ARG, "gru_fwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_gru_workspace_bytes(rows, d), MPNN_ERR_WORKSPACE, "gru_fwd: workspace");
  if (d <= 32) {
    const int DP = fused_dp(d);
    const int gpb = 256 / DP;
    long long want = (rows + gpb - 1) / gpb;
    int grid = (int)(want < 4LL * mpnn_num_sms() ? want : 4LL * mpnn_num_sms());
    size_t smem = (size_t)2 * d * 3 * d * sizeof(float);
    switch (DP) {
      case 8: k_gru_fwd_fused<8><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, b_ih, b_hh, rows, d, h_out, gates); break;
      case 16: k_gru_fwd_fused<16><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, b_ih, b_hh, rows, d, h_out, gates); break;
      default: k_gru_fwd_fused<32><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, b_ih, b_hh, rows, d, h_out, gates); break;
    }
    MPNN_CHECK_LAUNCH("k_gru_fwd_fused");
    return MPNN_OK;
  }
  if (mpnn_tc_gru_supported(d))  // widths 33..128: gate products + gate arithmetic fused on the tensor cores
    return mpnn_tc_gru_fwd(m, h, mask, W_ih, W_hh, b_ih, b_hh, rows, d, h_out, gates, workspace, workspace_bytes, stream);
  char* wp = (char*)workspace;
  float* gi = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  float* gh = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  int R = (int)rows;
  int rc;
  const int DP = mpnn_tc_dp(d, d);
  if (DP > 0) {
    // [rows, d] x [d, 3d] as three d x d N-blocks on the tcgen05 grouped-GEMM kernel (TF32 operands, fp32 accumulate)
    void* sub = wp;
    size_t sub_bytes = workspace_bytes - (size_t)(wp - (char*)workspace);
    rc = mpnn_tc_dense_gemm(m, rows, d, d, 1, 0, W_ih, 1, 3 * d, d, 0, 3, d, b_ih, gi, 3 * d, d, 0, DP, sub, sub_bytes, stream);
    if (rc) return rc;
    rc = mpnn_tc_dense_gemm(h, rows, d, d, 1, 0, W_hh, 1, 3 * d, d, 0, 3, d, b_hh, gh, 3 * d, d, 0, DP, sub, sub_bytes, stream);
    if (rc) return rc;
  } else {
    rc = mpnn_gemm(m, W_ih, gi, R, 3 * d, d, d, 1, 3 * d, 1, 3 * d, b_ih, 0, nullptr, 0, stream);
    if (rc) return rc;
    rc = mpnn_gemm(h, W_hh, gh, R, 3 * d, d, d, 1, 3 * d, 1, 3 * d, b_hh, 0, nullptr, 0, stream);
    if (rc) return rc;
  }
  k_gru_point_fwd<<<ceil_div(rows * d, 256), 256, 0, stream>>>(gi, gh, h, mask, rows, d, h_out, gates);
  MPNN_CHECK_LAUNCH("k_gru_point_fwd");
  return MPNN_OK;
}

// 1 (default): at widths 33..64 the pointwise GRU backward runs inside the weight-gradient kernel's producers; 0: as a
// separate launch (the form every other tensor-core width uses).  Returns the previous setting.
int mpnn_gru_bwd_one_pass(int enabled) {
  const int prev = g_point_in_param;
  g_point_in_param = enabled ? 1 : 0;
  return prev;
}

// GRU forward with the message aggregation folded in (tensor-core widths only, mpnn_gru_agg_supported): the messages
// are sum_{e in [row_ptr[i], row_ptr[i+1])} Y[e, :]; they are written to m_out [rows, d] as well (saved for backward).
int mpnn_gru_agg_supported(int d) { return mpnn_tc_gru_supported(d); }

int mpnn_gru_fwd_agg(const float* Y, const int* row_ptr, const float* h, const float* mask, const float* W_ih,
                     const float* W_hh, const float* b_ih, const float* b_hh, long long rows, int d, float* m_out,
                     float* h_out, float* gates, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && d > 0 && rows < (1ll << 31), MPNN_ERR_ARG, "gru_fwd_agg: bad dims");
  MPNN_REQUIRE(mpnn_tc_gru_supported(d), MPNN_ERR_UNSUPPORTED, "gru_fwd_agg: width %d not served", d);
  MPNN_REQUIRE(workspace_bytes >= mpnn_gru_workspace_bytes(rows, d), MPNN_ERR_WORKSPACE, "gru_fwd_agg: workspace");
  return mpnn_tc_gru_fwd_agg(Y, row_ptr, nullptr, h, mask, W_ih, W_hh, b_ih, b_hh, rows, d, m_out, h_out, gates,
                             workspace, workspace_bytes, stream);
}

// ---- shared-parameter form: the reference applies ONE GRUCell at every message-passing step (basic_model.py:50-58),
// so its four parameter gradients are sums over the T steps.  Each step's backward leaves its per-CTA partials in a
// caller-provided slab; one reduction over all slabs replaces T reductions + 4 (T-1) accumulate kernels of autograd.
// widths <= 32 (the fused CUDA-core kernels); returns 0 bytes otherwise.
size_t mpnn_gru_bwd_partial_bytes(long long rows, int d) {
  if (d > 32 || d <= 0 || rows <= 0) return 0;
  const int DP = fused_dp(d);
  return (size_t)gru_bwd_grid(rows, DP) * (6 * (size_t)d * d + 6 * d) * sizeof(float);
}

int mpnn_gru_bwd_data(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                      const float* gates, const float* dh_out, long long rows, int d, float* dm, float* dh,
                      float* partial, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && d > 0 && d <= 32 && rows < (1ll << 31), MPNN_ERR_ARG, "gru_bwd_data: bad dims");
  const int DP = fused_dp(d);
  const int grid = gru_bwd_grid(rows, DP);
  const size_t smem = gru_bwd_smem(d, DP);
  switch (DP) {
    case 8: k_gru_bwd_fused<8><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, gates, dh_out, rows, d, dm, dh, partial); break;
    case 16: k_gru_bwd_fused<16><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, gates, dh_out, rows, d, dm, dh, partial); break;
    default: k_gru_bwd_fused<32><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, gates, dh_out, rows, d, dm, dh, partial); break;
  }
  MPNN_CHECK_LAUNCH("k_gru_bwd_fused");
  return MPNN_OK;
}

// partial: `slabs` consecutive slabs of mpnn_gru_bwd_partial_bytes(rows, d) each (one per step, any order)
int mpnn_gru_bwd_params(const float* partial, int slabs, long long rows, int d, float* dW_ih, float* dW_hh,
                        float* db_ih, float* db_hh, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && d > 0 && d <= 32 && slabs > 0, MPNN_ERR_ARG, "gru_bwd_params: bad dims");
  const int grid = gru_bwd_grid(rows, fused_dp(d));
  k_gru_bwd_reduce<<<ceil_div(6 * d * d + 6 * d, 32), 256, 0, stream>>>(partial, slabs * grid, d, dW_ih, dW_hh, db_ih,
                                                                        db_hh);
  MPNN_CHECK_LAUNCH("k_gru_bwd_reduce");
  return MPNN_OK;
}

int mpnn_gru_bwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                 const float* gates, const float* dh_out, long long rows, int d, float* dm, float* dh, float* dW_ih,
                 float* dW_hh, float* db_ih, float* db_hh, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && d > 0 && rows < (1ll << 31), MPNN_ERR_ARG, "gru_bwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_gru_workspace_bytes(rows, d), MPNN_ERR_WORKSPACE, "gru_bwd: workspace");
  if (d <= 32) {
    const int DP = fused_dp(d);
    const int grid = gru_bwd_grid(rows, DP);
    const size_t smem = gru_bwd_smem(d, DP);
    float* partial = (float*)workspace;
    switch (DP) {
      case 8: k_gru_bwd_fused<8><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, gates, dh_out, rows, d, dm, dh, partial); break;
      case 16: k_gru_bwd_fused<16><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, gates, dh_out, rows, d, dm, dh, partial); break;
      default: k_gru_bwd_fused<32><<<grid, 256, smem, stream>>>(m, h, mask, W_ih, W_hh, gates, dh_out, rows, d, dm, dh, partial); break;
    }
    MPNN_CHECK_LAUNCH("k_gru_bwd_fused");
    k_gru_bwd_reduce<<<ceil_div(6 * d * d + 6 * d, 32), 256, 0, stream>>>(partial, grid, d, dW_ih, dW_hh, db_ih, db_hh);
    MPNN_CHECK_LAUNCH("k_gru_bwd_reduce");
    return MPNN_OK;
  }
  const int DP = mpnn_tc_dp(d, d);
  if (DP > 0 && ((dh - dm) & 3) == 0) {
    // ---- tensor-core path: three passes over ONE gate-gradient array ----
    //   1. k_gru_point_bwd5: dg [rows, 6d] = dar | daz | dan | dnh | hi(go z) | lo(go z), and the bias gradients
    //   2. one grouped product for both data gradients: (dm | dh) = dg Wc (block g of the output is buffer g)
    //   3. the weight gradients: widths <= 64 in one pass (m and h stacked into one M = 128 operand), else per product
    char* wp = (char*)workspace;
    float* dg = (float*)wp;
    wp += align_up((size_t)rows * 6 * d * sizeof(float), 256);
    const int nblk = 4 * mpnn_num_sms();
    float* bias_part = (float*)wp;
    wp += align_up((size_t)16 * mpnn_num_sms() * 4 * d * sizeof(float), 256);
    float* Wc = (float*)wp;
    wp += align_up((size_t)12 * d * d * sizeof(float), 256);
    void* sub = wp;
    size_t sub_bytes = workspace_bytes - (size_t)(wp - (char*)workspace);
    int rc;
    const bool one_pass = d <= 64 && g_point_in_param;
    if (one_pass) {
      // widths <= 64: the gate gradients never reach HBM.  The pointwise pass lives in the producers of BOTH tensor-core
      // kernels: the weight-gradient kernel (MN-major operand, + bias column sums) and the data-gradient kernel (K-major
      // operand, six stages = the six gate-gradient blocks); each reads the saved gates / h / dh' once.
      if ((rc = mpnn_tc_gru_param_point(m, h, mask, gates, dh_out, rows, d, nullptr, bias_part, dW_ih, dW_hh, sub,
                                        sub_bytes, stream)))
        return rc;
      k_gru_bias_final<<<ceil_div(4 * d, 32), 256, 0, stream>>>(bias_part, mpnn_tc_gru_param_bias_parts(), d, db_ih, db_hh);
      MPNN_CHECK_LAUNCH("k_gru_bias_final");
      k_gru_bwd_wcomb<<<ceil_div(12 * d * d, 256), 256, 0, stream>>>(W_ih, W_hh, d, Wc);
      MPNN_CHECK_LAUNCH("k_gru_bwd_wcomb");
      return mpnn_tc_gru_data_grad(gates, h, dh_out, mask, Wc, rows, d, dm, dh, sub, sub_bytes, stream);
    } else {
      k_gru_point_bwd5<<<nblk, 256, 0, stream>>>(gates, h, mask, dh_out, rows, d, dg, bias_part);
      MPNN_CHECK_LAUNCH("k_gru_point_bwd5");
      k_gru_bias_final<<<ceil_div(4 * d, 32), 256, 0, stream>>>(bias_part, nblk, d, db_ih, db_hh);
      MPNN_CHECK_LAUNCH("k_gru_bias_final");
    }
    k_gru_bwd_wcomb<<<ceil_div(12 * d * d, 256), 256, 0, stream>>>(W_ih, W_hh, d, Wc);
    MPNN_CHECK_LAUNCH("k_gru_bwd_wcomb");
    if (2 * d <= 256) {   // one N = 2d product, columns [0, d) -> dm, [d, 2d) -> dh: the gate gradients are read once
      const int DP2 = 2 * d <= 64 ? 64 : 2 * d <= 128 ? 128 : 256;
      rc = mpnn_tc_dense_gemm_ll(dg, rows, 6 * d, d, 6, d, Wc, d, 1, 0, (long long)2 * d * d, 1, 2 * d, nullptr, dm, d,
                                 (long long)(dh - dm), d, 0, DP2, sub, sub_bytes, stream);
    } else {              // two output groups
      rc = mpnn_tc_dense_gemm_ll(dg, rows, 6 * d, d, 6, d, Wc, d, 1, (long long)d * d, (long long)2 * d * d, 2, d, nullptr,
                                 dm, d, (long long)(dh - dm), 0, 0, DP, sub, sub_bytes, stream);
    }
    if (rc) return rc;
    if (one_pass) return MPNN_OK;
    if (d <= 64) return mpnn_tc_gru_param_grad(m, h, dg, 6 * d, rows, d, dW_ih, dW_hh, sub, sub_bytes, stream);
    if ((rc = mpnn_tc_dense_gemm_tn(m, rows, d, d, dg, 6 * d, d, 3, d, DP, dW_ih, d, 3 * d, sub, sub_bytes, stream))) return rc;
    if ((rc = mpnn_tc_dense_gemm_tn(h, rows, d, d, dg, 6 * d, d, 2, d, DP, dW_hh, d, 3 * d, sub, sub_bytes, stream))) return rc;
    return mpnn_tc_dense_gemm_tn(h, rows, d, d, dg + 3 * d, 6 * d, d, 1, d, DP, dW_hh + 2 * d, d, 3 * d, sub, sub_bytes, stream);
  }
  char* wp = (char*)workspace;
  float* dgi = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  float* dgh = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  void* sub = wp;
  size_t sub_bytes = workspace_bytes - (size_t)(wp - (char*)workspace);
  int R = (int)rows;
  k_gru_point_bwd<<<ceil_div(rows * d, 256), 256, 0, stream>>>(gates, h, mask, dh_out, rows, d, dgi, dgh, dh);
  MPNN_CHECK_LAUNCH("k_gru_point_bwd");
  int rc;
  // dm = dgi W_ih^T ; dh += dgh W_hh^T
  if ((rc = mpnn_gemm(dgi, W_ih, dm, R, d, 3 * d, 3 * d, 1, 1, 3 * d, d, nullptr, 0, nullptr, 0, stream))) return rc;
  if ((rc = mpnn_gemm(dgh, W_hh, dh, R, d, 3 * d, 3 * d, 1, 1, 3 * d, d, nullptr, 2, nullptr, 0, stream))) return rc;
  // dW_ih = m^T dgi ; dW_hh = h^T dgh   (split-K, fixed-order reduction)
  if ((rc = mpnn_gemm(m, dgi, dW_ih, d, 3 * d, R, 1, d, 3 * d, 1, 3 * d, nullptr, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_gemm(h, dgh, dW_hh, d, 3 * d, R, 1, d, 3 * d, 1, 3 * d, nullptr, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_colsum(dgi, nullptr, rows, 3 * d, 3 * d, 0, db_ih, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_colsum(dgh + 2 * d, nullptr, rows, d, 3 * d, 0, db_hh + 2 * d, 0, sub, sub_bytes, stream))) return rc;
  MPNN_CUDA(cudaMemcpyAsync(db_hh, db_ih, (size_t)2 * d * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  return MPNN_OK;
}

}  // extern "C"
