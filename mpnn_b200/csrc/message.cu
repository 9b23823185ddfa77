// Fused message function + neighbour aggregation (reference edge_network.py:42-52 + adjacent_message_agg.py:18).
//
// The reference materialises a d x d matrix per atom PAIR (edge_network.py:37-38, [B, N*mf, N*nf]) and
// multiplies it with the neighbour state.  Here, per receiver row i:
//
//     Z[i, p, l] = sum_{e in E(i)} alpha_e * x~_e[p] * g_e[l]  +  x~_0[p] * Q[i, l]
//     M[i, k]    = sum_{p,l} W~[k, l, p] * Z[i, p, l]  (+ beta[k])
//
// x~ = [x, 1] is the trunk output of the edge's bond row with a constant feature appended, so that
// W~[:, :, P] = the last Linear's bias and W~[:, :, :P] its weight (edge_network.py:21).  g_e is the
// sender state (afm[src_e], gather mode) or an explicit per-edge vector (attention-gated sender state,
// att_edge_network.py:26).  alpha_e is the aggregation weight (adj value / softmax weight / 1), Q the
// "virtual edge" that carries every non-bonded pair of the HEAD form (their bond row is all-zero, so
// they share x_0).  Z lives in shared memory only; neither per-edge matrices nor [B,N,N,mf] ever exist.
//
// This file is the fp32 CUDA-core implementation (exact-parity path).  Tile = TM receiver rows x 128 of
// the K = (P+1)*DP contraction axis per step; DP = feature width padded to a power of two.
#include "common.cuh"

extern "C" int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy,
                           float* out, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_colsum_workspace_bytes(long long rows, int width);

namespace {

constexpr int KC = 128;  // contraction elements per chunk

template <int DP>
struct Cfg {
  static constexpr int PC = KC / DP;             // p values per chunk
  static constexpr int CGM = DP / 4;             // float4 column groups of the output
  static constexpr int RGN = 256 / CGM;          // row groups
  static constexpr int RT = DP <= 16 ? 1 : DP / 16;  // rows per thread  (DP 32->2, 64->4, 128->8)
  static constexpr int TM = RT * RGN;            // receiver rows per tile
  static constexpr int KT = DP / 8;              // dW: kk per thread
  static constexpr int WS = DP + 4;              // smem row stride of W / dM tiles
  static constexpr int ZS = KC + 1;              // smem row stride of Z / dZ tiles
};

struct MsgArgs {
  const int* row_ptr;
  const int* edge_dst;
  const int* gidx;     // nullable: g_e = Gsrc[e]
  const int* xid;      // nullable: x_e = X[e]
  const float* alpha;  // nullable: 1
  const float* X;      // [R, ldx]
  const float* Gsrc;   // [*, ldg]
  const float* Q;      // nullable [nrows, nf]
  const float* Wt;     // [(P+1)*DP, DP]
  const float* beta;   // nullable [mf]
  int ldx, ldg, x0_row;
  int nrows, nf, mf, P;
};

// x~_e[p]
__device__ __forceinline__ float xt(const float* __restrict__ X, int ldx, int row, int p, int P) {
  return p < P ? __ldg(X + (size_t)row * ldx + p) : 1.0f;
}

template <int DP>
__device__ __forceinline__ void load_w_chunk(float* Ws, const float* __restrict__ Wt, int p0, int P) {
  using C = Cfg<DP>;
  const int total_rows = (P + 1) * DP;
  for (int idx = threadIdx.x; idx < KC * (DP / 4); idx += 256) {
    int kk = idx / (DP / 4), c4 = idx - kk * (DP / 4);
    int grow = p0 * DP + kk;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow < total_rows) v = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)grow * DP) + c4);
    *reinterpret_cast<float4*>(Ws + kk * C::WS + c4 * 4) = v;
  }
}

// Zs[r][kk] for the tile starting at receiver i0 and the chunk starting at p0
template <int DP>
__device__ __forceinline__ void build_z_chunk(float* Zs, const MsgArgs& a, int i0, int p0) {
  using C = Cfg<DP>;
  for (int idx = threadIdx.x; idx < C::TM * KC; idx += 256) {
    const int r = idx / KC, kk = idx - r * KC;  // a warp covers 32 consecutive kk of one row
    const int pc = kk / DP, l = kk - pc * DP;
    const int p = p0 + pc;
    const int i = i0 + r;
    float z = 0.f;
    if (i < a.nrows && l < a.nf && p <= a.P) {
      const int eb = a.row_ptr[i], ee = a.row_ptr[i + 1];
      for (int e = eb; e < ee; ++e) {
        const int xr = a.xid ? a.xid[e] : e;
        const int gr = a.gidx ? a.gidx[e] : e;
        float v = xt(a.X, a.ldx, xr, p, a.P) * __ldg(a.Gsrc + (size_t)gr * a.ldg + l);
        if (a.alpha) v *= a.alpha[e];
        z += v;
      }
      if (a.Q) z = fmaf(xt(a.X, a.ldx, a.x0_row, p, a.P), __ldg(a.Q + (size_t)i * a.nf + l), z);
    }
    Zs[r * C::ZS + kk] = z;
  }
}

template <int DP>
__device__ __forceinline__ bool tile_is_empty(const MsgArgs& a, int i0) {
  if (a.Q) return false;
  int i1 = i0 + Cfg<DP>::TM;
  if (i1 > a.nrows) i1 = a.nrows;
  return a.row_ptr[i0] == a.row_ptr[i1];
}

// ---------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(256) k_msg_fwd(MsgArgs a, float* __restrict__ M) {
  using C = Cfg<DP>;
  extern __shared__ __align__(16) float smem[];
  float* Zs = smem;                    // [TM][ZS]
  float* Ws = Zs + C::TM * C::ZS;      // [KC][WS]
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * C::TM;
  const int cg = tid % C::CGM, rg = tid / C::CGM;
  float acc[C::RT][4];
#pragma unroll
  for (int t = 0; t < C::RT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  if (!tile_is_empty<DP>(a, i0)) {
    const int nchunks = (a.P + 1 + C::PC - 1) / C::PC;
    for (int c = 0; c < nchunks; ++c) {
      const int p0 = c * C::PC;
      __syncthreads();
      load_w_chunk<DP>(Ws, a.Wt, p0, a.P);
      build_z_chunk<DP>(Zs, a, i0, p0);
      __syncthreads();
      const float* wp = Ws + cg * 4;
      const float* zp = Zs + (rg * C::RT) * C::ZS;
#pragma unroll 8
      for (int kk = 0; kk < KC; ++kk) {
        float4 w = *reinterpret_cast<const float4*>(wp + kk * C::WS);
#pragma unroll
        for (int t = 0; t < C::RT; ++t) {
          float z = zp[t * C::ZS + kk];
          acc[t][0] = fmaf(z, w.x, acc[t][0]);
          acc[t][1] = fmaf(z, w.y, acc[t][1]);
          acc[t][2] = fmaf(z, w.z, acc[t][2]);
          acc[t][3] = fmaf(z, w.w, acc[t][3]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < C::RT; ++t) {
    int i = i0 + rg * C::RT + t;
    if (i >= a.nrows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = cg * 4 + j;
      if (k < a.mf) M[(size_t)i * a.mf + k] = acc[t][j] + (a.beta ? a.beta[k] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, data side: dZ = dM W~^T per chunk, then the per-edge reductions.
//   T[e, p]  = sum_l dZ[i,p,l] g_e[l]            (un-scaled d x~_e; column P = d(const feature), used for d alpha)
//   dG[e, l] = alpha_e sum_p dZ[i,p,l] x~_e[p]
//   dQ[i, l] = sum_p dZ[i,p,l] x~_0[p]
//   dx0 partial[cta, p] = sum_{i,l} dZ[i,p,l] Q[i,l]
template <int DP>
__global__ void __launch_bounds__(256) k_msg_bwd_edges(MsgArgs a, const float* __restrict__ dM, float* __restrict__ T,
                                                       int ldt, float* __restrict__ dG, float* __restrict__ dQ,
                                                       float* __restrict__ dx0_partial) {
  using C = Cfg<DP>;
  extern __shared__ __align__(16) float smem[];
  float* dZs = smem;                      // [TM][ZS]
  float* Ws = dZs + C::TM * C::ZS;        // [KC][WS]
  float* dMs = Ws + KC * C::WS;           // [TM][WS]
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * C::TM;
  int i1 = i0 + C::TM;
  if (i1 > a.nrows) i1 = a.nrows;
  const int eb = a.row_ptr[i0], ee = a.row_ptr[i1];
  const int ne = ee - eb;
  const int nchunks = (a.P + 1 + C::PC - 1) / C::PC;
  if (ne == 0 && !a.Q) {
    if (dx0_partial)
      for (int p = tid; p <= a.P; p += 256) dx0_partial[(size_t)blockIdx.x * (a.P + 1) + p] = 0.f;
    return;
  }
  for (int idx = tid; idx < C::TM * DP; idx += 256) {
    int r = idx / DP, k = idx - r * DP;
    int i = i0 + r;
    dMs[r * C::WS + k] = (i < a.nrows && k < a.mf) ? dM[(size_t)i * a.mf + k] : 0.f;
  }
  for (int c = 0; c < nchunks; ++c) {
    const int p0 = c * C::PC;
    __syncthreads();
    load_w_chunk<DP>(Ws, a.Wt, p0, a.P);
    __syncthreads();
    // dZs[r][kk] = sum_k dMs[r][k] * Ws[kk][k]
    for (int idx = tid; idx < C::TM * KC; idx += 256) {
      const int r = idx / KC, kk = idx - r * KC;
      const float* dp = dMs + r * C::WS;
      const float* wp = Ws + kk * C::WS;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DP; k += 4) {
        float4 d4 = *reinterpret_cast<const float4*>(dp + k);
        float4 w4 = *reinterpret_cast<const float4*>(wp + k);
        s = fmaf(d4.x, w4.x, s);
        s = fmaf(d4.y, w4.y, s);
        s = fmaf(d4.z, w4.z, s);
        s = fmaf(d4.w, w4.w, s);
      }
      dZs[r * C::ZS + kk] = s;
    }
    __syncthreads();
    // T[e][p]: item (e, pc), e fastest
    for (int idx = tid; idx < ne * C::PC; idx += 256) {
      const int pc = idx / ne, el = idx - pc * ne;
      const int p = p0 + pc;
      if (p > a.P) continue;
      const int e = eb + el;
      const int r = a.edge_dst[e] - i0;
      const int gr = a.gidx ? a.gidx[e] : e;
      const float* g = a.Gsrc + (size_t)gr * a.ldg;
      const float* dz = dZs + r * C::ZS + pc * DP;
      float s = 0.f;
      for (int l = 0; l < a.nf; ++l) s = fmaf(dz[l], __ldg(g + l), s);
      T[(size_t)e * ldt + p] = s;
    }
    // dG[e][l]: item (e, l), l fastest
    for (int idx = tid; idx < ne * a.nf; idx += 256) {
      const int el = idx / a.nf, l = idx - el * a.nf;
      const int e = eb + el;
      const int r = a.edge_dst[e] - i0;
      const int xr = a.xid ? a.xid[e] : e;
      const float* dz = dZs + r * C::ZS + l;
      float s = 0.f;
#pragma unroll
      for (int pc = 0; pc < C::PC; ++pc) {
        int p = p0 + pc;
        if (p <= a.P) s = fmaf(dz[pc * DP], xt(a.X, a.ldx, xr, p, a.P), s);
      }
      if (a.alpha) s *= a.alpha[e];
      float* o = dG + (size_t)e * a.nf + l;
      *o = (c == 0) ? s : *o + s;
    }
    if (a.Q) {
      for (int idx = tid; idx < (i1 - i0) * a.nf; idx += 256) {
        const int r = idx / a.nf, l = idx - r * a.nf;
        const float* dz = dZs + r * C::ZS + l;
        float s = 0.f;
#pragma unroll
        for (int pc = 0; pc < C::PC; ++pc) {
          int p = p0 + pc;
          if (p <= a.P) s = fmaf(dz[pc * DP], xt(a.X, a.ldx, a.x0_row, p, a.P), s);
        }
        float* o = dQ + (size_t)(i0 + r) * a.nf + l;
        *o = (c == 0) ? s : *o + s;
      }
      // dx0 partial: one warp per pc, lanes over (r,l), fixed reduction order
      const int warp = tid >> 5, lane = tid & 31;
      for (int pc = warp; pc < C::PC; pc += 8) {
        int p = p0 + pc;
        if (p > a.P) continue;
        float s = 0.f;
        for (int q = lane; q < (i1 - i0) * a.nf; q += 32) {
          int r = q / a.nf, l = q - r * a.nf;
          s = fmaf(dZs[r * C::ZS + pc * DP + l], __ldg(a.Q + (size_t)(i0 + r) * a.nf + l), s);
        }
        s = warp_sum(s);
        if (lane == 0) dx0_partial[(size_t)blockIdx.x * (a.P + 1) + p] = s;
      }
    }
  }
}

// d alpha_e = sum_{p<=P} T[e,p] x~_e[p];  T[e,:] *= alpha_e  (T becomes d x_e)
__global__ void k_msg_alpha_bwd(const float* __restrict__ alpha, const int* __restrict__ xid,
                                const float* __restrict__ X, int ldx, int P, int E, float* __restrict__ T, int ldt,
                                float* __restrict__ dalpha) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= E) return;
  const int xr = xid ? xid[warp] : warp;
  float* t = T + (size_t)warp * ldt;
  const float al = alpha[warp];
  float s = 0.f;
  for (int p = lane; p <= P; p += 32) {
    float tv = t[p];
    s = fmaf(tv, xt(X, ldx, xr, p, P), s);
    t[p] = tv * al;
  }
  s = warp_sum(s);
  if (lane == 0) dalpha[warp] = s;
}

// T[x0_row, p] = sum over CTAs of dx0 partials (or 0)
__global__ void k_msg_dx0_reduce(const float* __restrict__ partial, int nparts, int P, float* __restrict__ trow) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > P) return;
  float s = 0.f;
  if (partial)
    for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * (P + 1) + p];
  trow[p] = s;
}

// ---------------------------------------------------------------------------------------------
// backward, weight side: dW~[p,l,k] = sum_i Z[i,p,l] dM[i,k].  grid = (chunks, splits); Z is rebuilt.
template <int DP>
__global__ void __launch_bounds__(256) k_msg_bwd_weights(MsgArgs a, const float* __restrict__ dM, int tiles_per_split,
                                                         float* __restrict__ partial /*[splits][chunks*KC][DP]*/) {
  using C = Cfg<DP>;
  extern __shared__ __align__(16) float smem[];
  float* Zs = smem;                     // [TM][ZS]
  float* dMs = Zs + C::TM * C::ZS;      // [TM][WS]
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, split = blockIdx.y;
  const int p0 = chunk * C::PC;
  const int ntiles = (a.nrows + C::TM - 1) / C::TM;
  const int t0 = split * tiles_per_split;
  int t1 = t0 + tiles_per_split;
  if (t1 > ntiles) t1 = ntiles;
  const int cg = tid % C::CGM, kg = tid / C::CGM;
  float acc[C::KT][4];
#pragma unroll
  for (int t = 0; t < C::KT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  for (int tile = t0; tile < t1; ++tile) {
    const int i0 = tile * C::TM;
    if (tile_is_empty<DP>(a, i0)) continue;  // block-uniform
    __syncthreads();
    build_z_chunk<DP>(Zs, a, i0, p0);
    for (int idx = tid; idx < C::TM * DP; idx += 256) {
      int r = idx / DP, k = idx - r * DP;
      int i = i0 + r;
      dMs[r * C::WS + k] = (i < a.nrows && k < a.mf) ? dM[(size_t)i * a.mf + k] : 0.f;
    }
    __syncthreads();
    const float* zp = Zs + kg * C::KT;
    const float* dp = dMs + cg * 4;
#pragma unroll 4
    for (int r = 0; r < C::TM; ++r) {
      float4 d4 = *reinterpret_cast<const float4*>(dp + r * C::WS);
#pragma unroll
      for (int t = 0; t < C::KT; ++t) {
        float z = zp[r * C::ZS + t];
        acc[t][0] = fmaf(z, d4.x, acc[t][0]);
        acc[t][1] = fmaf(z, d4.y, acc[t][1]);
        acc[t][2] = fmaf(z, d4.z, acc[t][2]);
        acc[t][3] = fmaf(z, d4.w, acc[t][3]);
      }
    }
  }
  float* out = partial + ((size_t)split * gridDim.x + chunk) * KC * DP;
#pragma unroll
  for (int t = 0; t < C::KT; ++t) {
    int kk = kg * C::KT + t;
    *reinterpret_cast<float4*>(out + (size_t)kk * DP + cg * 4) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
  }
}

// dW_last[(k*nf + l)*P + p], dB_last[k*nf + l] from the split partials (fixed summation order)
__global__ void k_msg_dw_reduce(const float* __restrict__ partial, int splits, int nchunks, int DP, int nf, int mf,
                                int P, float* __restrict__ dW, float* __restrict__ dB) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over (k, l, p<=P), p fastest
  int total = mf * nf * (P + 1);
  if (idx >= total) return;
  int p = idx % (P + 1);
  int kl = idx / (P + 1);
  int l = kl % nf, k = kl / nf;
  size_t row = (size_t)p * DP + l;  // global kk index: chunk*KC + kk == p*DP + l because KC == PC*DP
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)sp * nchunks * KC + row) * DP + k];
  if (p < P)
    dW[(size_t)kl * P + p] = s;
  else
    dB[kl] = s;
}

// Wt[(p*DP + l)*DP + k] = W_last[(k*nf + l)*P + p] (p < P) | B_last[k*nf + l] (p == P) | 0 (padding)
__global__ void k_msg_prepare(const float* __restrict__ W, const float* __restrict__ Bv, int nf, int mf, int P, int DP,
                              float* __restrict__ Wt) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)(P + 1) * DP * DP;
  if (idx >= total) return;
  int k = (int)(idx % DP);
  int l = (int)((idx / DP) % DP);
  int p = (int)(idx / ((long long)DP * DP));
  float v = 0.f;
  if (k < mf && l < nf) v = p < P ? W[((size_t)k * nf + l) * P + p] : Bv[k * nf + l];
  Wt[idx] = v;
}

int pick_dp(int nf, int mf) {
  int d = nf > mf ? nf : mf;
  return pow2_at_least(d, 8);
}

template <int DP>
size_t smem_fwd() {
  using C = Cfg<DP>;
  return (size_t)(C::TM * C::ZS + KC * C::WS) * sizeof(float);
}
template <int DP>
size_t smem_bwd_edges() {
  using C = Cfg<DP>;
  return (size_t)(C::TM * C::ZS + KC * C::WS + C::TM * C::WS) * sizeof(float);
}
template <int DP>
size_t smem_bwd_weights() {
  using C = Cfg<DP>;
  return (size_t)(C::TM * C::ZS + C::TM * C::WS) * sizeof(float);
}

struct BwdOut {
  float* T;
  int ldt;
  float* dG;
  float* dQ;
  float* dx0_partial;
  float* dw_partial;
  int splits, tiles_per_split;
};

template <int DP>
int launch_fwd(const MsgArgs& a, float* M, cudaStream_t stream) {
  using C = Cfg<DP>;
  size_t smem = smem_fwd<DP>();
  MPNN_CUDA(cudaFuncSetAttribute(k_msg_fwd<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_msg_fwd<DP><<<ceil_div(a.nrows, C::TM), 256, smem, stream>>>(a, M);
  MPNN_CHECK_LAUNCH("k_msg_fwd");
  return MPNN_OK;
}

template <int DP>
int launch_bwd(const MsgArgs& a, const float* dM, const BwdOut& o, cudaStream_t stream) {
  using C = Cfg<DP>;
  size_t s1 = smem_bwd_edges<DP>();
  MPNN_CUDA(cudaFuncSetAttribute(k_msg_bwd_edges<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
  k_msg_bwd_edges<DP><<<ceil_div(a.nrows, C::TM), 256, s1, stream>>>(a, dM, o.T, o.ldt, o.dG, o.dQ, o.dx0_partial);
  MPNN_CHECK_LAUNCH("k_msg_bwd_edges");
  size_t s2 = smem_bwd_weights<DP>();
  MPNN_CUDA(cudaFuncSetAttribute(k_msg_bwd_weights<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  int nchunks = ceil_div(a.P + 1, C::PC);
  dim3 grid(nchunks, o.splits);
  k_msg_bwd_weights<DP><<<grid, 256, s2, stream>>>(a, dM, o.tiles_per_split, o.dw_partial);
  MPNN_CHECK_LAUNCH("k_msg_bwd_weights");
  return MPNN_OK;
}

template <int DP>
int tile_rows() {
  return Cfg<DP>::TM;
}

int tm_for(int DP) {
  switch (DP) {
    case 8: return tile_rows<8>();
    case 16: return tile_rows<16>();
    case 32: return tile_rows<32>();
    case 64: return tile_rows<64>();
    default: return tile_rows<128>();
  }
}

void plan_splits(int nrows, int P, int DP, int* splits, int* tiles_per_split) {
  int ntiles = ceil_div(nrows, tm_for(DP));
  int nchunks = ceil_div(P + 1, KC / DP);
  int want = ceil_div(2 * mpnn_num_sms(), nchunks);
  if (want > ntiles) want = ntiles;
  if (want < 1) want = 1;
  int tps = ceil_div(ntiles, want);
  *tiles_per_split = tps;
  *splits = ceil_div(ntiles, tps);
}

}  // namespace

extern "C" {

// floats needed for the transposed/augmented last-layer weight
long long mpnn_message_wt_floats(int nf, int mf, int P) {
  int DP = pick_dp(nf, mf);
  if (DP > 128) return -1;
  return (long long)(P + 1) * DP * DP;
}

int mpnn_message_prepare(const float* W_last, const float* B_last, int nf, int mf, int P, float* Wt,
                         cudaStream_t stream) {
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(nf > 0 && mf > 0 && P > 0, MPNN_ERR_ARG, "message_prepare: bad dims");
  MPNN_REQUIRE(DP <= 128, MPNN_ERR_UNSUPPORTED, "message: feature width %d > 128 is not supported by the fp32 path",
               nf > mf ? nf : mf);
  long long total = (long long)(P + 1) * DP * DP;
  k_msg_prepare<<<ceil_div(total, 256), 256, 0, stream>>>(W_last, B_last, nf, mf, P, DP, Wt);
  MPNN_CHECK_LAUNCH("k_msg_prepare");
  return MPNN_OK;
}

int mpnn_message_fwd(const int* row_ptr, const int* edge_dst, const int* gidx, const int* xid, const float* alpha,
                     const float* X, int ldx, int x0_row, const float* Gsrc, int ldg, const float* Q, const float* Wt,
                     const float* beta, int nrows, int nf, int mf, int P, float* M, cudaStream_t stream) {
  MPNN_REQUIRE(nrows > 0 && nf > 0 && mf > 0 && P > 0, MPNN_ERR_ARG, "message_fwd: bad dims");
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(DP <= 128, MPNN_ERR_UNSUPPORTED, "message_fwd: feature width > 128 unsupported by the fp32 path");
  MsgArgs a = {row_ptr, edge_dst, gidx, xid, alpha, X, Gsrc, Q, Wt, beta, ldx, ldg, x0_row, nrows, nf, mf, P};
  switch (DP) {
    case 8: return launch_fwd<8>(a, M, stream);
    case 16: return launch_fwd<16>(a, M, stream);
    case 32: return launch_fwd<32>(a, M, stream);
    case 64: return launch_fwd<64>(a, M, stream);
    default: return launch_fwd<128>(a, M, stream);
  }
}

size_t mpnn_message_bwd_workspace_bytes(int nrows, int nf, int mf, int P) {
  int DP = pick_dp(nf, mf);
  if (DP > 128) return 0;
  int splits, tps;
  plan_splits(nrows, P, DP, &splits, &tps);
  int nchunks = ceil_div(P + 1, KC / DP);
  size_t dw = (size_t)splits * nchunks * KC * DP * sizeof(float);
  size_t dx0 = (size_t)ceil_div(nrows, tm_for(DP)) * (P + 1) * sizeof(float);
  size_t cs = mpnn_colsum_workspace_bytes(nrows, mf);
  return align_up(dw, 256) + align_up(dx0, 256) + align_up(cs, 256);
}

// Outputs (all written, not accumulated):
//   T      [n_edges + 1, ldt]  d x_e (column P is scratch); row x0_row-equivalent (index n_edges) = d x_0
//   dG     [n_edges, nf]       d g_e
//   dQ     [nrows, nf]         (only when Q != NULL)
//   dalpha [n_edges]           (only when alpha != NULL)
//   dW_last [mf*nf, P], dB_last [mf*nf], dbeta [mf] (only when beta != NULL)
int mpnn_message_bwd(const int* row_ptr, const int* edge_dst, const int* gidx, const int* xid, const float* alpha,
                     const float* X, int ldx, int x0_row, const float* Gsrc, int ldg, const float* Q, const float* Wt,
                     const float* beta, int nrows, int n_edges, int nf, int mf, int P, const float* dM, float* T,
                     int ldt, float* dG, float* dQ, float* dalpha, float* dW_last, float* dB_last, float* dbeta,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(nrows > 0 && nf > 0 && mf > 0 && P > 0 && n_edges >= 0, MPNN_ERR_ARG, "message_bwd: bad dims");
  MPNN_REQUIRE(ldt >= P + 1, MPNN_ERR_ARG, "message_bwd: ldt must be >= P+1");
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(DP <= 128, MPNN_ERR_UNSUPPORTED, "message_bwd: feature width > 128 unsupported by the fp32 path");
  MPNN_REQUIRE(workspace_bytes >= mpnn_message_bwd_workspace_bytes(nrows, nf, mf, P), MPNN_ERR_WORKSPACE,
               "message_bwd: workspace too small");
  MsgArgs a = {row_ptr, edge_dst, gidx, xid, alpha, X, Gsrc, Q, Wt, beta, ldx, ldg, x0_row, nrows, nf, mf, P};
  BwdOut o;
  plan_splits(nrows, P, DP, &o.splits, &o.tiles_per_split);
  int nchunks = ceil_div(P + 1, KC / DP);
  int ntiles = ceil_div(nrows, tm_for(DP));
  char* wp = (char*)workspace;
  o.dw_partial = (float*)wp;
  wp += align_up((size_t)o.splits * nchunks * KC * DP * sizeof(float), 256);
  o.dx0_partial = Q ? (float*)wp : nullptr;
  wp += align_up((size_t)ntiles * (P + 1) * sizeof(float), 256);
  void* cs_ws = wp;
  size_t cs_bytes = mpnn_colsum_workspace_bytes(nrows, mf);
  o.T = T;
  o.ldt = ldt;
  o.dG = dG;
  o.dQ = dQ;
  int rc;
  switch (DP) {
    case 8: rc = launch_bwd<8>(a, dM, o, stream); break;
    case 16: rc = launch_bwd<16>(a, dM, o, stream); break;
    case 32: rc = launch_bwd<32>(a, dM, o, stream); break;
    case 64: rc = launch_bwd<64>(a, dM, o, stream); break;
    default: rc = launch_bwd<128>(a, dM, o, stream); break;
  }
  if (rc) return rc;
  if (alpha && n_edges > 0) {
    k_msg_alpha_bwd<<<ceil_div((long long)n_edges * 32, 256), 256, 0, stream>>>(alpha, xid, X, ldx, P, n_edges, T, ldt,
                                                                               dalpha);
    MPNN_CHECK_LAUNCH("k_msg_alpha_bwd");
  }
  k_msg_dx0_reduce<<<ceil_div(P + 1, 128), 128, 0, stream>>>(o.dx0_partial, ntiles, P, T + (size_t)n_edges * ldt);
  MPNN_CHECK_LAUNCH("k_msg_dx0_reduce");
  k_msg_dw_reduce<<<ceil_div((long long)mf * nf * (P + 1), 256), 256, 0, stream>>>(o.dw_partial, o.splits, nchunks, DP,
                                                                                  nf, mf, P, dW_last, dB_last);
  MPNN_CHECK_LAUNCH("k_msg_dw_reduce");
  if (beta && dbeta) {
    rc = mpnn_colsum(dM, nullptr, nrows, mf, mf, 0, dbeta, 0, cs_ws, cs_bytes, stream);
    if (rc) return rc;
  }
  return MPNN_OK;
}

}  // extern "C"
