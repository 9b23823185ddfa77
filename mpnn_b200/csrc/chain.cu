// The whole T-step message-passing loop of the reference models as ONE persistent kernel each way (feature widths <= 32):
//
//     for t in range(T):  h = bn_t( uf( ma( mf_t(afm, bfm), adj ), h, mask ), mask )
//         models/normed_basic_model.py:56-59   (one EdgeNetwork per step, MaskBatchNorm)
//         models/basic_model.py:50-58          (shared EdgeNetwork, no batch norm)
//         models/normed_encoded_basic_model_ecfp.py:67-69  (one EdgeNetwork + one MaskBatchNorm1d per step)
//
// Per step and per receiver row, inside the kernel, nothing but the row's own state leaves the SM:
//     m  = sum_{e in E(i)} alpha_e T_t[uid_e]^T afm[src_e]     message function + aggregation  (edge_network.py:50-52 with
//                                                              the last Linear folded into the per-type table, csrc/typed.cu;
//                                                              adjacent_message_agg.py:18)
//     g  = GRU(m, h) * mask                                    gru_update.py:26-35,66-68
//     h' = BN(g)                                               mask_batch_norm.py:9-15 / :20-38; the batch statistics are the
//                                                              only coupling between rows: per-CTA (n, sum, M2) partials,
//                                                              ONE grid barrier per step, every CTA combines them in a fixed order
// Messages read the INPUT features at every step (normed_basic_model.py:58), so the gather has no dependency on the
// recurrence; rows are owned by a fixed CTA for the whole loop, so the recurrent state never crosses CTAs.
// The backward kernel walks the steps in reverse with the same ownership: batch-norm backward (one grid barrier per
// step for its two column sums), GRU backward, message gradients dM_t written for the table / sender gradients
// (csrc/typed.cu), the shared GRU cell's weight gradients accumulated in registers over all rows AND steps and reduced
// once in a fixed order behind a last barrier.  No float atomics: results are bit-reproducible for a fixed grid.
//
// Masks are the reference's 0/1 masks (pre_process/data_loader.py:18-21).
#include "common.cuh"

namespace {

constexpr int MAXT = 8;
constexpr int SS = 128;   // floats per step in the saved statistics: mean[32] | rstd[32] | var[32] | n, ...

struct BNDesc {
  int kind;       // 0 none, 1 MaskBatchNorm, 2 MaskBatchNorm1d
  int training;   // kind 2: batch statistics (1) or running statistics (0)
  float eps, momentum;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
};

struct Chain {
  const int* row_ptr;
  const int* edge_src;
  const int* uid;
  const float* alpha;
  int ecap, zero_type;
  const float* H0;       // [rows, d]  message input (afm)
  const float* h_init;   // [rows, d]  initial state
  const float* mask;     // [rows]
  const float* table[MAXT];
  const float* W_ih;
  const float* W_hh;
  const float* b_ih;
  const float* b_hh;
  BNDesc bn[MAXT];
  int T, rows, d;
  float* M;       // [T][rows][d]
  float* gates;   // [T][rows][4d]
  float* G;       // [T][rows][d]   GRU outputs before the batch norm
  float* stats;   // [T][SS]
  float* out;     // [rows][d]
  float* part;    // [T][grid][2*DP+2]
  unsigned* bar;  // [2] zero on entry, zero on exit
};

struct ChainB {
  Chain f;
  const float* dout;   // [rows, d]
  float* dM;           // [T][rows][d]
  float* dh_init;      // [rows][d] or null
  float* dY;           // [rows][d] scratch
  float* gpart;        // [grid][2*d*3d + 6d]
  float* dW_ih;
  float* dW_hh;
  float* db_ih;
  float* db_hh;
  float* dgamma[MAXT];
  float* dbeta[MAXT];
};

template <int DP>
__device__ __forceinline__ uint32_t grp_mask(int lane) {
  if constexpr (DP == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << DP) - 1u) << ((lane / DP) * DP);
  }
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the grid are co-resident (grid <= occupancy x SMs, checked on the host)
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire(bar) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void grid_exit(unsigned* bar) {
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(bar + 1, 1u);
    if (old == gridDim.x - 1) {   // every CTA is past its last barrier: leave the counters zero for the next launch
      bar[0] = 0u;
      bar[1] = 0u;
    }
  }
}

// fixed-order sum over the groups of a CTA of a per-(group, lane) value; every thread returns the column total
template <int DP>
__device__ __forceinline__ float block_colsum(float v, float* red, int grp, int c) {
  constexpr int GPB = 256 / DP;
  __syncthreads();
  red[grp * (DP + 1) + c] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < GPB; ++g) s += red[g * (DP + 1) + c];
  return s;
}

// fixed-order sum over the CTAs of the grid of part[cta * stride + idx]
template <int DP>
__device__ __forceinline__ float grid_colsum(const float* part, int stride, int idx, float* red, int grp, int c) {
  constexpr int GPB = 256 / DP;
  float s = 0.f;
  for (int cta = grp; cta < (int)gridDim.x; cta += GPB) s += __ldcg(part + (size_t)cta * stride + idx);
  return block_colsum<DP>(s, red, grp, c);
}

template <int DP>
__device__ __forceinline__ float message_row(const Chain& a, const float* __restrict__ table, int i, int k, uint32_t gm) {
  const int eb = min(a.row_ptr[i], a.ecap), ee = min(a.row_ptr[i + 1], a.ecap);
  float acc = 0.f;
  for (int e0 = eb; e0 < ee; e0 += DP) {
    const int cnt = min(DP, ee - e0);
    int jm = 0, um = 0;
    float am = 1.f;
    if (k < cnt) {
      jm = __ldg(a.edge_src + e0 + k);
      um = min(__ldg(a.uid + e0 + k), a.zero_type);
      if (a.alpha) am = __ldg(a.alpha + e0 + k);
    }
    int j = __shfl_sync(gm, jm, 0, DP);
    float hj = k < a.d ? __ldg(a.H0 + (size_t)j * a.d + k) : 0.f;
    for (int t = 0; t < cnt; ++t) {
      const int u = __shfl_sync(gm, um, t, DP);
      const float al = __shfl_sync(gm, am, t, DP);
      const float hcur = hj;
      if (t + 1 < cnt) {
        j = __shfl_sync(gm, jm, t + 1, DP);
        hj = k < a.d ? __ldg(a.H0 + (size_t)j * a.d + k) : 0.f;
      }
      const float* T = table + (size_t)u * DP * DP + k;
      float tv[DP];
#pragma unroll
      for (int l = 0; l < DP; ++l) tv[l] = __ldg(T + l * DP);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int l = 0; l < DP; l += 2) {
        s0 = fmaf(tv[l], __shfl_sync(gm, hcur, l, DP), s0);
        s1 = fmaf(tv[l + 1], __shfl_sync(gm, hcur, l + 1, DP), s1);
      }
      acc = fmaf(al, s0 + s1, acc);
    }
  }
  return acc;
}

// normalised value of a saved GRU output under the batch norm of its step: xh = (g - mean) * rstd * mu, and the module
// output y (kind 1: xh; kind 2: (gamma xh' + beta) mu); bnv = mean[DP] | rstd[DP] | gamma[DP] | beta[DP]
template <int DP>
__device__ __forceinline__ float bn_out(int kind, float g, float mu, const float* bnv, int c, float* xh_out) {
  if (kind == 0) {
    *xh_out = 0.f;
    return g;
  }
  const float xc = g - bnv[c];
  if (kind == 1) {
    const float y = (xc * mu) * bnv[DP + c];
    *xh_out = y;
    return y;
  }
  const float xn = xc * bnv[DP + c];
  *xh_out = xn * mu;
  return (bnv[2 * DP + c] * xn + bnv[3 * DP + c]) * mu;
}

template <int DP>
__device__ __forceinline__ void load_bnv(const Chain& a, int t, float* bnv) {
  // called by all threads between barriers; t < 0: identity
  __syncthreads();
  if (threadIdx.x < DP) {
    const int c = threadIdx.x;
    float mean = 0.f, rstd = 1.f, ga = 1.f, be = 0.f;
    if (t >= 0 && a.bn[t].kind) {
      const float* st = a.stats + (size_t)t * SS;
      mean = __ldcg(st + c);
      rstd = __ldcg(st + 32 + c);
      if (a.bn[t].kind == 2 && c < a.d) {
        if (a.bn[t].gamma) ga = a.bn[t].gamma[c];
        if (a.bn[t].beta) be = a.bn[t].beta[c];
      }
    }
    bnv[c] = mean;
    bnv[DP + c] = rstd;
    bnv[2 * DP + c] = ga;
    bnv[3 * DP + c] = be;
  }
  __syncthreads();
}

// ===================================================================================================================
// forward
// ===================================================================================================================
template <int DP>
__global__ void __launch_bounds__(256, 2) k_chain_fwd(Chain a) {
  constexpr int GPB = 256 / DP;
  constexpr int PS = 2 * DP + 2;
  extern __shared__ __align__(16) float sm[];
  const int d = a.d, d3 = 3 * d, rows = a.rows;
  float* Wi = sm;                        // [d][3d]
  float* Wh = Wi + d * d3;               // [d][3d]
  float* red = Wh + d * d3;              // [GPB][DP+1]
  float* bnv = red + GPB * (DP + 1);     // [4][DP]
  for (int i = threadIdx.x; i < d * d3; i += 256) {
    Wi[i] = __ldg(a.W_ih + i);
    Wh[i] = __ldg(a.W_hh + i);
  }
  const int lane = threadIdx.x & 31;
  const int c = lane % DP;
  const int grp = threadIdx.x / DP;
  const uint32_t gm = grp_mask<DP>(lane);
  const bool on = c < d;
  const int cc = on ? c : 0;
  const float bir = a.b_ih[cc], biz = a.b_ih[d + cc], bin = a.b_ih[2 * d + cc];
  const float bhr = a.b_hh[cc], bhz = a.b_hh[d + cc], bhn = a.b_hh[2 * d + cc];
  const int chunk = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = min(rows, (int)blockIdx.x * chunk), r1 = min(rows, r0 + chunk);
  unsigned nbar = 0;
  load_bnv<DP>(a, -1, bnv);   // also orders the weight staging before the first use
  int pkind = 0;
  for (int t = 0; t < a.T; ++t) {
    const BNDesc bn = a.bn[t];
    const float* table = a.table[t];
    const bool same_table = t > 0 && a.table[t] == a.table[t - 1];
    float* Mt = a.M + (size_t)t * rows * d;
    float* gt = a.gates + (size_t)t * rows * 4 * d;
    float* Gt = a.G + (size_t)t * rows * d;
    const float* Gp = t > 0 ? a.G + (size_t)(t - 1) * rows * d : nullptr;
    const float* Mp = t > 0 ? a.M + (size_t)(t - 1) * rows * d : nullptr;
    float lsum = 0.f, lcnt = 0.f;
    for (int base = r0; base < r1; base += GPB) {
      const int row = base + grp;
      if (row >= r1) continue;   // group-uniform
      const float mu = __ldg(a.mask + row);
      float mv;
      if (same_table) {
        mv = on ? Mp[(size_t)row * d + c] : 0.f;
      } else {
        mv = message_row<DP>(a, table, row, c, gm);
        if (!on) mv = 0.f;
      }
      float hv = 0.f, xh;
      if (on) {
        if (t == 0) hv = __ldg(a.h_init + (size_t)row * d + c);
        else hv = bn_out<DP>(pkind, Gp[(size_t)row * d + c], mu, bnv, c, &xh);
      }
      float ir = bir, iz = biz, in_ = bin, hr = bhr, hz = bhz, hn = bhn;
#pragma unroll 4
      for (int l = 0; l < d; ++l) {
        const float ml = __shfl_sync(gm, mv, l, DP);
        const float hl = __shfl_sync(gm, hv, l, DP);
        const float* wi = Wi + l * d3 + cc;
        const float* wh = Wh + l * d3 + cc;
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
        const float g = ((1.f - z) * n + z * hv) * mu;
        Gt[(size_t)row * d + c] = g;
        Mt[(size_t)row * d + c] = mv;
        float* gs = gt + (size_t)row * 4 * d;
        gs[c] = sr;
        gs[d + c] = sz;
        gs[2 * d + c] = tn;
        gs[3 * d + c] = hn;
        lsum += g * mu;
      }
      lcnt += mu;
    }
    pkind = bn.kind;
    if (bn.kind == 0) {
      load_bnv<DP>(a, -1, bnv);
      continue;
    }
    if (bn.kind == 2 && !bn.training) {   // running statistics: no coupling between rows
      __syncthreads();
      if (threadIdx.x < DP) {
        const int q = threadIdx.x;
        const float rm = q < d ? bn.running_mean[q] : 0.f;
        const float rv = q < d ? bn.running_var[q] : 1.f;
        const float rstd = 1.f / (sqrtf(rv) + bn.eps);
        bnv[q] = rm;
        bnv[DP + q] = rstd;
        bnv[2 * DP + q] = (q < d && bn.gamma) ? bn.gamma[q] : 1.f;
        bnv[3 * DP + q] = (q < d && bn.beta) ? bn.beta[q] : 0.f;
        if (blockIdx.x == 0) {
          float* st = a.stats + (size_t)t * SS;
          st[q] = rm;
          st[32 + q] = rstd;
          st[64 + q] = rv;
        }
      }
      __syncthreads();
      continue;
    }
    // ---- batch statistics: per-CTA (n, sum, M2), one grid barrier, fixed-order combination -------------------------
    const float csum = block_colsum<DP>(lsum, red, grp, c);
    const float ccnt = block_colsum<DP>(lcnt, red, grp, c);
    const float cmean = ccnt > 0.f ? csum / ccnt : 0.f;
    float lm2 = 0.f;
    if (on)
      for (int base = r0; base < r1; base += GPB) {
        const int row = base + grp;
        if (row >= r1) continue;
        const float dl = (Gt[(size_t)row * d + c] - cmean) * __ldg(a.mask + row);
        lm2 = fmaf(dl, dl, lm2);
      }
    const float cm2 = block_colsum<DP>(lm2, red, grp, c);
    float* part = a.part + ((size_t)t * gridDim.x + blockIdx.x) * PS;
    if (grp == 0) {
      part[c] = csum;
      part[DP + c] = cm2;
      if (c == 0) part[2 * DP] = ccnt;
    }
    grid_barrier(a.bar, (++nbar) * gridDim.x);
    const float* pt = a.part + (size_t)t * gridDim.x * PS;
    const float tot = grid_colsum<DP>(pt, PS, c, red, grp, c);
    const float n = grid_colsum<DP>(pt, PS, 2 * DP, red, grp, c);
    const float mean = tot / n;
    float s2 = 0.f;
    for (int cta = grp; cta < (int)gridDim.x; cta += GPB) {
      const float nc = __ldcg(pt + (size_t)cta * PS + 2 * DP);
      if (nc > 0.f) {
        const float dm = __ldcg(pt + (size_t)cta * PS + c) / nc - mean;
        s2 += __ldcg(pt + (size_t)cta * PS + DP + c) + nc * dm * dm;
      }
    }
    const float m2 = block_colsum<DP>(s2, red, grp, c);
    const float var = m2 / n;
    const float rstd = bn.kind == 1 ? 1.f / sqrtf(var + bn.eps) : 1.f / (sqrtf(var) + bn.eps);
    __syncthreads();
    if (grp == 0) {
      bnv[c] = mean;
      bnv[DP + c] = rstd;
      bnv[2 * DP + c] = (bn.kind == 2 && on && bn.gamma) ? bn.gamma[c] : 1.f;
      bnv[3 * DP + c] = (bn.kind == 2 && on && bn.beta) ? bn.beta[c] : 0.f;
      if (blockIdx.x == 0) {
        float* st = a.stats + (size_t)t * SS;
        st[c] = mean;
        st[32 + c] = rstd;
        st[64 + c] = var;
        if (c == 0) st[96] = n;
        if (bn.kind == 2 && on && bn.running_mean) {   // mask_batch_norm.py:30-33 (biased variance)
          bn.running_mean[c] = (1.f - bn.momentum) * bn.running_mean[c] + bn.momentum * mean;
          bn.running_var[c] = (1.f - bn.momentum) * bn.running_var[c] + bn.momentum * var;
        }
      }
    }
    __syncthreads();
  }
  // ---- output of the last step -----------------------------------------------------------------------------------
  {
    const float* Gl = a.G + (size_t)(a.T - 1) * rows * d;
    for (int base = r0; base < r1; base += GPB) {
      const int row = base + grp;
      if (row >= r1 || !on) continue;
      float xh;
      a.out[(size_t)row * d + c] = bn_out<DP>(pkind, Gl[(size_t)row * d + c], __ldg(a.mask + row), bnv, c, &xh);
    }
  }
  grid_exit(a.bar);
}

// ===================================================================================================================
// backward
// ===================================================================================================================
template <int DP>
__global__ void __launch_bounds__(256, 2) k_chain_bwd(ChainB b) {
  const Chain& a = b.f;
  constexpr int GPB = 256 / DP;
  constexpr int TR = GPB;
  constexpr int PS = 2 * DP + 2;
  constexpr int NACC = (6 * DP * DP + 255) / 256;
  extern __shared__ __align__(16) float sm[];
  const int d = a.d, d3 = 3 * d, rows = a.rows;
  const int ldt = d + 1;
  float* WiT = sm;                       // [3d][d+1]
  float* WhT = WiT + d3 * ldt;           // [3d][d+1]
  float* Ms = WhT + d3 * ldt;            // [TR][d]
  float* Hs = Ms + TR * d;               // [TR][d]
  float* Gi = Hs + TR * d;               // [TR][3d]
  float* Gh = Gi + TR * d3;              // [TR][3d]
  float* red = Gh + TR * d3;             // [GPB][DP+1]
  float* bnv = red + GPB * (DP + 1);     // BN of step t-1 (produces h_t):  mean | rstd | gamma | beta
  float* bnc = bnv + 4 * DP;             // BN of step t (being differentiated): mean | rstd | gamma | beta
  float* sv = bnc + 4 * DP;              // S1[DP] | S2[DP] of BN_t
  for (int i = threadIdx.x; i < d * d3; i += 256) {
    const int l = i / d3, g = i - l * d3;
    WiT[g * ldt + l] = __ldg(a.W_ih + i);
    WhT[g * ldt + l] = __ldg(a.W_hh + i);
  }
  const int lane = threadIdx.x & 31;
  const int c = lane % DP;
  const int grp = threadIdx.x / DP;
  const bool on = c < d;
  const int nW = d * d3;
  float acc[NACC];
  int pk[NACC];
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
  float accb = 0.f;
  const int chunk = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = min(rows, (int)blockIdx.x * chunk), r1 = min(rows, r0 + chunk);
  unsigned nbar = 0;
  float* bpart = a.part;   // [T][grid][PS]: S1 | S2 partials of the step's batch norm (the forward's slots, re-used)

  // prologue: the two column sums of the LAST batch norm, from the incoming gradient
  int T1 = a.T - 1;
  load_bnv<DP>(a, T1, bnc);
  if (a.bn[T1].kind) {
    const float* Gl = a.G + (size_t)T1 * rows * d;
    float s1 = 0.f, s2 = 0.f;
    if (on)
      for (int base = r0; base < r1; base += GPB) {
        const int row = base + grp;
        if (row >= r1) continue;
        const float mu = __ldg(a.mask + row);
        float xh;
        bn_out<DP>(a.bn[T1].kind, Gl[(size_t)row * d + c], mu, bnc, c, &xh);
        const float dy = __ldg(b.dout + (size_t)row * d + c);
        s1 = fmaf(dy, xh, s1);
        s2 = fmaf(dy, mu, s2);
      }
    const float c1 = block_colsum<DP>(s1, red, grp, c);
    const float c2 = block_colsum<DP>(s2, red, grp, c);
    float* part = bpart + ((size_t)T1 * gridDim.x + blockIdx.x) * PS;
    if (grp == 0) {
      part[c] = c1;
      part[DP + c] = c2;
    }
    grid_barrier(a.bar, (++nbar) * gridDim.x);
  }
  for (int t = T1; t >= 0; --t) {
    const BNDesc bn = a.bn[t];
    const int pkind = t > 0 ? a.bn[t - 1].kind : 0;
    // statistics of BN_t are in bnc; combine its column sums
    float Mn = 1.f;
    if (bn.kind) {
      const float* pt = bpart + (size_t)t * gridDim.x * PS;
      const float S1 = grid_colsum<DP>(pt, PS, c, red, grp, c);
      const float S2 = grid_colsum<DP>(pt, PS, DP + c, red, grp, c);
      Mn = __ldcg(a.stats + (size_t)t * SS + 96);
      __syncthreads();
      if (grp == 0) {
        sv[c] = S1;
        sv[DP + c] = S2;
        if (blockIdx.x == 0 && bn.kind == 2 && on) {
          if (b.dgamma[t]) b.dgamma[t][c] = S1;
          if (b.dbeta[t]) b.dbeta[t][c] = S2;
        }
      }
      __syncthreads();
    }
    load_bnv<DP>(a, t - 1, bnv);   // the batch norm that produced h_t
    const float* Gt = a.G + (size_t)t * rows * d;
    const float* gt = a.gates + (size_t)t * rows * 4 * d;
    const float* Mt = a.M + (size_t)t * rows * d;
    const float* Gp = t > 0 ? a.G + (size_t)(t - 1) * rows * d : nullptr;
    float* dMt = b.dM + (size_t)t * rows * d;
    const float* dyin = t == T1 ? b.dout : b.dY;
    float s1 = 0.f, s2 = 0.f;
    for (int base = r0; base < r1; base += TR) {
      const int row = base + grp;
      const bool live = row < r1;
      __syncthreads();   // previous tile's accumulation is done with the staging buffers
      float dar = 0.f, daz = 0.f, dan = 0.f, dnh = 0.f, dhd = 0.f, mv = 0.f, hv = 0.f, xhp = 0.f, mu = 0.f;
      if (live && on) {
        mu = __ldg(a.mask + row);
        const float dy = dyin[(size_t)row * d + c];
        // ---- batch norm backward: gradient w.r.t. the GRU output g_t --------------------------------------------
        float dg = dy;
        if (bn.kind == 1) {
          float xh;
          bn_out<DP>(1, Gt[(size_t)row * d + c], mu, bnc, c, &xh);
          dg = mu * bnc[DP + c] * (dy - (xh * sv[c] + sv[DP + c]) / Mn);
        } else if (bn.kind == 2) {
          const float ga = bnc[2 * DP + c], r = bnc[DP + c];
          if (bn.training) {
            float xh;
            bn_out<DP>(2, Gt[(size_t)row * d + c], mu, bnc, c, &xh);
            const float s = 1.f / r - bn.eps;   // sqrt(var)
            dg = mu * ga * (r * (dy - sv[DP + c] / Mn) - sv[c] * xh / (s * Mn));
          } else {
            dg = mu * ga * r * dy;
          }
        }
        // ---- GRU backward (gru_update.py:26-35) -------------------------------------------------------------------
        const float* g = gt + (size_t)row * 4 * d;
        const float sr = g[c], sz = g[d + c], tn = g[2 * d + c], nh = g[3 * d + c];
        mv = Mt[(size_t)row * d + c];
        if (t == 0) hv = __ldg(a.h_init + (size_t)row * d + c);
        else hv = bn_out<DP>(pkind, Gp[(size_t)row * d + c], mu, bnv, c, &xhp);
        const float r = sr * mu, z = sz * mu, n = tn * mu;
        const float go = dg * mu;
        const float dn = go * (1.f - z);
        const float dz = go * (hv - n);
        dan = dn * mu * (1.f - tn * tn);
        const float dr = dan * nh;
        dnh = dan * r;
        dar = dr * mu * sr * (1.f - sr);
        daz = dz * mu * sz * (1.f - sz);
        dhd = go * z;
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
        dMt[(size_t)row * d + c] = am;
        if (t > 0) {
          b.dY[(size_t)row * d + c] = ah;       // gradient w.r.t. h_t = output of BN_{t-1}
          s1 = fmaf(ah, xhp, s1);
          s2 = fmaf(ah, mu, s2);
        } else if (b.dh_init) {
          b.dh_init[(size_t)row * d + c] = ah;
        }
      }
#pragma unroll
      for (int q = 0; q < NACC; ++q) {
        if (pk[q] >= 0) {
          const float* xs = Ms + (pk[q] & 0xffff);
          const float* gs = Gi + (pk[q] >> 16);
          float v = acc[q];
#pragma unroll 8
          for (int r = 0; r < TR; ++r) v = fmaf(xs[r * d], gs[r * d3], v);
          acc[q] = v;
        }
      }
      if (threadIdx.x < 2 * d3) {
        const int which = threadIdx.x >= d3;
        const int g = threadIdx.x - which * d3;
        const float* gs = which ? Gh : Gi;
        float v = accb;
        for (int r = 0; r < TR; ++r) v += gs[r * d3 + g];
        accb = v;
      }
    }
    if (t > 0) {
      // bnv (statistics of BN_{t-1}) becomes the batch norm being differentiated
      __syncthreads();
      if (threadIdx.x < 4 * DP) bnc[threadIdx.x] = bnv[threadIdx.x];
      if (pkind) {
        const float c1 = block_colsum<DP>(s1, red, grp, c);
        const float c2 = block_colsum<DP>(s2, red, grp, c);
        float* part = bpart + ((size_t)(t - 1) * gridDim.x + blockIdx.x) * PS;
        if (grp == 0) {
          part[c] = c1;
          part[DP + c] = c2;
        }
        grid_barrier(a.bar, (++nbar) * gridDim.x);
      } else {
        __syncthreads();
      }
    }
  }
  // ---- the shared GRU cell's parameter gradients: per-CTA partials, reduced once in CTA order ---------------------
  const int total = 2 * nW + 2 * d3;
  float* gp = b.gpart + (size_t)blockIdx.x * total;
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    const int e = threadIdx.x + q * 256;
    if (e < 2 * nW) gp[e] = acc[q];
  }
  if (threadIdx.x < 2 * d3) gp[2 * nW + threadIdx.x] = accb;
  grid_barrier(a.bar, (++nbar) * gridDim.x);
  // element e is owned by CTA e % grid: 8 interleaved slices of the CTA list, combined in a fixed order
  {
    const int el = threadIdx.x >> 3, sl = threadIdx.x & 7;   // 32 elements per pass, 8 slices each
    for (int e0 = blockIdx.x * 32; e0 < total; e0 += gridDim.x * 32) {
      const int e = e0 + el;
      float s = 0.f;
      if (e < total)
        for (int p = sl; p < (int)gridDim.x; p += 8) s += __ldcg(b.gpart + (size_t)p * total + e);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (sl == 0 && e < total) {
        if (e < nW) b.dW_ih[e] = s;
        else if (e < 2 * nW) b.dW_hh[e - nW] = s;
        else if (e < 2 * nW + d3) b.db_ih[e - 2 * nW] = s;
        else b.db_hh[e - 2 * nW - d3] = s;
      }
    }
  }
  grid_exit(a.bar);
}

// co-resident grid (blocks_per_sm x SMs at most) whose CTAs each own a whole number of row tiles (256/DP rows)
int chain_grid(int rows, int DP, int blocks_per_sm) {
  const int per = 256 / DP;
  int tiles = ceil_div(rows, per);
  int cap = mpnn_num_sms() * blocks_per_sm;
  if (tiles < 1) tiles = 1;
  if (tiles <= cap) return tiles;
  const int tiles_per_cta = ceil_div(tiles, cap);
  return ceil_div(tiles, tiles_per_cta);
}

constexpr int CHAIN_BLOCKS_PER_SM = 2;

size_t fwd_smem(int d, int DP) {
  return (size_t)(2 * d * 3 * d + (256 / DP) * (DP + 1) + 4 * DP) * sizeof(float);
}
size_t bwd_smem(int d, int DP) {
  const int TR = 256 / DP;
  return (size_t)(2 * 3 * d * (d + 1) + 2 * TR * d + 2 * TR * 3 * d + TR * (DP + 1) + 10 * DP) * sizeof(float);
}

bool fill_chain(Chain* a, const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, int ecap,
                int zero_type, const float* H0, const float* h_init, const float* mask, const float* const* tables,
                int T, const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const int* bn_kind,
                const int* bn_training, const float* bn_eps, const float* bn_momentum, float* const* bn_ptrs,
                long long rows, int d, float* saved, float* out, void* workspace, int grid, int DP) {
  if (T < 1 || T > MAXT) return false;
  a->row_ptr = row_ptr;
  a->edge_src = edge_src;
  a->uid = uid;
  a->alpha = alpha;
  a->ecap = ecap;
  a->zero_type = zero_type;
  a->H0 = H0;
  a->h_init = h_init;
  a->mask = mask;
  a->W_ih = W_ih;
  a->W_hh = W_hh;
  a->b_ih = b_ih;
  a->b_hh = b_hh;
  a->T = T;
  a->rows = (int)rows;
  a->d = d;
  for (int t = 0; t < MAXT; ++t) {
    a->table[t] = t < T ? tables[t] : nullptr;
    BNDesc& b = a->bn[t];
    b.kind = t < T ? bn_kind[t] : 0;
    b.training = t < T ? bn_training[t] : 0;
    b.eps = t < T ? bn_eps[t] : 0.f;
    b.momentum = t < T ? bn_momentum[t] : 0.f;
    b.gamma = (t < T && bn_ptrs) ? bn_ptrs[4 * t + 0] : nullptr;
    b.beta = (t < T && bn_ptrs) ? bn_ptrs[4 * t + 1] : nullptr;
    b.running_mean = (t < T && bn_ptrs) ? bn_ptrs[4 * t + 2] : nullptr;
    b.running_var = (t < T && bn_ptrs) ? bn_ptrs[4 * t + 3] : nullptr;
    if (b.kind == 2 && !b.training && !(b.running_mean && b.running_var)) return false;
  }
  a->M = saved;
  a->gates = a->M + (size_t)T * rows * d;
  a->G = a->gates + (size_t)T * rows * 4 * d;
  a->stats = a->G + (size_t)T * rows * d;
  a->out = out;
  a->bar = (unsigned*)workspace;
  a->part = (float*)((char*)workspace + 256);
  (void)grid;
  (void)DP;
  return true;
}

}  // namespace

extern "C" {

int mpnn_chain_supported(int d, int T) {
  return d >= 1 && d <= 32 && T >= 1 && T <= MAXT ? 1 : 0;
}

// floats saved by the forward for the backward: messages, gates, GRU outputs, batch statistics
long long mpnn_chain_saved_floats(long long rows, int d, int T) {
  return (long long)T * rows * d * 6 + (long long)T * SS;
}

// the first 256 bytes (barrier counters) must be ZERO on entry and are zero again on exit; the rest is scratch
size_t mpnn_chain_workspace_bytes(long long rows, int d, int T) {
  const int DP = pow2_at_least(d, 8);
  const int grid = chain_grid((int)rows, DP, CHAIN_BLOCKS_PER_SM);
  size_t part = (size_t)T * grid * (2 * DP + 2) * sizeof(float);
  size_t gpart = (size_t)grid * (2 * d * 3 * d + 6 * d) * sizeof(float);
  size_t dy = (size_t)rows * d * sizeof(float);
  return 256 + align_up(part, 256) + align_up(gpart, 256) + align_up(dy, 256);
}

// bn_kind/bn_training/bn_eps/bn_momentum: host arrays [T]; bn_ptrs: host array [4T] of device pointers
// (gamma, beta, running_mean, running_var per step; NULL entries allowed); tables: host array [T] of device pointers.
int mpnn_chain_fwd(const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, int ecap, int zero_type,
                   const float* H0, const float* h_init, const float* mask, const float* const* tables, int T,
                   const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const int* bn_kind,
                   const int* bn_training, const float* bn_eps, const float* bn_momentum, float* const* bn_ptrs,
                   long long rows, int d, float* saved, float* out, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && rows < (1ll << 31) && mpnn_chain_supported(d, T), MPNN_ERR_UNSUPPORTED,
               "chain_fwd: unsupported dims (rows %lld, d %d, T %d)", rows, d, T);
  MPNN_REQUIRE(workspace_bytes >= mpnn_chain_workspace_bytes(rows, d, T), MPNN_ERR_WORKSPACE, "chain_fwd: workspace");
  const int DP = pow2_at_least(d, 8);
  const int grid = chain_grid((int)rows, DP, CHAIN_BLOCKS_PER_SM);
  Chain a;
  MPNN_REQUIRE(fill_chain(&a, row_ptr, edge_src, uid, alpha, ecap, zero_type, H0, h_init, mask, tables, T, W_ih, W_hh,
                          b_ih, b_hh, bn_kind, bn_training, bn_eps, bn_momentum, bn_ptrs, rows, d, saved, out, workspace,
                          grid, DP),
               MPNN_ERR_ARG, "chain_fwd: bad step description");
  const size_t smem = fwd_smem(d, DP);
  void* args[] = {&a};
  cudaError_t e;
  switch (DP) {
    case 8: e = cudaLaunchCooperativeKernel((void*)k_chain_fwd<8>, dim3(grid), dim3(256), args, smem, stream); break;
    case 16: e = cudaLaunchCooperativeKernel((void*)k_chain_fwd<16>, dim3(grid), dim3(256), args, smem, stream); break;
    default: e = cudaLaunchCooperativeKernel((void*)k_chain_fwd<32>, dim3(grid), dim3(256), args, smem, stream); break;
  }
  MPNN_REQUIRE(e == cudaSuccess, MPNN_ERR_CUDA, "chain_fwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return MPNN_OK;
}

// dM [T][rows][d] (message gradients per step), dh_init [rows][d] or NULL, the GRU cell's gradients, and per step the
// MaskBatchNorm1d gradients bn_grads[2t] (gamma) / bn_grads[2t+1] (beta) where non-NULL.  `saved`/`workspace`: the
// forward's buffers (the workspace's barrier words are zero between launches).
int mpnn_chain_bwd(const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, int ecap, int zero_type,
                   const float* H0, const float* h_init, const float* mask, const float* const* tables, int T,
                   const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const int* bn_kind,
                   const int* bn_training, const float* bn_eps, const float* bn_momentum, float* const* bn_ptrs,
                   long long rows, int d, float* saved, const float* dout, float* dM, float* dh_init, float* dW_ih,
                   float* dW_hh, float* db_ih, float* db_hh, float* const* bn_grads, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && rows < (1ll << 31) && mpnn_chain_supported(d, T), MPNN_ERR_UNSUPPORTED,
               "chain_bwd: unsupported dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_chain_workspace_bytes(rows, d, T), MPNN_ERR_WORKSPACE, "chain_bwd: workspace");
  const int DP = pow2_at_least(d, 8);
  const int grid = chain_grid((int)rows, DP, CHAIN_BLOCKS_PER_SM);
  ChainB b;
  MPNN_REQUIRE(fill_chain(&b.f, row_ptr, edge_src, uid, alpha, ecap, zero_type, H0, h_init, mask, tables, T, W_ih, W_hh,
                          b_ih, b_hh, bn_kind, bn_training, bn_eps, bn_momentum, bn_ptrs, rows, d, saved, nullptr,
                          workspace, grid, DP),
               MPNN_ERR_ARG, "chain_bwd: bad step description");
  b.dout = dout;
  b.dM = dM;
  b.dh_init = dh_init;
  const size_t part = align_up((size_t)T * grid * (2 * DP + 2) * sizeof(float), 256);
  const size_t gpart = align_up((size_t)grid * (2 * d * 3 * d + 6 * d) * sizeof(float), 256);
  b.gpart = (float*)((char*)workspace + 256 + part);
  b.dY = (float*)((char*)workspace + 256 + part + gpart);
  b.dW_ih = dW_ih;
  b.dW_hh = dW_hh;
  b.db_ih = db_ih;
  b.db_hh = db_hh;
  for (int t = 0; t < MAXT; ++t) {
    b.dgamma[t] = (t < T && bn_grads) ? bn_grads[2 * t] : nullptr;
    b.dbeta[t] = (t < T && bn_grads) ? bn_grads[2 * t + 1] : nullptr;
  }
  const size_t smem = bwd_smem(d, DP);
  void* args[] = {&b};
  cudaError_t e;
  switch (DP) {
    case 8: e = cudaLaunchCooperativeKernel((void*)k_chain_bwd<8>, dim3(grid), dim3(256), args, smem, stream); break;
    case 16: e = cudaLaunchCooperativeKernel((void*)k_chain_bwd<16>, dim3(grid), dim3(256), args, smem, stream); break;
    default: e = cudaLaunchCooperativeKernel((void*)k_chain_bwd<32>, dim3(grid), dim3(256), args, smem, stream); break;
  }
  MPNN_REQUIRE(e == cudaSuccess, MPNN_ERR_CUDA, "chain_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return MPNN_OK;
}

}  // extern "C"
