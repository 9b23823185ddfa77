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
//     h' = BN(g)                                               mask_batch_norm.py:9-15 / :20-38
// The batch statistics are the only coupling between rows: per-CTA (n, sum, M2) partials and ONE grid-wide reduction
// barrier per step.  Messages read the INPUT features at every step (normed_basic_model.py:58), so the gather of step
// t+1 has no dependency on the recurrence: it is computed while the barrier of step t is in flight.  Rows are owned by
// a fixed CTA for the whole loop, so the recurrent state never crosses CTAs.
// The backward kernel walks the steps in reverse with the same ownership: batch-norm backward (one reduction barrier per
// step for its two column sums; the step's operands are loaded while it is in flight), GRU backward, message gradients
// dM_t written for the table / sender gradients (csrc/typed.cu), the shared GRU cell's weight gradients accumulated in
// registers over all rows AND steps and reduced once in a fixed order behind a last barrier.  No float atomics: results
// are bit-reproducible for a fixed grid.
//
// The kernel is latency-bound at the reference's batch sizes (7 424 rows x 16 floats at BASELINE config 2): what it is
// built around is the number of DEPENDENT L2 round trips (~0.6 us each) per step, measured with %globaltimer stamps
// (tools/chain_phases.py) -- see the comments at grid_sums / the mailbox.
//
// Masks are the reference's 0/1 masks (pre_process/data_loader.py:18-21).
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int NT = 512;    // threads per CTA; ONE CTA per SM (half as many partials per reduction as 2 x 256, and 32 row
                           // groups of 16 lanes: the 28 real rows a CTA owns at BASELINE config 2 are a single tile)
constexpr int MAXT = 8;
constexpr int SS = 192;   // floats per step in the saved statistics: mean[32] | rstd[32] | var[32] | n .. | spare
constexpr int MAILW = 64; // mailbox words per step

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
  const int* real_list;  // optional: the rows with mask != 0 in increasing order (mpnn_real_rows), else null
  const int* real_count; // [1] their number
  const float* table[MAXT];
  const float* W_ih;
  const float* W_hh;
  const float* b_ih;
  const float* b_hh;
  BNDesc bn[MAXT];
  int T, rows, d;
  float* M;       // [T][rows][d]   (steps that share a table share the slot of the first of them)
  float* gates;   // [T][rows][4d]
  float* G;       // [T][rows][d]   GRU outputs before the batch norm
  float* stats;   // [T][SS]
  float* out;     // [rows][d]
  float* part;    // [T][grid][2*DP+4]
  unsigned* bar;  // [64]: arrivals, exits, ..., [32] release flag; zero on entry, zero on exit
  unsigned* mail; // [MAXT][MAILW] mailbox of the reduction barriers; zero (= empty) on entry and on exit
  long long* dbg; // optional: CTA 0 records %globaltimer at its phase boundaries (profiling aid), else null
};

struct ChainB {
  Chain f;
  const float* dout;   // [rows, d], row stride dout_ld floats (a column slice of the caller's wider gradient is fine)
  long long dout_ld;
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

__device__ __forceinline__ void dbg_stamp(const Chain& a, int& slot) {
  if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && slot < 40) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[slot] = t;
  }
  ++slot;
}

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
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- grid-wide reduction barrier ---------------------------------------------------------------------------------------
// All CTAs of the grid are co-resident (cooperative launch).  Every CTA publishes its partials and arrives (one atomic);
// the LAST CTA to arrive combines the partials in CTA order (the result does not depend on who is last) and posts the
// few result words in a MAILBOX; the others poll their own mailbox word.  A mailbox word holds ~bits(value), 0 = empty,
// so a word is its own ready flag: no separate flag read, no fence on the waiting side.  Dependent L2 round trips per
// barrier: arrive, load partials (one batch), post, poll = 4.  (Version 1 -- every CTA spinning on the arrival counter and
// then summing all partials itself with a load-add loop -- took 15 us per barrier on 232 CTAs.)
// Roles: with more than one CTA the LAST CTA of the grid is the REDUCER: it owns no rows, waits for the workers'
// arrivals and does every combination.  (When the last worker to arrive did the combination, its own next step started
// ~2 us late, it was last again at the next barrier, and the whole grid ran at its pace.)  Workers arrive with a
// fire-and-forget reduction (no round trip) and return true only when they are their own reducer (grid of one CTA).
__device__ __forceinline__ int n_workers() { return gridDim.x > 1 ? (int)gridDim.x - 1 : 1; }
__device__ __forceinline__ bool is_reducer() { return gridDim.x > 1 && blockIdx.x == gridDim.x - 1; }

__device__ __forceinline__ bool grid_arrive(unsigned* bar, unsigned epoch, int* s_last) {
  (void)s_last;
  __syncthreads();
  if (is_reducer()) {
    if (threadIdx.x == 0) {
      const unsigned target = epoch * (unsigned)n_workers();
      while (ld_acquire(bar) < target) __nanosleep(20);
      __threadfence();
    }
    __syncthreads();
    return true;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
  }
  return gridDim.x == 1;   // a single CTA: its partials are in place after the barrier above
}

__device__ __forceinline__ void mail_post(unsigned* slot, float v) {
  unsigned b = ~__float_as_uint(v);
  if (b == 0u) b = ~0x7fc00000u;   // the one NaN pattern that would read as "empty"
  st_relaxed(slot, b);
}

__device__ __forceinline__ float mail_wait(const unsigned* slot) {
  unsigned b;
  while ((b = ld_relaxed(slot)) == 0u) __nanosleep(20);
  return __uint_as_float(~b);
}

// plain barrier (flag = number of the last completed barrier); used once, in front of the final gradient reduction
__device__ __forceinline__ void grid_barrier(unsigned* bar, int* s_last) {   // all CTAs, own counter bar[2]
  (void)s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar + 2, 1u);
    while (ld_acquire(bar + 2) < gridDim.x) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}

// the last CTA to leave resets the counters and the mailboxes for the next launch
__device__ __forceinline__ void grid_exit(const Chain& a, int* s_last) {
  __syncthreads();
  if (threadIdx.x == 0) *s_last = atomicAdd(a.bar + 1, 1u) == gridDim.x - 1;
  __syncthreads();
  if (*s_last) {
    for (int i = threadIdx.x; i < MAXT * MAILW; i += NT) a.mail[i] = 0u;
    if (threadIdx.x < 64) a.bar[threadIdx.x] = 0u;
  }
}

// fixed-order sums over the groups of a CTA of per-(group, lane) values; every thread returns the column totals.
// Groups of one warp are combined with a shuffle tree, the NT/32 warp totals through shared memory (red: [NQ][NT/32][DP]).
template <int DP, int NQ>
__device__ __forceinline__ void block_colsums(float (&v)[NQ], float* red, int grp, int c) {
  constexpr int NW = NT / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  (void)grp;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int o = DP; o < 32; o <<= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  __syncthreads();
  if (lane < DP) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) red[(q * NW + warp) * DP + lane] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[(q * NW + w) * DP + c];
    v[q] = s;
  }
}

// The last CTA copies the grid's partials of one reduction (contiguous: grid x PS floats) into shared memory with
// coalesced 16-byte loads, ALL issued before the first store (<= 6 per thread), and combines them from there in CTA
// order.  (Per-lane scalar loads straight from L2 -- 60 per thread, 480 load instructions per SM -- took 2.9 us.)
constexpr int MAXGRID = 160;   // >= 1 CTA x 148 SMs

template <int DP>
__device__ __forceinline__ void stage_partials(const float* pt, float* pbuf) {
  constexpr int PS = 2 * DP + 4;
  constexpr int NV = (MAXGRID * PS / 4 + NT - 1) / NT;
  const int n4 = n_workers() * PS / 4;
  const float4* src = reinterpret_cast<const float4*>(pt);
  float4* dst = reinterpret_cast<float4*>(pbuf);
  float4 v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * NT;
    v[k] = i < n4 ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * NT;
    if (i < n4) dst[i] = v[k];
  }
  __syncthreads();
}

// fixed-order sums over the CTAs of two columns of the staged partials
template <int DP>
__device__ __forceinline__ void staged_sums(const float* pbuf, int i0, int i1, float (&out)[2], float* red, int grp, int c) {
  constexpr int GPB = NT / DP;
  constexpr int PS = 2 * DP + 4;
  out[0] = out[1] = 0.f;
  for (int cta = grp; cta < n_workers(); cta += GPB) {
    out[0] += pbuf[cta * PS + i0];
    out[1] += pbuf[cta * PS + i1];
  }
  block_colsums<DP, 2>(out, red, grp, c);
}

// combination of the per-CTA (sum, M2, n) partials of one batch norm (Chan et al.): mean, M2 total, n
template <int DP>
__device__ __forceinline__ void combine_stats(const float* pt, float* pbuf, float* red, int grp, int c, float* mean_out,
                                              float* m2_out, float* n_out) {
  constexpr int GPB = NT / DP;
  constexpr int PS = 2 * DP + 4;
  stage_partials<DP>(pt, pbuf);
  float tn[2];
  staged_sums<DP>(pbuf, c, 2 * DP, tn, red, grp, c);
  const float mean = tn[0] / tn[1];
  float s2[1] = {0.f};
  for (int cta = grp; cta < n_workers(); cta += GPB) {
    const float nc = pbuf[cta * PS + 2 * DP];
    if (nc > 0.f) {
      const float dm = pbuf[cta * PS + c] / nc - mean;
      s2[0] += pbuf[cta * PS + DP + c] + nc * dm * dm;
    }
  }
  block_colsums<DP, 1>(s2, red, grp, c);
  *mean_out = mean;
  *m2_out = s2[0];
  *n_out = tn[1];
}

// ---- message function + aggregation of one receiver row -------------------------------------------------------------
// Metadata of the first DP edges of a row (lane k holds edge eb + k) and the first sender state are loaded by the caller,
// for TWO rows at a time, before either row is contracted (three dependent global loads per row: row_ptr -> edge list ->
// sender state; issued back to back they cost one chain, not two).
struct RowMeta {
  int eb, ee, jm, um;
  float am, hj;
};

template <int DP>
__device__ __forceinline__ void meta_bounds(const Chain& a, int row, bool live, RowMeta& m) {
  m.eb = m.ee = 0;
  if (live) {
    m.eb = min(__ldg(a.row_ptr + row), a.ecap);
    m.ee = min(__ldg(a.row_ptr + row + 1), a.ecap);
  }
}
template <int DP>
__device__ __forceinline__ void meta_edges(const Chain& a, int k, RowMeta& m) {
  m.jm = m.um = 0;
  m.am = 1.f;
  if (k < m.ee - m.eb) {
    m.jm = __ldg(a.edge_src + m.eb + k);
    m.um = min(__ldg(a.uid + m.eb + k), a.zero_type);
    if (a.alpha) m.am = __ldg(a.alpha + m.eb + k);
  }
}
template <int DP>
__device__ __forceinline__ void meta_first(const Chain& a, int k, uint32_t gm, RowMeta& m) {
  const int j = __shfl_sync(gm, m.jm, 0, DP);
  m.hj = (m.ee > m.eb && k < a.d) ? __ldg(a.H0 + (size_t)j * a.d + k) : 0.f;
}

template <int DP>
__device__ __forceinline__ float message_row(const Chain& a, const float* __restrict__ table, RowMeta m, int k,
                                             uint32_t gm) {
  float acc = 0.f;
  int jm = m.jm, um = m.um;
  float am = m.am, hj = m.hj;
  for (int e0 = m.eb; e0 < m.ee; e0 += DP) {
    const int cnt = min(DP, m.ee - e0);
    if (e0 != m.eb) {   // further chunks of a row with more than DP edges
      jm = um = 0;
      am = 1.f;
      if (k < cnt) {
        jm = __ldg(a.edge_src + e0 + k);
        um = min(__ldg(a.uid + e0 + k), a.zero_type);
        if (a.alpha) am = __ldg(a.alpha + e0 + k);
      }
      const int j0 = __shfl_sync(gm, jm, 0, DP);
      hj = k < a.d ? __ldg(a.H0 + (size_t)j0 * a.d + k) : 0.f;
    }
    for (int t = 0; t < cnt; ++t) {
      const int u = __shfl_sync(gm, um, t, DP);
      const float al = __shfl_sync(gm, am, t, DP);
      const float hcur = hj;
      if (t + 1 < cnt) {
        const int j = __shfl_sync(gm, jm, t + 1, DP);
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

// messages of one step for the real rows of this CTA -> Mt (two rows of a group in flight at a time)
template <int DP>
__device__ __forceinline__ void messages_step(const Chain& a, const float* __restrict__ table, float* __restrict__ Mt,
                                              const int* rowlist, int nreal, int grp, int c, uint32_t gm) {
  constexpr int GPB = NT / DP;
  const int d = a.d;
  for (int base = 0; base < nreal; base += 2 * GPB) {
    const int iA = base + grp, iB = base + GPB + grp;
    const bool liveA = iA < nreal, liveB = iB < nreal;
    const int rowA = liveA ? rowlist[iA] : 0, rowB = liveB ? rowlist[iB] : 0;
    RowMeta mA, mB;
    meta_bounds<DP>(a, rowA, liveA, mA);
    meta_bounds<DP>(a, rowB, liveB, mB);
    meta_edges<DP>(a, c, mA);
    meta_edges<DP>(a, c, mB);
    meta_first<DP>(a, c, gm, mA);
    meta_first<DP>(a, c, gm, mB);
    if (liveA) {
      const float v = message_row<DP>(a, table, mA, c, gm);
      if (c < d) Mt[(size_t)rowA * d + c] = v;
    }
    if (liveB) {
      const float v = message_row<DP>(a, table, mB, c, gm);
      if (c < d) Mt[(size_t)rowB * d + c] = v;
    }
  }
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

// bnv <- (mean, rstd, gamma, beta) of step t's batch norm from the SAVED statistics (t < 0 or no batch norm: identity)
template <int DP>
__device__ __forceinline__ void load_bnv(const Chain& a, int t, float* bnv) {
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

__device__ __forceinline__ int m_slot(const Chain& a, int t) {   // steps that share a table share their messages
  while (t > 0 && a.table[t] == a.table[t - 1]) --t;
  return t;
}

// Each CTA keeps the list of the REAL rows (mask != 0) it owns in shared memory: padded rows (44 % of the rows at BASELINE
// config 2) cost nothing.  Padded rows are never read back: their outputs and gradients are exactly zero (the caller
// zero-fills them).  With the global list of real rows (mpnn_real_rows, small batches) every CTA gets the same number of
// them (+-1), i.e. one full tile per step at config 2; without it the rows are dealt round-robin (row = cta + k * grid),
// which spreads real and padded rows evenly for large batches.  Returns the number of rows owned.
__device__ __forceinline__ int build_rowlist(const Chain& a, int chunk, int* rowlist, int* s_cnt) {
  if (a.real_list) {
    const int total = is_reducer() ? 0 : __ldg(a.real_count);
    const int q = (total + n_workers() - 1) / n_workers();
    const int lo = min(total, (int)blockIdx.x * q), hi = min(total, lo + q);
    for (int i = threadIdx.x; i < hi - lo; i += NT) rowlist[i] = __ldg(a.real_list + lo + i);
    __syncthreads();
    return hi - lo;
  }
  if (threadIdx.x < 32) {
    int n = 0;
    for (int k0 = 0; k0 < chunk; k0 += 32) {
      const int k = k0 + threadIdx.x;
      const int row = blockIdx.x + k * n_workers();
      const bool real = !is_reducer() && k < chunk && row < a.rows && __ldg(a.mask + row) != 0.f;
      const unsigned bal = __ballot_sync(0xffffffffu, real);
      if (real) rowlist[n + __popc(bal & ((1u << threadIdx.x) - 1u))] = row;
      n += __popc(bal);
    }
    if (threadIdx.x == 0) *s_cnt = n;
  }
  __syncthreads();
  return *s_cnt;
}

// list of the rows with mask != 0 (one block of 1024 threads, rows <= SMALL_SCAN_MAX)
__global__ void __launch_bounds__(1024) k_real_rows(const float* __restrict__ mask, int rows, int* __restrict__ flags,
                                                    int* __restrict__ pos, int* __restrict__ list, int* __restrict__ count) {
  for (int i = threadIdx.x; i < rows; i += 1024) flags[i] = mask[i] != 0.f;
  __syncthreads();
  small_scan_block(flags, nullptr, rows, pos, nullptr);
  __syncthreads();
  for (int i = threadIdx.x; i < rows; i += 1024)
    if (flags[i]) list[pos[i]] = i;
  if (threadIdx.x == 0) count[0] = pos[rows];
}

// ===================================================================================================================
// forward
// ===================================================================================================================
template <int DP>
__global__ void __launch_bounds__(NT, 1) k_chain_fwd(Chain a) {
  constexpr int GPB = NT / DP;
  constexpr int PS = 2 * DP + 4;
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_last[1];
  const int d = a.d, d3 = 3 * d, rows = a.rows;
  float* pbuf = sm;                              // [grid][PS] staged partials (last arriver only)
  float4* Wi4 = reinterpret_cast<float4*>(sm + gridDim.x * PS);   // [d][DP]: (r, z, n) gate columns of input row l, column c
  float4* Wh4 = Wi4 + d * DP;
  float* red = reinterpret_cast<float*>(Wh4 + d * DP);   // [2][GPB][DP+1]
  float* bnv = red + 2 * (NT / 32) * DP;                 // [4][DP]
  int* rowlist = reinterpret_cast<int*>(bnv + 4 * DP);   // [chunk]
  for (int i = threadIdx.x; i < d * DP; i += NT) {
    const int l = i / DP, q = i - l * DP;
    float4 wi = make_float4(0.f, 0.f, 0.f, 0.f), wh = wi;
    if (q < d) {
      wi = make_float4(__ldg(a.W_ih + l * d3 + q), __ldg(a.W_ih + l * d3 + d + q), __ldg(a.W_ih + l * d3 + 2 * d + q), 0.f);
      wh = make_float4(__ldg(a.W_hh + l * d3 + q), __ldg(a.W_hh + l * d3 + d + q), __ldg(a.W_hh + l * d3 + 2 * d + q), 0.f);
    }
    Wi4[i] = wi;
    Wh4[i] = wh;
  }
  const int lane = threadIdx.x & 31;
  const int c = lane % DP;
  const int grp = threadIdx.x / DP;
  const uint32_t gm = grp_mask<DP>(lane);
  const bool on = c < d;
  const int cc = on ? c : 0;
  (void)d3;
  const float bir = a.b_ih[cc], biz = a.b_ih[d + cc], bin = a.b_ih[2 * d + cc];
  const float bhr = a.b_hh[cc], bhz = a.b_hh[d + cc], bhn = a.b_hh[2 * d + cc];
  const int chunk = (rows + n_workers() - 1) / n_workers();
  const int nreal = build_rowlist(a, chunk, rowlist, s_last);
  unsigned nbar = 0;
  int ds = 0;
  dbg_stamp(a, ds);
  messages_step<DP>(a, a.table[0], a.M, rowlist, nreal, grp, c, gm);
  load_bnv<DP>(a, -1, bnv);   // also orders the weight staging before the first use
  dbg_stamp(a, ds);
  int pkind = 0;
  for (int t = 0; t < a.T; ++t) {
    const BNDesc bn = a.bn[t];
    const float* Mt = a.M + (size_t)m_slot(a, t) * rows * d;
    float* gt = a.gates + (size_t)t * rows * 4 * d;
    float* Gt = a.G + (size_t)t * rows * d;
    const float* Hp = t > 0 ? a.G + (size_t)(t - 1) * rows * d : a.h_init;
    float lsum = 0.f, lcnt = 0.f;
    // ---- GRU over the rows of this CTA; the next row's operands are loaded before the current row is computed ----------
    float n_m = 0.f, n_h = 0.f, n_mu = 0.f;
    int n_row = 0;
    if (grp < nreal) {
      n_row = rowlist[grp];
      n_mu = __ldg(a.mask + n_row);
      if (on) {
        n_m = Mt[(size_t)n_row * d + c];
        n_h = Hp[(size_t)n_row * d + c];
      }
    }
    for (int base = 0; base < nreal; base += GPB) {
      const int idx = base + grp;
      const int row = n_row;
      const float mv = n_m, hraw = n_h, mu = n_mu;
      if (idx + GPB < nreal) {
        n_row = rowlist[idx + GPB];
        n_mu = __ldg(a.mask + n_row);
        if (on) {
          n_m = Mt[(size_t)n_row * d + c];
          n_h = Hp[(size_t)n_row * d + c];
        }
      }
      if (idx >= nreal) continue;   // group-uniform
      float hv = 0.f, xh;
      if (on) hv = t == 0 ? hraw : bn_out<DP>(pkind, hraw, mu, bnv, c, &xh);
      float ir = bir, iz = biz, in_ = bin, hr = bhr, hz = bhz, hn = bhn;
#pragma unroll 4
      for (int l = 0; l < d; ++l) {
        const float ml = __shfl_sync(gm, mv, l, DP);
        const float hl = __shfl_sync(gm, hv, l, DP);
        const float4 wi = Wi4[l * DP + c];
        const float4 wh = Wh4[l * DP + c];
        ir = fmaf(ml, wi.x, ir);
        iz = fmaf(ml, wi.y, iz);
        in_ = fmaf(ml, wi.z, in_);
        hr = fmaf(hl, wh.x, hr);
        hz = fmaf(hl, wh.y, hz);
        hn = fmaf(hl, wh.z, hn);
      }
      if (on) {
        const float sr = 1.f / (1.f + expf(-(ir + hr)));
        const float sz = 1.f / (1.f + expf(-(iz + hz)));
        const float r = sr * mu, z = sz * mu;
        const float tn = tanhf(in_ + r * hn);
        const float n = tn * mu;
        const float g = ((1.f - z) * n + z * hv) * mu;
        Gt[(size_t)row * d + c] = g;
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
    const bool more = t + 1 < a.T && a.table[t + 1] != a.table[t];
    float* Mn = a.M + (size_t)(t + 1) * rows * d;
    dbg_stamp(a, ds);
    if (bn.kind == 0 || (bn.kind == 2 && !bn.training)) {
      // no coupling between rows: identity, or running statistics (mask_batch_norm.py:26-28)
      __syncthreads();
      if (threadIdx.x < DP) {
        const int q = threadIdx.x;
        float mean = 0.f, rstd = 1.f, ga = 1.f, be = 0.f;
        if (bn.kind == 2) {
          mean = q < d ? bn.running_mean[q] : 0.f;
          const float rv = q < d ? bn.running_var[q] : 1.f;
          rstd = 1.f / (sqrtf(rv) + bn.eps);
          ga = (q < d && bn.gamma) ? bn.gamma[q] : 1.f;
          be = (q < d && bn.beta) ? bn.beta[q] : 0.f;
          if (blockIdx.x == 0) {
            float* st = a.stats + (size_t)t * SS;
            st[q] = mean;
            st[32 + q] = rstd;
            st[64 + q] = rv;
          }
        }
        bnv[q] = mean;
        bnv[DP + q] = rstd;
        bnv[2 * DP + q] = ga;
        bnv[3 * DP + q] = be;
      }
      __syncthreads();
      if (more) messages_step<DP>(a, a.table[t + 1], Mn, rowlist, nreal, grp, c, gm);
      dbg_stamp(a, ds);
      dbg_stamp(a, ds);
      dbg_stamp(a, ds);
      continue;
    }
    // ---- batch statistics: per-CTA (n, sum, M2), one reduction barrier ---------------------------------------------------
    float sc[2] = {lsum, lcnt};
    block_colsums<DP, 2>(sc, red, grp, c);
    const float cmean = sc[1] > 0.f ? sc[0] / sc[1] : 0.f;
    float lm2[1] = {0.f};
    if (on)
      for (int idx = grp; idx < nreal; idx += GPB) {
        const int row = rowlist[idx];
        const float dl = (Gt[(size_t)row * d + c] - cmean) * __ldg(a.mask + row);
        lm2[0] = fmaf(dl, dl, lm2[0]);
      }
    block_colsums<DP, 1>(lm2, red, grp, c);
    float* part = a.part + ((size_t)t * gridDim.x + blockIdx.x) * PS;
    if (grp == 0) {
      part[c] = sc[0];
      part[DP + c] = lm2[0];
      if (c == 0) part[2 * DP] = sc[1];
    }
    dbg_stamp(a, ds);
    ++nbar;
    unsigned* mail = a.mail + (size_t)t * MAILW;
    if (grid_arrive(a.bar, nbar, s_last)) {
      if (a.dbg && threadIdx.x == 0) {
        long long tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        a.dbg[40 + 3 * (nbar - 1)] = tt;
      }
      float mean, m2, n;
      combine_stats<DP>(a.part + (size_t)t * gridDim.x * PS, pbuf, red, grp, c, &mean, &m2, &n);
      if (a.dbg && threadIdx.x == 0) {
        long long tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        a.dbg[41 + 3 * (nbar - 1)] = tt;
        a.dbg[42 + 3 * (nbar - 1)] = blockIdx.x;
      }
      const float var = m2 / n;
      const float rstd = bn.kind == 1 ? 1.f / sqrtf(var + bn.eps) : 1.f / (sqrtf(var) + bn.eps);
      if (grp == 0) {
        mail_post(mail + c, mean);
        mail_post(mail + DP + c, rstd);
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
    dbg_stamp(a, ds);
    // the next step's messages do not depend on the statistics: computed while the reduction is in flight
    if (more) messages_step<DP>(a, a.table[t + 1], Mn, rowlist, nreal, grp, c, gm);
    dbg_stamp(a, ds);
    __syncthreads();
    if (threadIdx.x < 2 * DP) bnv[threadIdx.x] = mail_wait(mail + threadIdx.x);
    else if (threadIdx.x < 4 * DP) {
      const int q = threadIdx.x - 2 * DP;   // gamma[DP] | beta[DP]
      const int col = q % DP;
      float v = q < DP ? 1.f : 0.f;
      if (bn.kind == 2 && col < d) {
        if (q < DP && bn.gamma) v = bn.gamma[col];
        if (q >= DP && bn.beta) v = bn.beta[col];
      }
      bnv[threadIdx.x] = v;
    }
    __syncthreads();
    dbg_stamp(a, ds);
  }
  // ---- output of the last step -----------------------------------------------------------------------------------
  {
    const float* Gl = a.G + (size_t)(a.T - 1) * rows * d;
    for (int idx = grp; idx < nreal; idx += GPB) {   // padded rows stay zero (zero-filled by the caller)
      if (!on) continue;
      const int row = rowlist[idx];
      float xh;
      a.out[(size_t)row * d + c] = bn_out<DP>(pkind, Gl[(size_t)row * d + c], __ldg(a.mask + row), bnv, c, &xh);
    }
  }
  dbg_stamp(a, ds);
  grid_exit(a, s_last);
}

// ===================================================================================================================
// backward
// ===================================================================================================================
// this CTA's column sums (S1 = sum dY xh, S2 = sum dY mu) of step t's batch norm are published; the last CTA to arrive adds
// them up in CTA order and posts them in step t's mailbox (and writes the MaskBatchNorm1d parameter gradients)
template <int DP>
__device__ __forceinline__ void publish_bn_sums(const ChainB& b, int t, float c1, float c2, unsigned epoch, float* pbuf,
                                                float* red, int* s_last, int grp, int c) {
  const Chain& a = b.f;
  constexpr int PS = 2 * DP + 4;
  float* part = a.part + ((size_t)t * gridDim.x + blockIdx.x) * PS;
  if (grp == 0) {
    part[c] = c1;
    part[DP + c] = c2;
  }
  if (grid_arrive(a.bar, epoch, s_last)) {
    const float* pt = a.part + (size_t)t * gridDim.x * PS;
    float ss[2];
    stage_partials<DP>(pt, pbuf);
    staged_sums<DP>(pbuf, c, DP + c, ss, red, grp, c);
    if (grp == 0) {
      unsigned* mail = a.mail + (size_t)t * MAILW;
      mail_post(mail + c, ss[0]);
      mail_post(mail + DP + c, ss[1]);
      if (a.bn[t].kind == 2 && c < a.d) {
        if (b.dgamma[t]) b.dgamma[t][c] = ss[0];
        if (b.dbeta[t]) b.dbeta[t][c] = ss[1];
      }
    }
  }
}

struct BwdOps {   // operands of one row of one backward step (lane c)
  float mu, dy, g, sr, sz, tn, nh, mv, hraw;
};

template <int DP>
__device__ __forceinline__ void load_ops(const Chain& a, int t, const float* dyin, size_t dy_ld, const int* rowlist,
                                         int idx, int nreal, bool on, int c, BwdOps& o, int* row_out) {
  const bool live = idx < nreal && on;
  const int row = idx < nreal ? rowlist[idx] : 0;
  *row_out = row;
  o.mu = o.dy = o.g = o.sr = o.sz = o.tn = o.nh = o.mv = o.hraw = 0.f;
  if (!live) return;
  const int d = a.d;
  const size_t rows = a.rows;
  o.mu = __ldg(a.mask + row);
  o.dy = __ldcg(dyin + (size_t)row * dy_ld + c);
  if (a.bn[t].kind) o.g = __ldcg(a.G + ((size_t)t * rows + row) * d + c);
  const float* g = a.gates + ((size_t)t * rows + row) * 4 * d;
  o.sr = __ldcg(g + c);
  o.sz = __ldcg(g + d + c);
  o.tn = __ldcg(g + 2 * d + c);
  o.nh = __ldcg(g + 3 * d + c);
  o.mv = __ldcg(a.M + ((size_t)m_slot(a, t) * rows + row) * d + c);
  o.hraw = t == 0 ? __ldg(a.h_init + (size_t)row * d + c) : __ldcg(a.G + ((size_t)(t - 1) * rows + row) * d + c);
}

template <int DP>
__global__ void __launch_bounds__(NT, 1) k_chain_bwd(ChainB b) {
  const Chain& a = b.f;
  constexpr int GPB = NT / DP;
  constexpr int TR = GPB;
  constexpr int NACC = (6 * DP * DP + NT - 1) / NT;
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_last[1];
  const int d = a.d, d3 = 3 * d, rows = a.rows;
  const int ldt = d + 1;
  constexpr int PS = 2 * DP + 4;
  float* pbuf = sm;                        // [grid][PS] staged partials (last arriver only)
  float* WiT = sm + gridDim.x * PS;        // [3d][d+1]
  float* WhT = WiT + d3 * ldt;             // [3d][d+1]
  float* Ms = WhT + d3 * ldt;              // [TR][d]
  float* Hs = Ms + TR * d;                 // [TR][d]
  float* Gi = Hs + TR * d;                 // [TR][3d]
  float* Gh = Gi + TR * d3;                // [TR][3d]
  float* red = Gh + TR * d3;               // [2][GPB][DP+1]
  float* bnv = red + 2 * (NT / 32) * DP;   // BN of step t-1 (produces h_t):  mean | rstd | gamma | beta
  float* bnc = bnv + 4 * DP;               // BN of step t (being differentiated): mean | rstd | gamma | beta
  float* sv = bnc + 4 * DP;                // S1[DP] | S2[DP] of BN_t
  int* rowlist = reinterpret_cast<int*>(sv + 2 * DP);   // [chunk]
  for (int i = threadIdx.x; i < d * d3; i += NT) {
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
    const int e = threadIdx.x + q * NT;
    pk[q] = -1;
    if (e < 2 * nW) {
      const int which = e >= nW;
      const int ee = e - which * nW;
      const int l = ee / d3, g = ee - l * d3;
      pk[q] = (which * TR * d + l) | ((which * TR * d3 + g) << 16);
    }
  }
  float accb = 0.f;
  const int chunk = (rows + n_workers() - 1) / n_workers();
  const int nreal = build_rowlist(a, chunk, rowlist, s_last);
  unsigned nbar = 0;

  // prologue: the two column sums of the LAST batch norm, from the incoming gradient
  const int T1 = a.T - 1;
  load_bnv<DP>(a, T1, bnc);
  if (a.bn[T1].kind) {
    const float* Gl = a.G + (size_t)T1 * rows * d;
    float s12[2] = {0.f, 0.f};
    if (on)
      for (int idx = grp; idx < nreal; idx += GPB) {
        const int row = rowlist[idx];
        const float mu = __ldg(a.mask + row);
        float xh;
        bn_out<DP>(a.bn[T1].kind, __ldcg(Gl + (size_t)row * d + c), mu, bnc, c, &xh);
        const float dy = __ldg(b.dout + (size_t)row * (size_t)b.dout_ld + c);
        s12[0] = fmaf(dy, xh, s12[0]);
        s12[1] = fmaf(dy, mu, s12[1]);
      }
    block_colsums<DP, 2>(s12, red, grp, c);
    publish_bn_sums<DP>(b, T1, s12[0], s12[1], ++nbar, pbuf, red, s_last, grp, c);
  }
  for (int t = T1; t >= 0; --t) {
    const BNDesc bn = a.bn[t];
    const int pkind = t > 0 ? a.bn[t - 1].kind : 0;
    const float* dyin = t == T1 ? b.dout : b.dY;
    const size_t dy_ld = t == T1 ? (size_t)b.dout_ld : (size_t)d;
    // operands of the first tile and the statistics of the previous batch norm are loaded while the sums are in flight
    BwdOps nx;
    int nrow;
    load_ops<DP>(a, t, dyin, dy_ld, rowlist, grp, nreal, on, c, nx, &nrow);
    load_bnv<DP>(a, t - 1, bnv);   // the batch norm that produced h_t
    float Mn = 1.f;
    if (bn.kind) {
      Mn = __ldcg(a.stats + (size_t)t * SS + 96);
      if (threadIdx.x < 2 * DP) sv[threadIdx.x] = mail_wait(a.mail + (size_t)t * MAILW + threadIdx.x);
      __syncthreads();
    }
    float* dMt = b.dM + (size_t)t * rows * d;
    float s12[2] = {0.f, 0.f};
    for (int base = 0; base < nreal; base += TR) {
      const int row = nrow;
      const bool live = base + grp < nreal;
      const BwdOps o = nx;
      load_ops<DP>(a, t, dyin, dy_ld, rowlist, base + TR + grp, nreal, on, c, nx, &nrow);
      __syncthreads();   // previous tile's accumulation is done with the staging buffers
      float dar = 0.f, daz = 0.f, dan = 0.f, dnh = 0.f, dhd = 0.f, mv = 0.f, hv = 0.f, xhp = 0.f;
      const float mu = o.mu;
      if (live && on) {
        // ---- batch norm backward: gradient w.r.t. the GRU output g_t --------------------------------------------
        float dg = o.dy;
        if (bn.kind == 1) {
          float xh;
          bn_out<DP>(1, o.g, mu, bnc, c, &xh);
          dg = mu * bnc[DP + c] * (o.dy - (xh * sv[c] + sv[DP + c]) / Mn);
        } else if (bn.kind == 2) {
          const float ga = bnc[2 * DP + c], r = bnc[DP + c];
          if (bn.training) {
            float xh;
            bn_out<DP>(2, o.g, mu, bnc, c, &xh);
            const float s = 1.f / r - bn.eps;   // sqrt(var)
            dg = mu * ga * (r * (o.dy - sv[DP + c] / Mn) - sv[c] * xh / (s * Mn));
          } else {
            dg = mu * ga * r * o.dy;
          }
        }
        // ---- GRU backward (gru_update.py:26-35) -------------------------------------------------------------------
        const float sr = o.sr, sz = o.sz, tn = o.tn, nh = o.nh;
        mv = o.mv;
        hv = t == 0 ? o.hraw : bn_out<DP>(pkind, o.hraw, mu, bnv, c, &xhp);
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
          s12[0] = fmaf(ah, xhp, s12[0]);
          s12[1] = fmaf(ah, mu, s12[1]);
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
        block_colsums<DP, 2>(s12, red, grp, c);
        publish_bn_sums<DP>(b, t - 1, s12[0], s12[1], ++nbar, pbuf, red, s_last, grp, c);
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
    const int e = threadIdx.x + q * NT;
    if (e < 2 * nW) gp[e] = acc[q];
  }
  if (threadIdx.x < 2 * d3) gp[2 * nW + threadIdx.x] = accb;
  grid_barrier(a.bar, s_last);
  // one warp per element, lane = slice of the CTA list (all loads of a lane in flight together), fixed-order tree
  {
    const int warp = threadIdx.x >> 5;
    constexpr int NL = MAXGRID / 32;
    for (int e = blockIdx.x * (NT / 32) + warp; e < total; e += gridDim.x * (NT / 32)) {
      float v[NL];
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const int p = lane + 32 * i;
        v[i] = p < (int)gridDim.x ? __ldcg(b.gpart + (size_t)p * total + e) : 0.f;
      }
      float sacc = 0.f;
#pragma unroll
      for (int i = 0; i < NL; ++i) sacc += v[i];
      sacc = warp_sum(sacc);
      if (lane == 0) {
        if (e < nW) b.dW_ih[e] = sacc;
        else if (e < 2 * nW) b.dW_hh[e - nW] = sacc;
        else if (e < 2 * nW + d3) b.db_ih[e - 2 * nW] = sacc;
        else b.db_hh[e - 2 * nW - d3] = sacc;
      }
    }
  }
  grid_exit(a, s_last);
}

// co-resident grid: every SM gets the same number of CTAs whenever there is work for all of them (rows are dealt
// round-robin, so the CTAs are statistically identical)
int chain_grid(int rows, int DP, int blocks_per_sm) {
  const int per = NT / DP;
  int tiles = ceil_div(rows, per);
  int cap = mpnn_num_sms() * blocks_per_sm;
  if (cap > MAXGRID) cap = MAXGRID;
  if (tiles <= 1) return 1;                       // one CTA: worker and reducer in one
  return tiles + 1 < cap ? tiles + 1 : cap;       // workers + the reducer CTA
}
int chain_chunk(long long rows, int grid) { return ceil_div(rows, grid > 1 ? grid - 1 : 1); }

constexpr int CHAIN_BLOCKS_PER_SM = 1;

size_t fwd_smem(int d, int DP, int chunk, int grid) {
  return (size_t)(grid * (2 * DP + 4) + 2 * d * DP * 4 + 2 * (NT / 32) * DP + 4 * DP + chunk) * sizeof(float);
}
size_t bwd_smem(int d, int DP, int chunk, int grid) {
  const int TR = NT / DP;
  return (size_t)(grid * (2 * DP + 4) + 2 * 3 * d * (d + 1) + 2 * TR * d + 2 * TR * 3 * d + 2 * (NT / 32) * DP + 10 * DP +
                  chunk) * sizeof(float);
}

// workspace: [0, 256) barrier words | mailboxes | per-step partials | GRU gradient partials | dY scratch
constexpr size_t WS_HEAD = 256 + (size_t)MAXT * MAILW * sizeof(unsigned);

bool fill_chain(Chain* a, const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, int ecap,
                int zero_type, const float* H0, const float* h_init, const float* mask, const float* const* tables,
                int T, const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const int* bn_kind,
                const int* bn_training, const float* bn_eps, const float* bn_momentum, float* const* bn_ptrs,
                long long rows, int d, float* saved, float* out, void* workspace, int grid, int DP,
                const int* real_list) {
  if (T < 1 || T > MAXT) return false;
  a->real_list = real_list;
  a->real_count = real_list ? real_list + rows : nullptr;
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
  a->mail = (unsigned*)((char*)workspace + 256);
  a->part = (float*)((char*)workspace + WS_HEAD);
  a->dbg = nullptr;
  (void)grid;
  (void)DP;
  return true;
}

}  // namespace

static long long* g_chain_dbg = nullptr;

extern "C" {

// profiling aid: copies the 64 phase timestamps (ns) CTA 0 of the last forward recorded; synchronises the device
int mpnn_chain_debug(long long* out64) {
  if (!g_chain_dbg) return MPNN_ERR_ARG;
  MPNN_CUDA(cudaDeviceSynchronize());
  MPNN_CUDA(cudaMemcpy(out64, g_chain_dbg, (64 + 3 * 640) * sizeof(long long), cudaMemcpyDeviceToHost));
  return MPNN_OK;
}

// list [rows + 1]: the rows with mask != 0 in increasing order, list[rows] = their number (what mpnn_chain_* take as
// `real_list`); rows <= mpnn_real_rows_max(); workspace: 2 * (rows + 1) ints
int mpnn_real_rows_max(void) { return SMALL_SCAN_MAX; }

int mpnn_real_rows(const float* mask, long long rows, int* list, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && rows <= SMALL_SCAN_MAX, MPNN_ERR_UNSUPPORTED, "real_rows: %lld rows > %d", rows, SMALL_SCAN_MAX);
  MPNN_REQUIRE(workspace_bytes >= 2 * (size_t)(rows + 1) * sizeof(int), MPNN_ERR_WORKSPACE, "real_rows: workspace");
  int* flags = (int*)workspace;
  int* pos = flags + rows + 1;
  k_real_rows<<<1, 1024, 0, stream>>>(mask, (int)rows, flags, pos, list, list + rows);
  MPNN_CHECK_LAUNCH("k_real_rows");
  return MPNN_OK;
}

int mpnn_chain_supported(int d, int T) {
  return d >= 1 && d <= 32 && T >= 1 && T <= MAXT ? 1 : 0;
}

// floats saved by the forward for the backward: messages, gates, GRU outputs, batch statistics
long long mpnn_chain_saved_floats(long long rows, int d, int T) {
  return (long long)T * rows * d * 6 + (long long)T * SS;
}

// the first 256 + 2048 bytes (barrier words, mailboxes) must be ZERO on entry and are zero again on exit; the rest is scratch
size_t mpnn_chain_workspace_bytes(long long rows, int d, int T) {
  const int DP = pow2_at_least(d, 8);
  const int grid = chain_grid((int)rows, DP, CHAIN_BLOCKS_PER_SM);
  size_t part = (size_t)T * grid * (2 * DP + 4) * sizeof(float);
  size_t gpart = (size_t)grid * (2 * d * 3 * d + 6 * d) * sizeof(float);
  size_t dy = (size_t)rows * d * sizeof(float);
  return WS_HEAD + align_up(part, 256) + align_up(gpart, 256) + align_up(dy, 256);
}

// bn_kind/bn_training/bn_eps/bn_momentum: host arrays [T]; bn_ptrs: host array [4T] of device pointers
// (gamma, beta, running_mean, running_var per step; NULL entries allowed); tables: host array [T] of device pointers.
int mpnn_chain_fwd(const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, int ecap, int zero_type,
                   const float* H0, const float* h_init, const float* mask, const float* const* tables, int T,
                   const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const int* bn_kind,
                   const int* bn_training, const float* bn_eps, const float* bn_momentum, float* const* bn_ptrs,
                   long long rows, int d, const int* real_list, float* saved, float* out, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && rows < (1ll << 31) && mpnn_chain_supported(d, T), MPNN_ERR_UNSUPPORTED,
               "chain_fwd: unsupported dims (rows %lld, d %d, T %d)", rows, d, T);
  static long long* dbg_buf = nullptr;   // MPNN_B200_CHAIN_DEBUG=1: phase timestamps of CTA 0 (see mpnn_chain_debug)
  if (!dbg_buf && getenv("MPNN_B200_CHAIN_DEBUG")) cudaMalloc(&dbg_buf, (64 + 3 * 640) * sizeof(long long));
  MPNN_REQUIRE(workspace_bytes >= mpnn_chain_workspace_bytes(rows, d, T), MPNN_ERR_WORKSPACE, "chain_fwd: workspace");
  const int DP = pow2_at_least(d, 8);
  const int grid = chain_grid((int)rows, DP, CHAIN_BLOCKS_PER_SM);
  Chain a;
  MPNN_REQUIRE(fill_chain(&a, row_ptr, edge_src, uid, alpha, ecap, zero_type, H0, h_init, mask, tables, T, W_ih, W_hh,
                          b_ih, b_hh, bn_kind, bn_training, bn_eps, bn_momentum, bn_ptrs, rows, d, saved, out, workspace,
                          grid, DP, real_list),
               MPNN_ERR_ARG, "chain_fwd: bad step description");
  a.dbg = dbg_buf;
  g_chain_dbg = dbg_buf;
  const size_t smem = fwd_smem(d, DP, chain_chunk(rows, grid), grid);
  MPNN_REQUIRE(smem <= 200 * 1024, MPNN_ERR_UNSUPPORTED, "chain_fwd: %lld rows do not fit the per-CTA row list", rows);
  if (smem > 48 * 1024) {
    const void* f = DP == 8 ? (const void*)k_chain_fwd<8> : DP == 16 ? (const void*)k_chain_fwd<16> : (const void*)k_chain_fwd<32>;
    MPNN_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  void* args[] = {&a};
  cudaError_t e;
  switch (DP) {
    case 8: e = cudaLaunchCooperativeKernel((void*)k_chain_fwd<8>, dim3(grid), dim3(NT), args, smem, stream); break;
    case 16: e = cudaLaunchCooperativeKernel((void*)k_chain_fwd<16>, dim3(grid), dim3(NT), args, smem, stream); break;
    default: e = cudaLaunchCooperativeKernel((void*)k_chain_fwd<32>, dim3(grid), dim3(NT), args, smem, stream); break;
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
                   long long rows, int d, const int* real_list, float* saved, const float* dout, long long dout_ld,
                   float* dM, float* dh_init, float* dW_ih,
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
                          workspace, grid, DP, real_list),
               MPNN_ERR_ARG, "chain_bwd: bad step description");
  MPNN_REQUIRE(dout_ld >= d, MPNN_ERR_ARG, "chain_bwd: dout row stride smaller than d");
  b.dout = dout;
  b.dout_ld = dout_ld;
  b.dM = dM;
  b.dh_init = dh_init;
  const size_t part = align_up((size_t)T * grid * (2 * DP + 4) * sizeof(float), 256);
  const size_t gpart = align_up((size_t)grid * (2 * d * 3 * d + 6 * d) * sizeof(float), 256);
  b.gpart = (float*)((char*)workspace + WS_HEAD + part);
  b.dY = (float*)((char*)workspace + WS_HEAD + part + gpart);
  b.dW_ih = dW_ih;
  b.dW_hh = dW_hh;
  b.db_ih = db_ih;
  b.db_hh = db_hh;
  for (int t = 0; t < MAXT; ++t) {
    b.dgamma[t] = (t < T && bn_grads) ? bn_grads[2 * t] : nullptr;
    b.dbeta[t] = (t < T && bn_grads) ? bn_grads[2 * t + 1] : nullptr;
  }
  const size_t smem = bwd_smem(d, DP, chain_chunk(rows, grid), grid);
  MPNN_REQUIRE(smem <= 200 * 1024, MPNN_ERR_UNSUPPORTED, "chain_bwd: %lld rows do not fit the per-CTA row list", rows);
  if (smem > 48 * 1024) {
    const void* f = DP == 8 ? (const void*)k_chain_bwd<8> : DP == 16 ? (const void*)k_chain_bwd<16> : (const void*)k_chain_bwd<32>;
    MPNN_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  void* args[] = {&b};
  cudaError_t e;
  switch (DP) {
    case 8: e = cudaLaunchCooperativeKernel((void*)k_chain_bwd<8>, dim3(grid), dim3(NT), args, smem, stream); break;
    case 16: e = cudaLaunchCooperativeKernel((void*)k_chain_bwd<16>, dim3(grid), dim3(NT), args, smem, stream); break;
    default: e = cudaLaunchCooperativeKernel((void*)k_chain_bwd<32>, dim3(grid), dim3(NT), args, smem, stream); break;
  }
  MPNN_REQUIRE(e == cudaSuccess, MPNN_ERR_CUDA, "chain_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return MPNN_OK;
}

}  // extern "C"
