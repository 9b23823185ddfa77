// Set2Vec (reference mpnn_functions/readout/set2vec.py:93-151) as TWO persistent cooperative kernels: the `steps`
// (default 100) strictly sequential LSTM-attention iterations of the forward, and the backward through time.
//
// Every iteration is latency-bound (gates from a [B,2F] vector, a query, B*N attention energies, ONE softmax across the
// whole batch -- set2vec.py:139, dim=0 -- and the per-graph read-out), so the per-step launches of set2vec.cu cost
// ~26 us per iteration (six launches forward, six backward).  Here a CTA owns G = ceil(B / grid) whole graphs for the
// whole loop:
//   * the LSTM weights [2F,4F] (forward) / their transpose (backward), the query weights, the CTA's rows of X and its
//     recurrent state live in shared memory for all iterations;
//   * everything except the softmax is graph-local.  The softmax needs two batch-wide numbers (max and sum; the backward
//     one: sum att*datt): each CTA publishes its local pair in a 16-byte slot tagged with the iteration number and reads
//     everybody else's slots -- an all-gather through L2 with the flag inside the data, no reducer, no atomics; slots
//     are double-buffered by iteration parity (a CTA can only be one barrier ahead of the slowest one).  The slots are
//     combined in CTA order by every CTA: bit-reproducible.
//   * the backward reads what the forward saved (gates, tanh c, c_prev, q, att of the iteration) through a cp.async
//     double buffer one iteration ahead; tanh(q + X) is recomputed.
// The parameter gradients stay what they were: stacks of per-iteration dq / pw / dpre, contracted once after the loop
// (set2vec.cu).  Widths F <= 64 (the weights must fit in shared memory) and batches of up to ~24 graphs per SM; anything
// else keeps the per-step path.
#include "common.cuh"

namespace {

constexpr float BIG_NEGATIVE = -1e8f;  // set2vec.py:10
constexpr int NT = 512;
constexpr int SMEM_LIMIT = 227 * 1024;

struct S2VFwd {
  const float *X, *mask, *Wcat, *bcat, *Wq, *we, *m0, *c0;
  float *saved, *out;
  uint4* slots;   // [2][grid]
  unsigned long long* dbg;
  int B, N, F, steps, G, x_smem, KS, GS;
};

struct S2VBwd {
  const float *X, *Wcat, *Wq, *we, *saved, *dout, *c0;
  float *dX, *dqS, *pwS, *dpreS, *dm0, *dc0;
  uint4* slots;
  unsigned long long* dbg;
  int B, N, F, steps, G, x_smem, KS, GS, fast;
};

__host__ __device__ inline size_t step_stride(int B, int N, int F) { return (size_t)B * 9 * F + (size_t)B * N; }

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define S2V_STAMP(base, it, ph)                                                              \
  do {                                                                                       \
    if (a.dbg && blockIdx.x == 0 && tid == 0 && (it) < 4) a.dbg[(base) + (it) * 16 + (ph)] = gtime(); \
  } while (0)

__device__ __forceinline__ uint4 ld_slot(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_slot(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// row stride of X in shared memory: the smallest value >= F that is 8 mod 32, so that the eight threads that share a
// row (features sub, sub + 8, ...) and the four rows of a warp fall into 32 different banks
__host__ __device__ inline int x_stride(int F) { return (F + 23) / 32 * 32 + 8; }
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& w) {
  acc.x = fmaf(s, w.x, acc.x);
  acc.y = fmaf(s, w.y, acc.y);
  acc.z = fmaf(s, w.z, acc.z);
  acc.w = fmaf(s, w.w, acc.w);
}

// out[g, c4..c4+3] = sum_{k0 <= k < k1} v[g, k] Wm[k, c4..c4+3] for the NB graphs g = gb, gb + GS, ...
template <int NB>
__device__ __forceinline__ void gemv4(const float* __restrict__ Wm, int ldw, int c4, int k0, int k1,
                                      const float* __restrict__ v, int ldv, int gb, int GS, float* __restrict__ out,
                                      int ldo) {
  float4 acc[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const float4 w = *reinterpret_cast<const float4*>(Wm + k * ldw + c4);
#pragma unroll
    for (int b = 0; b < NB; ++b) fma4(acc[b], v[(gb + b * GS) * ldv + k], w);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) *reinterpret_cast<float4*>(out + (gb + b * GS) * ldo + c4) = acc[b];
}
__device__ __forceinline__ void gemv4_graphs(const float* __restrict__ Wm, int ldw, int c4, int k0, int k1,
                                             const float* __restrict__ v, int ldv, int gs, int GS, int Gc,
                                             float* __restrict__ out, int ldo) {
  for (int gb = gs; gb < Gc; gb += 4 * GS) {
    const int nb = (Gc - gb + GS - 1) / GS;
    if (nb >= 4) gemv4<4>(Wm, ldw, c4, k0, k1, v, ldv, gb, GS, out, ldo);
    else if (nb == 3) gemv4<3>(Wm, ldw, c4, k0, k1, v, ldv, gb, GS, out, ldo);
    else if (nb == 2) gemv4<2>(Wm, ldw, c4, k0, k1, v, ldv, gb, GS, out, ldo);
    else gemv4<1>(Wm, ldw, c4, k0, k1, v, ldv, gb, GS, out, ldo);
  }
}

// every lane reads its slots of the all-gather (CTAs lane, lane + 32, ...; at most MAXSLOT) with all loads in flight at
// once and polls until each carries the iteration tag
constexpr int MAXSLOT = 5;   // grids of up to 160 CTAs
__device__ __forceinline__ void gather_slots(const uint4* sl, int grid, int lane, uint32_t tag, uint4 (&v)[MAXSLOT]) {
  bool ok[MAXSLOT];
#pragma unroll
  for (int i = 0; i < MAXSLOT; ++i) ok[i] = lane + 32 * i >= grid;
  long long spins = 0;
  bool all;
  do {
#pragma unroll
    for (int i = 0; i < MAXSLOT; ++i)
      if (!ok[i]) v[i] = ld_slot(sl + lane + 32 * i);
    all = true;
#pragma unroll
    for (int i = 0; i < MAXSLOT; ++i) {
      if (!ok[i]) ok[i] = v[i].x == tag && v[i].w == tag;
      all = all && ok[i];
    }
    if (++spins > (1ll << 26)) __trap();   // a protocol bug ends in a CUDA error, never in a hung GPU
  } while (!all);
}

// Work split used by every per-graph product of both kernels.  A thread is (column c, part p); the NP parts of a column
// are KS slices of the reduction axis times GS groups of graphs (KS * GS <= NP; one graph per SM: KS = NP, GS = 1, many
// graphs per SM: KS = 1, GS = NP); slice partial sums go through shared memory and are summed in slice order.

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) k_s2v_fwd(S2VFwd a) {
  extern __shared__ __align__(16) float sm[];
  const int F = a.F, N = a.N, G = a.G, F2 = 2 * F, F4 = 4 * F, KS = a.KS, GS = a.GS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.x * G;
  const int Gc = min(G, a.B - g0);       // graphs of this CTA (>= 1)
  const int rows = Gc * N;
  float* W = sm;                         // [2F][4F]
  float* preH = W + F2 * F4;             // [KS][G][4F] slice partial sums of the gate pre-activations, h rows of W
  float* preR = preH + KS * G * F4;      // [KS][G][4F] the same, read-out rows of W
  float* WqT = preR + KS * G * F4;       // [F][F]: WqT[k][f] = Wq[f][k]
  float* bc = WqT + F * F;               // [4F]
  float* we = bc + F4;                   // [F]
  float* m = we + F;                     // [G][2F]
  float* c = m + G * F2;                 // [G][F]
  float* qp = c + G * F;                 // [KS][G][F] slice partial sums of the query, then of the read-out
  float* q = qp + KS * G * F;            // [G][F]
  float* e = q + G * F;                  // [G*N]
  float* neg = e + G * N;                // [G*N]: (1 - mask) * BIG_NEGATIVE
  float* red = neg + G * N;              // [64]
  float* Xs = red + 64;                  // [G*N][x_stride] (optional)

  for (int i = tid; i < F2 * F4; i += NT) W[i] = __ldg(a.Wcat + i);
  for (int i = tid; i < F * F; i += NT) {
    const int f = i / F, k = i - f * F;
    WqT[k * F + f] = __ldg(a.Wq + i);
  }
  for (int i = tid; i < F4; i += NT) bc[i] = __ldg(a.bcat + i);
  for (int i = tid; i < F; i += NT) we[i] = __ldg(a.we + i);
  for (int i = tid; i < Gc * F2; i += NT) m[i] = a.m0 ? __ldg(a.m0 + (size_t)g0 * F2 + i) : 0.f;
  for (int i = tid; i < Gc * F; i += NT) c[i] = a.c0 ? __ldg(a.c0 + (size_t)g0 * F + i) : 0.f;
  const float* Xg = a.X + (size_t)g0 * N * F;
  const int ldx = a.x_smem ? x_stride(F) : F;
  if (a.x_smem)
    for (int i = tid; i < rows * F; i += NT) Xs[(i / F) * ldx + (i % F)] = __ldg(Xg + i);
  for (int r = tid; r < rows; r += NT) neg[r] = a.mask ? (1.f - __ldg(a.mask + (size_t)g0 * N + r)) * BIG_NEGATIVE : 0.f;
  const float* Xr = a.x_smem ? Xs : Xg;
  __syncthreads();

  const size_t stride = step_stride(a.B, N, F);
  const int grid = gridDim.x;
  const int col = tid % F, part = tid / F;            // part >= KS * GS: no product work
  const int ks = part % KS, gs = part / KS;
  const bool worker = gs < GS;
  const int kper = (F + KS - 1) / KS, k0 = min(F, ks * kper), k1 = min(F, k0 + kper);   // this thread's slice of [0, F)
  // gate pre-activations pre[g, j] = sum_k m[g, k] W[k, j] (thread = 4 columns, slice of k, group of graphs), in two
  // halves: the rows of W that meet h (known as soon as the LSTM cell is done) and the rows that meet the read-out
  if (worker) {
    gemv4_graphs(W, F4, 4 * col, k0, k1, m, F2, gs, GS, Gc, preH + ks * G * F4, F4);
    gemv4_graphs(W + F * F4, F4, 4 * col, k0, k1, m + F, F2, gs, GS, Gc, preR + ks * G * F4, F4);
  }
  __syncthreads();
  for (int s = 0; s < a.steps; ++s) {
    float* sv = a.saved + (size_t)s * stride;
    float* sv_m = sv;
    float* sv_c = sv_m + (size_t)a.B * F2;
    float* sv_g = sv_c + (size_t)a.B * F;
    float* sv_tc = sv_g + (size_t)a.B * F4;
    float* sv_q = sv_tc + (size_t)a.B * F;
    float* sv_att = sv_q + (size_t)a.B * F;
    S2V_STAMP(0, s, 0);
    // ---- LSTM cell (set2vec.py:68-75) ----
    for (int i = tid; i < Gc * F; i += NT) {
      const int g = i / F, f = i - g * F;
      float p0 = bc[f], p1 = bc[F + f], p2 = bc[2 * F + f], p3 = bc[3 * F + f];
      for (int k = 0; k < KS; ++k) {
        const float* ph = preH + (k * G + g) * F4;
        const float* pr = preR + (k * G + g) * F4;
        p0 += ph[f] + pr[f];
        p1 += ph[F + f] + pr[F + f];
        p2 += ph[2 * F + f] + pr[2 * F + f];
        p3 += ph[3 * F + f] + pr[3 * F + f];
      }
      const float ig = 1.f / (1.f + expf(-p0));
      const float fg = 1.f / (1.f + expf(-p1));
      const float gg = tanhf(p2);
      const float og = 1.f / (1.f + expf(-p3));
      const float cn = fg * c[i] + ig * gg;
      const float th = tanhf(cn);
      const float h = og * th;
      c[i] = cn;
      m[g * F2 + f] = h;
      const size_t b = (size_t)(g0 + g);
      float* gsv = sv_g + b * F4;
      gsv[f] = ig;
      gsv[F + f] = fg;
      gsv[2 * F + f] = gg;
      gsv[3 * F + f] = og;
      sv_c[b * F + f] = cn;
      sv_tc[b * F + f] = th;
      sv_m[b * F2 + f] = h;
    }
    __syncthreads();
    S2V_STAMP(0, s, 1);
    // ---- query: q[g, f] = sum_k h[g, k] Wq[f, k] ----
    if (worker) {
      for (int g = gs; g < Gc; g += GS) {
        const float* h = m + g * F2;
        float acc = 0.f;
#pragma unroll 8
        for (int k = k0; k < k1; ++k) acc = fmaf(h[k], WqT[k * F + col], acc);
        qp[(ks * G + g) * F + col] = acc;
      }
    }
    __syncthreads();
    for (int i = tid; i < Gc * F; i += NT) {
      const int g = i / F, f = i - g * F;
      float acc = 0.f;
      for (int k = 0; k < KS; ++k) acc += qp[(k * G + g) * F + f];
      q[i] = acc;
      sv_q[(size_t)(g0 + g) * F + f] = acc;
    }
    __syncthreads();
    S2V_STAMP(0, s, 2);
    // ---- energies: e[r] = sum_f we[f] tanh(q[g, f] + X[r, f]) + (1 - mask[r]) BIG_NEGATIVE; eight threads per row ----
    for (int base = 0; base < rows * 8; base += NT) {
      const int job = base + tid, r = job >> 3, sub = job & 7;
      float acc = 0.f;
      if (r < rows) {
        const float* qg = q + (r / N) * F;
        const float* xr = Xr + (size_t)r * ldx;
        for (int f = sub; f < F; f += 8) acc = fmaf(we[f], tanhf(qg[f] + xr[f]), acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (r < rows && sub == 0) e[r] = acc + neg[r];
    }
    __syncthreads();
    S2V_STAMP(0, s, 3);
    // ---- the softmax across the whole batch (set2vec.py:139).  Every warp finds the CTA's largest energy lm; warp 0
    // publishes (lm, sum exp(e - lm)) and, LAST, collects everybody's pair.  Meanwhile all threads do what does not
    // need the batch-wide numbers: the read-out with weights exp(e - lm) and the h half of the next gate product ----
    float lm = -INFINITY;
    for (int r = lane; r < rows; r += 32) lm = fmaxf(lm, e[r]);
    lm = warp_max(lm);
    const uint32_t tag = (uint32_t)(s + 1);
    uint4* sl = a.slots + (size_t)(s & 1) * grid;
    if (warp == 0) {
      float ls = 0.f;
      for (int r = lane; r < rows; r += 32) ls += expf(e[r] - lm);
      ls = warp_sum(ls);
      if (lane == 0) {
        st_slot(sl + blockIdx.x, make_uint4(tag, __float_as_uint(lm), __float_as_uint(ls), tag));
        if (a.dbg && s == 2) a.dbg[128 + 2 * blockIdx.x] = gtime();
      }
    }
    S2V_STAMP(0, s, 4);
    if (worker) {
      for (int g = gs; g < Gc; g += GS) {
        float acc = 0.f;
        for (int n = ks; n < N; n += KS)
          acc = fmaf(expf(e[g * N + n] - lm), Xr[(size_t)(g * N + n) * ldx + col], acc);
        qp[(ks * G + g) * F + col] = acc;
      }
      gemv4_graphs(W, F4, 4 * col, k0, k1, m, F2, gs, GS, Gc, preH + ks * G * F4, F4);
    }
    S2V_STAMP(0, s, 5);
    if (warp == 0) {
      uint4 v[MAXSLOT];
      gather_slots(sl, grid, lane, tag, v);
      float gm = -INFINITY;
#pragma unroll
      for (int i = 0; i < MAXSLOT; ++i)
        if (lane + 32 * i < grid) gm = fmaxf(gm, __uint_as_float(v[i].y));
      gm = warp_max(gm);
      float z = 0.f;
#pragma unroll
      for (int i = 0; i < MAXSLOT; ++i)
        if (lane + 32 * i < grid) z += __uint_as_float(v[i].z) * expf(__uint_as_float(v[i].y) - gm);
      z = warp_sum(z);
      if (lane == 0) {
        red[0] = expf(lm - gm) / z;     // exp(e - lm) * red[0] = exp(e - gm) / Z
        if (a.dbg && s == 2) a.dbg[129 + 2 * blockIdx.x] = gtime();
      }
      S2V_STAMP(0, s, 6);
    }
    __syncthreads();
    {
      const float sc = red[0];
      for (int r = tid; r < rows; r += NT) sv_att[(size_t)g0 * N + r] = expf(e[r] - lm) * sc;
      for (int i = tid; i < Gc * F; i += NT) {
        const int g = i / F, f = i - g * F;
        float acc = 0.f;
        for (int k = 0; k < KS; ++k) acc += qp[(k * G + g) * F + f];
        acc *= sc;
        m[g * F2 + F + f] = acc;
        sv_m[(size_t)(g0 + g) * F2 + F + f] = acc;
      }
    }
    __syncthreads();
    S2V_STAMP(0, s, 7);
    if (worker) gemv4_graphs(W + F * F4, F4, 4 * col, k0, k1, m + F, F2, gs, GS, Gc, preR + ks * G * F4, F4);
    __syncthreads();
    S2V_STAMP(0, s, 8);
  }
  for (int i = tid; i < Gc * F2; i += NT) a.out[(size_t)g0 * F2 + i] = m[i];
}

// ---------------------------------------------------------------------------------------------------
// backward through time
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) k_s2v_bwd(S2VBwd a) {
  extern __shared__ __align__(16) float sm[];
  const int F = a.F, N = a.N, G = a.G, F2 = 2 * F, F4 = 4 * F, KS = a.KS, GS = a.GS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.x * G;
  const int Gc = min(G, a.B - g0);
  const int rows = Gc * N;
  const int SG = 7 * F + N;              // staged floats per graph and iteration: gates 4F | tc F | c_prev F | q F | att N
  float* WT = sm;                        // [4F][2F]: WT[j][k] = Wcat[k][j]
  float* part = WT + F4 * F2;            // [2 KS][G][2F] slice partial sums (dm_prev); energy / dh phases use a prefix
  float* Wq = part + 2 * KS * G * F2;    // [F][F]
  float* we = Wq + F * F;                // [F]
  float* dm = we + F;                    // [G][2F]
  float* dc = dm + G * F2;               // [G][F]
  float* dq = dc + G * F;                // [G][F]
  float* dpre = dq + G * F;              // [G][4F]
  float* stage = dpre + G * F4;          // [2][G][SG]
  float* datt = stage + 2 * G * SG;      // [G*N]
  float* red = datt + G * N;             // [64]
  float* Xs = red + 64;                        // [G*N][x_stride]   (optional)
  float* dXs = Xs + (size_t)G * N * x_stride(F);   // [G*N][x_stride]   (optional)

  for (int i = tid; i < F2 * F4; i += NT) {
    const int k = i / F4, j = i - k * F4;
    WT[j * F2 + k] = __ldg(a.Wcat + i);
  }
  for (int i = tid; i < F * F; i += NT) Wq[i] = __ldg(a.Wq + i);
  for (int i = tid; i < F; i += NT) we[i] = __ldg(a.we + i);
  for (int i = tid; i < Gc * F2; i += NT) dm[i] = __ldg(a.dout + (size_t)g0 * F2 + i);
  for (int i = tid; i < Gc * F; i += NT) dc[i] = 0.f;
  const float* Xg = a.X + (size_t)g0 * N * F;
  float* dXg = a.dX + (size_t)g0 * N * F;
  const int ldx = a.x_smem ? x_stride(F) : F;
  if (a.x_smem) {
    for (int i = tid; i < rows * F; i += NT) {
      Xs[(i / F) * ldx + (i % F)] = __ldg(Xg + i);
      dXs[(i / F) * ldx + (i % F)] = 0.f;
    }
  }
  const float* Xr = a.x_smem ? Xs : Xg;
  float* dXr = a.x_smem ? dXs : dXg;     // (global accumulation: the host zeroed dX)

  const size_t stride = step_stride(a.B, N, F);
  const int grid = gridDim.x;
  const size_t sb_F = (size_t)a.B * F;
  const int col = tid % F, prt = tid / F;
  const int ks = prt % KS, gs = prt / KS;
  const bool worker = gs < GS;
  // the dm_prev product has 2F outputs per graph, four per thread: F/2 column groups, twice the parts
  const int H = F / 2;
  const int col2 = tid % H, prt2 = tid / H;
  const int KS2 = 2 * KS;
  const int ks2 = prt2 % KS2, gs2 = prt2 / KS2;
  const bool worker2 = gs2 < GS;

  // saved values of iteration s -> staging buffer (s & 1)
  auto prefetch = [&](int s) {
    const float* sv = a.saved + (size_t)s * stride;
    const float* sv_g = sv + (size_t)a.B * 3 * F;
    const float* sv_tc = sv_g + (size_t)a.B * F4;
    const float* sv_q = sv_tc + sb_F;
    const float* sv_att = sv_q + sb_F;
    const float* cprev = s > 0 ? a.saved + (size_t)(s - 1) * stride + (size_t)a.B * F2 : a.c0;
    float* st = stage + (s & 1) * G * SG;
    for (int i = tid; i < Gc * SG; i += NT) {
      const int g = i / SG, o = i - g * SG;
      const size_t b = (size_t)(g0 + g);
      const float* src;
      if (o < F4) src = sv_g + b * F4 + o;
      else if (o < 5 * F) src = sv_tc + b * F + (o - F4);
      else if (o < 6 * F) src = cprev ? cprev + b * F + (o - 5 * F) : nullptr;
      else if (o < 7 * F) src = sv_q + b * F + (o - 6 * F);
      else src = sv_att + b * N + (o - 7 * F);
      if (src) cp_async4(st + i, src);
      else st[i] = 0.f;
    }
    cp_async_commit();
  };
  prefetch(a.steps - 1);

  for (int s = a.steps - 1; s >= 0; --s) {
    const int it = a.steps - 1 - s;
    cp_async_wait_all();
    __syncthreads();
    S2V_STAMP(64, it, 0);
    if (s > 0) prefetch(s - 1);
    const float* st = stage + (s & 1) * G * SG;
    // ---- datt[r] = sum_f dread[g, f] X[r, f] ----
    for (int base = 0; base < rows * 8; base += NT) {
      const int job = base + tid, r = job >> 3, sub = job & 7;
      float acc = 0.f;
      if (r < rows) {
        const float* dr = dm + (r / N) * F2 + F;
        const float* xr = Xr + (size_t)r * ldx;
        for (int f = sub; f < F; f += 8) acc = fmaf(dr[f], xr[f], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (r < rows && sub == 0) datt[r] = acc;
    }
    __syncthreads();
    S2V_STAMP(64, it, 1);
    // ---- batch-wide S = sum att * datt (softmax backward): warp 0 publishes the CTA's share, and collects LAST ----
    const uint32_t tag = (uint32_t)(a.steps - s);
    uint4* sl = a.slots + (size_t)(s & 1) * grid;
    if (warp == 0) {
      float ls = 0.f;
      for (int r = lane; r < rows; r += 32) {
        const int g = r / N;
        ls = fmaf(st[g * SG + 7 * F + (r - g * N)], datt[r], ls);
      }
      ls = warp_sum(ls);
      if (lane == 0) st_slot(sl + blockIdx.x, make_uint4(tag, __float_as_uint(ls), 0u, tag));
    }
    if (a.fast) {
      // Everything downstream is affine in S.  While the slots travel: tanh(q + X) and, per (graph, feature),
      //   dq = w (A1 - S A2),  A1 = sum_n att datt (1 - th^2),  A2 = sum_n att (1 - th^2)
      //   pw =    B1 - S B2,   B1 = sum_n att datt th,          B2 = sum_n att th
      //   dX += att (dread + datt w (1 - th^2))  -  S * c2,     c2 = att w (1 - th^2)   (kept in registers)
      float c2r[8];                                     // (host: one graph per thread, at most 8 rows of it)
      if (worker && gs < Gc) {
        const int g = gs;
        const float qv = st[g * SG + 6 * F + col], w = we[col], dread = dm[g * F2 + F + col];
        const float* att = st + g * SG + 7 * F;
        float A1 = 0.f, A2 = 0.f, B1 = 0.f, B2 = 0.f;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const int n = ks + ni * KS;
          c2r[ni] = 0.f;
          if (n < N) {
            const int r = g * N + n;
            const float th = tanhf(qv + Xr[(size_t)r * ldx + col]);
            const float at = att[n], da = datt[r], u = 1.f - th * th;
            const float au = at * u, ad = at * da;
            A1 = fmaf(ad, u, A1);
            A2 += au;
            B1 = fmaf(ad, th, B1);
            B2 = fmaf(at, th, B2);
            dXr[(size_t)r * ldx + col] += at * dread + ad * w * u;
            c2r[ni] = au * w;
          }
        }
        part[((0 * KS + ks) * G + g) * F + col] = A1;
        part[((1 * KS + ks) * G + g) * F + col] = A2;
        part[((2 * KS + ks) * G + g) * F + col] = B1;
        part[((3 * KS + ks) * G + g) * F + col] = B2;
      }
      if (warp == 0) {
        uint4 v[MAXSLOT];
        gather_slots(sl, grid, lane, tag, v);
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < MAXSLOT; ++i)
          if (lane + 32 * i < grid) tot += __uint_as_float(v[i].y);
        tot = warp_sum(tot);
        if (lane == 0) red[0] = tot;
      }
      __syncthreads();
      S2V_STAMP(64, it, 2);
      const float S = red[0];
      if (worker && gs < Gc) {
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          const int n = ks + ni * KS;
          if (n < N) dXr[(size_t)(gs * N + n) * ldx + col] -= S * c2r[ni];
        }
      }
      for (int i = tid; i < Gc * F; i += NT) {
        const int g = i / F, f = i - g * F;
        float A1 = 0.f, A2 = 0.f, B1 = 0.f, B2 = 0.f;
        for (int k = 0; k < KS; ++k) {
          A1 += part[((0 * KS + k) * G + g) * F + f];
          A2 += part[((1 * KS + k) * G + g) * F + f];
          B1 += part[((2 * KS + k) * G + g) * F + f];
          B2 += part[((3 * KS + k) * G + g) * F + f];
        }
        const float sq = we[f] * (A1 - S * A2), sw = B1 - S * B2;
        dq[i] = sq;
        const size_t o = ((size_t)s * a.B + g0 + g) * F + f;
        a.dqS[o] = sq;
        a.pwS[o] = sw;
      }
      __syncthreads();
      S2V_STAMP(64, it, 3);
    } else {
      if (warp == 0) {
        uint4 v[MAXSLOT];
        gather_slots(sl, grid, lane, tag, v);
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < MAXSLOT; ++i)
          if (lane + 32 * i < grid) tot += __uint_as_float(v[i].y);
        tot = warp_sum(tot);
        if (lane == 0) red[0] = tot;
      }
      __syncthreads();
      S2V_STAMP(64, it, 2);
      // ---- de[r] = att[r] (datt[r] - S) ----
      {
        const float S = red[0];
        for (int r = tid; r < rows; r += NT) {
          const int g = r / N;
          datt[r] = st[g * SG + 7 * F + (r - g * N)] * (datt[r] - S);
        }
      }
      __syncthreads();
      // ---- energy backward: thread = (feature, slice of the rows, group of graphs) ----
      if (worker) {
        for (int g = gs; g < Gc; g += GS) {
          const float qv = st[g * SG + 6 * F + col], w = we[col], dread = dm[g * F2 + F + col];
          const float* att = st + g * SG + 7 * F;
          float sq = 0.f, sw = 0.f;
          for (int n = ks; n < N; n += KS) {
            const int r = g * N + n;
            const float th = tanhf(qv + Xr[(size_t)r * ldx + col]);
            const float d = datt[r];
            const float dp = d * w * (1.f - th * th);
            sq += dp;
            sw = fmaf(d, th, sw);
            dXr[(size_t)r * ldx + col] += att[n] * dread + dp;
          }
          part[(ks * G + g) * F + col] = sq;
          part[((KS + ks) * G + g) * F + col] = sw;
        }
      }
      __syncthreads();
      S2V_STAMP(64, it, 3);
      for (int i = tid; i < Gc * F; i += NT) {
        const int g = i / F, f = i - g * F;
        float sq = 0.f, sw = 0.f;
        for (int k = 0; k < KS; ++k) {
          sq += part[(k * G + g) * F + f];
          sw += part[((KS + k) * G + g) * F + f];
        }
        dq[i] = sq;
        const size_t o = ((size_t)s * a.B + g0 + g) * F + f;
        a.dqS[o] = sq;
        a.pwS[o] = sw;
      }
      __syncthreads();
    }
    // ---- dh = dq Wq + dm[:, :F] (slices of f), then the LSTM cell backward ----
    if (worker) {
      const int fper = (F + KS - 1) / KS, f0 = ks * fper, f1 = min(F, f0 + fper);
      for (int g = gs; g < Gc; g += GS) {
        float acc = 0.f;
#pragma unroll 8
        for (int f = f0; f < f1; ++f) acc = fmaf(dq[g * F + f], Wq[f * F + col], acc);
        part[(ks * G + g) * F + col] = acc;
      }
    }
    __syncthreads();
    S2V_STAMP(64, it, 4);
    for (int i = tid; i < Gc * F; i += NT) {
      const int g = i / F, k = i - g * F;
      float dh = dm[g * F2 + k];
      for (int q_ = 0; q_ < KS; ++q_) dh += part[(q_ * G + g) * F + k];
      const float* sg = st + g * SG;
      const float ig = sg[k], fg = sg[F + k], gg = sg[2 * F + k], og = sg[3 * F + k];
      const float th = sg[F4 + k], cp = sg[5 * F + k];
      const float dcn = dh * og * (1.f - th * th) + dc[i];
      float4 dp;
      dp.x = dcn * gg * ig * (1.f - ig);
      dp.y = dcn * cp * fg * (1.f - fg);
      dp.z = dcn * ig * (1.f - gg * gg);
      dp.w = dh * th * og * (1.f - og);
      dc[i] = dcn * fg;
      float* d_ = dpre + g * F4;
      d_[k] = dp.x;
      d_[F + k] = dp.y;
      d_[2 * F + k] = dp.z;
      d_[3 * F + k] = dp.w;
      float* o = a.dpreS + ((size_t)s * a.B + g0 + g) * F4;
      o[k] = dp.x;
      o[F + k] = dp.y;
      o[2 * F + k] = dp.z;
      o[3 * F + k] = dp.w;
    }
    __syncthreads();
    S2V_STAMP(64, it, 5);
    if (s > 0 || a.dm0) {
      // ---- dm_prev[g, k] = sum_j dpre[g, j] Wcat[k, j]; thread = (4 outputs k, slice of j, group of graphs) ----
      if (worker2) {
        const int jper = (F4 + KS2 - 1) / KS2, j0 = ks2 * jper, j1 = min(F4, j0 + jper);
        gemv4_graphs(WT, F2, 4 * col2, j0, j1, dpre, F4, gs2, GS, Gc, part + ks2 * G * F2, F2);
      }
      __syncthreads();
      for (int i = tid; i < Gc * F2; i += NT) {
        const int g = i / F2, k = i - g * F2;
        float v = 0.f;
        for (int q_ = 0; q_ < KS2; ++q_) v += part[(q_ * G + g) * F2 + k];
        if (s > 0) dm[i] = v;
        else a.dm0[(size_t)g0 * F2 + i] = v;
      }
    }
    S2V_STAMP(64, it, 6);
  }
  __syncthreads();
  if (a.x_smem)
    for (int i = tid; i < rows * F; i += NT) dXg[i] = dXs[(i / F) * ldx + (i % F)];
  if (a.dc0)
    for (int i = tid; i < Gc * F; i += NT) a.dc0[(size_t)g0 * F + i] = dc[i];
}

int g_persist = 1;
unsigned long long* g_dbg = nullptr;

struct Plan {
  int G, grid, KS, GS;
};
// graphs per CTA, grid and the work split; false = not served
bool plan(int B, int N, int F, Plan* p) {
  if (!g_persist || F > 64 || F < 2 || (F & 1) || B < 1) return false;
  const int sms = mpnn_num_sms();
  p->G = (B + sms - 1) / sms;
  p->grid = (B + p->G - 1) / p->G;
  const int NP = NT / F;                 // parts per column (>= 8)
  int ksl = NP / p->G;                   // slices of the reduction axis: as many as leave one part per graph
  if (ksl < 1) ksl = 1;
  if (ksl > 8) ksl = 8;
  p->KS = ksl;
  p->GS = NP / ksl;
  if (p->GS > p->G) p->GS = p->G;
  return true;
}
size_t fwd_floats(const Plan& p, int N, int F, bool x) {
  const size_t G = p.G, KG = (size_t)p.KS * p.G;
  return (size_t)8 * F * F + 2 * KG * 4 * F + (size_t)F * F + 5 * F + G * (4 * F) + KG * F + 2 * G * N + 64 +
         (x ? G * N * x_stride(F) : 0);
}
size_t bwd_floats(const Plan& p, int N, int F, bool x) {
  const size_t G = p.G, KG = (size_t)p.KS * p.G;
  return (size_t)8 * F * F + 2 * KG * 2 * F + (size_t)F * F + F + G * (2 * F + F + F + 4 * F) +
         2 * G * (7 * F + N) + G * N + 64 + (x ? 2 * G * N * x_stride(F) : 0);
}

}  // namespace

// Called by set2vec.cu.  Return 1 = served (rc holds the status), 0 = shape not served by the persistent kernels.
size_t s2v_persist_slot_bytes() { return 2 * 160 * sizeof(uint4) + 256; }

int s2v_persist_fwd(const float* X, const float* mask, const float* Wcat, const float* bcat, const float* Wq,
                    const float* we, const float* m0, const float* c0, int B, int N, int F, int steps, float* out,
                    float* saved, void* slots, cudaStream_t stream, int* rc) {
  Plan p;
  if (!plan(B, N, F, &p)) return 0;
  if (fwd_floats(p, N, F, false) * 4 > (size_t)SMEM_LIMIT) return 0;
  const int G = p.G, grid = p.grid;
  S2VFwd a;
  a.X = X; a.mask = mask; a.Wcat = Wcat; a.bcat = bcat; a.Wq = Wq; a.we = we; a.m0 = m0; a.c0 = c0;
  a.saved = saved; a.out = out; a.slots = (uint4*)slots;
  a.B = B; a.N = N; a.F = F; a.steps = steps; a.G = G; a.KS = p.KS; a.GS = p.GS; a.dbg = g_dbg;
  a.x_smem = fwd_floats(p, N, F, true) * 4 <= (size_t)SMEM_LIMIT;
  const size_t smem = fwd_floats(p, N, F, a.x_smem) * 4;
  *rc = MPNN_ERR_CUDA;
  if (cudaFuncSetAttribute(k_s2v_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    mpnn_set_error("set2vec_fwd: cannot reserve %zu bytes of shared memory", smem);
    return 1;
  }
  if (cudaMemsetAsync(slots, 0, 2 * (size_t)grid * sizeof(uint4), stream) != cudaSuccess) {
    mpnn_set_error("set2vec_fwd: memset failed");
    return 1;
  }
  void* args[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_s2v_fwd, dim3(grid), dim3(NT), args, smem, stream);
  if (e != cudaSuccess) {
    mpnn_set_error("set2vec_fwd: cooperative launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  *rc = MPNN_OK;
  return 1;
}

int s2v_persist_bwd(const float* X, const float* Wcat, const float* Wq, const float* we, const float* c0,
                    const float* saved, const float* dout, int B, int N, int F, int steps, float* dX, float* dqS,
                    float* pwS, float* dpreS, float* dm0, float* dc0, void* slots, cudaStream_t stream, int* rc) {
  Plan p;
  if (!plan(B, N, F, &p)) return 0;
  if (bwd_floats(p, N, F, false) * 4 > (size_t)SMEM_LIMIT) return 0;
  const int G = p.G, grid = p.grid;
  S2VBwd a;
  a.X = X; a.Wcat = Wcat; a.Wq = Wq; a.we = we; a.saved = saved; a.dout = dout; a.c0 = c0;
  a.dX = dX; a.dqS = dqS; a.pwS = pwS; a.dpreS = dpreS; a.dm0 = dm0; a.dc0 = dc0; a.slots = (uint4*)slots;
  a.B = B; a.N = N; a.F = F; a.steps = steps; a.G = G; a.KS = p.KS; a.GS = p.GS; a.dbg = g_dbg;
  a.fast = p.GS >= G && (N + p.KS - 1) / p.KS <= 8;   // one graph per thread, its rows fit the register path
  a.x_smem = bwd_floats(p, N, F, true) * 4 <= (size_t)SMEM_LIMIT;
  const size_t smem = bwd_floats(p, N, F, a.x_smem) * 4;
  *rc = MPNN_ERR_CUDA;
  if (cudaFuncSetAttribute(k_s2v_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    mpnn_set_error("set2vec_bwd: cannot reserve %zu bytes of shared memory", smem);
    return 1;
  }
  if (cudaMemsetAsync(slots, 0, 2 * (size_t)grid * sizeof(uint4), stream) != cudaSuccess ||
      (!a.x_smem && cudaMemsetAsync(dX, 0, (size_t)B * N * F * sizeof(float), stream) != cudaSuccess)) {
    mpnn_set_error("set2vec_bwd: memset failed");
    return 1;
  }
  void* args[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_s2v_bwd, dim3(grid), dim3(NT), args, smem, stream);
  if (e != cudaSuccess) {
    mpnn_set_error("set2vec_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  *rc = MPNN_OK;
  return 1;
}

extern "C" {
// 1 (default): Set2Vec's loop runs in the persistent kernels where they serve the shape; 0: per-step launches.  Returns
// the previous setting.
int mpnn_set2vec_set_persistent(int enabled) {
  const int prev = g_persist;
  g_persist = enabled ? 1 : 0;
  return prev;
}
// profiling aid: CTA 0 writes %globaltimer stamps of the phases of its first four iterations into buf (448 x u64:
// forward at [0,64), backward at [64,128), 16 per iteration); NULL switches it off
void mpnn_set2vec_debug(unsigned long long* buf) { g_dbg = buf; }
}
