// "Typed" message path: the edge network evaluated once per DISTINCT bond row (csrc/dedup.cu), its last Linear
// (edge_network.py:21) folded into a table of U+1 small matrices, and the message function + neighbour
// aggregation (edge_network.py:42-52 + adjacent_message_agg.py:18) as one gather kernel over the CSR:
//
//     T[u][l][k] = B_last[k*nf+l] + sum_p W_last[k*nf+l, p] * x_u[p]          (x_u = trunk output of distinct row u)
//     M[i, k]    = sum_{e in E(i)} alpha_e * sum_l T[uid_e][l][k] * H[src_e, l]
//                  (+ sum_l T[zero][l][k] * (S_b[l] - sum_{e in E(i)} H[src_e, l]) + beta[k]      HEAD form)
//
// This is the same contraction as csrc/message.cu (edge-embedding (x) neighbour-state against the shared weight),
// with the contraction over p hoisted out of the per-edge work: exact, because x depends on the bond row only.
// Per-edge d x d matrices are never formed; the table has one matrix per distinct bond row (a few dozen).
// Feature widths up to 32 run here on CUDA cores (HBM/L1-bound gather); wider states use csrc/tc_message.cu (tcgen05).
//
// All reductions have a fixed order (CSC lists, type-sorted chunks): results are bit-reproducible.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// per-graph column sums: out[b, c] = sum_i X[b, i, c]
// ---------------------------------------------------------------------------------------------------
__global__ void k_graph_sum(const float* __restrict__ X, int B, int N, int w, float* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * w) return;
  int b = t / w, c = t - b * w;
  const float* p = X + (size_t)b * N * w + c;
  float s = 0.f;
  for (int i = 0; i < N; ++i) s += p[(size_t)i * w];
  out[t] = s;
}

struct TMsg {
  const int* row_ptr;   // CSR by receiver [n_rows+1]
  const int* edge_src;  // [E]
  const int* edge_dst;  // [E]
  const int* uid;       // [E]
  const float* alpha;   // [E] or null (1)
  const float* H;       // [n_rows, nf] sender states
  const float* table;   // [(ucap+1)][DP][DP]  T[u][l][k]
  const float* tableT;  // [(ucap+1)][DP][DP]  T[u][k][l]
  const float* S;       // [B, nf] per-graph sums of H (HEAD form) or null
  const float* beta;    // [mf] or null
  int n_rows, N, nf, mf, zero_type;
  int n_src;            // rows of H (== n_rows when H holds node states; the edge count for per-edge sender vectors)
};

template <int DP>
__device__ __forceinline__ uint32_t group_mask(int lane) {
  if constexpr (DP == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << DP) - 1u) << ((lane / DP) * DP);
  }
}

// ---------------------------------------------------------------------------------------------------
// forward: one group of DP lanes per receiver row, lane k owns M[i, k]
// ---------------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(256) k_tmsg_fwd(TMsg a, float* __restrict__ M) {
  const int lane = threadIdx.x & 31;
  const int k = lane % DP;
  const uint32_t gm = group_mask<DP>(lane);
  const int groups_per_block = 256 / DP;
  for (int i = blockIdx.x * groups_per_block + threadIdx.x / DP; i < a.n_rows; i += gridDim.x * groups_per_block) {
    const int eb = a.row_ptr[i], ee = a.row_ptr[i + 1];
    float acc = 0.f, hs = 0.f;
    for (int e0 = eb; e0 < ee; e0 += DP) {
      // lane t of the group fetches the metadata of edge e0 + t (coalesced), then the group walks the edges
      const int cnt = min(DP, ee - e0);
      int jm = 0, um = 0;
      float am = 1.f;
      if (k < cnt) {
        jm = __ldg(a.edge_src + e0 + k);
        um = __ldg(a.uid + e0 + k);
        if (a.alpha) am = __ldg(a.alpha + e0 + k);
      }
      int j = __shfl_sync(gm, jm, 0, DP);
      float hj = k < a.nf ? __ldg(a.H + (size_t)j * a.nf + k) : 0.f;
      for (int t = 0; t < cnt; ++t) {
        const int u = __shfl_sync(gm, um, t, DP);
        const float al = __shfl_sync(gm, am, t, DP);
        const float hcur = hj;
        if (t + 1 < cnt) {  // prefetch the next sender state while this edge is contracted
          j = __shfl_sync(gm, jm, t + 1, DP);
          hj = k < a.nf ? __ldg(a.H + (size_t)j * a.nf + k) : 0.f;
        }
        hs += hcur;
        const float* T = a.table + (size_t)u * DP * DP + k;
        float tv[DP];
#pragma unroll
        for (int l = 0; l < DP; ++l) tv[l] = __ldg(T + l * DP);   // DP independent loads (padding rows are zero)
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int l = 0; l < DP; l += 2) {
          s0 = fmaf(tv[l], __shfl_sync(gm, hcur, l, DP), s0);
          s1 = fmaf(tv[l + 1], __shfl_sync(gm, hcur, l + 1, DP), s1);
        }
        acc = fmaf(al, s0 + s1, acc);
      }
    }
    if (a.S) {  // all non-bonded pairs of the row share the zero bond row (HEAD form, edge_network.py:50)
      const int b = i / a.N;
      const float q = (k < a.nf ? __ldg(a.S + (size_t)b * a.nf + k) : 0.f) - hs;
      const float* T = a.table + (size_t)a.zero_type * DP * DP + k;
      float tv[DP];
#pragma unroll
      for (int l = 0; l < DP; ++l) tv[l] = __ldg(T + l * DP);
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < DP; ++l) s = fmaf(tv[l], __shfl_sync(gm, q, l, DP), s);
      acc += s;
    }
    if (k < a.mf) M[(size_t)i * a.mf + k] = acc + (a.beta ? a.beta[k] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------
// backward w.r.t. the sender states: one group per sender row j (CSC), lane l owns dH[j, l]
//   dH[j,l] = sum_{e: src_e = j} alpha_e sum_k (T[u_e] - [HEAD] T[zero])[l][k] dM[dst_e, k]  (+ [HEAD] dS_b[l])
// ---------------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(256) k_tmsg_bwd_src(TMsg a, const int* __restrict__ col_ptr,
                                                      const int* __restrict__ csc_eid, const float* __restrict__ dM,
                                                      const float* __restrict__ Dsum /*[B, mf] or null*/,
                                                      float* __restrict__ dH) {
  const int lane = threadIdx.x & 31;
  const int l = lane % DP;
  const uint32_t gm = group_mask<DP>(lane);
  const int groups_per_block = 256 / DP;
  const bool head = a.S != nullptr;
  const float* T0 = a.tableT + (size_t)a.zero_type * DP * DP + l;
  for (int j = blockIdx.x * groups_per_block + threadIdx.x / DP; j < a.n_src; j += gridDim.x * groups_per_block) {
    const int cb = col_ptr[j], ce = col_ptr[j + 1];
    float acc = 0.f;
    for (int c0 = cb; c0 < ce; c0 += DP) {
      const int cnt = min(DP, ce - c0);
      int im = 0, um = 0;
      float am = 1.f;
      if (l < cnt) {
        const int e = __ldg(csc_eid + c0 + l);
        im = __ldg(a.edge_dst + e);
        um = __ldg(a.uid + e);
        if (a.alpha) am = __ldg(a.alpha + e);
      }
      int i = __shfl_sync(gm, im, 0, DP);
      float dmn = l < a.mf ? __ldg(dM + (size_t)i * a.mf + l) : 0.f;
      for (int t = 0; t < cnt; ++t) {
        const int u = __shfl_sync(gm, um, t, DP);
        const float al = __shfl_sync(gm, am, t, DP);
        const float dm = dmn;
        if (t + 1 < cnt) {
          i = __shfl_sync(gm, im, t + 1, DP);
          dmn = l < a.mf ? __ldg(dM + (size_t)i * a.mf + l) : 0.f;
        }
        const float* T = a.tableT + (size_t)u * DP * DP + l;
        float tv[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) tv[k] = __ldg(T + k * DP);
        if (head) {
#pragma unroll
          for (int k = 0; k < DP; ++k) tv[k] -= __ldg(T0 + k * DP);
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < DP; k += 2) {
          s0 = fmaf(tv[k], __shfl_sync(gm, dm, k, DP), s0);
          s1 = fmaf(tv[k + 1], __shfl_sync(gm, dm, k + 1, DP), s1);
        }
        acc = fmaf(al, s0 + s1, acc);
      }
    }
    if (head) {
      const int b = j / a.N;
      const float ds = l < a.mf ? __ldg(Dsum + (size_t)b * a.mf + l) : 0.f;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DP; ++k) s = fmaf(__ldg(T0 + k * DP), __shfl_sync(gm, ds, k, DP), s);
      acc += s;
    }
    if (l < a.nf) dH[(size_t)j * a.nf + l] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------
// backward w.r.t. the table, level 1: chunks of CH edges in type-sorted order; partial of the (chunk c,
// type u) pair goes to slot c + u (pairs are strictly increasing in c + u along the sorted list).
//   dT[u][l][k] = sum_{e in type u} alpha_e H[src_e, l] dM[dst_e, k]
// ---------------------------------------------------------------------------------------------------
constexpr int SB = 64;   // edges staged per round (chunk sizes are multiples of SB)

template <int DP>
__global__ void __launch_bounds__(256) k_tmsg_bwd_table(TMsg a, const int* __restrict__ counts, int cap, int ch,
                                                        const int* __restrict__ type_eid, const float* __restrict__ dM,
                                                        float* __restrict__ part, long long dm_stride,
                                                        long long part_stride) {
  constexpr int OPT = (DP * DP + 255) / 256;  // outputs per thread (DP=32: 4, else 1)
  dM += (size_t)blockIdx.y * dm_stride;       // blockIdx.y: message-passing step (several gradients, one launch)
  part += (size_t)blockIdx.y * part_stride;
  __shared__ float g[SB][DP + 1];
  __shared__ float m[SB][DP + 1];
  __shared__ int ty[SB];
  const int E = min(counts[0], cap);
  const int p0 = blockIdx.x * ch;
  if (p0 >= E) return;
  const int p1 = min(p0 + ch, E);
  const int tid = threadIdx.x;
  // thread -> (l, k0..k0+OPT): consecutive threads walk k first
  const int o0 = tid * OPT;
  const int l = o0 / DP, k0 = o0 % DP;
  const bool active = o0 < DP * DP;
  float acc[OPT];
#pragma unroll
  for (int q = 0; q < OPT; ++q) acc[q] = 0.f;
  int cur = -1;
  for (int base = p0; base < p1; base += SB) {
    const int nb = min(SB, p1 - base);
    __syncthreads();
    for (int idx = tid; idx < nb * DP; idx += 256) {
      const int s = idx / DP, c = idx - s * DP;
      const int e = type_eid[base + s];
      const float al = a.alpha ? a.alpha[e] : 1.f;
      g[s][c] = c < a.nf ? al * a.H[(size_t)a.edge_src[e] * a.nf + c] : 0.f;
      m[s][c] = c < a.mf ? dM[(size_t)a.edge_dst[e] * a.mf + c] : 0.f;
      if (c == 0) ty[s] = a.uid[e];
    }
    __syncthreads();
    if (active) {
      for (int s = 0; s < nb; ++s) {
        const int u = ty[s];
        if (u != cur) {
          if (cur >= 0) {
            float* o = part + (size_t)(blockIdx.x + cur) * DP * DP + o0;
#pragma unroll
            for (int q = 0; q < OPT; ++q) {
              o[q] = acc[q];
              acc[q] = 0.f;
            }
          }
          cur = u;
        }
        const float gv = g[s][l];
#pragma unroll
        for (int q = 0; q < OPT; ++q) acc[q] = fmaf(gv, m[s][k0 + q], acc[q]);
      }
    }
  }
  if (active && cur >= 0) {
    float* o = part + (size_t)(blockIdx.x + cur) * DP * DP + o0;
#pragma unroll
    for (int q = 0; q < OPT; ++q) o[q] = acc[q];
  }
}

// level 2: one block per type; sums the type's chunk partials in chunk order.  dT [(ucap+1)][DP][DP].
template <int DP>
__global__ void __launch_bounds__(256) k_tmsg_bwd_table_reduce(const int* __restrict__ type_ptr,
                                                               const float* __restrict__ part, int zero_type, int ch,
                                                               float* __restrict__ dT, long long part_stride,
                                                               long long dt_stride) {
  const int u = blockIdx.x;
  part += (size_t)blockIdx.y * part_stride;
  dT += (size_t)blockIdx.y * dt_stride;
  float* out = dT + (size_t)u * DP * DP;
  int b = 0, e = 0;
  if (u < zero_type) {
    b = type_ptr[u];
    e = type_ptr[u + 1];
  }
  for (int o = threadIdx.x; o < DP * DP; o += 256) {
    float s = 0.f;
    if (e > b) {
      const int c0 = b / ch, c1 = (e - 1) / ch;
      const float* p = part + (size_t)(c0 + u) * DP * DP + o;
#pragma unroll 4
      for (int c = c0; c <= c1; ++c, p += DP * DP) s += *p;
    }
    out[o] = s;
  }
}

// HEAD form: dT[zero][l][k] = sum_b S[b,l] Dsum[b,k] - sum_{u < zero} dT[u][l][k]     (one block)
template <int DP>
__global__ void __launch_bounds__(256) k_tmsg_bwd_table_zero(const float* __restrict__ S, const float* __restrict__ Dsum,
                                                             int B, int nf, int mf, int zero_type,
                                                             float* __restrict__ dT) {
  for (int o = threadIdx.x; o < DP * DP; o += 256) {
    const int l = o / DP, k = o % DP;
    float s = 0.f;
    if (l < nf && k < mf)
      for (int b = 0; b < B; ++b) s = fmaf(S[(size_t)b * nf + l], Dsum[(size_t)b * mf + k], s);
    float t = 0.f;
    for (int u = 0; u < zero_type; ++u) t += dT[(size_t)u * DP * DP + o];
    dT[(size_t)zero_type * DP * DP + o] = s - t;
  }
}

// ---------------------------------------------------------------------------------------------------
// table <-> flat last-layer output.  flat[u, k*nf + l] (what Linear(P, nf*mf) returns, edge_network.py:21,37)
// ---------------------------------------------------------------------------------------------------
__global__ void k_table_from_flat(const float* __restrict__ flat, int R, int nf, int mf, int DP,
                                  float* __restrict__ table, float* __restrict__ tableT) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * DP * DP) return;
  int k = (int)(t % DP), l = (int)((t / DP) % DP);
  int u = (int)(t / ((long long)DP * DP));
  float v = (k < mf && l < nf) ? flat[(size_t)u * mf * nf + (size_t)k * nf + l] : 0.f;
  table[t] = v;
  tableT[((size_t)u * DP + k) * DP + l] = v;
}

__global__ void k_table_to_flat(const float* __restrict__ dT, int R, int nf, int mf, int DP,
                                float* __restrict__ dflat) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * mf * nf) return;
  int l = (int)(t % nf), k = (int)((t / nf) % mf);
  int u = (int)(t / ((long long)mf * nf));
  dflat[t] = dT[((size_t)u * DP + l) * DP + k];
}

// ===================================================================================================
// Fused edge network on the distinct rows, P <= 64:  growth layers -> 50 tied layers -> table.
// One distinct row per CTA at a time (the rows are few and the 50 layers are a serial chain, so the kernel is
// latency-bound: every dot product is split over 4 lanes and finished with two shuffles, one barrier per
// layer).  The tied weight stays in shared memory for all layers.
// ===================================================================================================
constexpr int PW = 64;    // max padded trunk width handled here
constexpr int MAXG = 4;

struct ENet {
  const float* rows;    // [R, ef]
  const float* gw[MAXG];
  const float* gb[MAXG];
  const float* w_tied;  // [P, P]
  const float* w_last;  // [mf*nf, P]
  const float* b_last;  // [mf*nf]
  int R, ef, G, P, L, nf, mf, DP;
  int gin[MAXG], gout[MAXG];
};

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// saved layout (floats): acts[(G + L + 1)][R][PW] : slot 0 = input rows (zero padded), 1..G growth outputs,
// G+1..G+L tied outputs (slot G+L = x)
__device__ __forceinline__ void enet_fwd_body(const ENet& n, int fch, float* __restrict__ acts,
                                              float* __restrict__ table, float* __restrict__ tableT) {
  extern __shared__ __align__(16) float Wl[];   // [fch][P | 1] staged rows of the last Linear
  __shared__ __align__(16) float A[2][PW];
  int staged_f0 = -1;
  const int tid = threadIdx.x;
  const int o = tid >> 2, q = tid & 3;
  const int P = n.P;
  // A table that fits one chunk: the last Linear's weight goes to shared memory with one wave of async copies issued
  // HERE, so its L2 / HBM latency hides behind the serial chain of tied layers instead of following it.
  bool pre_staged = false;
  if (n.mf * n.nf <= fch && blockIdx.x < n.R) {
    const int tot = n.mf * n.nf * P, LPp = P | 1;
    if (LPp == P && (tot & 3) == 0 && (reinterpret_cast<size_t>(n.w_last) & 15) == 0) {
      for (int i4 = tid; i4 < tot / 4; i4 += 256) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(Wl + i4 * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(n.w_last + i4 * 4) : "memory");
      }
    } else {
      for (int i = tid; i < tot; i += 256) {
        const int f = i / P, pp = i - f * P;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(Wl + f * LPp + pp);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(n.w_last + i) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pre_staged = true;
    staged_f0 = 0;
  }
  // this thread's slice of row o of the tied weight, kept in REGISTERS for all layers and rows:
  // wreg[4j + t] = W[o][16j + 4q + t]   (the 16-byte chunk q of every 64-byte group: conflict-free LDS.128)
  float wreg[16];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = 16 * j + 4 * q + t;
      const float w = __ldg(n.w_tied + (size_t)min(o, P - 1) * P + min(i, P - 1));
      wreg[4 * j + t] = (o < P && i < P) ? w : 0.f;
    }
  for (int row = blockIdx.x; row < n.R; row += gridDim.x) {
    __syncthreads();  // previous row done with A
    if (tid < PW) {
      float v = tid < n.ef ? n.rows[(size_t)row * n.ef + tid] : 0.f;
      A[0][tid] = v;
      acts[((size_t)0 * n.R + row) * PW + tid] = v;
    }
    __syncthreads();
    int cur = 0, slot = 1;
    for (int g = 0; g < n.G; ++g, ++slot) {
      float acc = 0.f;
      if (o < n.gout[g]) {
        const float* w = n.gw[g] + (size_t)o * n.gin[g];
        for (int i = q; i < n.gin[g]; i += 4) acc = fmaf(__ldg(w + i), A[cur][i], acc);
      }
      acc = quad_sum(acc);
      if (q == 0) {
        acc = o < n.gout[g] ? fmaxf(acc + n.gb[g][o], 0.f) : 0.f;
        A[cur ^ 1][o] = acc;
        acts[((size_t)slot * n.R + row) * PW + o] = acc;
      }
      __syncthreads();
      cur ^= 1;
    }
    for (int l = 0; l < n.L; ++l, ++slot) {
      const float4* ap = reinterpret_cast<const float4*>(&A[cur][4 * q]);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a4 = ap[4 * j];
        a0 = fmaf(wreg[4 * j + 0], a4.x, a0);
        a1 = fmaf(wreg[4 * j + 1], a4.y, a1);
        a0 = fmaf(wreg[4 * j + 2], a4.z, a0);
        a1 = fmaf(wreg[4 * j + 3], a4.w, a1);
      }
      float acc = quad_sum(a0 + a1);
      if (q == 0) {
        acc = fmaxf(acc, 0.f);
        A[cur ^ 1][o] = acc;
        acts[((size_t)slot * n.R + row) * PW + o] = acc;
      }
      __syncthreads();
      cur ^= 1;
    }
    // table of this distinct row: T[u][l][k] = B[k*nf+l] + W_last[k*nf+l, :] . x_u
    // W_last is staged through shared memory in chunks of `fch` output rows (coalesced, all loads in flight),
    // then one thread per output walks its row (odd row stride: conflict-free) against the broadcast x_u.
    const int DP = n.DP;
    const int warp = tid >> 5, lane = tid & 31;
    const int nout = n.mf * n.nf;
    const int LP = P | 1;
    if (pre_staged) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      pre_staged = false;
    }
    for (int f0 = 0; f0 < nout; f0 += fch) {
      const int cnt = min(fch, nout - f0);
      if (f0 != staged_f0) {     // single-chunk tables are staged once per CTA
        __syncthreads();
        for (int fb = warp * 8; fb < cnt; fb += 64) {
          float w0[8], w1[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float* w = n.w_last + (size_t)(f0 + min(fb + j, cnt - 1)) * P;
            w0[j] = __ldg(w + min(lane, P - 1));
            w1[j] = __ldg(w + min(lane + 32, P - 1));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (fb + j < cnt) {
              if (lane < P) Wl[(fb + j) * LP + lane] = w0[j];
              if (lane + 32 < P) Wl[(fb + j) * LP + lane + 32] = w1[j];
            }
          }
        }
        staged_f0 = f0;
        __syncthreads();
      }
      for (int t = tid; t < cnt; t += 256) {
        const int f = f0 + t;
        const float* wr = Wl + t * LP;
        float a0 = n.b_last[f], a1 = 0.f;
        int pp = 0;
        for (; pp + 1 < P; pp += 2) {
          a0 = fmaf(wr[pp], A[cur][pp], a0);
          a1 = fmaf(wr[pp + 1], A[cur][pp + 1], a1);
        }
        if (pp < P) a0 = fmaf(wr[pp], A[cur][pp], a0);
        const float v = a0 + a1;
        const int k = f / n.nf, l = f - k * n.nf;
        table[((size_t)row * DP + l) * DP + k] = v;
        tableT[((size_t)row * DP + k) * DP + l] = v;
      }
    }
    // zero the padding of the table (feature widths that are not a power of two)
    if (n.nf < DP || n.mf < DP) {
      for (int e = tid; e < DP * DP; e += 256) {
        const int l = e / DP, k = e - l * DP;
        if (l >= n.nf || k >= n.mf) {
          table[((size_t)row * DP + l) * DP + k] = 0.f;
          tableT[((size_t)row * DP + k) * DP + l] = 0.f;
        }
      }
    }
  }
}

// backward of the fused edge network.  Per distinct row: dx = dT . W_last, the tied layers (dW_tied partial in
// registers: thread owns a 4 x 4 micro-tile of the 64 x 64 gradient, accumulated over the CTA's rows), the growth
// layers; partials per CTA.
// partial layout per CTA (floats): [PW*PW tied] then for g = G-1 .. 0: [gout*gin weights][gout bias]
__device__ __forceinline__ void enet_bwd_body(const ENet& n, const float* __restrict__ acts,
                                              const float* __restrict__ dT, float* __restrict__ partial,
                                              int partial_stride, float* __restrict__ d_rows /*[R, ef] or null*/,
                                              int stage_wl /*1: W_last + the row's dT staged in shared memory*/) {
  __shared__ __align__(16) float D[2][PW];       // delta of the current layer
  __shared__ __align__(16) float Ap[2][PW];      // input activation of the current layer
  __shared__ __align__(16) float dAs[PW];        // un-masked gradient w.r.t. the tied input
  __shared__ float red4[4][PW];
  extern __shared__ __align__(16) float As[];    // [G + L + 1][PW]: every saved activation of the current row
  const int tid = threadIdx.x;
  const int i_ = tid >> 2, q = tid & 3;          // (output index, split lane) of the dA dot products
  const int og = tid >> 4, ig = tid & 15;        // dW micro-tile
  const int P = n.P, DP = n.DP;
  // wreg[4j + t] = W[16j + 4q + t][i_]: this thread's slice of COLUMN i_ of the tied weight, in registers
  float wreg[16];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int oo = 16 * j + 4 * q + t;
      const float w = __ldg(n.w_tied + (size_t)min(oo, P - 1) * P + min(i_, P - 1));
      wreg[4 * j + t] = (oo < P && i_ < P) ? w : 0.f;
    }
  float accW[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) accW[a][b] = 0.f;
  float* part = partial + (size_t)blockIdx.x * partial_stride;
  // growth partials accumulate over the CTA's rows in global memory (own slice): the first row writes them, a CTA
  // without rows zeroes them
  if ((int)blockIdx.x >= n.R)
    for (int e = PW * PW + tid; e < partial_stride; e += 256) part[e] = 0.f;
  const int nslots = n.G + n.L + 1;
  // W_last [mf*nf, P] -> shared memory behind the saved activations with one wave of async copies (the dx pre-pass below
  // would otherwise walk it in eight dependent batches of L2 loads per row)
  float* Wls = As + (size_t)nslots * PW;
  float* dTs = Wls + (((size_t)n.mf * n.nf * P + 3) & ~(size_t)3);
  if (stage_wl && (int)blockIdx.x < n.R) {
    const int tot = n.mf * n.nf * P;
    if ((tot & 3) == 0 && (reinterpret_cast<size_t>(n.w_last) & 15) == 0) {
      for (int i4 = tid; i4 < tot / 4; i4 += 256) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(Wls + i4 * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(n.w_last + i4 * 4) : "memory");
      }
    } else {
      for (int i = tid; i < tot; i += 256) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(Wls + i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(n.w_last + i) : "memory");
      }
    }
  }
  for (int row = blockIdx.x; row < n.R; row += gridDim.x) {
    const bool first_row = row == (int)blockIdx.x;
    __syncthreads();
    // all saved activations of the row -> shared memory with one wave of 16-byte async copies (the per-layer
    // loads of the 52-layer chain would otherwise pay one L2/HBM latency each); overlapped with the dx pre-pass
    for (int idx = tid; idx < nslots * (PW / 4); idx += 256) {
      const int slot = idx / (PW / 4), c4 = idx - slot * (PW / 4);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(As + slot * PW + c4 * 4);
      const float* src = acts + ((size_t)slot * n.R + row) * PW + c4 * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    if (stage_wl)
      for (int i4 = tid; i4 < DP * DP / 4; i4 += 256) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(dTs + i4 * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(dT + (size_t)row * DP * DP + i4 * 4)
                     : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // dx[p] = sum_f dT[row][l][k] W_last[f, p]   (f = k*nf + l), f split over the 4 warp pairs
    if (stage_wl) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const int p = tid & 63, part_f = tid >> 6;
      float acc0 = 0.f, acc1 = 0.f;
      if (p < P) {
        const int nout = n.mf * n.nf;
        int k = part_f / n.nf, l = part_f - k * n.nf;
        int f = part_f;
#pragma unroll 4
        for (; f + 4 < nout; f += 8) {
          const float t0 = dTs[l * DP + k];
          const float w0 = Wls[(size_t)f * P + p];
          l += 4;
          while (l >= n.nf) {
            l -= n.nf;
            ++k;
          }
          const float t1 = dTs[l * DP + k];
          const float w1 = Wls[(size_t)(f + 4) * P + p];
          l += 4;
          while (l >= n.nf) {
            l -= n.nf;
            ++k;
          }
          acc0 = fmaf(t0, w0, acc0);
          acc1 = fmaf(t1, w1, acc1);
        }
        for (; f < nout; f += 4) {
          acc0 = fmaf(dTs[l * DP + k], Wls[(size_t)f * P + p], acc0);
          l += 4;
          while (l >= n.nf) {
            l -= n.nf;
            ++k;
          }
        }
      }
      red4[part_f][p] = acc0 + acc1;
    } else {
      const int p = tid & 63, part_f = tid >> 6;
      float acc = 0.f;
      if (p < P) {
        const float* dt = dT + (size_t)row * DP * DP;
        const int nout = n.mf * n.nf;
        int k = part_f / n.nf, l = part_f - k * n.nf;   // f = k*nf + l, advanced without divisions
#pragma unroll 8
        for (int f = part_f; f < nout; f += 4) {
          acc = fmaf(__ldg(dt + l * DP + k), __ldg(n.w_last + (size_t)f * P + p), acc);
          l += 4;
          while (l >= n.nf) {
            l -= n.nf;
            ++k;
          }
        }
      }
      red4[part_f][p] = acc;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (tid < PW) {
      const float g = (red4[0][tid] + red4[1][tid]) + (red4[2][tid] + red4[3][tid]);
      const float aout = As[(nslots - 1) * PW + tid];
      D[0][tid] = aout > 0.f ? g : 0.f;
      Ap[0][tid] = As[(nslots - 2) * PW + tid];
    }
    __syncthreads();
    int cur = 0;
    for (int l = n.L; l >= 1; --l) {
      // slot of this layer's input: G + l - 1; prefetch the input of the NEXT (lower) layer: slot G + l - 2
      float pre = 0.f;
      if (q == 0 && l >= 2) pre = As[(n.G + l - 2) * PW + i_];
      {
        float4 d4 = *reinterpret_cast<const float4*>(&D[cur][og * 4]);
        float4 a4 = *reinterpret_cast<const float4*>(&Ap[cur][ig * 4]);
        float dv[4] = {d4.x, d4.y, d4.z, d4.w};
        float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) accW[a][b] = fmaf(dv[a], av[b], accW[a][b]);
      }
      const float4* dp = reinterpret_cast<const float4*>(&D[cur][4 * q]);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 d4 = dp[4 * j];
        a0 = fmaf(wreg[4 * j + 0], d4.x, a0);
        a1 = fmaf(wreg[4 * j + 1], d4.y, a1);
        a0 = fmaf(wreg[4 * j + 2], d4.z, a0);
        a1 = fmaf(wreg[4 * j + 3], d4.w, a1);
      }
      const float g = quad_sum(a0 + a1);   // gradient w.r.t. this layer's input a_{l-1}[i_]
      if (q == 0) {
        if (l >= 2) {
          D[cur ^ 1][i_] = Ap[cur][i_] > 0.f ? g : 0.f;
          Ap[cur ^ 1][i_] = pre;
        } else {
          dAs[i_] = g;
        }
      }
      __syncthreads();
      cur ^= 1;
    }
    // growth layers, last to first; dAs = gradient w.r.t. the output of growth layer G-1 (or the input rows)
    size_t poff = (size_t)PW * PW;
    for (int g = n.G - 1; g >= 0; --g) {
      const int gin = n.gin[g], gout = n.gout[g];
      // delta and input activation of this layer
      if (tid < PW) {
        const float aout = As[(g + 1) * PW + tid];
        D[0][tid] = (tid < gout && aout > 0.f) ? dAs[tid] : 0.f;
        Ap[0][tid] = As[g * PW + tid];
      }
      __syncthreads();
      for (int e = tid; e < gout * gin; e += 256) {
        const int oo = e / gin, ii = e - oo * gin;
        const float v = D[0][oo] * Ap[0][ii];
        part[poff + e] = first_row ? v : part[poff + e] + v;
      }
      for (int e = tid; e < gout; e += 256) {
        const size_t at = poff + (size_t)gout * gin + e;
        part[at] = first_row ? D[0][e] : part[at] + D[0][e];
      }
      poff += (size_t)gout * gin + gout;
      float acc = 0.f;
      if (tid < gin) {
        const float* w = n.gw[g] + tid;
        for (int oo = 0; oo < gout; ++oo) acc = fmaf(D[0][oo], __ldg(w + (size_t)oo * gin), acc);
      }
      __syncthreads();
      if (tid < PW) dAs[tid] = tid < gin ? acc : 0.f;
      __syncthreads();
    }
    if (d_rows && tid < n.ef) d_rows[(size_t)row * n.ef + tid] = dAs[tid];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
    *reinterpret_cast<float4*>(part + (og * 4 + a) * PW + ig * 4) =
        make_float4(accW[a][0], accW[a][1], accW[a][2], accW[a][3]);
}

// Everything that follows k_enet_bwd, as ONE launch: blocks [0, nb_red) reduce the per-CTA partials (fixed order) and
// write each sum straight to its destination -- the tied weight's [P][P] gradient out of the padded [PW][PW] block, the
// growth layers' weight / bias gradients -- and blocks [nb_red, ...) compute the last Linear's gradient, which does not
// depend on the trunk backward at all.  (Was reduce + unpack + two copies per growth layer + last-layer kernel: five
// dependent launches at the tail of the backward pass' critical path.)
struct ENetDst {
  float* tied;
  float* gw[MAXG];
  float* gb[MAXG];
  int off[MAXG], wn[MAXG], bn[MAXG];   // region of growth layer g inside the reduced vector: [off, off+wn) | [.., +bn)
  int G, P;
};

__device__ __forceinline__ void enet_finish_body(const float* __restrict__ partial, int nparts, int stride,
                                                 const ENetDst& d, int nb_red, const float* __restrict__ acts_x,
                                                 const float* __restrict__ dT, int R, int nf, int mf, int DP,
                                                 float* __restrict__ dW, float* __restrict__ dB) {
  if ((int)blockIdx.x < nb_red) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= stride) return;
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < nparts; ++c) s += partial[(size_t)c * stride + idx];
    if (idx < PW * PW) {
      const int o = idx / PW, i = idx - o * PW;
      if (o < d.P && i < d.P) d.tied[o * d.P + i] = s;
      return;
    }
#pragma unroll
    for (int g = 0; g < MAXG; ++g) {
      if (g < d.G) {
        const int r = idx - d.off[g];
        if (r >= 0 && r < d.wn[g]) d.gw[g][r] = s;
        else if (r >= d.wn[g] && r < d.wn[g] + d.bn[g]) d.gb[g][r - d.wn[g]] = s;
      }
    }
    return;
  }
  const int P = d.P;
  const int idx = (blockIdx.x - nb_red) * blockDim.x + threadIdx.x;  // over (f, p), p fastest, p in [0, P]
  const int total = mf * nf * (P + 1);
  if (idx >= total) return;
  const int p = idx % (P + 1), f = idx / (P + 1);
  const int l = f % nf, k = f / nf;
  float s = 0.f;
  const float* dcol = dT + (size_t)l * DP + k;
  const float* xcol = acts_x + min(p, P - 1);
  for (int u0 = 0; u0 < R; u0 += 8) {
    float dv[8], xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // 16 independent clamped loads
      const int u = min(u0 + j, R - 1);
      dv[j] = __ldg(dcol + (size_t)u * DP * DP);
      xv[j] = __ldg(xcol + (size_t)u * PW);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (u0 + j < R) s = fmaf(dv[j], p < P ? xv[j] : 1.f, s);
  }
  if (p < P)
    dW[(size_t)f * P + p] = s;
  else
    dB[f] = s;
}

// ---- kernels: one edge network per launch, or K sibling networks (one EdgeNetwork per message-passing step,
// normed_basic_model.py:24-27: same layer plan and the same distinct bond rows, different weights) as ONE launch with
// the network index in blockIdx.y -- their 52-layer chains are independent and latency-bound, so K of them cost one.
constexpr int MAXNET = 8;

__global__ void __launch_bounds__(256) k_enet_fwd(ENet n, int fch, float* __restrict__ acts,
                                                  float* __restrict__ table, float* __restrict__ tableT) {
  enet_fwd_body(n, fch, acts, table, tableT);
}
__global__ void __launch_bounds__(256) k_enet_bwd(ENet n, const float* __restrict__ acts, const float* __restrict__ dT,
                                                  float* __restrict__ partial, int partial_stride,
                                                  float* __restrict__ d_rows, int stage_wl) {
  enet_bwd_body(n, acts, dT, partial, partial_stride, d_rows, stage_wl);
}
__global__ void k_enet_finish(const float* __restrict__ partial, int nparts, int stride, ENetDst d, int nb_red,
                              const float* __restrict__ acts_x, const float* __restrict__ dT, int R, int nf, int mf,
                              int DP, float* __restrict__ dW, float* __restrict__ dB) {
  enet_finish_body(partial, nparts, stride, d, nb_red, acts_x, dT, R, nf, mf, DP, dW, dB);
}

struct ENetFwdMulti {
  ENet n[MAXNET];
  float* acts[MAXNET];
  float* table[MAXNET];
  float* tableT[MAXNET];
};
struct ENetBwdMulti {
  ENet n[MAXNET];
  const float* acts[MAXNET];
  const float* dT[MAXNET];
  float* partial[MAXNET];
  float* d_rows[MAXNET];
};
struct ENetFinMulti {
  ENetDst d[MAXNET];
  const float* partial[MAXNET];
  const float* acts_x[MAXNET];
  const float* dT[MAXNET];
  float* dW[MAXNET];
  float* dB[MAXNET];
};

__global__ void __launch_bounds__(256) k_enet_fwd_multi(const __grid_constant__ ENetFwdMulti m, int fch) {
  const int y = blockIdx.y;
  enet_fwd_body(m.n[y], fch, m.acts[y], m.table[y], m.tableT[y]);
}
__global__ void __launch_bounds__(256) k_enet_bwd_multi(const __grid_constant__ ENetBwdMulti m, int partial_stride,
                                                        int stage_wl) {
  const int y = blockIdx.y;
  enet_bwd_body(m.n[y], m.acts[y], m.dT[y], m.partial[y], partial_stride, m.d_rows[y], stage_wl);
}
__global__ void k_enet_finish_multi(const __grid_constant__ ENetFinMulti m, int nparts, int stride, int nb_red, int R,
                                    int nf, int mf, int DP) {
  const int y = blockIdx.y;
  enet_finish_body(m.partial[y], nparts, stride, m.d[y], nb_red, m.acts_x[y], m.dT[y], R, nf, mf, DP, m.dW[y], m.dB[y]);
}

int pick_dp(int nf, int mf) {
  int d = nf > mf ? nf : mf;
  return pow2_at_least(d, 8);
}

// edges per chunk of the table-gradient pass: enough chunks to fill the machine, few enough partials to reduce
int table_chunk(int edge_capacity) {
  int ch = SB;
  while (ch < 1024 && (long long)ch * 4 * mpnn_num_sms() < edge_capacity) ch <<= 1;
  return ch;
}

int msg_grid(int n_rows, int DP) {
  int per = 256 / DP;
  int want = ceil_div(n_rows, per);
  int cap = mpnn_num_sms() * 8;
  return want < cap ? (want > 0 ? want : 1) : cap;
}

bool fill_enet(ENet* n, const float* rows, int R, int ef, int G, const float* const* gw, const float* const* gb,
               const float* w_tied, int P, int L, const float* w_last, const float* b_last, int nf, int mf) {
  if (G > MAXG || G < 0 || P > PW) return false;
  n->rows = rows;
  n->w_tied = w_tied;
  n->w_last = w_last;
  n->b_last = b_last;
  n->R = R;
  n->ef = ef;
  n->G = G;
  n->P = P;
  n->L = L;
  n->nf = nf;
  n->mf = mf;
  n->DP = pick_dp(nf, mf);
  int w = ef;
  for (int g = 0; g < MAXG; ++g) {
    n->gw[g] = nullptr;
    n->gb[g] = nullptr;
    n->gin[g] = n->gout[g] = 0;
  }
  for (int g = 0; g < G; ++g) {
    n->gw[g] = gw[g];
    n->gb[g] = gb ? gb[g] : nullptr;
    n->gin[g] = w;
    n->gout[g] = w * w;
    w = w * w;
  }
  return w == P && ef <= PW;
}

int enet_grid(int R) {
  int cap = 2 * mpnn_num_sms();
  return R < cap ? (R > 0 ? R : 1) : cap;
}

int enet_partial_stride(const ENet& n) {
  size_t s = (size_t)PW * PW;
  for (int g = 0; g < n.G; ++g) s += (size_t)n.gout[g] * n.gin[g] + n.gout[g];
  return (int)((s + 3) & ~(size_t)3);
}

}  // namespace

extern "C" {

// ---- table geometry -----------------------------------------------------------------------------------
// padded feature width of the table (power of two >= max(nf, mf), >= 8); -1 if this file cannot serve it
int mpnn_typed_dp(int nf, int mf) {
  int DP = pick_dp(nf, mf);
  return DP <= 32 ? DP : -1;
}

int mpnn_graph_sum(const float* X, int B, int N, int width, float* out, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && width > 0, MPNN_ERR_ARG, "graph_sum: bad dims");
  k_graph_sum<<<ceil_div((long long)B * width, 256), 256, 0, stream>>>(X, B, N, width, out);
  MPNN_CHECK_LAUNCH("k_graph_sum");
  return MPNN_OK;
}

// flat [R, mf*nf] (output of the last Linear on the distinct rows) -> table / tableT [R][DP][DP]
int mpnn_table_from_flat(const float* flat, int R, int nf, int mf, float* table, float* tableT, cudaStream_t stream) {
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(R > 0 && nf > 0 && mf > 0, MPNN_ERR_ARG, "table_from_flat: bad dims");
  k_table_from_flat<<<ceil_div((long long)R * DP * DP, 256), 256, 0, stream>>>(flat, R, nf, mf, DP, table, tableT);
  MPNN_CHECK_LAUNCH("k_table_from_flat");
  return MPNN_OK;
}

int mpnn_table_to_flat(const float* dT, int R, int nf, int mf, float* dflat, cudaStream_t stream) {
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(R > 0 && nf > 0 && mf > 0, MPNN_ERR_ARG, "table_to_flat: bad dims");
  k_table_to_flat<<<ceil_div((long long)R * mf * nf, 256), 256, 0, stream>>>(dT, R, nf, mf, DP, dflat);
  MPNN_CHECK_LAUNCH("k_table_to_flat");
  return MPNN_OK;
}

// ---- fused edge network on the distinct rows (P <= 64) -------------------------------------------------
int mpnn_enet_supported(int ef, int n_growth, int P) { return (P <= PW && ef <= PW && n_growth <= MAXG) ? 1 : 0; }
// widest padded table (DP) the fused kernel writes; wider tables go through the generic trunk + last Linear
int mpnn_enet_max_dp(void) { return 64; }

long long mpnn_enet_saved_floats(int R, int n_growth, int n_tied) {
  return (long long)(n_growth + n_tied + 1) * R * PW;
}

size_t mpnn_enet_workspace_bytes(int R, int ef, int n_growth, int P) {
  ENet n;
  const float* dummy[MAXG] = {nullptr, nullptr, nullptr, nullptr};
  if (!fill_enet(&n, nullptr, R, ef, n_growth, dummy, dummy, nullptr, P, 1, nullptr, nullptr, 1, 1)) return 0;
  size_t stride = (size_t)enet_partial_stride(n);
  return ((size_t)enet_grid(R) + 1) * stride * sizeof(float);
}

int mpnn_enet_fwd(const float* rows, int R, int ef, int n_growth, const float* const* growth_w,
                  const float* const* growth_b, const float* w_tied, int P, int n_tied, const float* w_last,
                  const float* b_last, int nf, int mf, float* saved, float* table, float* tableT,
                  cudaStream_t stream) {
  ENet n;
  MPNN_REQUIRE(R > 0 && n_tied >= 1 && nf > 0 && mf > 0, MPNN_ERR_ARG, "enet_fwd: bad dims");
  MPNN_REQUIRE(fill_enet(&n, rows, R, ef, n_growth, growth_w, growth_b, w_tied, P, n_tied, w_last, b_last, nf, mf),
               MPNN_ERR_UNSUPPORTED, "enet_fwd: layer plan ef=%d growth=%d P=%d not supported by the fused kernel", ef,
               n_growth, P);
  MPNN_REQUIRE(n.DP <= 64, MPNN_ERR_UNSUPPORTED, "enet_fwd: feature width > 64");
  int fch = nf * mf < 512 ? nf * mf : 512;
  size_t smem = (size_t)fch * (P | 1) * sizeof(float);
  MPNN_CUDA(cudaFuncSetAttribute(k_enet_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_enet_fwd<<<enet_grid(R), 256, smem, stream>>>(n, fch, saved, table, tableT);
  MPNN_CHECK_LAUNCH("k_enet_fwd");
  return MPNN_OK;
}

// dT [R][DP][DP] -> d_growth_w/b, d_w_tied, d_w_last, d_b_last (all written), optional d_rows [R, ef]
int mpnn_enet_bwd(const float* rows, int R, int ef, int n_growth, const float* const* growth_w, const float* w_tied,
                  int P, int n_tied, const float* w_last, int nf, int mf, const float* saved, const float* dT,
                  float* const* d_growth_w, float* const* d_growth_b, float* d_w_tied, float* d_w_last,
                  float* d_b_last, float* d_rows, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ENet n;
  MPNN_REQUIRE(R > 0 && n_tied >= 1 && nf > 0 && mf > 0, MPNN_ERR_ARG, "enet_bwd: bad dims");
  MPNN_REQUIRE(fill_enet(&n, rows, R, ef, n_growth, growth_w, nullptr, w_tied, P, n_tied, w_last, nullptr, nf, mf),
               MPNN_ERR_UNSUPPORTED, "enet_bwd: layer plan not supported by the fused kernel");
  MPNN_REQUIRE(workspace_bytes >= mpnn_enet_workspace_bytes(R, ef, n_growth, P), MPNN_ERR_WORKSPACE,
               "enet_bwd: workspace too small");
  const int stride = enet_partial_stride(n);
  const int nparts = enet_grid(R);
  float* partial = (float*)workspace;
  const size_t act_smem = (size_t)(n_growth + n_tied + 1) * PW * sizeof(float);
  MPNN_REQUIRE(act_smem <= 160 * 1024, MPNN_ERR_UNSUPPORTED, "enet_bwd: too many layers for the shared-memory stage");
  // the last Linear's weight and the row's table gradient ride behind the activations when they fit
  const size_t wl_smem = ((((size_t)nf * mf * P + 3) & ~(size_t)3) + (size_t)n.DP * n.DP) * sizeof(float);
  const int stage_wl = act_smem + wl_smem <= 160 * 1024 ? 1 : 0;
  const size_t bwd_smem_bytes = act_smem + (stage_wl ? wl_smem : 0);
  MPNN_CUDA(cudaFuncSetAttribute(k_enet_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes));
  k_enet_bwd<<<nparts, 256, bwd_smem_bytes, stream>>>(n, saved, dT, partial, stride, d_rows, stage_wl);
  MPNN_CHECK_LAUNCH("k_enet_bwd");
  ENetDst dst;
  memset(&dst, 0, sizeof(dst));
  dst.tied = d_w_tied;
  dst.G = n_growth;
  dst.P = P;
  size_t off = (size_t)PW * PW;
  for (int g = n_growth - 1; g >= 0; --g) {   // order of the partial vector: last growth layer first
    dst.gw[g] = d_growth_w[g];
    dst.gb[g] = d_growth_b[g];
    dst.off[g] = (int)off;
    dst.wn[g] = n.gout[g] * n.gin[g];
    dst.bn[g] = n.gout[g];
    off += (size_t)dst.wn[g] + dst.bn[g];
  }
  const float* x = saved + (size_t)(n_growth + n_tied) * R * PW;
  const int nb_red = ceil_div(stride, 256);
  const int nb_last = ceil_div((long long)mf * nf * (P + 1), 256);
  k_enet_finish<<<nb_red + nb_last, 256, 0, stream>>>(partial, nparts, stride, dst, nb_red, x, dT, R, nf, mf, n.DP,
                                                      d_w_last, d_b_last);
  MPNN_CHECK_LAUNCH("k_enet_finish");
  return MPNN_OK;
}

// ---- K sibling edge networks in one launch each way ---------------------------------------------------------------
// All K networks share the layer plan (ef, n_growth, P, n_tied, nf, mf) and the distinct rows; per-network arrays are
// HOST arrays of device pointers: growth_w / growth_b [K * n_growth] (network-major), the rest [K].
int mpnn_enet_max_nets(void) { return MAXNET; }

int mpnn_enet_fwd_multi(int K, const float* rows, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* const* growth_b, const float* const* w_tied, int P, int n_tied,
                        const float* const* w_last, const float* const* b_last, int nf, int mf, float* const* saved,
                        float* const* table, float* const* tableT, cudaStream_t stream) {
  MPNN_REQUIRE(K >= 1 && K <= MAXNET && R > 0 && n_tied >= 1 && nf > 0 && mf > 0, MPNN_ERR_ARG, "enet_fwd_multi: bad dims");
  ENetFwdMulti m;
  memset(&m, 0, sizeof(m));
  for (int k = 0; k < K; ++k) {
    MPNN_REQUIRE(fill_enet(&m.n[k], rows, R, ef, n_growth, growth_w + (size_t)k * n_growth,
                           growth_b + (size_t)k * n_growth, w_tied[k], P, n_tied, w_last[k], b_last[k], nf, mf),
                 MPNN_ERR_UNSUPPORTED, "enet_fwd_multi: layer plan ef=%d growth=%d P=%d not supported", ef, n_growth, P);
    m.acts[k] = saved[k];
    m.table[k] = table[k];
    m.tableT[k] = tableT[k];
  }
  MPNN_REQUIRE(m.n[0].DP <= 64, MPNN_ERR_UNSUPPORTED, "enet_fwd_multi: feature width > 64");
  int fch = nf * mf < 512 ? nf * mf : 512;
  size_t smem = (size_t)fch * (P | 1) * sizeof(float);
  MPNN_CUDA(cudaFuncSetAttribute(k_enet_fwd_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_enet_fwd_multi<<<dim3(enet_grid(R), K), 256, smem, stream>>>(m, fch);
  MPNN_CHECK_LAUNCH("k_enet_fwd_multi");
  return MPNN_OK;
}

// workspace: K x mpnn_enet_workspace_bytes.  d_rows [K] entries may be NULL (no gradient w.r.t. the distinct rows).
int mpnn_enet_bwd_multi(int K, const float* rows, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* const* w_tied, int P, int n_tied, const float* const* w_last, int nf, int mf,
                        const float* const* saved, const float* const* dT, float* const* d_growth_w,
                        float* const* d_growth_b, float* const* d_w_tied, float* const* d_w_last,
                        float* const* d_b_last, float* const* d_rows, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream) {
  MPNN_REQUIRE(K >= 1 && K <= MAXNET && R > 0 && n_tied >= 1 && nf > 0 && mf > 0, MPNN_ERR_ARG, "enet_bwd_multi: bad dims");
  const size_t ws1 = mpnn_enet_workspace_bytes(R, ef, n_growth, P);
  MPNN_REQUIRE(workspace_bytes >= (size_t)K * ws1, MPNN_ERR_WORKSPACE, "enet_bwd_multi: workspace too small");
  ENetBwdMulti m;
  ENetFinMulti f;
  memset(&m, 0, sizeof(m));
  memset(&f, 0, sizeof(f));
  const size_t act_smem = (size_t)(n_growth + n_tied + 1) * PW * sizeof(float);
  MPNN_REQUIRE(act_smem <= 160 * 1024, MPNN_ERR_UNSUPPORTED, "enet_bwd_multi: too many layers for the shared-memory stage");
  for (int k = 0; k < K; ++k) {
    MPNN_REQUIRE(fill_enet(&m.n[k], rows, R, ef, n_growth, growth_w + (size_t)k * n_growth, nullptr, w_tied[k], P, n_tied,
                           w_last[k], nullptr, nf, mf),
                 MPNN_ERR_UNSUPPORTED, "enet_bwd_multi: layer plan not supported by the fused kernel");
    m.acts[k] = saved[k];
    m.dT[k] = dT[k];
    m.partial[k] = (float*)((char*)workspace + (size_t)k * ws1);
    m.d_rows[k] = d_rows ? d_rows[k] : nullptr;
    ENetDst& dst = f.d[k];
    dst.tied = d_w_tied[k];
    dst.G = n_growth;
    dst.P = P;
    size_t off = (size_t)PW * PW;
    for (int g = n_growth - 1; g >= 0; --g) {
      dst.gw[g] = d_growth_w[(size_t)k * n_growth + g];
      dst.gb[g] = d_growth_b[(size_t)k * n_growth + g];
      dst.off[g] = (int)off;
      dst.wn[g] = m.n[k].gout[g] * m.n[k].gin[g];
      dst.bn[g] = m.n[k].gout[g];
      off += (size_t)dst.wn[g] + dst.bn[g];
    }
    f.partial[k] = m.partial[k];
    f.acts_x[k] = saved[k] + (size_t)(n_growth + n_tied) * R * PW;
    f.dT[k] = dT[k];
    f.dW[k] = d_w_last[k];
    f.dB[k] = d_b_last[k];
  }
  const int stride = enet_partial_stride(m.n[0]);
  const int nparts = enet_grid(R);
  const size_t wl_smem = ((((size_t)nf * mf * P + 3) & ~(size_t)3) + (size_t)m.n[0].DP * m.n[0].DP) * sizeof(float);
  const int stage_wl = act_smem + wl_smem <= 160 * 1024 ? 1 : 0;
  const size_t bwd_smem_bytes = act_smem + (stage_wl ? wl_smem : 0);
  MPNN_CUDA(cudaFuncSetAttribute(k_enet_bwd_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes));
  k_enet_bwd_multi<<<dim3(nparts, K), 256, bwd_smem_bytes, stream>>>(m, stride, stage_wl);
  MPNN_CHECK_LAUNCH("k_enet_bwd_multi");
  const int nb_red = ceil_div(stride, 256);
  const int nb_last = ceil_div((long long)mf * nf * (P + 1), 256);
  k_enet_finish_multi<<<dim3(nb_red + nb_last, K), 256, 0, stream>>>(f, nparts, stride, nb_red, R, nf, mf, m.n[0].DP);
  MPNN_CHECK_LAUNCH("k_enet_finish_multi");
  return MPNN_OK;
}

// ---- typed message + aggregation ------------------------------------------------------------------------
size_t mpnn_tmsg_bwd_workspace_bytes(int edge_capacity, int unique_capacity, int nf, int mf, int B) {
  int DP = pick_dp(nf, mf);
  size_t chunks = (size_t)ceil_div(edge_capacity > 0 ? edge_capacity : 1, table_chunk(edge_capacity));
  return (chunks + unique_capacity + 2) * DP * DP * sizeof(float) + align_up((size_t)B * mf * sizeof(float), 256);
}

// M[n_rows, mf].  S = per-graph sums of H ([B, nf], mpnn_graph_sum) selects the HEAD form (all pairs, zero-row
// type = zero_type); S == NULL is the documented per-pair form fused with AdjMsgAgg (alpha = adj value).
int mpnn_tmsg_fwd(const int* row_ptr, const int* edge_src, const int* uid, const float* alpha, const float* H,
                  const float* table, const float* S, const float* beta, int n_rows, int N, int nf, int mf,
                  int zero_type, float* M, cudaStream_t stream) {
  MPNN_REQUIRE(n_rows > 0 && N > 0 && nf > 0 && mf > 0, MPNN_ERR_ARG, "tmsg_fwd: bad dims");
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(DP <= 32, MPNN_ERR_UNSUPPORTED, "tmsg_fwd: feature width > 32 is served by the tensor-core path");
  TMsg a = {row_ptr, edge_src, nullptr, uid, alpha, H, table, nullptr, S, beta, n_rows, N, nf, mf, zero_type, n_rows};
  int grid = msg_grid(n_rows, DP);
  switch (DP) {
    case 8: k_tmsg_fwd<8><<<grid, 256, 0, stream>>>(a, M); break;
    case 16: k_tmsg_fwd<16><<<grid, 256, 0, stream>>>(a, M); break;
    default: k_tmsg_fwd<32><<<grid, 256, 0, stream>>>(a, M); break;
  }
  MPNN_CHECK_LAUNCH("k_tmsg_fwd");
  return MPNN_OK;
}

// dH [n_rows, nf], dT [(unique_capacity+1)][DP][DP] (both written).  counts = device {E, U, ..} of the edge list.
int mpnn_tmsg_bwd(const int* row_ptr, const int* col_ptr, const int* csc_eid, const int* edge_src, const int* edge_dst,
                  const int* uid, const int* type_ptr, const int* type_eid, const int* counts, const float* alpha,
                  const float* H, const float* table, const float* tableT, const float* S, int n_rows, int n_src_rows,
                  int B, int N, int nf, int mf, int edge_capacity, int unique_capacity, const float* dM, float* dH,
                  float* dT, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(n_rows > 0 && N > 0 && nf > 0 && mf > 0 && B > 0 && n_src_rows >= 0, MPNN_ERR_ARG, "tmsg_bwd: bad dims");
  const int n_src = n_src_rows > 0 ? n_src_rows : n_rows;
  MPNN_REQUIRE(!(S && n_src != n_rows), MPNN_ERR_UNSUPPORTED, "tmsg_bwd: the HEAD form needs node-state senders");
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(DP <= 32, MPNN_ERR_UNSUPPORTED, "tmsg_bwd: feature width > 32 is served by the tensor-core path");
  MPNN_REQUIRE(workspace_bytes >= mpnn_tmsg_bwd_workspace_bytes(edge_capacity, unique_capacity, nf, mf, B),
               MPNN_ERR_WORKSPACE, "tmsg_bwd: workspace too small");
  const int zero_type = unique_capacity;
  TMsg a = {row_ptr, edge_src, edge_dst, uid, alpha, H, table, tableT, S, nullptr, n_rows, N, nf, mf, zero_type, n_src};
  char* wp = (char*)workspace;
  float* Dsum = (float*)wp;
  wp += align_up((size_t)B * mf * sizeof(float), 256);
  float* part = (float*)wp;
  if (S) {
    int rc = mpnn_graph_sum(dM, B, N, mf, Dsum, stream);
    if (rc) return rc;
  }
  const int grid = msg_grid(n_src, DP);
  const int ch = table_chunk(edge_capacity);
  const int chunks = ceil_div(edge_capacity > 0 ? edge_capacity : 1, ch);
  // dH == NULL / dT == NULL skip that half: the sender-state gradient feeds the main backward chain, the table gradient
  // only the edge network's parameter gradients (the caller may enqueue the two halves on different streams)
#define MPNN_TMSG_BWD(DPV)                                                                                        \
  do {                                                                                                            \
    if (dH) k_tmsg_bwd_src<DPV><<<grid, 256, 0, stream>>>(a, col_ptr, csc_eid, dM, S ? Dsum : nullptr, dH);        \
    if (dT) {                                                                                                     \
      k_tmsg_bwd_table<DPV><<<chunks, 256, 0, stream>>>(a, counts, edge_capacity, ch, type_eid, dM, part, 0, 0);  \
      k_tmsg_bwd_table_reduce<DPV><<<unique_capacity + 1, 256, 0, stream>>>(type_ptr, part, zero_type, ch, dT, 0, \
                                                                            0);                                  \
      if (S) k_tmsg_bwd_table_zero<DPV><<<1, 256, 0, stream>>>(S, Dsum, B, nf, mf, zero_type, dT);                \
    }                                                                                                             \
  } while (0)
  switch (DP) {
    case 8: MPNN_TMSG_BWD(8); break;
    case 16: MPNN_TMSG_BWD(16); break;
    default: MPNN_TMSG_BWD(32); break;
  }
#undef MPNN_TMSG_BWD
  MPNN_CHECK_LAUNCH("k_tmsg_bwd");
  return MPNN_OK;
}

// table gradients of K message-passing steps that share the edge list and the sender states (documented per-pair form,
// no HEAD terms) as one launch pair: dM [K][n_rows][mf] -> dT [K][(unique_capacity+1)][DP][DP].
// workspace: K x mpnn_tmsg_bwd_workspace_bytes.
int mpnn_tmsg_bwd_table_multi(int K, const int* edge_src, const int* edge_dst, const int* uid, const int* type_ptr,
                              const int* type_eid, const int* counts, const float* alpha, const float* H, int n_rows,
                              int nf, int mf, int edge_capacity, int unique_capacity, const float* dM, float* dT,
                              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(K >= 1 && n_rows > 0 && nf > 0 && mf > 0, MPNN_ERR_ARG, "tmsg_bwd_table_multi: bad dims");
  int DP = pick_dp(nf, mf);
  MPNN_REQUIRE(DP <= 32, MPNN_ERR_UNSUPPORTED, "tmsg_bwd_table_multi: feature width > 32");
  const size_t ws1 = mpnn_tmsg_bwd_workspace_bytes(edge_capacity, unique_capacity, nf, mf, 1);
  MPNN_REQUIRE(workspace_bytes >= (size_t)K * ws1, MPNN_ERR_WORKSPACE, "tmsg_bwd_table_multi: workspace too small");
  const int zero_type = unique_capacity;
  TMsg a = {nullptr, edge_src, edge_dst, uid, alpha, H, nullptr, nullptr, nullptr, nullptr, n_rows, 1, nf, mf, zero_type,
            n_rows};
  float* part = (float*)workspace;
  const long long part_stride = (long long)(ws1 / sizeof(float));
  const int ch = table_chunk(edge_capacity);
  const int chunks = ceil_div(edge_capacity > 0 ? edge_capacity : 1, ch);
  const long long dm_stride = (long long)n_rows * mf, dt_stride = (long long)(unique_capacity + 1) * DP * DP;
#define MPNN_TMSG_TM(DPV)                                                                                            \
  do {                                                                                                               \
    k_tmsg_bwd_table<DPV><<<dim3(chunks, K), 256, 0, stream>>>(a, counts, edge_capacity, ch, type_eid, dM, part,      \
                                                                dm_stride, part_stride);                             \
    k_tmsg_bwd_table_reduce<DPV><<<dim3(unique_capacity + 1, K), 256, 0, stream>>>(type_ptr, part, zero_type, ch, dT, \
                                                                                   part_stride, dt_stride);          \
  } while (0)
  switch (DP) {
    case 8: MPNN_TMSG_TM(8); break;
    case 16: MPNN_TMSG_TM(16); break;
    default: MPNN_TMSG_TM(32); break;
  }
#undef MPNN_TMSG_TM
  MPNN_CHECK_LAUNCH("k_tmsg_bwd_table_multi");
  return MPNN_OK;
}

}  // extern "C"
