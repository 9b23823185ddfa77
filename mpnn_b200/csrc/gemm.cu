// Small strided fp32 GEMM used for the plumbing GEMMs of the path (edge-map growth layers, GRU gate
// pre-activations, readout projections, weight gradients).  C[m,n] = sum_k A(m,k) * B(k,n) with
// arbitrary element strides, so NN / NT / TN are the same kernel.  K can be split across CTAs; the
// partial tiles are then summed in a fixed order (deterministic, no float atomics).
//
// The heavy contraction of the path (message step) does NOT go through this file.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS_PAD = 4;

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int M, N, K;
  long long sam, sak, sbk, sbn, ldc;
  int k_per_split;
  int flags;     // 1 relu, 2 accumulate into C
  float* partial;  // [splits][M][N] when splits > 1
};

__global__ void __launch_bounds__(256) k_gemm(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + LDS_PAD];
  __shared__ __align__(16) float Bs[BK][BN + LDS_PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  const bool a_kfast = (g.sak == 1);
  const bool b_nfast = (g.sbn == 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int it = 0; it < (BM * BK) / 256; ++it) {
      int i = tid + it * 256;
      int m, k;
      if (a_kfast) {
        k = i % BK;
        m = i / BK;
      } else {
        m = i % BM;
        k = i / BM;
      }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < kend) v = g.A[(long long)gm * g.sam + (long long)gk * g.sak];
      As[k][m] = v;
    }
#pragma unroll
    for (int it = 0; it < (BN * BK) / 256; ++it) {
      int i = tid + it * 256;
      int n, k;
      if (b_nfast) {
        n = i % BN;
        k = i / BN;
      } else {
        k = i % BK;
        n = i / BK;
      }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = g.B[(long long)gk * g.sbk + (long long)gn * g.sbn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.partial) {
        g.partial[((size_t)blockIdx.z * g.M + gm) * g.N + gn] = v;
      } else {
        if (g.bias) v += g.bias[gn];
        float* c = g.C + (long long)gm * g.ldc + gn;
        if (g.flags & 2) v += *c;
        if (g.flags & 1) v = fmaxf(v, 0.f);
        *c = v;
      }
    }
  }
}

__global__ void k_gemm_reduce(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ C,
                              long long ldc, const float* __restrict__ bias, int flags) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  int m = idx / N, n = idx - m * N;
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += partial[(size_t)s * M * N + idx];  // fixed order
  if (bias) v += bias[n];
  float* c = C + (long long)m * ldc + n;
  if (flags & 2) v += *c;
  if (flags & 1) v = fmaxf(v, 0.f);
  *c = v;
}

// ---------------------------------------------------------------------------------------------------
// Skinny GEMM: M <= 64 rows against a large [K, N] operand -- the wide edge-network trunk on the DISTINCT bond rows
// (P = 4096 when hidden >= 128: 50 layers of [~33, 4096] x [4096, 4096], edge_network.py:20) and its last Linear.
// The 64 x 64 tile kernel above wastes half its FMAs on M ~ 33 and runs 64 CTAs; here a CTA owns 64 output
// columns and ALL rows, the A tile is read as warp-broadcast float4 along k, B as float2 along n:
// (RPT + 4) shared loads per 8*RPT FMAs.  fp32 FFMA on purpose (the trunk must stay fp32-accurate).
// ---------------------------------------------------------------------------------------------------
constexpr int SK_BN = 64, SK_BK = 32;

template <int RPT>   // rows per thread: rows tm, tm+8, ... (tm = warp index), M <= 8*RPT
__global__ void __launch_bounds__(256) k_gemm_skinny(GemmArgs g) {
  __shared__ __align__(16) float As[2][8 * RPT][SK_BK + 4];
  __shared__ __align__(16) float Bs[2][SK_BK][SK_BN + 2];
  const int tid = threadIdx.x;
  const int tn = tid & 31, tm = tid >> 5;
  const int n0 = blockIdx.x * SK_BN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  const bool b_kfast = (g.sbk == 1);
  float acc[RPT][2];
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = 0.f;

  auto stage = [&](int buf, int k0) {
    // A tile [8*RPT rows][32 k]: k fastest (sak == 1 required)
    for (int i = tid; i < 8 * RPT * SK_BK; i += 256) {
      const int k = i % SK_BK, m = i / SK_BK;
      float v = 0.f;
      if (m < g.M && k0 + k < kend) v = __ldg(g.A + (long long)m * g.sam + (k0 + k));
      As[buf][m][k] = v;
    }
    // B tile [32 k][64 n]
    for (int i = tid; i < SK_BK * SK_BN; i += 256) {
      int k, n;
      if (b_kfast) {
        k = i % SK_BK;
        n = i / SK_BK;
      } else {
        n = i % SK_BN;
        k = i / SK_BN;
      }
      float v = 0.f;
      if (n0 + n < g.N && k0 + k < kend) v = __ldg(g.B + (long long)(k0 + k) * g.sbk + (long long)(n0 + n) * g.sbn);
      Bs[buf][k][n] = v;
    }
  };

  int buf = 0;
  if (kbeg < kend) stage(0, kbeg);
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += SK_BK) {
    if (k0 + SK_BK < kend) stage(buf ^ 1, k0 + SK_BK);   // next tile's loads overlap this tile's FMAs
#pragma unroll
    for (int kk = 0; kk < SK_BK; kk += 4) {
      float2 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float2*>(&Bs[buf][kk + j][2 * tn]);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][tm + 8 * r][kk]);
        acc[r][0] = fmaf(a.x, w[0].x, acc[r][0]);
        acc[r][1] = fmaf(a.x, w[0].y, acc[r][1]);
        acc[r][0] = fmaf(a.y, w[1].x, acc[r][0]);
        acc[r][1] = fmaf(a.y, w[1].y, acc[r][1]);
        acc[r][0] = fmaf(a.z, w[2].x, acc[r][0]);
        acc[r][1] = fmaf(a.z, w[2].y, acc[r][1]);
        acc[r][0] = fmaf(a.w, w[3].x, acc[r][0]);
        acc[r][1] = fmaf(a.w, w[3].y, acc[r][1]);
      }
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int m = tm + 8 * r;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + 2 * tn + j;
      if (n >= g.N) continue;
      float v = acc[r][j];
      if (g.partial) {
        g.partial[((size_t)blockIdx.z * g.M + m) * g.N + n] = v;
      } else {
        if (g.bias) v += g.bias[n];
        float* c = g.C + (long long)m * g.ldc + n;
        if (g.flags & 2) v += *c;
        if (g.flags & 1) v = fmaxf(v, 0.f);
        *c = v;
      }
    }
  }
}

// k-contiguous B (out[m, n] = sum_k A[m, k] B[n, k]: the forward layers, and the data gradient against the transposed
// weight), everything 16-byte aligned: A and B tiles go global -> shared with cp.async through a 4-stage ring, so the
// loads of three tiles are in flight while one is multiplied; B rows n = tn and tn + 32 per thread (row stride 36
// floats: the float4 reads of a quarter-warp cover all 32 banks).
constexpr int SKA_STAGES = 4;
// KFAST = false: n-contiguous B (out[m, n] = sum_k A[m, k] B[k, n]): the B tile is kept [32 k][64 n] and read as scalars.
template <int RPT, bool KFAST>
__global__ void __launch_bounds__(256) k_gemm_skinny_async(GemmArgs g) {
  extern __shared__ __align__(16) float sk_smem[];
  constexpr int AS = 8 * RPT * (SK_BK + 4);      // floats per A stage
  constexpr int BS = KFAST ? SK_BN * (SK_BK + 4) : SK_BK * (SK_BN + 4);   // floats per B stage
  float* As = sk_smem;                           // [stages][8*RPT][36]
  float* Bs = sk_smem + SKA_STAGES * AS;         // [stages][64][36]
  const int tid = threadIdx.x;
  const int tn = tid & 31, tm = tid >> 5;
  const int n0 = blockIdx.x * SK_BN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  const int ntiles = kend > kbeg ? (kend - kbeg + SK_BK - 1) / SK_BK : 0;
  float acc[RPT][2];
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = 0.f;

  auto issue = [&](int tile) {
    if (tile < ntiles) {
      const int k0 = kbeg + tile * SK_BK;
      float* a = As + (tile % SKA_STAGES) * AS;
      float* b = Bs + (tile % SKA_STAGES) * BS;
      for (int i = tid; i < 8 * RPT * (SK_BK / 4); i += 256) {
        const int c = i % (SK_BK / 4), m = i / (SK_BK / 4);
        const bool ok = m < g.M && k0 + 4 * c < kend;
        const float* src = ok ? g.A + (long long)m * g.sam + k0 + 4 * c : g.A;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(a + m * (SK_BK + 4) + 4 * c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
      if (KFAST) {
        for (int i = tid; i < SK_BN * (SK_BK / 4); i += 256) {
          const int c = i % (SK_BK / 4), n = i / (SK_BK / 4);
          const bool ok = n0 + n < g.N && k0 + 4 * c < kend;
          const float* src = ok ? g.B + (long long)(n0 + n) * g.sbn + k0 + 4 * c : g.B;
          const unsigned dst = (unsigned)__cvta_generic_to_shared(b + n * (SK_BK + 4) + 4 * c);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
      } else {
        for (int i = tid; i < SK_BK * (SK_BN / 4); i += 256) {
          const int c = i % (SK_BN / 4), k = i / (SK_BN / 4);
          const bool ok = k0 + k < kend && n0 + 4 * c < g.N;
          const float* src = ok ? g.B + (long long)(k0 + k) * g.sbk + n0 + 4 * c : g.B;
          const unsigned dst = (unsigned)__cvta_generic_to_shared(b + k * (SK_BN + 4) + 4 * c);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

#pragma unroll
  for (int t = 0; t < SKA_STAGES - 1; ++t) issue(t);
  for (int t = 0; t < ntiles; ++t) {
    asm volatile("cp.async.wait_group %0;" ::"n"(SKA_STAGES - 2) : "memory");
    __syncthreads();                       // tile t has landed for everyone; stage (t-1)%S is free again
    issue(t + SKA_STAGES - 1);
    const float* a = As + (t % SKA_STAGES) * AS;
    const float* b = Bs + (t % SKA_STAGES) * BS;
#pragma unroll
    for (int kk = 0; kk < SK_BK; kk += 4) {
      float4 w0, w1;
      if (KFAST) {
        w0 = *reinterpret_cast<const float4*>(b + tn * (SK_BK + 4) + kk);
        w1 = *reinterpret_cast<const float4*>(b + (tn + 32) * (SK_BK + 4) + kk);
      } else {
        const float* bp = b + kk * (SK_BN + 4) + tn;
        w0 = make_float4(bp[0], bp[SK_BN + 4], bp[2 * (SK_BN + 4)], bp[3 * (SK_BN + 4)]);
        w1 = make_float4(bp[32], bp[SK_BN + 4 + 32], bp[2 * (SK_BN + 4) + 32], bp[3 * (SK_BN + 4) + 32]);
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(a + (tm + 8 * r) * (SK_BK + 4) + kk);
        acc[r][0] = fmaf(x.x, w0.x, acc[r][0]);
        acc[r][1] = fmaf(x.x, w1.x, acc[r][1]);
        acc[r][0] = fmaf(x.y, w0.y, acc[r][0]);
        acc[r][1] = fmaf(x.y, w1.y, acc[r][1]);
        acc[r][0] = fmaf(x.z, w0.z, acc[r][0]);
        acc[r][1] = fmaf(x.z, w1.z, acc[r][1]);
        acc[r][0] = fmaf(x.w, w0.w, acc[r][0]);
        acc[r][1] = fmaf(x.w, w1.w, acc[r][1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int m = tm + 8 * r;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tn + 32 * j;
      if (n >= g.N) continue;
      float v = acc[r][j];
      if (g.partial) {
        g.partial[((size_t)blockIdx.z * g.M + m) * g.N + n] = v;
      } else {
        if (g.bias) v += g.bias[n];
        float* c = g.C + (long long)m * g.ldc + n;
        if (g.flags & 2) v += *c;
        if (g.flags & 1) v = fmaxf(v, 0.f);
        *c = v;
      }
    }
  }
}
template <int RPT>
int launch_skinny_async(const GemmArgs& g, dim3 grid, cudaStream_t stream) {
  if (g.sbk == 1) {
    const size_t smem = (size_t)SKA_STAGES * (8 * RPT + SK_BN) * (SK_BK + 4) * sizeof(float);
    if (cudaFuncSetAttribute(k_gemm_skinny_async<RPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return -1;
    k_gemm_skinny_async<RPT, true><<<grid, 256, smem, stream>>>(g);
  } else {
    const size_t smem = (size_t)SKA_STAGES * (8 * RPT * (SK_BK + 4) + SK_BK * (SK_BN + 4)) * sizeof(float);
    if (cudaFuncSetAttribute(k_gemm_skinny_async<RPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return -1;
    k_gemm_skinny_async<RPT, false><<<grid, 256, smem, stream>>>(g);
  }
  return 0;
}

// split K so that the skinny kernel fills the machine (N/64 column blocks x splits CTAs)
int skinny_splits(int N, int K) {
  const int cols = ceil_div(N, SK_BN);
  int s = ceil_div(2 * mpnn_num_sms(), cols);
  const int by_k = K / 512 > 0 ? K / 512 : 1;
  if (s > by_k) s = by_k;
  return s < 1 ? 1 : s;
}
bool skinny_ok(int M, int N, int K, long long sak) { return M <= 64 && sak == 1 && N >= 256 && K >= 128; }

// ---------------------------------------------------------------------------------------------------
// Small GEMM: problems whose 64 x 64 tiling would leave most SMs idle (Set2Vec's per-step products: 128 x 256 x 128
// and smaller, set2vec.py:69-72,128, six of them in each of the 100 attention steps).  32 x 32 tiles (one CTA each,
// 2 x 2 outputs per thread), 32-deep k-slabs whose global loads are issued one slab ahead of the FMAs.
// ---------------------------------------------------------------------------------------------------
constexpr int SM_T = 32, SM_K = 32;

__global__ void __launch_bounds__(256) k_gemm_small(GemmArgs g) {
  __shared__ float As[2][SM_K][SM_T + 1];   // [k][m]
  __shared__ float Bs[2][SM_K][SM_T + 1];   // [k][n]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // outputs (m = ty*2 + {0,1}, n = tx*2 + {0,1})
  const int m0 = blockIdx.y * SM_T, n0 = blockIdx.x * SM_T;
  const bool a_kfast = (g.sak == 1), b_nfast = (g.sbn == 1);
  // this thread's 4 + 4 elements of a slab: index i = tid + 256*it over the 32 x 32 tile
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = tid + it * 256;
      const int ka = a_kfast ? (i & 31) : (i >> 5), ma = a_kfast ? (i >> 5) : (i & 31);
      const int kb = b_nfast ? (i >> 5) : (i & 31), nb = b_nfast ? (i & 31) : (i >> 5);
      ra[it] = (m0 + ma < g.M && k0 + ka < g.K) ? __ldg(g.A + (long long)(m0 + ma) * g.sam + (long long)(k0 + ka) * g.sak) : 0.f;
      rb[it] = (n0 + nb < g.N && k0 + kb < g.K) ? __ldg(g.B + (long long)(k0 + kb) * g.sbk + (long long)(n0 + nb) * g.sbn) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = tid + it * 256;
      const int ka = a_kfast ? (i & 31) : (i >> 5), ma = a_kfast ? (i >> 5) : (i & 31);
      const int kb = b_nfast ? (i >> 5) : (i & 31), nb = b_nfast ? (i & 31) : (i >> 5);
      As[buf][ka][ma] = ra[it];
      Bs[buf][kb][nb] = rb[it];
    }
  };
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < g.K; k0 += SM_K) {
    const bool more = k0 + SM_K < g.K;
    if (more) fetch(k0 + SM_K);            // in flight while this slab is multiplied
#pragma unroll
    for (int kk = 0; kk < SM_K; ++kk) {
      const float a0 = As[buf][kk][ty * 2], a1 = As[buf][kk][ty * 2 + 1];
      const float b0 = Bs[buf][kk][tx * 2], b1 = Bs[buf][kk][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty * 2 + i, n = n0 + tx * 2 + j;
      if (m >= g.M || n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      float* c = g.C + (long long)m * g.ldc + n;
      if (g.flags & 2) v += *c;
      if (g.flags & 1) v = fmaxf(v, 0.f);
      *c = v;
    }
}
bool small_ok(int M, int N, int K) {
  if (K == 0 || K > 2048) return false;
  const long long big_tiles = (long long)ceil_div(M, BM) * ceil_div(N, BN);
  const long long small_tiles = (long long)ceil_div(M, SM_T) * ceil_div(N, SM_T);
  return big_tiles * 4 <= mpnn_num_sms() && small_tiles <= 4LL * mpnn_num_sms();
}

int choose_splits(int M, int N, int K) {
  long long tiles = (long long)ceil_div(M, BM) * ceil_div(N, BN);
  int sms = mpnn_num_sms();
  if (K < 1024 || tiles >= 2 * sms) return 1;
  long long want = (2ll * sms + tiles - 1) / tiles;
  int by_k = ceil_div(K, 256);
  int s = (int)(want < by_k ? want : by_k);
  return s < 1 ? 1 : s;
}

}  // namespace

extern "C" {

size_t mpnn_gemm_workspace_bytes(int M, int N, int K) {
  int s = choose_splits(M, N, K);
  if (M <= 64 && N >= 256 && K >= 128) {
    const int sk = skinny_splits(N, K);
    if (sk > s) s = sk;
  }
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
              long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
              size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(M >= 0 && N >= 0 && K >= 0, MPNN_ERR_ARG, "gemm: negative dims");
  if (M == 0 || N == 0) return MPNN_OK;
  if (skinny_ok(M, N, K, sak)) {
    int splits = skinny_splits(N, K);
    while (splits > 1 && (size_t)splits * M * N * sizeof(float) > workspace_bytes) --splits;
    GemmArgs g;
    g.A = A;
    g.B = B;
    g.C = C;
    g.bias = bias;
    g.M = M;
    g.N = N;
    g.K = K;
    g.sam = sam;
    g.sak = sak;
    g.sbk = sbk;
    g.sbn = sbn;
    g.ldc = ldc;
    g.flags = flags;
    g.k_per_split = ceil_div(ceil_div(K, splits), SK_BK) * SK_BK;
    g.partial = splits > 1 ? (float*)workspace : nullptr;
    dim3 grid(ceil_div(N, SK_BN), 1, splits);
    const int rpt = ceil_div(M, 8);
    const bool aligned = (K & 3) == 0 && (sam & 3) == 0 && ((((uintptr_t)A) | ((uintptr_t)B)) & 15) == 0;
    const bool async_ok = aligned && ((sbk == 1 && (sbn & 3) == 0) || (sbn == 1 && (sbk & 3) == 0 && (N & 3) == 0));
    if (async_ok) {
      int rc = 0;
      switch (rpt) {
        case 1: rc = launch_skinny_async<1>(g, grid, stream); break;
        case 2: rc = launch_skinny_async<2>(g, grid, stream); break;
        case 3: rc = launch_skinny_async<3>(g, grid, stream); break;
        case 4: rc = launch_skinny_async<4>(g, grid, stream); break;
        case 5: rc = launch_skinny_async<5>(g, grid, stream); break;
        case 6: rc = launch_skinny_async<6>(g, grid, stream); break;
        case 7: rc = launch_skinny_async<7>(g, grid, stream); break;
        default: rc = launch_skinny_async<8>(g, grid, stream); break;
      }
      MPNN_REQUIRE(rc == 0, MPNN_ERR_CUDA, "gemm: shared-memory attribute (skinny)");
    } else
    switch (rpt) {
      case 1: k_gemm_skinny<1><<<grid, 256, 0, stream>>>(g); break;
      case 2: k_gemm_skinny<2><<<grid, 256, 0, stream>>>(g); break;
      case 3: k_gemm_skinny<3><<<grid, 256, 0, stream>>>(g); break;
      case 4: k_gemm_skinny<4><<<grid, 256, 0, stream>>>(g); break;
      case 5: k_gemm_skinny<5><<<grid, 256, 0, stream>>>(g); break;
      case 6: k_gemm_skinny<6><<<grid, 256, 0, stream>>>(g); break;
      case 7: k_gemm_skinny<7><<<grid, 256, 0, stream>>>(g); break;
      default: k_gemm_skinny<8><<<grid, 256, 0, stream>>>(g); break;
    }
    MPNN_CHECK_LAUNCH("k_gemm_skinny");
    if (splits > 1) {
      k_gemm_reduce<<<ceil_div((long long)M * N, 256), 256, 0, stream>>>(g.partial, splits, M, N, C, ldc, bias, flags);
      MPNN_CHECK_LAUNCH("k_gemm_reduce");
    }
    return MPNN_OK;
  }
  if (small_ok(M, N, K)) {
    GemmArgs g;
    g.A = A;
    g.B = B;
    g.C = C;
    g.bias = bias;
    g.M = M;
    g.N = N;
    g.K = K;
    g.sam = sam;
    g.sak = sak;
    g.sbk = sbk;
    g.sbn = sbn;
    g.ldc = ldc;
    g.flags = flags;
    g.k_per_split = K;
    g.partial = nullptr;
    dim3 grid(ceil_div(N, SM_T), ceil_div(M, SM_T), 1);
    k_gemm_small<<<grid, 256, 0, stream>>>(g);
    MPNN_CHECK_LAUNCH("k_gemm_small");
    return MPNN_OK;
  }
  int splits = choose_splits(M, N, K);
  while (splits > 1 && (size_t)splits * M * N * sizeof(float) > workspace_bytes) --splits;
  GemmArgs g;
  g.A = A;
  g.B = B;
  g.C = C;
  g.bias = bias;
  g.M = M;
  g.N = N;
  g.K = K;
  g.sam = sam;
  g.sak = sak;
  g.sbk = sbk;
  g.sbn = sbn;
  g.ldc = ldc;
  g.flags = flags;
  g.k_per_split = K == 0 ? 1 : ceil_div(ceil_div(K, splits), BK) * BK;
  g.partial = splits > 1 ? (float*)workspace : nullptr;
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
  MPNN_REQUIRE(grid.y <= 65535 || splits == 1, MPNN_ERR_UNSUPPORTED, "gemm: M too large for split-K");
  if (grid.y > 65535) {
    // fold M over several launches
    int rows_per = 65535 * BM;
    for (int m = 0; m < M; m += rows_per) {
      GemmArgs h = g;
      h.A = A + (long long)m * sam;
      h.C = C + (long long)m * ldc;
      h.M = (M - m < rows_per) ? M - m : rows_per;
      dim3 gr(ceil_div(N, BN), ceil_div(h.M, BM), 1);
      k_gemm<<<gr, 256, 0, stream>>>(h);
    }
  } else {
    k_gemm<<<grid, 256, 0, stream>>>(g);
  }
  MPNN_CHECK_LAUNCH("k_gemm");
  if (splits > 1) {
    k_gemm_reduce<<<ceil_div((long long)M * N, 256), 256, 0, stream>>>(g.partial, splits, M, N, C, ldc, bias, flags);
    MPNN_CHECK_LAUNCH("k_gemm_reduce");
  }
  return MPNN_OK;
}

}  // extern "C"
