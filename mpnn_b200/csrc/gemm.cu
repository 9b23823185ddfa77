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
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
              long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
              size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(M >= 0 && N >= 0 && K >= 0, MPNN_ERR_ARG, "gemm: negative dims");
  if (M == 0 || N == 0) return MPNN_OK;
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
