// Edge-network trunk (reference mpnn_functions/message/edge_network.py:14-21): growth layers
// Linear(in,in^2)+ReLU, then 50 applications of ONE bias-free Linear(P,P)+ReLU (the reference repeats the
// same module object 50 times), evaluated on the COMPACTED bond rows only (edges + the single all-zero row),
// never on the dense B*N*N grid the reference feeds it (edge_network.py:36-37).
//
// The tied P x P weight stays resident in shared memory for all 50 layers and the activation tile
// ping-pongs between two shared buffers; each layer's output is also streamed to HBM once because the
// backward pass needs it (dW_t = sum_k delta_k^T a_{k-1}).  fp32 FFMA on purpose: bf16 through 50 tied layers
// is 7.5 % off (hard parts).
//
// saved-buffer layout (floats), R rows, widths padded to a multiple of 4:
//   [growth_0 out | growth_1 out | ... | growth_{G-1} out (= tied input, width PP)]   (G == 0: padded input copy)
//   [tied_1 | tied_2 | ... | tied_L]   each [R, PP];   x = tied_L is what the message kernel reads.
#include "common.cuh"

extern "C" int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
                         long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_gemm_workspace_bytes(int M, int N, int K);
extern "C" int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy,
                           float* out, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_colsum_workspace_bytes(long long rows, int width);

extern "C" size_t mpnn_tc_dense_grad_workspace_bytes(int G, int DP);
extern "C" int mpnn_tc_dense_gemm_tn(const float* X, long long rows, int ldx, int M, const float* D, int ldd, int dcol,
                                     int G, int N, int DP, float* out, long long o_sg, long long o_sl, void* workspace,
                                     size_t workspace_bytes, cudaStream_t stream);

extern "C" int mpnn_tensor_cores_enabled(void);

namespace {

constexpr int TR = 64;        // rows per tile
constexpr int TRP = TR + 4;   // transposed-tile row stride (== 4 mod 32: conflict-free float4 stores)
constexpr int MAX_GROWTH = 4;

struct Layout {
  int G, P, PP, L;
  int gin[MAX_GROWTH], gout[MAX_GROWTH], gld[MAX_GROWTH];
  size_t goff[MAX_GROWTH];
  size_t tied_in_off;  // offset of the tied input [R, PP]
  size_t tied_off;     // offset of tied_1
  size_t total;
};

bool make_layout(int R, int ef, int n_growth, int P, int n_tied, Layout* lo) {
  if (n_growth > MAX_GROWTH || n_growth < 0) return false;
  lo->G = n_growth;
  lo->P = P;
  lo->PP = pad4(P);
  lo->L = n_tied;
  size_t off = 0;
  int w = ef;
  for (int g = 0; g < n_growth; ++g) {
    lo->gin[g] = w;
    lo->gout[g] = w * w;
    lo->gld[g] = pad4(w * w);
    lo->goff[g] = off;
    off += (size_t)R * lo->gld[g];
    w = w * w;
  }
  if (w != P) return false;
  if (n_growth == 0) {
    lo->tied_in_off = off;
    off += (size_t)R * lo->PP;
  } else {
    lo->tied_in_off = lo->goff[n_growth - 1];
  }
  lo->tied_off = off;
  off += (size_t)n_tied * R * lo->PP;
  lo->total = off;
  return true;
}

__global__ void k_pad_copy(const float* __restrict__ src, int R, int w, float* __restrict__ dst, int ld) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * ld) return;
  int r = (int)(t / ld), c = (int)(t - (long long)r * ld);
  dst[t] = c < w ? src[(long long)r * w + c] : 0.f;
}

// out[c][r] = in[r][c] for a square [P, P] matrix (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) k_transpose_sq(const float* __restrict__ in, int P, float* __restrict__ out) {
  __shared__ float t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (by + j < P && bx + tx < P) t[j][tx] = in[(size_t)(by + j) * P + bx + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (bx + j < P && by + tx < P) out[(size_t)(bx + j) * P + by + tx] = t[tx][j];
}

__global__ void k_relu_mask(float* __restrict__ d, const float* __restrict__ a, long long n) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && !(a[t] > 0.f)) d[t] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// forward: L tied layers, W resident in smem
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tied_fwd(const float* __restrict__ in, const float* __restrict__ W, int R,
                                                  int P, int PP, int L, float* __restrict__ out /*[L][R][PP]*/) {
  extern __shared__ __align__(16) float smem[];
  float* Wt = smem;                    // [PP][PP]  Wt[i][o] = W[o][i]
  float* A0 = Wt + PP * PP;            // [PP][TRP]
  float* A1 = A0 + PP * TRP;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < PP * PP; idx += 256) {
    int o = idx / PP, i = idx - o * PP;   // coalesced over i
    Wt[i * PP + o] = (o < P && i < P) ? W[o * P + i] : 0.f;
  }
  const int CG = PP >> 2;
  const int tasks = CG * (TR / 4);
  const int ntiles = (R + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * TR;
    __syncthreads();  // Wt ready / previous tile done with A buffers
    for (int idx = tid; idx < TR * PP; idx += 256) {
      int r = idx / PP, i = idx - r * PP;
      int row = row0 + r;
      A0[i * TRP + r] = row < R ? in[(size_t)row * PP + i] : 0.f;
    }
    __syncthreads();
    float* Ain = A0;
    float* Aout = A1;
    for (int l = 0; l < L; ++l) {
      float* og = out + ((size_t)l * R) * PP;
      for (int task = tid; task < tasks; task += 256) {
        const int cg = task % CG, rg = task / CG;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        const float* ap = Ain + rg * 4;
        const float* wp = Wt + cg;
#pragma unroll 4
        for (int i = 0; i < P; ++i) {
          float4 a4 = *reinterpret_cast<const float4*>(ap + i * TRP);
          float w0 = wp[i * PP], w1 = wp[i * PP + CG], w2 = wp[i * PP + 2 * CG], w3 = wp[i * PP + 3 * CG];
          float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            acc[a][0] = fmaf(av[a], w0, acc[a][0]);
            acc[a][1] = fmaf(av[a], w1, acc[a][1]);
            acc[a][2] = fmaf(av[a], w2, acc[a][2]);
            acc[a][3] = fmaf(av[a], w3, acc[a][3]);
          }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float4 v = make_float4(fmaxf(acc[0][b], 0.f), fmaxf(acc[1][b], 0.f), fmaxf(acc[2][b], 0.f),
                                 fmaxf(acc[3][b], 0.f));
          *reinterpret_cast<float4*>(Aout + (cg + b * CG) * TRP + rg * 4) = v;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          int row = row0 + rg * 4 + a;
          if (row < R) {
#pragma unroll
            for (int b = 0; b < 4; ++b) og[(size_t)row * PP + cg + b * CG] = fmaxf(acc[a][b], 0.f);
          }
        }
      }
      __syncthreads();
      float* t = Ain;
      Ain = Aout;
      Aout = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward through the L tied layers.  NT = ceil((PP/4)^2 / 256) dW micro-tiles per thread (registers).
// ---------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) k_tied_bwd(const float* __restrict__ tied_in, const float* __restrict__ tied_out,
                                                  const float* __restrict__ W, const float* __restrict__ dx, int lddx,
                                                  int R, int P, int PP, int L, float* __restrict__ d_in /*[R][PP]*/,
                                                  float* __restrict__ dW_partial /*[grid][PP*PP]*/) {
  extern __shared__ __align__(16) float smem[];
  const int PPs = PP + 4;
  float* Ws = smem;                 // [PP][PP] natural W[o][i]
  float* Dl = Ws + PP * PP;         // delta  [TR][PPs]
  float* Ap = Dl + TR * PPs;        // a_{k-1}[TR][PPs]
  float* dA = Ap + TR * PPs;        // dA     [TR][PPs]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < PP * PP; idx += 256) {
    int o = idx / PP, i = idx - o * PP;
    Ws[idx] = (o < P && i < P) ? W[o * P + i] : 0.f;
  }
  const int G4 = PP >> 2;
  const int wtasks = G4 * G4;
  float accW[NT][4][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) accW[t][a][b] = 0.f;

  const int ntiles = (R + TR - 1) / TR;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * TR;
    __syncthreads();
    // dA <- dx tile ; Ap <- a_L tile (mask source of the first iteration)
    for (int idx = tid; idx < TR * PP; idx += 256) {
      int r = idx / PP, c = idx - r * PP;
      int row = row0 + r;
      bool ok = row < R;
      dA[r * PPs + c] = (ok && c < P) ? dx[(size_t)row * lddx + c] : 0.f;
      Ap[r * PPs + c] = ok ? tied_out[((size_t)(L - 1) * R + row) * PP + c] : 0.f;
    }
    __syncthreads();
    for (int l = L; l >= 1; --l) {
      // delta = dA * (a_l > 0) with a_l currently in Ap; then Ap <- a_{l-1}
      const float* prev = (l == 1) ? tied_in : tied_out + ((size_t)(l - 2) * R) * PP;
      for (int idx = tid; idx < TR * PP; idx += 256) {
        int r = idx / PP, c = idx - r * PP;
        int row = row0 + r;
        float a = Ap[r * PPs + c];
        Dl[r * PPs + c] = a > 0.f ? dA[r * PPs + c] : 0.f;
        Ap[r * PPs + c] = row < R ? prev[(size_t)row * PP + c] : 0.f;
      }
      __syncthreads();
      // dW[o][i] += sum_r delta[r][o] * a_{l-1}[r][i]
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        int task = tid + t * 256;
        if (task < wtasks) {
          const int ig = task % G4, og = task / G4;
          const float* dp = Dl + og * 4;
          const float* ap = Ap + ig * 4;
#pragma unroll 4
          for (int r = 0; r < TR; ++r) {
            float4 d4 = *reinterpret_cast<const float4*>(dp + r * PPs);
            float4 a4 = *reinterpret_cast<const float4*>(ap + r * PPs);
            float dv[4] = {d4.x, d4.y, d4.z, d4.w};
            float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) accW[t][a][b] = fmaf(dv[a], av[b], accW[t][a][b]);
          }
        }
      }
      // dA_prev[r][i] = sum_o delta[r][o] * W[o][i]   (written to dA; dA is not read again this layer)
      const int atasks = G4 * (TR / 4);
      for (int task = tid; task < atasks; task += 256) {
        const int ig = task % G4, rg = task / G4;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        const float* dp = Dl + (rg * 4) * PPs;
        const float* wp = Ws + ig * 4;
#pragma unroll 4
        for (int o = 0; o < P; ++o) {
          float4 w4 = *reinterpret_cast<const float4*>(wp + o * PP);
          float d0 = dp[o], d1 = dp[PPs + o], d2 = dp[2 * PPs + o], d3 = dp[3 * PPs + o];
          float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acc[0][b] = fmaf(d0, wv[b], acc[0][b]);
            acc[1][b] = fmaf(d1, wv[b], acc[1][b]);
            acc[2][b] = fmaf(d2, wv[b], acc[2][b]);
            acc[3][b] = fmaf(d3, wv[b], acc[3][b]);
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
          *reinterpret_cast<float4*>(dA + (rg * 4 + a) * PPs + ig * 4) =
              make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      }
      __syncthreads();
    }
    for (int idx = tid; idx < TR * PP; idx += 256) {
      int r = idx / PP, c = idx - r * PP;
      int row = row0 + r;
      if (row < R) d_in[(size_t)row * PP + c] = dA[r * PPs + c];
    }
  }
  float* part = dW_partial + (size_t)blockIdx.x * PP * PP;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    int task = tid + t * 256;
    if (task < wtasks) {
      const int ig = task % G4, og = task / G4;
#pragma unroll
      for (int a = 0; a < 4; ++a)
        *reinterpret_cast<float4*>(part + (og * 4 + a) * PP + ig * 4) =
            make_float4(accW[t][a][0], accW[t][a][1], accW[t][a][2], accW[t][a][3]);
    }
  }
}

__global__ void k_tied_dw_reduce(const float* __restrict__ partial, int nparts, int P, int PP,
                                 float* __restrict__ dW) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * P) return;
  int o = idx / P, i = idx - o * P;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * PP * PP + o * PP + i];
  dW[idx] = s;
}

int tied_grid(int R, size_t smem_bytes) {
  int tiles = ceil_div(R, TR);
  int per_sm = (int)((220 * 1024) / (smem_bytes + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int g = mpnn_num_sms() * per_sm;
  return tiles < g ? tiles : g;
}

size_t fwd_smem(int PP) { return ((size_t)PP * PP + 2 * (size_t)PP * TRP) * sizeof(float); }
size_t bwd_smem(int PP) { return ((size_t)PP * PP + 3 * (size_t)TR * (PP + 4)) * sizeof(float); }
constexpr int MAX_PP_SMEM = 128;
constexpr int MAX_TIED = 64;   // tied layers the wide-trunk backward can stack (the reference uses 50)
constexpr int WIDE_STACK_ROWS = 256;   // the delta stack is kept only for few rows (typed path: distinct bond rows)
// widths whose tied-weight gradient runs as ONE stacked X^T D GEMM on the tensor cores (TF32 operands)
inline int wide_block(int P) { return P % 256 == 0 ? 256 : (P % 128 == 0 ? 128 : 64); }
inline bool wide_dw_on_tc(int P) { return mpnn_tensor_cores_enabled() && P % 64 == 0 && P >= 256; }

}  // namespace

extern "C" {

// Number of floats the caller must allocate for `saved`; *x_offset receives the offset of x = tied_L, *ldx = PP.
long long mpnn_edge_trunk_saved_floats(int R, int ef, int n_growth, int P, int n_tied, long long* x_offset, int* ldx) {
  Layout lo;
  if (!make_layout(R, ef, n_growth, P, n_tied, &lo)) return -1;
  if (x_offset) *x_offset = (long long)(lo.tied_off + (size_t)(n_tied - 1) * R * lo.PP);
  if (ldx) *ldx = lo.PP;
  return (long long)lo.total;
}

size_t mpnn_edge_trunk_workspace_bytes(int R, int ef, int n_growth, int P) {
  int PP = pad4(P);
  size_t g = mpnn_gemm_workspace_bytes(P, P, R);  // largest split-K user (weight grads)
  size_t c = mpnn_colsum_workspace_bytes(R, PP);
  size_t partial = PP <= MAX_PP_SMEM ? (size_t)2 * mpnn_num_sms() * PP * PP * sizeof(float) : 0;
  size_t dA = 2 * (size_t)R * PP * sizeof(float);
  size_t wide = 0;
  if (PP > MAX_PP_SMEM) {
    // wide trunks: transposed weight + the stacked deltas of all tied layers (+ the tensor-core X^T D workspace)
    size_t skinny = mpnn_gemm_workspace_bytes(R, P, P);
    if (skinny > g) g = skinny;
    wide = align_up((size_t)P * P * sizeof(float), 256) + 512;
    if (R <= WIDE_STACK_ROWS) {
      wide += align_up((size_t)MAX_TIED * R * PP * sizeof(float), 256);
      if (wide_dw_on_tc(P)) wide += align_up(mpnn_tc_dense_grad_workspace_bytes(P / wide_block(P), wide_block(P)), 256);
    }
  }
  return align_up(g > c ? g : c, 256) + align_up(partial, 256) + align_up(dA, 256) + wide + 1024;
}

int mpnn_edge_trunk_fwd(const float* rows_in, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* const* growth_b, const float* w_tied, int P, int n_tied, float* saved,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  Layout lo;
  MPNN_REQUIRE(R > 0 && n_tied >= 1, MPNN_ERR_ARG, "edge_trunk_fwd: bad dims R=%d n_tied=%d", R, n_tied);
  MPNN_REQUIRE(make_layout(R, ef, n_growth, P, n_tied, &lo), MPNN_ERR_ARG,
               "edge_trunk_fwd: ef=%d with %d growth layers does not reach P=%d", ef, n_growth, P);
  const int PP = lo.PP;
  float* tied_in = saved + lo.tied_in_off;
  if (n_growth == 0) {
    k_pad_copy<<<ceil_div((long long)R * PP, 256), 256, 0, stream>>>(rows_in, R, ef, tied_in, PP);
    MPNN_CHECK_LAUNCH("k_pad_copy");
  } else {
    const float* a = rows_in;
    long long lda = ef;
    for (int g = 0; g < n_growth; ++g) {
      float* o = saved + lo.goff[g];
      if (lo.gld[g] != lo.gout[g]) MPNN_CUDA(cudaMemsetAsync(o, 0, (size_t)R * lo.gld[g] * sizeof(float), stream));
      // a_g = relu(a_{g-1} W_g^T + b_g), W_g is [out, in] row-major (nn.Linear)
      int rc = mpnn_gemm(a, growth_w[g], o, R, lo.gout[g], lo.gin[g], lda, 1, 1, lo.gin[g], lo.gld[g], growth_b[g], 1,
                         workspace, workspace_bytes, stream);
      if (rc) return rc;
      a = o;
      lda = lo.gld[g];
    }
  }
  float* tied_out = saved + lo.tied_off;
  if (PP <= MAX_PP_SMEM) {
    size_t smem = fwd_smem(PP);
    MPNN_CUDA(cudaFuncSetAttribute(k_tied_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tied_fwd<<<tied_grid(R, smem), 256, smem, stream>>>(tied_in, w_tied, R, P, PP, n_tied, tied_out);
    MPNN_CHECK_LAUNCH("k_tied_fwd");
  } else {
    // wide trunks (P = 256, 625, 4096 ...): layer-by-layer GEMM, weight streamed from L2
    const float* a = tied_in;
    for (int l = 0; l < n_tied; ++l) {
      float* o = tied_out + (size_t)l * R * PP;
      if (PP != P) MPNN_CUDA(cudaMemsetAsync(o, 0, (size_t)R * PP * sizeof(float), stream));
      int rc = mpnn_gemm(a, w_tied, o, R, P, P, PP, 1, 1, P, PP, nullptr, 1, workspace, workspace_bytes, stream);
      if (rc) return rc;
      a = o;
    }
  }
  return MPNN_OK;
}

// dx: [R, lddx] gradient w.r.t. x = tied_L (columns >= P are ignored).  Gradients are written (not accumulated).
int mpnn_edge_trunk_bwd(const float* rows_in, int R, int ef, int n_growth, const float* const* growth_w,
                        const float* w_tied, int P, int n_tied, const float* saved, const float* dx, int lddx,
                        float* const* d_growth_w, float* const* d_growth_b, float* d_w_tied, float* d_rows_in,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  Layout lo;
  MPNN_REQUIRE(R > 0 && n_tied >= 1 && lddx >= P, MPNN_ERR_ARG, "edge_trunk_bwd: bad dims");
  MPNN_REQUIRE(make_layout(R, ef, n_growth, P, n_tied, &lo), MPNN_ERR_ARG, "edge_trunk_bwd: inconsistent layer plan");
  MPNN_REQUIRE(workspace_bytes >= mpnn_edge_trunk_workspace_bytes(R, ef, n_growth, P), MPNN_ERR_WORKSPACE,
               "edge_trunk_bwd: workspace too small");
  const int PP = lo.PP;
  char* wp = (char*)workspace;
  size_t gbytes = mpnn_gemm_workspace_bytes(P, P, R);
  if (PP > MAX_PP_SMEM) {
    const size_t skinny = mpnn_gemm_workspace_bytes(R, P, P);
    if (skinny > gbytes) gbytes = skinny;
  }
  size_t cbytes = mpnn_colsum_workspace_bytes(R, PP);
  size_t sub_bytes = align_up(gbytes > cbytes ? gbytes : cbytes, 256);
  void* sub = wp;
  wp += sub_bytes;
  float* partial = (float*)wp;
  wp += align_up(PP <= MAX_PP_SMEM ? (size_t)2 * mpnn_num_sms() * PP * PP * sizeof(float) : 0, 256);
  float* dA = (float*)wp;  // [R, PP] grad w.r.t. tied input (then reused down the growth layers)
  float* dB = dA + (size_t)R * PP;

  const float* tied_in = saved + lo.tied_in_off;
  const float* tied_out = saved + lo.tied_off;
  if (PP <= MAX_PP_SMEM) {
    size_t smem = bwd_smem(PP);
    int grid = tied_grid(R, smem);
    int wtasks = (PP / 4) * (PP / 4);
    if (wtasks <= 256) {
      MPNN_CUDA(cudaFuncSetAttribute(k_tied_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tied_bwd<1><<<grid, 256, smem, stream>>>(tied_in, tied_out, w_tied, dx, lddx, R, P, PP, n_tied, dA, partial);
    } else if (wtasks <= 512) {
      MPNN_CUDA(cudaFuncSetAttribute(k_tied_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tied_bwd<2><<<grid, 256, smem, stream>>>(tied_in, tied_out, w_tied, dx, lddx, R, P, PP, n_tied, dA, partial);
    } else {
      MPNN_CUDA(cudaFuncSetAttribute(k_tied_bwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tied_bwd<4><<<grid, 256, smem, stream>>>(tied_in, tied_out, w_tied, dx, lddx, R, P, PP, n_tied, dA, partial);
    }
    MPNN_CHECK_LAUNCH("k_tied_bwd");
    k_tied_dw_reduce<<<ceil_div(P * P, 256), 256, 0, stream>>>(partial, grid, P, PP, d_w_tied);
    MPNN_CHECK_LAUNCH("k_tied_dw_reduce");
  } else {
    // Wide trunk (P = 256, 625, 4096: hidden >= 128 in the reference's models), R = distinct bond rows (tens):
    //   delta_l = d_l * relu'(a_l) is kept for every layer in one stack [L*R, PP];
    //   d_{l-1} = delta_l W runs on the skinny GEMM against a TRANSPOSED copy of W (k-contiguous operand: 4x faster
    //   than the strided form), and the tied-weight gradient sum_l delta_l^T a_{l-1} is ONE GEMM over the stacked
    //   rows -- on the tensor cores when P is a multiple of 64, else layer by layer in fp32.
    const bool stacked = R <= WIDE_STACK_ROWS && n_tied <= MAX_TIED;
    char* xp = (char*)(dB + (size_t)R * PP);
    xp = (char*)align_up((size_t)xp, 256);
    float* Wt = (float*)xp;
    xp += align_up((size_t)P * P * sizeof(float), 256);
    float* stack = (float*)xp;
    xp += align_up((size_t)MAX_TIED * R * PP * sizeof(float), 256);
    void* tcws = xp;
    const size_t tcws_bytes = wide_dw_on_tc(P) ? mpnn_tc_dense_grad_workspace_bytes(P / wide_block(P), wide_block(P)) : 0;
    {
      dim3 tg(ceil_div(P, 32), ceil_div(P, 32));
      k_transpose_sq<<<tg, 256, 0, stream>>>(w_tied, P, Wt);
      MPNN_CHECK_LAUNCH("k_transpose_sq");
    }
    const size_t slab = (size_t)R * PP;
    // layer l's delta lives in slot(l): its own slab of the stack, or (many rows: per-edge path) one of two buffers
    auto slot = [&](int l) { return stacked ? stack + (size_t)(l - 1) * slab : (((n_tied - l) & 1) ? dB : dA); };
    float* top = slot(n_tied);
    if (PP != P) {
      if (stacked) MPNN_CUDA(cudaMemsetAsync(stack, 0, (size_t)n_tied * slab * sizeof(float), stream));
      MPNN_CUDA(cudaMemsetAsync(dA, 0, 2 * slab * sizeof(float), stream));
    }
    MPNN_CUDA(cudaMemcpy2DAsync(top, (size_t)PP * sizeof(float), dx, (size_t)lddx * sizeof(float), (size_t)P * sizeof(float),
                                R, cudaMemcpyDeviceToDevice, stream));
    const bool contiguous_acts = lo.tied_in_off + slab == lo.tied_off;
    const bool tc_dw = stacked && wide_dw_on_tc(P) && contiguous_acts;
    if (!tc_dw) MPNN_CUDA(cudaMemsetAsync(d_w_tied, 0, (size_t)P * P * sizeof(float), stream));
    for (int l = n_tied; l >= 1; --l) {
      float* cur = slot(l);
      const float* al = tied_out + (size_t)(l - 1) * slab;
      const float* ap = (l == 1) ? tied_in : tied_out + (size_t)(l - 2) * slab;
      k_relu_mask<<<ceil_div((long long)slab, 256), 256, 0, stream>>>(cur, al, (long long)slab);
      MPNN_CHECK_LAUNCH("k_relu_mask");
      int rc;
      if (!tc_dw) {
        rc = mpnn_gemm(cur, ap, d_w_tied, P, P, R, 1, PP, PP, 1, P, nullptr, 2, sub, sub_bytes, stream);
        if (rc) return rc;
      }
      float* nxt = (l == 1) ? (stacked ? dA : (cur == dA ? dB : dA)) : slot(l - 1);
      // d_{l-1}[r, i] = sum_o delta[r, o] W[o, i] = sum_o delta[r, o] Wt[i, o]: operand Wt is k(=o)-contiguous
      rc = mpnn_gemm(cur, Wt, nxt, R, P, P, PP, 1, 1, P, PP, nullptr, 0, sub, sub_bytes, stream);
      if (rc) return rc;
    }
    if (!stacked) {  // the gradient w.r.t. the tied input must end up in dA
      float* last = (slot(1) == dA) ? dB : dA;
      if (last != dA) MPNN_CUDA(cudaMemcpyAsync(dA, last, slab * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    }
    if (tc_dw) {
      // dW[o, i] = sum_{l, r} delta_l[r, o] a_{l-1}[r, i]: X = the delta stack, D = the saved activations (tied input
      // followed by the outputs of layers 1..L-1 are contiguous in `saved`), K = L*R stacked rows
      const int blk = wide_block(P);
      const long long krows = (long long)n_tied * R;
      for (int mb = 0; mb < P / blk; ++mb) {
        int rc = mpnn_tc_dense_gemm_tn(stack + (size_t)mb * blk, krows, PP, blk, tied_in, PP, blk, P / blk, blk, blk,
                                       d_w_tied + (size_t)mb * blk * P, blk, P, tcws, tcws_bytes, stream);
        if (rc) return rc;
      }
    }
  }

  // growth layers, last to first: dA holds d(output of growth g) with row stride gld[g]
  float* cur = dA;
  float* nxt = dB;
  for (int g = n_growth - 1; g >= 0; --g) {
    const float* ag = saved + lo.goff[g];
    const float* ap = g == 0 ? rows_in : saved + lo.goff[g - 1];
    long long ldp = g == 0 ? ef : lo.gld[g - 1];
    int ld = lo.gld[g];
    k_relu_mask<<<ceil_div((long long)R * ld, 256), 256, 0, stream>>>(cur, ag, (long long)R * ld);
    MPNN_CHECK_LAUNCH("k_relu_mask");
    int rc = mpnn_gemm(cur, ap, d_growth_w[g], lo.gout[g], lo.gin[g], R, 1, ld, ldp, 1, lo.gin[g], nullptr, 0, sub,
                       sub_bytes, stream);
    if (rc) return rc;
    rc = mpnn_colsum(cur, nullptr, R, lo.gout[g], ld, 0, d_growth_b[g], 0, sub, sub_bytes, stream);
    if (rc) return rc;
    if (g > 0 || d_rows_in) {
      float* o = g == 0 ? d_rows_in : nxt;
      long long ldo = g == 0 ? ef : lo.gld[g - 1];
      if (g > 0 && lo.gld[g - 1] != lo.gout[g - 1])
        MPNN_CUDA(cudaMemsetAsync(o, 0, (size_t)R * ldo * sizeof(float), stream));
      rc = mpnn_gemm(cur, growth_w[g], o, R, lo.gin[g], lo.gout[g], ld, 1, lo.gin[g], 1, ldo, nullptr, 0, sub,
                     sub_bytes, stream);
      if (rc) return rc;
    }
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  if (n_growth == 0 && d_rows_in) {
    MPNN_CUDA(cudaMemcpy2DAsync(d_rows_in, (size_t)ef * sizeof(float), dA, (size_t)PP * sizeof(float),
                                (size_t)ef * sizeof(float), R, cudaMemcpyDeviceToDevice, stream));
  }
  return MPNN_OK;
}

}  // extern "C"
