// (1) Row softmax fused with a gating multiply: the per-pair attention gate of AttEdgeNetwork
//     (reference att_edge_network.py:18-26: softmax over the nf features of Linear(cat(h_i, bond)) times the
//     sender state) and the row softmax of WAdjMsgAgg (weighted_adjacent_message_agg.py:20).
// (2) Dense weighted aggregation over senders for messages that really are a dense [B,N,N,mf] tensor
//     (the stand-alone contract of the three aggregators: adjacent_message_agg.py:18,
//     weighted_adjacent_message_agg.py:20, attention_message_agg.py:24).  Pure HBM streaming.
#include "common.cuh"

namespace {

constexpr int KMAX = 32;  // width <= 1024

// gate = softmax(logits[row,:]); out = gate * (V ? V[row,:] : 1).  One warp per row.
__global__ void __launch_bounds__(256) k_softmax_mul_fwd(const float* __restrict__ logits, const float* __restrict__ V,
                                                         long long rows, int n, float* __restrict__ gate,
                                                         float* __restrict__ out) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nk = (n + 31) / 32;
  const float* lr = logits + row * n;
  float s[KMAX];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < n) {
      s[k] = lr[o];
      mx = fmaxf(mx, s[k]);
    }
  }
  mx = warp_max(mx);
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < n) {
      s[k] = expf(s[k] - mx);
      den += s[k];
    }
  }
  den = warp_sum(den);
  const float inv = 1.f / den;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < n) {
      float g = s[k] * inv;
      if (gate) gate[row * n + o] = g;
      out[row * n + o] = V ? g * V[row * n + o] : g;
    }
  }
}

// dV = dout*gate ; dlogits = gate * (dgate - sum(dgate*gate)), dgate = dout * (V ? V : 1)
__global__ void __launch_bounds__(256) k_softmax_mul_bwd(const float* __restrict__ gate, const float* __restrict__ V,
                                                         const float* __restrict__ dout, long long rows, int n,
                                                         float* __restrict__ dlogits, float* __restrict__ dV) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nk = (n + 31) / 32;
  float g[KMAX], dg[KMAX];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < n) {
      g[k] = gate[row * n + o];
      float d = dout[row * n + o];
      dg[k] = V ? d * V[row * n + o] : d;
      if (dV) dV[row * n + o] = d * g[k];
      dot += dg[k] * g[k];
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    int o = lane + 32 * k;
    if (k < nk && o < n) dlogits[row * n + o] = g[k] * (dg[k] - dot);
  }
}

// Rows of at most 32 values (the attention gate at nf <= 32, the per-atom readout of config 4 at 16): LANES lanes per
// row (8, 16 or 32), so a warp handles 32 / LANES rows and nothing is predicated over the 32-slice general form above
// (102 400 rows of 16 took 137 + 58 us there, one half-empty warp per row walking 32 predicated slices).
template <int LANES>
__global__ void __launch_bounds__(256) k_softmax_mul_fwd_small(const float* __restrict__ logits, const float* __restrict__ V,
                                                               long long rows, int n, float* __restrict__ gate,
                                                               float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t / LANES;
  const int o = (int)(t % LANES);
  const bool ok = row < rows && o < n;
  float v = ok ? logits[row * n + o] : -INFINITY;
  float mx = v;
#pragma unroll
  for (int w = LANES / 2; w > 0; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
  const float e = ok ? expf(v - mx) : 0.f;
  float den = e;
#pragma unroll
  for (int w = LANES / 2; w > 0; w >>= 1) den += __shfl_xor_sync(0xffffffffu, den, w);
  if (ok) {
    const float g = e / den;
    if (gate) gate[row * n + o] = g;
    out[row * n + o] = V ? g * V[row * n + o] : g;
  }
}

template <int LANES>
__global__ void __launch_bounds__(256) k_softmax_mul_bwd_small(const float* __restrict__ gate, const float* __restrict__ V,
                                                               const float* __restrict__ dout, long long rows, int n,
                                                               float* __restrict__ dlogits, float* __restrict__ dV) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t / LANES;
  const int o = (int)(t % LANES);
  const bool ok = row < rows && o < n;
  float g = 0.f, dg = 0.f;
  if (ok) {
    g = gate[row * n + o];
    const float d = dout[row * n + o];
    dg = V ? d * V[row * n + o] : d;
    if (dV) dV[row * n + o] = d * g;
  }
  float dot = dg * g;
#pragma unroll
  for (int w = LANES / 2; w > 0; w >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, w);
  if (ok) dlogits[row * n + o] = g * (dg - dot);
}

// out[r,k] = sum_j w[r,j] * m[r,j,k]
__global__ void k_dense_agg_fwd(const float* __restrict__ m, const float* __restrict__ w, long long R, int N, int mf,
                                float* __restrict__ out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * mf) return;
  long long r = t / mf;
  int k = (int)(t - r * mf);
  const float* mr = m + r * N * mf + k;
  const float* wr = w + r * N;
  float s = 0.f;
  for (int j = 0; j < N; ++j) s = fmaf(wr[j], mr[(size_t)j * mf], s);
  out[t] = s;
}
// dm[r,j,k] = w[r,j] * dout[r,k]
__global__ void k_dense_agg_bwd_m(const float* __restrict__ w, const float* __restrict__ dout, long long R, int N,
                                  int mf, float* __restrict__ dm) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * N * mf) return;
  long long rj = t / mf;
  int k = (int)(t - rj * mf);
  long long r = rj / N;
  dm[t] = w[rj] * dout[r * mf + k];
}
// dw[r,j] = sum_k m[r,j,k] * dout[r,k]   (one warp per (r,j))
__global__ void k_dense_agg_bwd_w(const float* __restrict__ m, const float* __restrict__ dout, long long RN, int N,
                                  int mf, float* __restrict__ dw) {
  const long long rj = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (rj >= RN) return;
  const long long r = rj / N;
  float s = 0.f;
  for (int k = lane; k < mf; k += 32) s = fmaf(m[rj * mf + k], dout[r * mf + k], s);
  s = warp_sum(s);
  if (lane == 0) dw[rj] = s;
}

}  // namespace

extern "C" {

int mpnn_softmax_mul_fwd(const float* logits, const float* V, long long rows, int n, float* gate, float* out,
                         cudaStream_t stream) {
  MPNN_REQUIRE(rows >= 0 && n > 0 && n <= 32 * KMAX, MPNN_ERR_ARG, "softmax_mul_fwd: bad dims rows=%lld n=%d", rows, n);
  if (rows == 0) return MPNN_OK;
  if (n <= 8) k_softmax_mul_fwd_small<8><<<ceil_div(rows * 8, 256), 256, 0, stream>>>(logits, V, rows, n, gate, out);
  else if (n <= 16) k_softmax_mul_fwd_small<16><<<ceil_div(rows * 16, 256), 256, 0, stream>>>(logits, V, rows, n, gate, out);
  else if (n <= 32) k_softmax_mul_fwd_small<32><<<ceil_div(rows * 32, 256), 256, 0, stream>>>(logits, V, rows, n, gate, out);
  else k_softmax_mul_fwd<<<ceil_div(rows * 32, 256), 256, 0, stream>>>(logits, V, rows, n, gate, out);
  MPNN_CHECK_LAUNCH("k_softmax_mul_fwd");
  return MPNN_OK;
}

int mpnn_softmax_mul_bwd(const float* gate, const float* V, const float* dout, long long rows, int n, float* dlogits,
                         float* dV, cudaStream_t stream) {
  MPNN_REQUIRE(rows >= 0 && n > 0 && n <= 32 * KMAX, MPNN_ERR_ARG, "softmax_mul_bwd: bad dims");
  if (rows == 0) return MPNN_OK;
  if (n <= 8) k_softmax_mul_bwd_small<8><<<ceil_div(rows * 8, 256), 256, 0, stream>>>(gate, V, dout, rows, n, dlogits, dV);
  else if (n <= 16) k_softmax_mul_bwd_small<16><<<ceil_div(rows * 16, 256), 256, 0, stream>>>(gate, V, dout, rows, n, dlogits, dV);
  else if (n <= 32) k_softmax_mul_bwd_small<32><<<ceil_div(rows * 32, 256), 256, 0, stream>>>(gate, V, dout, rows, n, dlogits, dV);
  else k_softmax_mul_bwd<<<ceil_div(rows * 32, 256), 256, 0, stream>>>(gate, V, dout, rows, n, dlogits, dV);
  MPNN_CHECK_LAUNCH("k_softmax_mul_bwd");
  return MPNN_OK;
}

int mpnn_dense_agg_fwd(const float* messages, const float* weights, long long R, int N, int mf, float* out,
                       cudaStream_t stream) {
  MPNN_REQUIRE(R > 0 && N > 0 && mf > 0, MPNN_ERR_ARG, "dense_agg_fwd: bad dims");
  k_dense_agg_fwd<<<ceil_div(R * mf, 256), 256, 0, stream>>>(messages, weights, R, N, mf, out);
  MPNN_CHECK_LAUNCH("k_dense_agg_fwd");
  return MPNN_OK;
}

int mpnn_dense_agg_bwd(const float* messages, const float* weights, const float* dout, long long R, int N, int mf,
                       float* dmessages, float* dweights, cudaStream_t stream) {
  MPNN_REQUIRE(R > 0 && N > 0 && mf > 0, MPNN_ERR_ARG, "dense_agg_bwd: bad dims");
  if (dmessages) k_dense_agg_bwd_m<<<ceil_div(R * N * mf, 256), 256, 0, stream>>>(weights, dout, R, N, mf, dmessages);
  if (dweights)
    k_dense_agg_bwd_w<<<ceil_div(R * N * 32, 256), 256, 0, stream>>>(messages, dout, R * N, N, mf, dweights);
  MPNN_CHECK_LAUNCH("k_dense_agg_bwd");
  return MPNN_OK;
}

}  // extern "C"
