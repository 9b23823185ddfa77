// Set2Vec readout (reference mpnn_functions/readout/set2vec.py:93-151, inner_prod="default") with its
// input-less LSTM cell (LSTMCellHidden, set2vec.py:68-75).  `steps` (default 100) strictly sequential
// iterations; each one: LSTM gates from the previous [m | read] vector, a query, additive attention
// energies over ALL B*N rows, a softmax across the whole batch (set2vec.py:139, dim=0), and the per-graph
// read-out.  The whole loop (forward or backward-through-time) is ONE C call that enqueues every launch;
// the host never synchronises.  Reductions have a fixed order (bit-reproducible).
//
// Wcat = [w_hi | w_hf | w_hg | w_ho] ([2F, 4F]), bcat likewise ([4F]); gate order i, f, g, o.
// saved, per step: m [B,2F] | c [B,F] | gates [B,4F] (activated) | tanh(c) [B,F] | q [B,F] | att [B*N]
#include "common.cuh"

extern "C" int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
                         long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_gemm_workspace_bytes(int M, int N, int K);
extern "C" int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy,
                           float* out, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_colsum_workspace_bytes(long long rows, int width);

// the persistent kernels (s2v_persist.cu): return 1 if they served the call (*rc = status), 0 if the shape is not theirs
size_t s2v_persist_slot_bytes();
int s2v_persist_fwd(const float* X, const float* mask, const float* Wcat, const float* bcat, const float* Wq,
                    const float* we, const float* m0, const float* c0, int B, int N, int F, int steps, float* out,
                    float* saved, void* slots, cudaStream_t stream, int* rc);
int s2v_persist_bwd(const float* X, const float* Wcat, const float* Wq, const float* we, const float* c0,
                    const float* saved, const float* dout, int B, int N, int F, int steps, float* dX, float* dqS,
                    float* pwS, float* dpreS, float* dm0, float* dc0, void* slots, cudaStream_t stream, int* rc);

namespace {

constexpr float BIG_NEGATIVE = -1e8f;  // set2vec.py:10

struct StepPtrs {
  float *m, *c, *gates, *tc, *q, *att;
};

__host__ __device__ inline size_t step_stride(int B, int N, int F) { return (size_t)B * 9 * F + (size_t)B * N; }

inline StepPtrs step_ptrs(float* saved, int s, int B, int N, int F) {
  float* p = saved + (size_t)s * step_stride(B, N, F);
  StepPtrs r;
  r.m = p;
  r.c = r.m + (size_t)B * 2 * F;
  r.gates = r.c + (size_t)B * F;
  r.tc = r.gates + (size_t)B * 4 * F;
  r.q = r.tc + (size_t)B * F;
  r.att = r.q + (size_t)B * F;
  return r;
}

// pre [B,4F] -> activated gates, c', h' (written to m[:, :F], row stride ldm)
__global__ void k_lstm_fwd(const float* __restrict__ pre, const float* __restrict__ cprev, int B, int F,
                           float* __restrict__ gates, float* __restrict__ c, float* __restrict__ tc,
                           float* __restrict__ m, int ldm) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * F) return;
  int b = t / F, f = t - b * F;
  const float* p = pre + (size_t)b * 4 * F;
  float i = 1.f / (1.f + expf(-p[f]));
  float fg = 1.f / (1.f + expf(-p[F + f]));
  float g = tanhf(p[2 * F + f]);
  float o = 1.f / (1.f + expf(-p[3 * F + f]));
  float cp = cprev ? cprev[t] : 0.f;
  float cn = fg * cp + i * g;
  float th = tanhf(cn);
  float* gs = gates + (size_t)b * 4 * F;
  gs[f] = i;
  gs[F + f] = fg;
  gs[2 * F + f] = g;
  gs[3 * F + f] = o;
  c[t] = cn;
  tc[t] = th;
  m[(size_t)b * ldm + f] = o * th;
}

// e[r] = sum_f we[f] * tanh(q[b,f] + X[r,f]) + (1-mask[r]) * BIG_NEGATIVE.   One warp per row.
__global__ void k_energy(const float* __restrict__ X, const float* __restrict__ q, const float* __restrict__ we,
                         const float* __restrict__ mask, int rows, int N, int F, float* __restrict__ e) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  int b = row / N;
  float s = 0.f;
  for (int f = lane; f < F; f += 32) s = fmaf(we[f], tanhf(q[(size_t)b * F + f] + X[(size_t)row * F + f]), s);
  s = warp_sum(s);
  if (lane == 0) e[row] = mask ? s + (1.f - mask[row]) * BIG_NEGATIVE : s;
}

__device__ float block_reduce_1024(float v, float* red, bool is_max) {
  // fixed-order tree over 1024 threads
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      float a = red[threadIdx.x], b = red[threadIdx.x + o];
      red[threadIdx.x] = is_max ? fmaxf(a, b) : a + b;
    }
    __syncthreads();
  }
  float r = red[0];
  __syncthreads();
  return r;
}

// att = softmax(e) over all rows (single block of 1024 threads)
__global__ void __launch_bounds__(1024) k_global_softmax(const float* __restrict__ e, int rows,
                                                         float* __restrict__ att) {
  __shared__ float red[1024];
  float mx = -INFINITY;
  for (int r = threadIdx.x; r < rows; r += 1024) mx = fmaxf(mx, e[r]);
  mx = block_reduce_1024(mx, red, true);
  float s = 0.f;
  for (int r = threadIdx.x; r < rows; r += 1024) s += expf(e[r] - mx);
  s = block_reduce_1024(s, red, false);
  float inv = 1.f / s;
  for (int r = threadIdx.x; r < rows; r += 1024) att[r] = expf(e[r] - mx) * inv;
}

// read[b,f] = sum_i att[b,i] X[b,i,f]  -> m[b, F+f]
__global__ void k_read(const float* __restrict__ X, const float* __restrict__ att, int B, int N, int F,
                       float* __restrict__ m) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * F) return;
  int b = t / F, f = t - b * F;
  float s = 0.f;
  for (int i = 0; i < N; ++i) s = fmaf(att[(size_t)b * N + i], X[((size_t)b * N + i) * F + f], s);
  m[(size_t)b * 2 * F + F + f] = s;
}

// ---- backward pieces ----
// datt[r] = sum_f dm[b, F+f] * X[r,f]
__global__ void k_datt(const float* __restrict__ X, const float* __restrict__ dm, int rows, int N, int F,
                       float* __restrict__ datt) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  int b = row / N;
  float s = 0.f;
  for (int f = lane; f < F; f += 32) s = fmaf(dm[(size_t)b * 2 * F + F + f], X[(size_t)row * F + f], s);
  s = warp_sum(s);
  if (lane == 0) datt[row] = s;
}
// de[r] = att[r] * (datt[r] - sum att*datt)   (single block)
__global__ void __launch_bounds__(1024) k_global_softmax_bwd(const float* __restrict__ att,
                                                             const float* __restrict__ datt, int rows,
                                                             float* __restrict__ de) {
  __shared__ float red[1024];
  float s = 0.f;
  for (int r = threadIdx.x; r < rows; r += 1024) s = fmaf(att[r], datt[r], s);
  s = block_reduce_1024(s, red, false);
  for (int r = threadIdx.x; r < rows; r += 1024) de[r] = att[r] * (datt[r] - s);
}
// one block per graph, thread per feature (looped): dq[b,f], pw[b,f] (partial d we), dX += att*dread + dpre
__global__ void k_energy_bwd(const float* __restrict__ X, const float* __restrict__ q, const float* __restrict__ we,
                             const float* __restrict__ att, const float* __restrict__ de,
                             const float* __restrict__ dm, int N, int F, float* __restrict__ dq,
                             float* __restrict__ pw, float* __restrict__ dX) {
  int b = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float qv = q[(size_t)b * F + f], w = we[f], dread = dm[(size_t)b * 2 * F + F + f];
    float sq = 0.f, sw = 0.f;
    for (int i = 0; i < N; ++i) {
      size_t r = (size_t)b * N + i;
      float th = tanhf(qv + X[r * F + f]);
      float d = de[r];
      float dpre = d * w * (1.f - th * th);
      sq += dpre;
      sw = fmaf(d, th, sw);
      dX[r * F + f] += att[r] * dread + dpre;
    }
    dq[(size_t)b * F + f] = sq;
    pw[(size_t)b * F + f] = sw;
  }
}
// dh [B,F] (already = dq Wq + dm[:, :F]) and dc_next -> dpre [B,4F], dc_prev
__global__ void k_lstm_bwd(const float* __restrict__ gates, const float* __restrict__ tc,
                           const float* __restrict__ cprev, const float* __restrict__ dh,
                           const float* __restrict__ dm, float* __restrict__ dc, int B, int F,
                           float* __restrict__ dpre) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * F) return;
  int b = t / F, f = t - b * F;
  const float* gs = gates + (size_t)b * 4 * F;
  float i = gs[f], fg = gs[F + f], g = gs[2 * F + f], o = gs[3 * F + f];
  float th = tc[t];
  float dhv = dh[t] + (dm ? dm[(size_t)b * 2 * F + f] : 0.f);   // dh = dq Wq + dm[:, :F]
  float dcn = dhv * o * (1.f - th * th) + dc[t];
  float cp = cprev ? cprev[t] : 0.f;
  float* dp = dpre + (size_t)b * 4 * F;
  dp[f] = dcn * g * i * (1.f - i);
  dp[F + f] = dcn * cp * fg * (1.f - fg);
  dp[2 * F + f] = dcn * i * (1.f - g * g);
  dp[3 * F + f] = dhv * th * o * (1.f - o);
  dc[t] = dcn * fg;
}
}  // namespace

extern "C" {

long long mpnn_set2vec_saved_floats(int B, int N, int F, int steps) {
  return (long long)steps * (long long)step_stride(B, N, F);
}

size_t mpnn_set2vec_workspace_bytes(int B, int N, int F) {
  size_t rows = (size_t)B * N;
  size_t fl = (size_t)B * 4 * F      // pre / dpre
              + 3 * rows             // e / datt / de
              + (size_t)B * 2 * F * 2  // dm ping-pong
              + (size_t)B * F * 4;   // dc, dq, pw, dh
  size_t g = mpnn_gemm_workspace_bytes(2 * F, 4 * F, B);
  size_t c = mpnn_colsum_workspace_bytes(B, 4 * F);
  return s2v_persist_slot_bytes() + align_up(fl * sizeof(float), 256) + align_up(g > c ? g : c, 256) + 256;
}

// The backward keeps every step's dq / pw / dpre and a packed copy of the saved m so that the four parameter gradients
// are ONE product (or column sum) over steps*B rows each instead of one small accumulate-launch per step.
size_t mpnn_set2vec_bwd_workspace_bytes(int B, int N, int F, int steps) {
  size_t rows = (size_t)B * N;
  size_t sb = (size_t)steps * B;
  size_t fl = 3 * rows                 // datt / de / spare
              + (size_t)B * 2 * F * 2  // dm ping-pong
              + (size_t)B * F * 2      // dc, dh
              + sb * F * 2             // dq, pw stacks
              + sb * 4 * F             // dpre stack
              + sb * 2 * F;            // packed m
  size_t g = mpnn_gemm_workspace_bytes(2 * F, 4 * F, (int)sb);
  size_t g2 = mpnn_gemm_workspace_bytes(F, F, (int)sb);
  size_t c = mpnn_colsum_workspace_bytes((int)sb, 4 * F);
  size_t m = g > g2 ? g : g2;
  return s2v_persist_slot_bytes() + align_up(fl * sizeof(float), 256) + align_up(m > c ? m : c, 256) + 256;
}

int mpnn_set2vec_fwd(const float* X, const float* mask, const float* Wcat, const float* bcat, const float* Wq,
                     const float* we, const float* m0, const float* c0, int B, int N, int F, int steps, float* out,
                     float* saved, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && F > 0 && steps > 0, MPNN_ERR_ARG, "set2vec_fwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_set2vec_workspace_bytes(B, N, F), MPNN_ERR_WORKSPACE, "set2vec_fwd: workspace");
  const int rows = B * N;
  {
    int rc = MPNN_OK;
    if (s2v_persist_fwd(X, mask, Wcat, bcat, Wq, we, m0, c0, B, N, F, steps, out, saved, workspace, stream, &rc))
      return rc;
  }
  float* pre = (float*)((char*)workspace + s2v_persist_slot_bytes());
  float* e = pre + (size_t)B * 4 * F;
  for (int s = 0; s < steps; ++s) {
    StepPtrs cur = step_ptrs(saved, s, B, N, F);
    int rc;
    if (s == 0 && !m0) {
      // m_prev = 0: pre-activations are just the biases
      rc = mpnn_gemm(nullptr, nullptr, pre, B, 4 * F, 0, 0, 0, 0, 0, 4 * F, bcat, 0, nullptr, 0, stream);
    } else if (s == 0) {
      // caller-supplied initial state (set2vec.py:111-117): m0 = cat(mprev, 0) [B, 2F]
      rc = mpnn_gemm(m0, Wcat, pre, B, 4 * F, 2 * F, 2 * F, 1, 4 * F, 1, 4 * F, bcat, 0, nullptr, 0, stream);
    } else {
      StepPtrs prev = step_ptrs(saved, s - 1, B, N, F);
      rc = mpnn_gemm(prev.m, Wcat, pre, B, 4 * F, 2 * F, 2 * F, 1, 4 * F, 1, 4 * F, bcat, 0, nullptr, 0, stream);
    }
    if (rc) return rc;
    const float* cprev = s == 0 ? c0 : step_ptrs(saved, s - 1, B, N, F).c;
    k_lstm_fwd<<<ceil_div(B * F, 256), 256, 0, stream>>>(pre, cprev, B, F, cur.gates, cur.c, cur.tc, cur.m, 2 * F);
    // q = h Wq^T  (h = m[:, :F], row stride 2F; Wq is nn.Linear weight [F_out, F_in])
    if ((rc = mpnn_gemm(cur.m, Wq, cur.q, B, F, F, 2 * F, 1, 1, F, F, nullptr, 0, nullptr, 0, stream))) return rc;
    k_energy<<<ceil_div((long long)rows * 32, 256), 256, 0, stream>>>(X, cur.q, we, mask, rows, N, F, e);
    k_global_softmax<<<1, 1024, 0, stream>>>(e, rows, cur.att);
    k_read<<<ceil_div(B * F, 256), 256, 0, stream>>>(X, cur.att, B, N, F, cur.m);
    MPNN_CHECK_LAUNCH("set2vec_fwd step");
  }
  StepPtrs last = step_ptrs(saved, steps - 1, B, N, F);
  MPNN_CUDA(cudaMemcpyAsync(out, last.m, (size_t)B * 2 * F * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  return MPNN_OK;
}

// dWcat [2F,4F], dbcat [4F], dWq [F,F], dwe [F], dX [B,N,F] are written; dm0 [B,2F] / dc0 [B,F] when non-NULL
// (gradients of the caller-supplied initial state m0 / c0, NULL = the reference's zero initial state).
int mpnn_set2vec_bwd(const float* X, const float* mask, const float* Wcat, const float* Wq, const float* we,
                     const float* m0, const float* c0, const float* saved_c, const float* dout, int B, int N, int F,
                     int steps, float* dX, float* dWcat, float* dbcat, float* dWq, float* dwe, float* dm0, float* dc0,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && F > 0 && steps > 0, MPNN_ERR_ARG, "set2vec_bwd: bad dims");
  MPNN_REQUIRE((long long)steps * B < (1ll << 31), MPNN_ERR_UNSUPPORTED, "set2vec_bwd: steps*B too large");
  MPNN_REQUIRE(workspace_bytes >= mpnn_set2vec_bwd_workspace_bytes(B, N, F, steps), MPNN_ERR_WORKSPACE,
               "set2vec_bwd: workspace");
  (void)mask;
  float* saved = const_cast<float*>(saved_c);
  const int rows = B * N;
  const size_t sb = (size_t)steps * B;
  float* datt = (float*)((char*)workspace + s2v_persist_slot_bytes());
  float* de = datt + rows;
  float* spare = de + rows;
  float* dmA = spare + rows;
  float* dmB = dmA + (size_t)B * 2 * F;
  float* dc = dmB + (size_t)B * 2 * F;
  float* dh = dc + (size_t)B * F;
  float* dqS = dh + (size_t)B * F;
  float* pwS = dqS + sb * F;
  float* dpreS = pwS + sb * F;
  float* mS = dpreS + sb * 4 * F;
  float* fl_end = mS + sb * 2 * F;
  char* sub = (char*)datt + align_up((size_t)(fl_end - datt) * sizeof(float), 256);
  size_t sub_bytes = workspace_bytes - (size_t)(sub - (char*)workspace);
  // packed m: step s at rows [s*B, (s+1)*B)
  MPNN_CUDA(cudaMemcpy2DAsync(mS, (size_t)B * 2 * F * sizeof(float), saved, step_stride(B, N, F) * sizeof(float),
                              (size_t)B * 2 * F * sizeof(float), steps, cudaMemcpyDeviceToDevice, stream));
  int rc;
  int prc = MPNN_OK;
  const bool persistent = s2v_persist_bwd(X, Wcat, Wq, we, c0, saved, dout, B, N, F, steps, dX, dqS, pwS, dpreS,
                                          m0 ? dm0 : nullptr, dc0, workspace, stream, &prc) != 0;
  if (persistent && prc) return prc;
  if (!persistent) {
  MPNN_CUDA(cudaMemsetAsync(dX, 0, (size_t)rows * F * sizeof(float), stream));
  MPNN_CUDA(cudaMemsetAsync(dc, 0, (size_t)B * F * sizeof(float), stream));
  MPNN_CUDA(cudaMemcpyAsync(dmA, dout, (size_t)B * 2 * F * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  float* dm = dmA;
  float* dm_prev = dmB;
  for (int s = steps - 1; s >= 0; --s) {
    StepPtrs cur = step_ptrs(saved, s, B, N, F);
    float* dq = dqS + (size_t)s * B * F;
    float* pw = pwS + (size_t)s * B * F;
    float* dpre = dpreS + (size_t)s * B * 4 * F;
    k_datt<<<ceil_div((long long)rows * 32, 256), 256, 0, stream>>>(X, dm, rows, N, F, datt);
    k_global_softmax_bwd<<<1, 1024, 0, stream>>>(cur.att, datt, rows, de);
    k_energy_bwd<<<B, 128, 0, stream>>>(X, cur.q, we, cur.att, de, dm, N, F, dq, pw, dX);
    MPNN_CHECK_LAUNCH("set2vec_bwd energy");
    // dh = dq Wq (+ dm[:, :F], added inside k_lstm_bwd)
    if ((rc = mpnn_gemm(dq, Wq, dh, B, F, F, F, 1, F, 1, F, nullptr, 0, nullptr, 0, stream))) return rc;
    const float* cprev = s == 0 ? c0 : step_ptrs(saved, s - 1, B, N, F).c;
    k_lstm_bwd<<<ceil_div(B * F, 256), 256, 0, stream>>>(cur.gates, cur.tc, cprev, dh, dm, dc, B, F, dpre);
    MPNN_CHECK_LAUNCH("set2vec_bwd lstm");
    if (s > 0) {
      // dm_prev = dpre Wcat^T
      if ((rc = mpnn_gemm(dpre, Wcat, dm_prev, B, 2 * F, 4 * F, 4 * F, 1, 1, 4 * F, 2 * F, nullptr, 0, nullptr, 0,
                          stream)))
        return rc;
      float* t = dm;
      dm = dm_prev;
      dm_prev = t;
    } else if (m0 && dm0) {
      if ((rc = mpnn_gemm(dpre, Wcat, dm0, B, 2 * F, 4 * F, 4 * F, 1, 1, 4 * F, 2 * F, nullptr, 0, nullptr, 0, stream)))
        return rc;
    }
  }
  if (dc0) MPNN_CUDA(cudaMemcpyAsync(dc0, dc, (size_t)B * F * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  // parameter gradients over all steps at once
  if ((rc = mpnn_colsum(pwS, nullptr, (int)sb, F, F, 0, dwe, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_colsum(dpreS, nullptr, (int)sb, 4 * F, 4 * F, 0, dbcat, 0, sub, sub_bytes, stream))) return rc;
  // dWq = sum_s dq_s^T h_s
  if ((rc = mpnn_gemm(dqS, mS, dWq, F, F, (int)sb, 1, F, 2 * F, 1, F, nullptr, 0, sub, sub_bytes, stream))) return rc;
  // dWcat = sum_{s>=1} m_{s-1}^T dpre_s  (step 0 has no recurrent input)
  if (steps > 1) {
    if ((rc = mpnn_gemm(mS, dpreS + (size_t)B * 4 * F, dWcat, 2 * F, 4 * F, (int)(sb - B), 1, 2 * F, 4 * F, 1, 4 * F,
                        nullptr, 0, sub, sub_bytes, stream)))
      return rc;
  } else {
    MPNN_CUDA(cudaMemsetAsync(dWcat, 0, (size_t)2 * F * 4 * F * sizeof(float), stream));
  }
  if (m0) {  // step 0 read the caller's initial state: dWcat += m0^T dpre_0
    if ((rc = mpnn_gemm(m0, dpreS, dWcat, 2 * F, 4 * F, B, 1, 2 * F, 4 * F, 1, 4 * F, nullptr, 2, sub, sub_bytes,
                        stream)))
      return rc;
  }
  return MPNN_OK;
}

// ---- the input-less LSTM cell alone (set2vec.py:68-75): pre [B,4F] = hprev Wcat + bcat is the caller's GEMM --------
// gates [B,4F] (activated i,f,g,o), c [B,F], tc [B,F] = tanh(c), h [B,F] are written.
int mpnn_lstm_hidden_fwd(const float* pre, const float* cprev, int B, int F, float* gates, float* c, float* tc,
                         float* h, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && F > 0, MPNN_ERR_ARG, "lstm_hidden_fwd: bad dims");
  k_lstm_fwd<<<ceil_div(B * F, 256), 256, 0, stream>>>(pre, cprev, B, F, gates, c, tc, h, F);
  MPNN_CHECK_LAUNCH("lstm_hidden_fwd");
  return MPNN_OK;
}

// dh, dc_next [B,F] -> dpre [B,4F], dc_prev [B,F]
int mpnn_lstm_hidden_bwd(const float* gates, const float* tc, const float* cprev, const float* dh, const float* dc_next,
                         int B, int F, float* dpre, float* dc_prev, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && F > 0, MPNN_ERR_ARG, "lstm_hidden_bwd: bad dims");
  MPNN_CUDA(cudaMemcpyAsync(dc_prev, dc_next, (size_t)B * F * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  k_lstm_bwd<<<ceil_div(B * F, 256), 256, 0, stream>>>(gates, tc, cprev, dh, nullptr, dc_prev, B, F, dpre);
  MPNN_CHECK_LAUNCH("lstm_hidden_bwd");
  return MPNN_OK;
}

}  // extern "C"
