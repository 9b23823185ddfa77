// Masked GRU node update (reference mpnn_functions/update/gru_update.py:26-35, 66-68).
//   [ri|zi|ni] = m W_ih + b_ih ; [rh|zh|nh] = h W_hh + b_hh        (weights stored [d, 3d], gate order r,z,n)
//   r = sigmoid(ri+rh)*mu ; z = sigmoid(zi+zh)*mu ; n = tanh(ni + r*nh)*mu ; h' = ((1-z)*n + z*h)*mu
// Saved for backward: gates[rows, 4d] = (sigmoid_r, sigmoid_z, tanh_n, nh) -- un-masked activations, so any
// float mask value is differentiated exactly like the reference's autograd graph.
#include "common.cuh"

extern "C" int mpnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak,
                         long long sbk, long long sbn, long long ldc, const float* bias, int flags, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_gemm_workspace_bytes(int M, int N, int K);
extern "C" int mpnn_colsum(const float* X, const float* Y, long long rows, int width, long long ldx, long long ldy,
                           float* out, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
extern "C" size_t mpnn_colsum_workspace_bytes(long long rows, int width);

namespace {

__global__ void k_gru_point_fwd(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h,
                                const float* __restrict__ mask, long long rows, int d, float* __restrict__ hout,
                                float* __restrict__ gates) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * d) return;
  long long row = t / d;
  int c = (int)(t - row * d);
  const float mu = mask[row];
  const float* gir = gi + row * 3 * d;
  const float* ghr = gh + row * 3 * d;
  float sr = 1.f / (1.f + expf(-(gir[c] + ghr[c])));
  float sz = 1.f / (1.f + expf(-(gir[d + c] + ghr[d + c])));
  float nh = ghr[2 * d + c];
  float r = sr * mu, z = sz * mu;
  float tn = tanhf(gir[2 * d + c] + r * nh);
  float n = tn * mu;
  float hv = h[t];
  hout[t] = ((1.f - z) * n + z * hv) * mu;
  float* g = gates + row * 4 * d;
  g[c] = sr;
  g[d + c] = sz;
  g[2 * d + c] = tn;
  g[3 * d + c] = nh;
}

__global__ void k_gru_point_bwd(const float* __restrict__ gates, const float* __restrict__ h,
                                const float* __restrict__ mask, const float* __restrict__ dhout, long long rows, int d,
                                float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * d) return;
  long long row = t / d;
  int c = (int)(t - row * d);
  const float mu = mask[row];
  const float* g = gates + row * 4 * d;
  float sr = g[c], sz = g[d + c], tn = g[2 * d + c], nh = g[3 * d + c];
  float r = sr * mu, z = sz * mu, n = tn * mu;
  float go = dhout[t] * mu;
  float dn = go * (1.f - z);
  float dz = go * (h[t] - n);
  float dan = dn * mu * (1.f - tn * tn);
  float dr = dan * nh;
  float dnh = dan * r;
  float dar = dr * mu * sr * (1.f - sr);
  float daz = dz * mu * sz * (1.f - sz);
  float* a = dgi + row * 3 * d;
  float* b = dgh + row * 3 * d;
  a[c] = dar;
  a[d + c] = daz;
  a[2 * d + c] = dan;
  b[c] = dar;
  b[d + c] = daz;
  b[2 * d + c] = dnh;
  dh[t] = go * z;
}

}  // namespace

extern "C" {

size_t mpnn_gru_workspace_bytes(long long rows, int d) {
  size_t pre = 2 * align_up((size_t)rows * 3 * d * sizeof(float), 256);
  size_t g = mpnn_gemm_workspace_bytes(d, 3 * d, (int)rows);
  size_t c = mpnn_colsum_workspace_bytes(rows, 3 * d);
  return pre + align_up(g > c ? g : c, 256);
}

int mpnn_gru_fwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                 const float* b_ih, const float* b_hh, long long rows, int d, float* h_out, float* gates,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && d > 0 && rows < (1ll << 31), MPNN_ERR_ARG, "gru_fwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_gru_workspace_bytes(rows, d), MPNN_ERR_WORKSPACE, "gru_fwd: workspace");
  char* wp = (char*)workspace;
  float* gi = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  float* gh = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  int R = (int)rows;
  int rc = mpnn_gemm(m, W_ih, gi, R, 3 * d, d, d, 1, 3 * d, 1, 3 * d, b_ih, 0, nullptr, 0, stream);
  if (rc) return rc;
  rc = mpnn_gemm(h, W_hh, gh, R, 3 * d, d, d, 1, 3 * d, 1, 3 * d, b_hh, 0, nullptr, 0, stream);
  if (rc) return rc;
  k_gru_point_fwd<<<ceil_div(rows * d, 256), 256, 0, stream>>>(gi, gh, h, mask, rows, d, h_out, gates);
  MPNN_CHECK_LAUNCH("k_gru_point_fwd");
  return MPNN_OK;
}

int mpnn_gru_bwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                 const float* gates, const float* dh_out, long long rows, int d, float* dm, float* dh, float* dW_ih,
                 float* dW_hh, float* db_ih, float* db_hh, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
  MPNN_REQUIRE(rows > 0 && d > 0 && rows < (1ll << 31), MPNN_ERR_ARG, "gru_bwd: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_gru_workspace_bytes(rows, d), MPNN_ERR_WORKSPACE, "gru_bwd: workspace");
  char* wp = (char*)workspace;
  float* dgi = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  float* dgh = (float*)wp;
  wp += align_up((size_t)rows * 3 * d * sizeof(float), 256);
  void* sub = wp;
  size_t sub_bytes = workspace_bytes - (size_t)(wp - (char*)workspace);
  int R = (int)rows;
  k_gru_point_bwd<<<ceil_div(rows * d, 256), 256, 0, stream>>>(gates, h, mask, dh_out, rows, d, dgi, dgh, dh);
  MPNN_CHECK_LAUNCH("k_gru_point_bwd");
  int rc;
  // dm = dgi W_ih^T ; dh += dgh W_hh^T
  if ((rc = mpnn_gemm(dgi, W_ih, dm, R, d, 3 * d, 3 * d, 1, 1, 3 * d, d, nullptr, 0, nullptr, 0, stream))) return rc;
  if ((rc = mpnn_gemm(dgh, W_hh, dh, R, d, 3 * d, 3 * d, 1, 1, 3 * d, d, nullptr, 2, nullptr, 0, stream))) return rc;
  // dW_ih = m^T dgi ; dW_hh = h^T dgh   (split-K, fixed-order reduction)
  if ((rc = mpnn_gemm(m, dgi, dW_ih, d, 3 * d, R, 1, d, 3 * d, 1, 3 * d, nullptr, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_gemm(h, dgh, dW_hh, d, 3 * d, R, 1, d, 3 * d, 1, 3 * d, nullptr, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_colsum(dgi, nullptr, rows, 3 * d, 3 * d, 0, db_ih, 0, sub, sub_bytes, stream))) return rc;
  if ((rc = mpnn_colsum(dgh, nullptr, rows, 3 * d, 3 * d, 0, db_hh, 0, sub, sub_bytes, stream))) return rc;
  return MPNN_OK;
}

}  // extern "C"
