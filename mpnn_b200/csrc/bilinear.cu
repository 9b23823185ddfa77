// BiLiniearEdgeNetwork (reference mpnn_functions/message/bilinear_edge_network.py:25-37): a parameter-free message
// function.  The bond row of a pair, viewed as an [nf, nf, nf] tensor X (so ef == nf^3), contracts with the sender
// state on its first index and the receiver state on its last:
//     Y[e, p] = sum_{a, q} h[src_e, a] * X_e[a, p, q] * h[dst_e, q]
// Pairs with an all-zero bond row give exactly 0, so the compacted edge list carries the whole (dense, [B,N,N,nf])
// reference output.  fp32 CUDA cores; the op reads nf^3 floats per edge and is HBM-bound on the bond rows.
#include "common.cuh"

namespace {

// one thread per (edge, p)
__global__ void __launch_bounds__(256) k_bil_fwd(const int* __restrict__ edge_src, const int* __restrict__ edge_dst,
                                                 const float* __restrict__ X, long long ldx,
                                                 const float* __restrict__ H, long long E, int nf,
                                                 float* __restrict__ Y) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * nf) return;
  const long long e = t / nf;
  const int p = (int)(t - e * nf);
  const float* hs = H + (size_t)edge_src[e] * nf;
  const float* hd = H + (size_t)edge_dst[e] * nf;
  const float* x = X + e * ldx + (size_t)p * nf;
  float acc = 0.f;
  for (int a = 0; a < nf; ++a) {
    float s = 0.f;
    const float* xa = x + (size_t)a * nf * nf;
    for (int q = 0; q < nf; ++q) s = fmaf(xa[q], hd[q], s);
    acc = fmaf(hs[a], s, acc);
  }
  Y[t] = acc;
}

// dX_e[a, p, q] = dY[e, p] h[src, a] h[dst, q]: one thread per element of the bond row
__global__ void __launch_bounds__(256) k_bil_bwd_x(const int* __restrict__ edge_src, const int* __restrict__ edge_dst,
                                                   const float* __restrict__ H, const float* __restrict__ dY,
                                                   long long E, int nf, float* __restrict__ dX, long long lddx) {
  const int n3 = nf * nf * nf;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * n3) return;
  const long long e = t / n3;
  const int r = (int)(t - e * n3);
  const int a = r / (nf * nf), p = (r / nf) % nf, q = r % nf;
  dX[e * lddx + r] = dY[e * nf + p] * H[(size_t)edge_src[e] * nf + a] * H[(size_t)edge_dst[e] * nf + q];
}

// dH[i, c] = sum_{e in E(i)} sum_{a,p} dY[e,p] h[src_e,a] X_e[a,p,c]              (receiver role, CSR)
//          + sum_{e: src_e = i} sum_{p,q} dY[e,p] X_e[c,p,q] h[dst_e,q]           (sender role, CSC)
// one thread per (node row, c): gather-only, fixed order
__global__ void __launch_bounds__(256) k_bil_bwd_h(const int* __restrict__ row_ptr, const int* __restrict__ col_ptr,
                                                   const int* __restrict__ csc_eid, const int* __restrict__ edge_src,
                                                   const int* __restrict__ edge_dst, const float* __restrict__ X,
                                                   long long ldx, const float* __restrict__ H,
                                                   const float* __restrict__ dY, int n_rows, int nf,
                                                   float* __restrict__ dH) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_rows * nf) return;
  const int i = (int)(t / nf), c = (int)(t - (long long)i * nf);
  float acc = 0.f;
  for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
    const float* hs = H + (size_t)edge_src[e] * nf;
    const float* x = X + (size_t)e * ldx;
    const float* dy = dY + (size_t)e * nf;
    for (int a = 0; a < nf; ++a) {
      float s = 0.f;
      for (int p = 0; p < nf; ++p) s = fmaf(dy[p], x[((size_t)a * nf + p) * nf + c], s);
      acc = fmaf(hs[a], s, acc);
    }
  }
  for (int k = col_ptr[i]; k < col_ptr[i + 1]; ++k) {
    const int e = csc_eid[k];
    const float* hd = H + (size_t)edge_dst[e] * nf;
    const float* x = X + (size_t)e * ldx + (size_t)c * nf * nf;
    const float* dy = dY + (size_t)e * nf;
    for (int p = 0; p < nf; ++p) {
      float s = 0.f;
      for (int q = 0; q < nf; ++q) s = fmaf(x[p * nf + q], hd[q], s);
      acc = fmaf(dy[p], s, acc);
    }
  }
  dH[t] = acc;
}

}  // namespace

extern "C" {

// Y [E, nf] per-edge messages.  X [E(+1), ldx] bond rows (ldx >= nf^3), H [n_rows, nf]
int mpnn_bilinear_fwd(const int* edge_src, const int* edge_dst, const float* X, long long ldx, const float* H,
                      long long E, int nf, float* Y, cudaStream_t stream) {
  MPNN_REQUIRE(nf > 0 && E >= 0 && ldx >= (long long)nf * nf * nf, MPNN_ERR_ARG, "bilinear_fwd: bad dims");
  if (E == 0) return MPNN_OK;
  k_bil_fwd<<<ceil_div(E * nf, 256), 256, 0, stream>>>(edge_src, edge_dst, X, ldx, H, E, nf, Y);
  MPNN_CHECK_LAUNCH("k_bil_fwd");
  return MPNN_OK;
}

// dH [n_rows, nf] (written), dX [E, lddx] (written when non-NULL)
int mpnn_bilinear_bwd(const int* row_ptr, const int* col_ptr, const int* csc_eid, const int* edge_src,
                      const int* edge_dst, const float* X, long long ldx, const float* H, const float* dY, int n_rows,
                      long long E, int nf, float* dH, float* dX, long long lddx, cudaStream_t stream) {
  MPNN_REQUIRE(nf > 0 && E >= 0 && n_rows > 0, MPNN_ERR_ARG, "bilinear_bwd: bad dims");
  if (dH) {
    k_bil_bwd_h<<<ceil_div((long long)n_rows * nf, 256), 256, 0, stream>>>(row_ptr, col_ptr, csc_eid, edge_src, edge_dst, X,
                                                                          ldx, H, dY, n_rows, nf, dH);
    MPNN_CHECK_LAUNCH("k_bil_bwd_h");
  }
  if (dX && E > 0) {
    k_bil_bwd_x<<<ceil_div(E * nf * nf * nf, 256), 256, 0, stream>>>(edge_src, edge_dst, H, dY, E, nf, dX, lddx);
    MPNN_CHECK_LAUNCH("k_bil_bwd_x");
  }
  return MPNN_OK;
}

}  // extern "C"
