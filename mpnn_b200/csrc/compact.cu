// On-device compaction of the reference's padded batch (afm/bfm/adj/mask as produced by
// collate_2d_graphs, reference pre_process/data_loader.py:50-70) to a per-batch edge list.
//
// Edge set (bit-exact contract): every (b,i,j) with adj[b,i,j] != 0 or any
// bfm[b,i,j,:] != 0, in row-major order (what torch.nonzero returns).  Outputs:
//   CSR by receiver: row_ptr[B*N+1], edge_src[e] = b*N+j, edge_dst[e] = b*N+i, edge_w[e] = adj value,
//                    edge_x[e,:] = bfm row
// adj may be NULL (edge set from the bond rows only, edge_w = 0).
//   CSC by sender  : col_ptr[B*N+1], csc_eid[k] (edge ids of column (b,j) in increasing i)
// Pass 1 reads the dense tensors once (HBM-bound: 4*B*N*N*(ef+1) bytes) and leaves a bitmask per
// row, so passes 2/3 touch only the rows that survive.
#include "common.cuh"

namespace {

// one warp per padded row r = b*N+i
__global__ void k_count(const float* __restrict__ bfm, const float* __restrict__ adj, int rows, int N, int ef,
                        int words, uint32_t* __restrict__ bitmask, int* __restrict__ row_cnt,
                        int* __restrict__ col_cnt) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* brow = bfm + (size_t)warp * N * ef;
  const float* arow = adj ? adj + (size_t)warp * N : nullptr;
  int gbase = (warp / N) * N;
  int cnt = 0;
  for (int w = 0; w < words; ++w) {
    int j = w * 32 + lane;
    bool keep = false;
    if (j < N) {
      keep = arow ? (arow[j] != 0.0f) : false;
      const float* x = brow + (size_t)j * ef;
      for (int f = 0; f < ef; ++f) keep |= (x[f] != 0.0f);
    }
    uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) bitmask[(size_t)warp * words + w] = m;
    if (keep) atomicAdd(&col_cnt[gbase + j], 1);  // integer add: result independent of order
    cnt += __popc(m);
  }
  if (lane == 0) row_cnt[warp] = cnt;
}

// exclusive scan of up to two int arrays of length n (n+1 outputs each).  Three phases; phase 2 is
// a single block over the per-block sums.
constexpr int SCAN_T = 256;
constexpr int SCAN_PER_T = 8;
constexpr int SCAN_BLK = SCAN_T * SCAN_PER_T;

__global__ void __launch_bounds__(1024) k_scan_small(const int* __restrict__ a, const int* __restrict__ b, int n,
                                                     int* __restrict__ outa, int* __restrict__ outb) {
  small_scan_block(a, b, n, outa, outb);
}

__global__ void k_scan_blocksum(const int* __restrict__ a, const int* __restrict__ b, int n, int* __restrict__ sa,
                                int* __restrict__ sb) {
  __shared__ int red[2][SCAN_T / 32];
  int base = blockIdx.x * SCAN_BLK;
  int va = 0, vb = 0;
  for (int i = threadIdx.x; i < SCAN_BLK; i += SCAN_T) {
    int idx = base + i;
    if (idx < n) {
      va += a[idx];
      vb += b[idx];
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    va += __shfl_xor_sync(0xffffffffu, va, o);
    vb += __shfl_xor_sync(0xffffffffu, vb, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = va;
    red[1][threadIdx.x >> 5] = vb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ta = 0, tb = 0;
    for (int w = 0; w < SCAN_T / 32; ++w) {
      ta += red[0][w];
      tb += red[1][w];
    }
    sa[blockIdx.x] = ta;
    sb[blockIdx.x] = tb;
  }
}

__global__ void k_scan_top(int* __restrict__ sa, int* __restrict__ sb, int nblk) {
  // single thread block; sequential over chunks of blockDim (nblk is small: rows / 2048)
  __shared__ int carry[2];
  __shared__ int buf[2][SCAN_T];
  if (threadIdx.x == 0) carry[0] = carry[1] = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += SCAN_T) {
    int idx = base + threadIdx.x;
    int va = idx < nblk ? sa[idx] : 0;
    int vb = idx < nblk ? sb[idx] : 0;
    buf[0][threadIdx.x] = va;
    buf[1][threadIdx.x] = vb;
    __syncthreads();
    for (int o = 1; o < SCAN_T; o <<= 1) {  // Hillis-Steele inclusive
      int ta = threadIdx.x >= o ? buf[0][threadIdx.x - o] : 0;
      int tb = threadIdx.x >= o ? buf[1][threadIdx.x - o] : 0;
      __syncthreads();
      buf[0][threadIdx.x] += ta;
      buf[1][threadIdx.x] += tb;
      __syncthreads();
    }
    if (idx < nblk) {
      sa[idx] = carry[0] + buf[0][threadIdx.x] - va;  // exclusive
      sb[idx] = carry[1] + buf[1][threadIdx.x] - vb;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry[0] += buf[0][SCAN_T - 1];
      carry[1] += buf[1][SCAN_T - 1];
    }
    __syncthreads();
  }
}

__global__ void k_scan_final(const int* __restrict__ a, const int* __restrict__ b, int n, const int* __restrict__ sa,
                             const int* __restrict__ sb, int* __restrict__ outa, int* __restrict__ outb) {
  // each thread owns SCAN_PER_T consecutive elements
  __shared__ int wsum[2][SCAN_T / 32];
  int base = blockIdx.x * SCAN_BLK + threadIdx.x * SCAN_PER_T;
  int la[SCAN_PER_T], lb[SCAN_PER_T];
  int ta = 0, tb = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER_T; ++i) {
    int idx = base + i;
    la[i] = idx < n ? a[idx] : 0;
    lb[i] = idx < n ? b[idx] : 0;
    ta += la[i];
    tb += lb[i];
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int ia = ta, ib = tb;
  for (int o = 1; o < 32; o <<= 1) {
    int xa = __shfl_up_sync(0xffffffffu, ia, o);
    int xb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) {
      ia += xa;
      ib += xb;
    }
  }
  if (lane == 31) {
    wsum[0][warp] = ia;
    wsum[1][warp] = ib;
  }
  __syncthreads();
  int oa = sa[blockIdx.x], ob = sb[blockIdx.x];
  for (int w = 0; w < warp; ++w) {
    oa += wsum[0][w];
    ob += wsum[1][w];
  }
  oa += ia - ta;
  ob += ib - tb;
#pragma unroll
  for (int i = 0; i < SCAN_PER_T; ++i) {
    int idx = base + i;
    if (idx < n) {
      outa[idx] = oa;
      outb[idx] = ob;
    }
    oa += la[i];
    ob += lb[i];
    if (idx == n - 1) {  // totals
      outa[n] = oa;
      outb[n] = ob;
    }
  }
}

// one warp per padded row: writes the row's edges at row_ptr[r]..
__global__ void k_fill(const float* __restrict__ bfm, const float* __restrict__ adj, const uint32_t* __restrict__ bitmask,
                       const int* __restrict__ row_ptr, int rows, int N, int ef, int words, int capacity,
                       int* __restrict__ edge_src, int* __restrict__ edge_dst, float* __restrict__ edge_w,
                       float* __restrict__ edge_x) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  int pos = row_ptr[warp];
  int gbase = (warp / N) * N;
  const float* brow = bfm + (size_t)warp * N * ef;
  const float* arow = adj ? adj + (size_t)warp * N : nullptr;
  for (int w = 0; w < words; ++w) {
    uint32_t m = bitmask[(size_t)warp * words + w];
    if (m == 0) continue;
    int j = w * 32 + lane;
    if ((m >> lane) & 1u) {
      int e = pos + __popc(m & ((1u << lane) - 1u));
      if (e < capacity) {
        edge_src[e] = gbase + j;
        edge_dst[e] = warp;
        edge_w[e] = arow ? arow[j] : 0.0f;
        const float* x = brow + (size_t)j * ef;
        float* o = edge_x + (size_t)e * ef;
        for (int f = 0; f < ef; ++f) o[f] = x[f];
      }
    }
    pos += __popc(m);
  }
}

// one warp per column (b,j): lists the edges whose sender is (b,j), in increasing receiver order
__global__ void k_fill_csc(const uint32_t* __restrict__ bitmask, const int* __restrict__ row_ptr,
                           const int* __restrict__ col_ptr, int rows, int N, int words, int capacity,
                           int* __restrict__ csc_eid) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  int b = warp / N, j = warp - b * N;
  int wj = j >> 5;
  uint32_t bit = 1u << (j & 31);
  int pos = col_ptr[warp];
  for (int i0 = 0; i0 < N; i0 += 32) {
    int i = i0 + lane;
    bool has = false;
    int e = 0;
    if (i < N) {
      int r = b * N + i;
      const uint32_t* bm = bitmask + (size_t)r * words;
      uint32_t m = bm[wj];
      has = (m & bit) != 0;
      if (has) {
        e = row_ptr[r] + __popc(m & (bit - 1u));
        for (int w = 0; w < wj; ++w) e += __popc(bm[w]);
      }
    }
    uint32_t hm = __ballot_sync(0xffffffffu, has);
    if (has) {
      int k = pos + __popc(hm & ((1u << lane) - 1u));
      if (k < capacity) csc_eid[k] = e;
    }
    pos += __popc(hm);
  }
}

// d_bfm[b,i,j,:] = d_edge_x[e,:] for every compacted edge (the dense tensor is zero-initialised by the caller)
__global__ void k_scatter_rows(const float* __restrict__ dx, const int* __restrict__ edge_dst,
                               const int* __restrict__ edge_src, int E, int N, int ef, float* __restrict__ dense) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)E * ef) return;
  int e = (int)(t / ef), f = (int)(t - (long long)e * ef);
  int j = edge_src[e] % N;
  dense[((size_t)edge_dst[e] * N + j) * ef + f] = dx[t];
}

// ---------------------------------------------------------------------------------------------------
// Device-side collate: the host ships the batch RAGGED -- the real atoms' feature rows and the
// edge list -- and this kernel writes the reference's padded layout (collate_2d_graphs, data_loader.py:50-70:
// afm [B,N,Fa], bfm [B,N,N,ef], adj [B,N,N], mask [B,N,1]) into pre-zeroed buffers.  At config-5 size the padded bfm
// is 757 MB of mostly zeros; the ragged form is ~130 MB.
//   atom_row[a] = b*N + i (flat padded row of real atom a); edge_dst[e] = b*N + i, edge_j[e] = j.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_collate_atoms(const int* __restrict__ atom_row, const float* __restrict__ afm_cat,
                                                       long long n, int Fa, float* __restrict__ afm,
                                                       float* __restrict__ mask) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * Fa) return;
  const long long a = t / Fa;
  const int f = (int)(t - a * Fa);
  const int row = __ldg(atom_row + a);
  afm[(size_t)row * Fa + f] = afm_cat[t];
  if (f == 0) mask[row] = 1.0f;
}

__global__ void __launch_bounds__(256) k_collate_edges(const int* __restrict__ edge_dst, const int* __restrict__ edge_j,
                                                       const float* __restrict__ edge_w, const float* __restrict__ edge_x,
                                                       long long E, int ef, int N, float* __restrict__ bfm,
                                                       float* __restrict__ adj) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * (ef + 1)) return;
  const long long e = t / (ef + 1);
  const int f = (int)(t - e * (ef + 1));
  const size_t pair = (size_t)__ldg(edge_dst + e) * N + __ldg(edge_j + e);
  if (f < ef)
    bfm[pair * ef + f] = edge_x[e * ef + f];
  else
    adj[pair] = edge_w[e];
}

// capacity mode: consumers walk row_ptr / col_ptr ranges, so the pointers themselves are clamped to the allocated edge
// slots (an overflowing batch then computes on a truncated edge list instead of reading out of bounds); the true edge
// count is kept for the overflow flag
__global__ void k_clamp_ptrs(int* __restrict__ row_ptr, int* __restrict__ col_ptr, int n, int cap, int* __restrict__ e_true) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) e_true[0] = row_ptr[n];
  if (i <= n) {
    const int a = row_ptr[i], b = col_ptr[i];
    if (a > cap) row_ptr[i] = cap;
    if (b > cap) col_ptr[i] = cap;
  }
}

}  // namespace

extern "C" {

int mpnn_scatter_edge_rows(const float* d_edge_x, const int* edge_dst, const int* edge_src, int E, int N, int ef,
                           float* dense, cudaStream_t stream) {
  if (E <= 0) return MPNN_OK;
  k_scatter_rows<<<ceil_div((long long)E * ef, 256), 256, 0, stream>>>(d_edge_x, edge_dst, edge_src, E, N, ef, dense);
  MPNN_CHECK_LAUNCH("k_scatter_rows");
  return MPNN_OK;
}


size_t mpnn_compact_workspace_bytes(int B, int N) {
  size_t rows = (size_t)B * N;
  size_t words = (size_t)(N + 31) / 32;
  size_t nblk = (rows + SCAN_BLK - 1) / SCAN_BLK;
  // bitmask | row_cnt | col_cnt | block sums a | block sums b
  return align_up(rows * words * 4, 256) + 2 * align_up(rows * 4, 256) + 2 * align_up(nblk * 4 + 4, 256);
}

// Phase 1: predicate + counts + both exclusive scans.  row_ptr[B*N] (device) holds the edge count E.
int mpnn_compact_count(const float* bfm, const float* adj, int B, int N, int ef, int* row_ptr, int* col_ptr,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && ef > 0, MPNN_ERR_ARG, "compact_count: bad dims B=%d N=%d ef=%d", B, N, ef);
  MPNN_REQUIRE((long long)B * N < (1ll << 31), MPNN_ERR_ARG, "compact_count: B*N overflows int32");
  MPNN_REQUIRE(workspace_bytes >= mpnn_compact_workspace_bytes(B, N), MPNN_ERR_WORKSPACE,
               "compact_count: workspace too small");
  int rows = B * N;
  int words = (N + 31) / 32;
  int nblk = ceil_div(rows, SCAN_BLK);
  char* p = (char*)workspace;
  uint32_t* bitmask = (uint32_t*)p;
  p += align_up((size_t)rows * words * 4, 256);
  int* row_cnt = (int*)p;
  p += align_up((size_t)rows * 4, 256);
  int* col_cnt = (int*)p;
  p += align_up((size_t)rows * 4, 256);
  int* sa = (int*)p;
  p += align_up((size_t)nblk * 4 + 4, 256);
  int* sb = (int*)p;
  MPNN_CUDA(cudaMemsetAsync(col_cnt, 0, (size_t)rows * 4, stream));
  k_count<<<ceil_div((long long)rows * 32, 256), 256, 0, stream>>>(bfm, adj, rows, N, ef, words, bitmask, row_cnt,
                                                                   col_cnt);
  MPNN_CHECK_LAUNCH("k_count");
  if (rows <= SMALL_SCAN_MAX) {
    k_scan_small<<<1, 1024, 0, stream>>>(row_cnt, col_cnt, rows, row_ptr, col_ptr);
  } else {
    k_scan_blocksum<<<nblk, SCAN_T, 0, stream>>>(row_cnt, col_cnt, rows, sa, sb);
    k_scan_top<<<1, SCAN_T, 0, stream>>>(sa, sb, nblk);
    k_scan_final<<<nblk, SCAN_T, 0, stream>>>(row_cnt, col_cnt, rows, sa, sb, row_ptr, col_ptr);
  }
  MPNN_CHECK_LAUNCH("k_scan");
  return MPNN_OK;
}

// Phase 2: fill the CSR/CSC arrays (capacity = allocated edge slots; edges beyond it are dropped, the
// caller compares row_ptr[B*N] with its capacity).  `workspace` is the one phase 1 filled.
int mpnn_compact_fill(const float* bfm, const float* adj, int B, int N, int ef, const int* row_ptr, const int* col_ptr,
                      int capacity, int* edge_src, int* edge_dst, float* edge_w, float* edge_x, int* csc_eid,
                      const void* workspace, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && ef > 0 && capacity >= 0, MPNN_ERR_ARG, "compact_fill: bad dims");
  int rows = B * N;
  int words = (N + 31) / 32;
  const uint32_t* bitmask = (const uint32_t*)workspace;
  if (capacity == 0) return MPNN_OK;
  k_fill<<<ceil_div((long long)rows * 32, 256), 256, 0, stream>>>(bfm, adj, bitmask, row_ptr, rows, N, ef, words,
                                                                  capacity, edge_src, edge_dst, edge_w, edge_x);
  MPNN_CHECK_LAUNCH("k_fill");
  if (csc_eid) {
    k_fill_csc<<<ceil_div((long long)rows * 32, 256), 256, 0, stream>>>(bitmask, row_ptr, col_ptr, rows, N, words,
                                                                        capacity, csc_eid);
    MPNN_CHECK_LAUNCH("k_fill_csc");
  }
  return MPNN_OK;
}

// Ragged batch -> the reference's padded layout (collate_2d_graphs, data_loader.py:50-70).  All four outputs are
// zero-filled here and then scattered into; n real atoms, E edges (pairs with adj != 0 or a non-zero bond row).
int mpnn_collate_ragged(const int* atom_row, const float* afm_cat, long long n, int Fa, const int* edge_dst,
                        const int* edge_j, const float* edge_w, const float* edge_x, long long E, int ef, int B, int N,
                        float* afm, float* bfm, float* adj, float* mask, cudaStream_t stream) {
  MPNN_REQUIRE(B > 0 && N > 0 && Fa > 0 && ef > 0 && n >= 0 && E >= 0, MPNN_ERR_ARG, "collate_ragged: bad dims");
  MPNN_REQUIRE(afm && bfm && adj && mask, MPNN_ERR_ARG, "collate_ragged: null output");
  const size_t rows = (size_t)B * N;
  MPNN_CUDA(cudaMemsetAsync(afm, 0, rows * Fa * sizeof(float), stream));
  MPNN_CUDA(cudaMemsetAsync(mask, 0, rows * sizeof(float), stream));
  MPNN_CUDA(cudaMemsetAsync(adj, 0, rows * N * sizeof(float), stream));
  MPNN_CUDA(cudaMemsetAsync(bfm, 0, rows * N * ef * sizeof(float), stream));
  if (n > 0) {
    k_collate_atoms<<<ceil_div(n * Fa, 256), 256, 0, stream>>>(atom_row, afm_cat, n, Fa, afm, mask);
    MPNN_CHECK_LAUNCH("k_collate_atoms");
  }
  if (E > 0) {
    k_collate_edges<<<ceil_div(E * (ef + 1), 256), 256, 0, stream>>>(edge_dst, edge_j, edge_w, edge_x, E, ef, N, bfm, adj);
    MPNN_CHECK_LAUNCH("k_collate_edges");
  }
  return MPNN_OK;
}

// clamps row_ptr / col_ptr [n_rows + 1] to `capacity` in place and stores the true edge count in e_true[0]
int mpnn_compact_clamp(int* row_ptr, int* col_ptr, int n_rows, int capacity, int* e_true, cudaStream_t stream) {
  MPNN_REQUIRE(n_rows >= 0 && capacity >= 0, MPNN_ERR_ARG, "compact_clamp: bad dims");
  // (single pass: entry n is read by thread 0 of block 0 before any block can clamp it only if it is in block 0's range;
  //  read it first in its own launch to keep the order explicit)
  k_clamp_ptrs<<<1, 32, 0, stream>>>(row_ptr + n_rows, col_ptr + n_rows, 0, 0x7fffffff, e_true);
  k_clamp_ptrs<<<ceil_div(n_rows + 1, 256), 256, 0, stream>>>(row_ptr, col_ptr, n_rows, capacity, e_true + 1);
  MPNN_CHECK_LAUNCH("k_clamp_ptrs");
  return MPNN_OK;
}

}  // extern "C"
