// Exact de-duplication of the compacted bond rows and a stable grouping of the edges by distinct row.
//
// Molecular bond features are categorical (reference mol_graph/mol_graph.py:74-90: bond type one-hot + a few
// flags), so the E compacted rows of a batch hold only a few dozen DISTINCT rows.  The edge network
// (edge_network.py:14-21) is a pure function of the row, hence evaluating it once per distinct row is exact:
// the trunk runs on U+1 rows instead of E (or the reference's B*N*N), and its last Linear turns into a table
// of U+1 matrices (csrc/typed.cu).  Rows are compared by their BIT patterns (no float equality subtleties).
//
// Everything here is deterministic and host-sync free:
//   rep[e]  = smallest edge index holding the same row          (open-addressing table, atomicMin on the slot)
//   uid[e]  = rank of rep[e] among the representatives          (distinct rows numbered by first occurrence)
//   urows   = the U distinct rows, then one all-zero row x_0    (row U; the caller pre-zeroes the buffer)
//   counts  = {E, U, overflow flag, 0}
// and a stable counting sort of the edges by uid (type_ptr / type_eid / type_pos), which gives every
// per-type reduction of the backward pass a fixed summation order.
#include "common.cuh"

namespace {

constexpr uint32_t SLOT_EMPTY = 0x7f7f7f7fu;  // memset(0x7f)

__device__ __forceinline__ uint32_t hash_row(const uint32_t* __restrict__ x, int ef) {
  uint32_t h = 0x9e3779b9u;
  for (int f = 0; f < ef; ++f) {
    uint32_t k = x[f] * 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    h ^= k;
    h = (h << 13) | (h >> 19);
    h = h * 5u + 0xe6546b64u;
  }
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

__device__ __forceinline__ bool same_row(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int ef) {
  bool eq = true;
  for (int f = 0; f < ef; ++f) eq &= (a[f] == b[f]);
  return eq;
}

// pass 1: claim / join a slot; the slot ends up holding the smallest edge index of its row value.
// Bond rows are categorical: a batch of 10^5 edges holds a few dozen distinct rows, so a naive insert is 10^5 atomics
// on a few dozen addresses.  Two filters keep the atomics to a handful per warp: (1) lanes of a warp that carry the
// same row elect their lowest lane (= smallest edge index) to insert for all of them; (2) a slot is read before it is
// touched -- once it holds a smaller index of the same row value there is nothing to do (slot values only decrease,
// and always within one row value, so a stale read is still a valid witness).
__global__ void k_dedup_insert(const uint32_t* __restrict__ rows, const int* __restrict__ n_edges_ptr, int cap,
                               int ef, uint32_t* __restrict__ table, uint32_t mask, int* __restrict__ slot_of) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int E = min(*n_edges_ptr, cap);
  const bool valid = e < E;
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const int lane = threadIdx.x & 31;
  const uint32_t* x = rows + (size_t)e * ef;
  const uint32_t h = hash_row(x, ef);
  const unsigned grp = __match_any_sync(act, h);
  const int leader = __ffs(grp) - 1;
  // equal hash is not equal row: lanes whose row differs from their leader's insert on their own
  const int e_lead = __shfl_sync(grp, e, leader);
  const bool follows = (lane != leader) && same_row(rows + (size_t)e_lead * ef, x, ef);
  uint32_t s = h & mask;
  if (!follows) {
    while (true) {
      uint32_t cur = *((volatile uint32_t*)&table[s]);
      if (cur == SLOT_EMPTY) {
        cur = atomicCAS(&table[s], SLOT_EMPTY, (uint32_t)e);
        if (cur == SLOT_EMPTY) break;
      }
      if (same_row(rows + (size_t)cur * ef, x, ef)) {
        if (cur > (uint32_t)e) atomicMin(&table[s], (uint32_t)e);
        break;
      }
      s = (s + 1) & mask;
    }
  }
  const uint32_t s_lead = __shfl_sync(grp, s, leader);
  slot_of[e] = (int)(follows ? s_lead : s);
}

// pass 2: rep[e] and the representative flags (0 beyond E so that the scan can run over the capacity)
__global__ void k_dedup_flag(const uint32_t* __restrict__ table, const int* __restrict__ slot_of,
                             const int* __restrict__ n_edges_ptr, int cap, int* __restrict__ rep,
                             int* __restrict__ flag) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cap) return;
  int E = min(*n_edges_ptr, cap);
  int r = e, f = 0;
  if (e < E) {
    r = (int)table[slot_of[e]];
    f = (r == e);
  }
  rep[e] = r;
  flag[e] = f;
}

// exclusive scan of one int array (three phases, phase 2 is one block)
constexpr int ST = 256, SPT = 8, SBLK = ST * SPT;

__global__ void k_scan1_blocksum(const int* __restrict__ a, int n, int* __restrict__ sa) {
  __shared__ int red[ST / 32];
  int base = blockIdx.x * SBLK;
  int v = 0;
  for (int i = threadIdx.x; i < SBLK; i += ST) {
    int idx = base + i;
    if (idx < n) v += a[idx];
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < ST / 32; ++w) t += red[w];
    sa[blockIdx.x] = t;
  }
}

__global__ void k_scan1_top(int* __restrict__ sa, int nblk) {
  __shared__ int carry;
  __shared__ int buf[ST];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += ST) {
    int idx = base + threadIdx.x;
    int v = idx < nblk ? sa[idx] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < ST; o <<= 1) {
      int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (idx < nblk) sa[idx] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += buf[ST - 1];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) k_scan1_small(const int* __restrict__ a, int n, int* __restrict__ out) {
  small_scan_block(a, nullptr, n, out, nullptr);
}

__global__ void k_scan1_final(const int* __restrict__ a, int n, const int* __restrict__ sa, int* __restrict__ out) {
  __shared__ int wsum[ST / 32];
  int base = blockIdx.x * SBLK + threadIdx.x * SPT;
  int l[SPT];
  int t = 0;
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    int idx = base + i;
    l[i] = idx < n ? a[idx] : 0;
    t += l[i];
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = t;
  for (int o = 1; o < 32; o <<= 1) {
    int x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int off = sa[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += wsum[w];
  off += inc - t;
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    int idx = base + i;
    if (idx < n) out[idx] = off;
    off += l[i];
    if (idx == n - 1) out[n] = off;
  }
}

// pass 3: uid, distinct rows, counts
__global__ void k_dedup_assign(const uint32_t* __restrict__ rows, const int* __restrict__ n_edges_ptr, int cap, int ef,
                               int ucap, const int* __restrict__ rep, const int* __restrict__ pos /*[cap+1]*/,
                               int* __restrict__ uid, uint32_t* __restrict__ urows, int* __restrict__ counts) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  int Eraw = *n_edges_ptr;
  int E = min(Eraw, cap);
  int U = pos[cap];
  if (e == 0) {
    counts[0] = Eraw;
    counts[1] = U;
    if (Eraw > cap || U > ucap) counts[2] = 1;   // sticky: cleared only by the host (GraphedStep.check)
    counts[3] = 0;
  }
  if (e >= E) return;
  int r = rep[e];
  int u = pos[r];
  uid[e] = u;
  if (r == e && u < ucap) {
    const uint32_t* x = rows + (size_t)e * ef;
    uint32_t* o = urows + (size_t)u * ef;
    for (int f = 0; f < ef; ++f) o[f] = x[f];
  }
}

// ---- stable counting sort of the edges by uid ----------------------------------------------------------
constexpr int TB = 1024;  // edges per block

__global__ void __launch_bounds__(TB) k_type_hist(const int* __restrict__ uid, const int* __restrict__ counts, int cap,
                                                  int ucap, int* __restrict__ blk_hist /*[nblk][ucap]*/) {
  extern __shared__ int hist[];
  int E = min(counts[0], cap);
  for (int u = threadIdx.x; u < ucap; u += TB) hist[u] = 0;
  __syncthreads();
  int e = blockIdx.x * TB + threadIdx.x;
  if (e < E) {
    int u = uid[e];
    if (u < ucap) atomicAdd(&hist[u], 1);
  }
  __syncthreads();
  for (int u = threadIdx.x; u < ucap; u += TB) blk_hist[(size_t)blockIdx.x * ucap + u] = hist[u];
}

// one block: per-type exclusive scan over the blocks (in place), then the scan over the types
__global__ void __launch_bounds__(TB) k_type_scan(int* __restrict__ blk_hist, int nblk, int ucap,
                                                  int* __restrict__ type_ptr /*[ucap+1]*/) {
  __shared__ int buf[TB];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int u0 = 0; u0 < ucap; u0 += TB) {
    int u = u0 + threadIdx.x;
    int tot = 0;
    if (u < ucap) {
      for (int b = 0; b < nblk; ++b) {
        int v = blk_hist[(size_t)b * ucap + u];
        blk_hist[(size_t)b * ucap + u] = tot;
        tot += v;
      }
    }
    buf[threadIdx.x] = tot;
    __syncthreads();
    for (int o = 1; o < TB; o <<= 1) {
      int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (u < ucap) type_ptr[u] = carry + buf[threadIdx.x] - tot;
    __syncthreads();
    if (threadIdx.x == 0) carry += buf[TB - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) type_ptr[ucap] = carry;
}

__global__ void __launch_bounds__(TB) k_type_fill(const int* __restrict__ uid, const int* __restrict__ counts, int cap,
                                                  int ucap, const int* __restrict__ blk_base,
                                                  const int* __restrict__ type_ptr, int* __restrict__ type_eid,
                                                  int* __restrict__ type_pos) {
  extern __shared__ int cnt[];
  int E = min(counts[0], cap);
  for (int u = threadIdx.x; u < ucap; u += TB) cnt[u] = 0;
  __syncthreads();
  int e = blockIdx.x * TB + threadIdx.x;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool live = e < E;
  int u = live ? uid[e] : -1;
  if (u >= ucap) {
    live = false;
    u = -1;
  }
  // rank among the earlier lanes of the warp holding the same type
  uint32_t peers = __match_any_sync(0xffffffffu, u);
  int rank_w = __popc(peers & ((1u << lane) - 1u));
  bool leader = (peers >> lane) == 1u;  // highest lane of its group
  int rank = 0;
  for (int w = 0; w < TB / 32; ++w) {
    if (warp == w && live) {
      rank = cnt[u] + rank_w;
    }
    __syncthreads();
    if (warp == w && live && leader) cnt[u] += __popc(peers);
    __syncthreads();
  }
  if (live) {
    int p = type_ptr[u] + blk_base[(size_t)blockIdx.x * ucap + u] + rank;
    type_eid[p] = e;
    type_pos[e] = p;
  }
}

uint32_t table_slots(int cap) {
  uint32_t s = 64;
  while (s < 2u * (uint32_t)(cap > 0 ? cap : 1)) s <<= 1;
  return s;
}

}  // namespace

extern "C" {

size_t mpnn_dedup_workspace_bytes(int edge_capacity, int unique_capacity) {
  size_t cap = (size_t)(edge_capacity > 0 ? edge_capacity : 1);
  size_t nblk = (cap + SBLK - 1) / SBLK;
  size_t tblk = (cap + TB - 1) / TB;
  size_t ucap = (size_t)(unique_capacity > 0 ? unique_capacity : 1);
  // table | slot_of | rep | flag | pos[cap+1] | scan sums | blk_hist
  return align_up((size_t)table_slots(edge_capacity) * 4, 256) + 3 * align_up(cap * 4, 256) +
         align_up((cap + 1) * 4, 256) + align_up(nblk * 4 + 4, 256) + align_up(tblk * ucap * 4, 256);
  // (the last term is only touched when sort != 0; callers that de-duplicate with a large unique_capacity and
  //  sort == 0 may pass mpnn_dedup_workspace_bytes(edge_capacity, 1))
}

size_t mpnn_type_sort_workspace_bytes(int edge_capacity, int unique_capacity) {
  size_t cap = (size_t)(edge_capacity > 0 ? edge_capacity : 1);
  size_t tblk = (cap + TB - 1) / TB;
  return tblk * (size_t)(unique_capacity > 0 ? unique_capacity : 1) * 4;
}

// Stable counting sort of the edges by uid: type_ptr [unique_capacity+1], type_eid / type_pos [edge_capacity]
// (edges grouped by uid, increasing edge index inside a group; type_pos is the inverse permutation).
int mpnn_type_sort(const int* uid, const int* counts, int edge_capacity, int unique_capacity, int* type_ptr,
                   int* type_eid, int* type_pos, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(edge_capacity >= 0 && unique_capacity > 0, MPNN_ERR_ARG, "type_sort: bad dims");
  MPNN_REQUIRE(unique_capacity <= 8192, MPNN_ERR_UNSUPPORTED, "type_sort: more than 8192 types");
  MPNN_REQUIRE(workspace_bytes >= mpnn_type_sort_workspace_bytes(edge_capacity, unique_capacity), MPNN_ERR_WORKSPACE,
               "type_sort: workspace too small");
  const int cap = edge_capacity > 0 ? edge_capacity : 1;
  const int tblk = ceil_div(cap, TB);
  int* blk_hist = (int*)workspace;
  size_t sm = (size_t)unique_capacity * sizeof(int);
  k_type_hist<<<tblk, TB, sm, stream>>>(uid, counts, edge_capacity, unique_capacity, blk_hist);
  MPNN_CHECK_LAUNCH("k_type_hist");
  k_type_scan<<<1, TB, 0, stream>>>(blk_hist, tblk, unique_capacity, type_ptr);
  MPNN_CHECK_LAUNCH("k_type_scan");
  k_type_fill<<<tblk, TB, sm, stream>>>(uid, counts, edge_capacity, unique_capacity, blk_hist, type_ptr, type_eid,
                                        type_pos);
  MPNN_CHECK_LAUNCH("k_type_fill");
  return MPNN_OK;
}

// rows [edge_capacity(+1), ef] compacted bond rows; n_edges_ptr: DEVICE pointer to the edge count (row_ptr + B*N).
// Outputs: uid [edge_capacity]; urows [unique_capacity+1, ef] PRE-ZEROED by the caller (row U stays the zero row);
// counts [4] = {E, U, overflow, 0}.  If sort != 0 also type_ptr [unique_capacity+1], type_eid / type_pos
// [edge_capacity]: edges grouped by uid, increasing edge index inside a group.
int mpnn_dedup_rows(const float* rows, const int* n_edges_ptr, int edge_capacity, int ef, int unique_capacity,
                    int* uid, float* urows, int* counts, int sort, int* type_ptr, int* type_eid, int* type_pos,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(edge_capacity >= 0 && ef > 0 && unique_capacity > 0, MPNN_ERR_ARG, "dedup_rows: bad dims");
  MPNN_REQUIRE(workspace_bytes >= mpnn_dedup_workspace_bytes(edge_capacity, sort ? unique_capacity : 1),
               MPNN_ERR_WORKSPACE, "dedup_rows: workspace too small");
  const int cap = edge_capacity > 0 ? edge_capacity : 1;
  const uint32_t slots = table_slots(edge_capacity);
  const int nblk = ceil_div(cap, SBLK);
  const int tblk = ceil_div(cap, TB);
  char* p = (char*)workspace;
  uint32_t* table = (uint32_t*)p;
  p += align_up((size_t)slots * 4, 256);
  int* slot_of = (int*)p;
  p += align_up((size_t)cap * 4, 256);
  int* rep = (int*)p;
  p += align_up((size_t)cap * 4, 256);
  int* flag = (int*)p;
  p += align_up((size_t)cap * 4, 256);
  int* pos = (int*)p;
  p += align_up((size_t)(cap + 1) * 4, 256);
  int* sums = (int*)p;
  p += align_up((size_t)nblk * 4 + 4, 256);
  int* blk_hist = (int*)p;

  MPNN_CUDA(cudaMemsetAsync(table, 0x7f, (size_t)slots * 4, stream));
  const uint32_t* urows_in = (const uint32_t*)rows;
  k_dedup_insert<<<ceil_div(cap, 256), 256, 0, stream>>>(urows_in, n_edges_ptr, edge_capacity, ef, table, slots - 1,
                                                         slot_of);
  MPNN_CHECK_LAUNCH("k_dedup_insert");
  k_dedup_flag<<<ceil_div(cap, 256), 256, 0, stream>>>(table, slot_of, n_edges_ptr, edge_capacity, rep, flag);
  MPNN_CHECK_LAUNCH("k_dedup_flag");
  if (cap <= SMALL_SCAN_MAX) {
    k_scan1_small<<<1, 1024, 0, stream>>>(flag, cap, pos);
  } else {
    k_scan1_blocksum<<<nblk, ST, 0, stream>>>(flag, cap, sums);
    k_scan1_top<<<1, ST, 0, stream>>>(sums, nblk);
    k_scan1_final<<<nblk, ST, 0, stream>>>(flag, cap, sums, pos);
  }
  MPNN_CHECK_LAUNCH("k_scan1");
  k_dedup_assign<<<ceil_div(cap, 256), 256, 0, stream>>>(urows_in, n_edges_ptr, edge_capacity, ef, unique_capacity, rep,
                                                         pos, uid, (uint32_t*)urows, counts);
  MPNN_CHECK_LAUNCH("k_dedup_assign");
  if (sort) {
    int rc = mpnn_type_sort(uid, counts, edge_capacity, unique_capacity, type_ptr, type_eid, type_pos, blk_hist,
                            (size_t)tblk * unique_capacity * 4, stream);
    if (rc) return rc;
  }
  return MPNN_OK;
}

}  // extern "C"
