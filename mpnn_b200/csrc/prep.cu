// Edge compaction + exact de-duplication of the bond rows of a padded batch as ONE cooperative launch (small batches).
//
// Same contract as csrc/compact.cu + csrc/dedup.cu (the edge set is every (b,i,j) with adj != 0 or any
// bfm != 0, in row-major = torch.nonzero order; distinct bond rows numbered by first occurrence; CSC lists by sender),
// which take eight dependent launches (count, scan, fill, CSC fill, hash insert, flag, scan, assign: 30 us of single-block
// scans and tiny grids in front of every forward pass at BASELINE config 2).  Here every CTA owns whole graphs, so the
// CSR and the CSC of its graphs are local to it and ONE grid barrier is enough:
//   phase A  per row: edge predicate -> row / column counts (kept in row_ptr / col_ptr), every edge's bond row inserted
//            into a global open-addressing table (claim by CAS, priority = first position by atomicMax), CTA edge total;
//   barrier  (co-resident grid: cooperative launch)
//   phase B  CTA offset = sum of the totals in front of it; local scans -> row_ptr / col_ptr; the table's occupied slots
//            ranked by first position -> type id of every slot (every CTA, redundantly: a few dozen entries); edges written
//            in order with their type ids; CSC lists from the graph's bit matrix in shared memory; CTA 0 writes the
//            distinct rows and the counts.
// Integer work only; results are bit-identical to the multi-launch path (tests/test_gpu_typed.py).
#include "common.cuh"

namespace {

constexpr int PT = 512;           // threads per CTA
constexpr int PW_ = PT / 32;      // warps
constexpr int MAXN = 512;         // atoms per graph the shared-memory bit matrix holds
constexpr int MAXH = 4096;        // hash slots

struct Prep {
  const float* bfm;
  const float* adj;
  int B, N, ef, ecap, ucap, H;
  int* row_ptr;
  int* col_ptr;
  int* edge_src;
  int* edge_dst;
  float* edge_w;
  int* csc_eid;
  int* uid;
  float* urows;       // [ucap + 1][ef]
  int* counts;        // {E, U, overflow (sticky), 0}
  unsigned* claim;    // [H] representative pair index + 1 (0 = empty)
  unsigned* prio;     // [H] 0x7fffffff - first pair index (0 = empty)
  int* part;          // [grid] edges per CTA
  unsigned* bar;      // [4] arrivals, exits
};

__device__ __forceinline__ uint32_t hash_bits(const uint32_t* __restrict__ x, int ef) {   // == dedup.cu hash_row
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

__device__ __forceinline__ bool same_bits(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int ef) {
  bool eq = true;
  for (int f = 0; f < ef; ++f) eq &= (a[f] == b[f]);
  return eq;
}

__device__ __forceinline__ bool edge_pred(const Prep& p, size_t pair) {
  bool keep = p.adj ? (p.adj[pair] != 0.0f) : false;
  const float* x = p.bfm + pair * p.ef;
  for (int f = 0; f < p.ef; ++f) keep |= (x[f] != 0.0f);
  return keep;
}

// slot of the bond row of `pair` (inserted by phase A; the probe sequence always ends at it)
__device__ __forceinline__ int find_slot(const Prep& p, size_t pair) {
  const uint32_t* x = reinterpret_cast<const uint32_t*>(p.bfm) + pair * p.ef;
  uint32_t s = hash_bits(x, p.ef) & (uint32_t)(p.H - 1);
  while (true) {
    const unsigned c = __ldcg(p.claim + s);
    if (c == 0u) return -1;   // cannot happen for an inserted row
    if (same_bits(reinterpret_cast<const uint32_t*>(p.bfm) + (size_t)(c - 1u) * p.ef, x, p.ef)) return (int)s;
    s = (s + 1) & (uint32_t)(p.H - 1);
  }
}

__device__ __forceinline__ unsigned ld_acq(const unsigned* q) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
  return v;
}

__global__ void __launch_bounds__(PT, 1) k_prep(Prep p) {
  __shared__ int s_red[PW_];
  __shared__ int s_carry[2];
  __shared__ uint32_t s_bits[MAXN * (MAXN / 32)];   // bit matrix of the current graph: [N][words]
  __shared__ unsigned s_lprio[1024 + 32];           // priorities of the occupied slots
  __shared__ short s_uid[MAXH];                     // type id of every slot
  __shared__ int s_U;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, ef = p.ef, words = (N + 31) >> 5;
  const int G = gridDim.x;
  const int gper = (p.B + G - 1) / G;
  const int g0 = min(p.B, (int)blockIdx.x * gper), g1 = min(p.B, g0 + gper);
  const int r0 = g0 * N, r1 = g1 * N;
  // ---------------- phase A ----------------
  for (int r = r0 + tid; r < r1; r += PT) p.col_ptr[r] = 0;
  __syncthreads();
  int my_edges = 0;
  for (int r = r0 + warp; r < r1; r += PW_) {
    const int gbase = (r / N) * N;
    int cnt = 0;
    for (int w = 0; w < words; ++w) {
      const int j = w * 32 + lane;
      const size_t pair = (size_t)r * N + j;
      const bool keep = j < N && edge_pred(p, pair);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      cnt += __popc(m);
      if (keep) {
        atomicAdd(p.col_ptr + gbase + j, 1);   // integer, CTA-local range
        // insert the bond row: claim an empty slot or join the slot that holds the same row; priority = first position
        const uint32_t* x = reinterpret_cast<const uint32_t*>(p.bfm) + pair * ef;
        uint32_t s = hash_bits(x, ef) & (uint32_t)(p.H - 1);
        while (true) {
          unsigned c = *((volatile unsigned*)(p.claim + s));
          if (c == 0u) {
            c = atomicCAS(p.claim + s, 0u, (unsigned)pair + 1u);
            if (c == 0u) break;
          }
          if (same_bits(reinterpret_cast<const uint32_t*>(p.bfm) + (size_t)(c - 1u) * ef, x, ef)) break;
          s = (s + 1) & (uint32_t)(p.H - 1);
        }
        const unsigned pr = 0x7fffffffu - (unsigned)pair;
        if (*((volatile unsigned*)(p.prio + s)) < pr) atomicMax(p.prio + s, pr);
      }
    }
    if (lane == 0) p.row_ptr[r] = cnt;
    my_edges += cnt;
  }
  if (lane == 0) s_red[warp] = my_edges;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < PW_; ++w) t += s_red[w];
    p.part[blockIdx.x] = t;
    __threadfence();
    atomicAdd(p.bar, 1u);
    while (ld_acq(p.bar) < (unsigned)G) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
  // ---------------- phase B ----------------
  // (1) offset of this CTA's edges, total
  int base = 0, total = 0;
  {
    int b_ = 0, t_ = 0;
    for (int c = tid; c < G; c += PT) {
      const int v = __ldcg(p.part + c);
      t_ += v;
      if (c < (int)blockIdx.x) b_ += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      b_ += __shfl_xor_sync(0xffffffffu, b_, o);
      t_ += __shfl_xor_sync(0xffffffffu, t_, o);
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = b_;
    __syncthreads();
    for (int w = 0; w < PW_; ++w) base += s_red[w];
    __syncthreads();
    if (lane == 0) s_red[warp] = t_;
    __syncthreads();
    for (int w = 0; w < PW_; ++w) total += s_red[w];
    __syncthreads();
  }
  // (2) exclusive scans of the row / column counts of this CTA's rows (chunks of PT with a carry)
  if (tid == 0) s_carry[0] = s_carry[1] = base;
  __syncthreads();
  for (int c0 = r0; c0 < r1; c0 += PT) {
    const int r = c0 + tid;
    const int va = r < r1 ? p.row_ptr[r] : 0, vb = r < r1 ? p.col_ptr[r] : 0;
    int ia = va, ib = vb;
    for (int o = 1; o < 32; o <<= 1) {
      const int xa = __shfl_up_sync(0xffffffffu, ia, o), xb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) {
        ia += xa;
        ib += xb;
      }
    }
    __shared__ int s_wa[PW_], s_wb[PW_];
    if (lane == 31) {
      s_wa[warp] = ia;
      s_wb[warp] = ib;
    }
    __syncthreads();
    int oa = s_carry[0], ob = s_carry[1];
    for (int w = 0; w < warp; ++w) {
      oa += s_wa[w];
      ob += s_wb[w];
    }
    if (r < r1) {   // clamped to the allocated slots: consumers walk these ranges
      p.row_ptr[r] = min(oa + ia - va, p.ecap);
      p.col_ptr[r] = min(ob + ib - vb, p.ecap);
    }
    __syncthreads();
    if (tid == PT - 1) {
      s_carry[0] = oa + ia;
      s_carry[1] = ob + ib;
    }
    __syncthreads();
  }
  if (blockIdx.x == G - 1 && tid == 0) {
    p.row_ptr[p.B * N] = min(total, p.ecap);
    p.col_ptr[p.B * N] = min(total, p.ecap);
  }
  // (3) type id of every slot: occupied slots ranked by first position (every CTA computes the same map)
  if (tid == 0) s_U = 0;
  __syncthreads();
  for (int s0 = 0; s0 < p.H; s0 += PT) {
    const int s = s0 + tid;
    const unsigned pr = s < p.H ? __ldcg(p.prio + s) : 0u;
    const bool occ = pr != 0u;
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    int wbase = 0;
    if (lane == 0 && m) wbase = atomicAdd(&s_U, __popc(m));   // order of the list does not matter (ranks are by value)
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (occ) {
      const int k = wbase + __popc(m & ((1u << lane) - 1u));
      if (k < 1024 + 32) s_lprio[k] = pr;
    }
  }
  __syncthreads();
  const int U = s_U;
  const int Ul = min(U, 1024 + 32);
  for (int s = tid; s < p.H; s += PT) {
    const unsigned pr = __ldcg(p.prio + s);
    int rank = 0;
    if (pr != 0u)
      for (int k = 0; k < Ul; ++k) rank += (s_lprio[k] > pr);
    s_uid[s] = (short)min(rank, p.ucap);   // beyond the capacity: the zero type (flagged as overflow)
  }
  __syncthreads();
  // (4) CTA 0: distinct rows in first-occurrence order, zero rows behind them, counts
  if (blockIdx.x == 0) {
    for (int s = tid; s < p.H; s += PT) {
      const unsigned pr = __ldcg(p.prio + s);
      if (pr != 0u && s_uid[s] < p.ucap) {
        const size_t pair = (size_t)(0x7fffffffu - pr);
        for (int f = 0; f < ef; ++f) p.urows[(size_t)s_uid[s] * ef + f] = p.bfm[pair * ef + f];
      }
    }
    for (int i = tid; i < (p.ucap + 1 - min(U, p.ucap)) * ef; i += PT) p.urows[(size_t)min(U, p.ucap) * ef + i] = 0.f;
    if (tid == 0) {
      p.counts[0] = total;
      p.counts[1] = U;
      if (total > p.ecap || U > p.ucap) p.counts[2] = 1;   // sticky: cleared only by the host (GraphedStep.check)
      p.counts[3] = 0;
    }
  }
  // (5) edges of this CTA's graphs in order, then their CSC lists
  for (int g = g0; g < g1; ++g) {
    const int gbase = g * N;
    __syncthreads();   // s_bits free
    for (int i = warp; i < N; i += PW_) {
      const int r = gbase + i;
      int pos = p.row_ptr[r];
      for (int w = 0; w < words; ++w) {
        const int j = w * 32 + lane;
        const size_t pair = (size_t)r * N + j;
        const bool keep = j < N && edge_pred(p, pair);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_bits[i * words + w] = m;
        if (keep) {
          const int e = pos + __popc(m & ((1u << lane) - 1u));
          if (e < p.ecap) {
            p.edge_src[e] = gbase + j;
            p.edge_dst[e] = r;
            p.edge_w[e] = p.adj ? p.adj[pair] : 0.0f;
            const int s = find_slot(p, pair);
            p.uid[e] = s >= 0 ? (int)s_uid[s] : p.ucap;
          }
        }
        pos += __popc(m);
      }
    }
    __syncthreads();
    // CSC of column j: the edges (i, j) in increasing i; edge id = row_ptr[i] + rank of j inside row i
    for (int j = warp; j < N; j += PW_) {
      int cpos = p.col_ptr[gbase + j];
      const int wj = j >> 5;
      const unsigned below = (1u << (j & 31)) - 1u;
      for (int i0 = 0; i0 < N; i0 += 32) {
        const int i = i0 + lane;
        bool has = false;
        int e = 0;
        if (i < N) {
          const uint32_t* row = s_bits + i * words;
          has = (row[wj] >> (j & 31)) & 1u;
          if (has) {
            e = p.row_ptr[gbase + i] + __popc(row[wj] & below);
            for (int w = 0; w < wj; ++w) e += __popc(row[w]);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (has) {
          const int k = cpos + __popc(m & ((1u << lane) - 1u));
          if (k < p.ecap) p.csc_eid[k] = e;
        }
        cpos += __popc(m);
      }
    }
  }
  // ---------------- exit: the last CTA leaves the table and the barrier words empty for the next launch --------------
  __syncthreads();
  __shared__ int s_lastcta;
  if (tid == 0) {
    __threadfence();
    s_lastcta = atomicAdd(p.bar + 1, 1u) == (unsigned)G - 1u;
  }
  __syncthreads();
  if (s_lastcta) {
    for (int s = tid; s < p.H; s += PT) {
      p.claim[s] = 0u;
      p.prio[s] = 0u;
    }
    if (tid < 4) p.bar[tid] = 0u;
  }
}

int prep_hash_slots(int ucap) {
  int h = 256;
  while (h < 4 * (ucap + 1)) h <<= 1;
  return h;
}

}  // namespace

extern "C" {

// 1 when the single-launch path serves the shape (whole graphs per CTA, bit matrix in shared memory, <= 8 graphs per CTA)
int mpnn_prep_supported(int B, int N, int ef, int unique_capacity) {
  return (B >= 1 && N >= 1 && N <= MAXN && ef >= 1 && ef <= 64 && unique_capacity >= 1 && unique_capacity <= 1024 &&
          prep_hash_slots(unique_capacity) <= MAXH && (long long)N * ((N + 31) / 32) <= (long long)MAXN * (MAXN / 32) &&
          B <= 8 * mpnn_num_sms())
             ? 1
             : 0;
}

// workspace: ZERO on first use, left zero by every launch (hash table, barrier words); the rest is scratch
size_t mpnn_prep_workspace_bytes(int B, int unique_capacity) {
  (void)B;
  return 256 + 2 * (size_t)prep_hash_slots(unique_capacity) * sizeof(unsigned) + (size_t)mpnn_num_sms() * sizeof(int) + 256;
}

// Capacity-mode compaction + de-duplication (see csrc/compact.cu, csrc/dedup.cu for the contract).  counts[2] is only
// ever SET here (sticky overflow flag).  uid is clamped to the zero type (= unique_capacity) on overflow.
int mpnn_prep_edges(const float* bfm, const float* adj, int B, int N, int ef, int edge_capacity, int unique_capacity,
                    int* row_ptr, int* col_ptr, int* edge_src, int* edge_dst, float* edge_w, int* csc_eid, int* uid,
                    float* urows, int* counts, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(mpnn_prep_supported(B, N, ef, unique_capacity), MPNN_ERR_UNSUPPORTED, "prep_edges: unsupported shape");
  MPNN_REQUIRE(workspace_bytes >= mpnn_prep_workspace_bytes(B, unique_capacity), MPNN_ERR_WORKSPACE, "prep_edges: workspace");
  MPNN_REQUIRE((long long)B * N * N < (1ll << 30), MPNN_ERR_UNSUPPORTED, "prep_edges: too many atom pairs");
  Prep p;
  p.bfm = bfm;
  p.adj = adj;
  p.B = B;
  p.N = N;
  p.ef = ef;
  p.ecap = edge_capacity;
  p.ucap = unique_capacity;
  p.H = prep_hash_slots(unique_capacity);
  p.row_ptr = row_ptr;
  p.col_ptr = col_ptr;
  p.edge_src = edge_src;
  p.edge_dst = edge_dst;
  p.edge_w = edge_w;
  p.csc_eid = csc_eid;
  p.uid = uid;
  p.urows = urows;
  p.counts = counts;
  char* w = (char*)workspace;
  p.bar = (unsigned*)w;
  p.claim = (unsigned*)(w + 256);
  p.prio = p.claim + p.H;
  p.part = (int*)(p.prio + p.H);
  int grid = B < mpnn_num_sms() ? B : mpnn_num_sms();
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_prep, dim3(grid), dim3(PT), args, 0, stream);
  MPNN_REQUIRE(e == cudaSuccess, MPNN_ERR_CUDA, "prep_edges: cooperative launch failed: %s", cudaGetErrorString(e));
  return MPNN_OK;
}

}  // extern "C"
