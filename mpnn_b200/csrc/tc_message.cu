// Typed message path on the 5th-generation tensor cores (tcgen05 + TMEM), feature widths 33..256.
//
// Reference: edge_network.py:42-52 (message function) + adjacent_message_agg.py:18 (aggregation).  On the typed
// path (exact de-duplication of bond rows, csrc/dedup.cu) the per-pair matrix A(bfm[b,i,j]) is one of U table
// entries T[u] (U ~ tens), so with the edges grouped by u the message function is a GROUPED GEMM:
//
//     Y[e, :]  = alpha_e * T[u_e] h[src_e]                (forward;  A = gathered sender states,  B = T[u])
//     dG[e, :] = alpha_e * T[u_e]^T dM[dst_e]             (backward; A = gathered message grads,  B = T[u]^T)
//     dT[u]    = sum_{e of type u} alpha_e h[src_e] (x) dM[dst_e]      (backward, K = the edges of the type)
//
// followed by the fixed-order CSR / CSC segmented sums (mpnn_segment_sum) -> no float atomics, bit-reproducible.
//
// Kernel anatomy (one persistent CTA per SM, 288 threads):
//   warps 0-3  producers: gather 128-byte row pieces with 16-byte cp.async (8 lanes per row, fully coalesced) straight
//              into the canonical 128B-swizzled UMMA layout in shared memory -- no register staging, so a 4-stage
//              ring keeps ~64 KB of gathers in flight per SM; completion is tracked by the stage's mbarrier
//              (cp.async.mbarrier.arrive.noinc).  The B operand (the type's matrix) comes from a pre-swizzled image
//              of the table with ONE bulk async copy per stage (cp.async.bulk + complete_tx).
//   warp  4    one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=DP, K=8) with the accumulator in TMEM,
//              tcgen05.commit releases the stage (empty[stage]) and publishes the accumulator (acc_full).
//   warps 5-8  epilogue: tcgen05.ld the accumulator (one TMEM lane = one edge row per thread), scale by alpha_e,
//              store.  The accumulator is double buffered in TMEM so the epilogue overlaps the next tile's MMAs.
// Operands are fp32 in memory and are read by the tensor core as TF32 (10-bit mantissa); accumulation is fp32.
#include "common.cuh"

namespace {

constexpr int TILE = 128;      // edges per tile = UMMA M
constexpr int KB = 32;         // fp32 elements per 128-byte swizzle line
constexpr int PRODUCERS = 128;
constexpr int THREADS = 288;
constexpr int MMA_WARP = 4;

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol bug must end in a trap (a CUDA error on the host), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
// 16-byte asynchronous global->shared copy; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// contiguous global->shared bulk copy (TMA engine, no tensor map); bytes are credited to the mbarrier's tx count
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// mbarrier arrive once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA instruction descriptor, kind::tf32: D fp32 [4,6)=1, A/B format TF32 [7,10)=[10,13)=2, A/B major bits 15/16
// (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// UMMA shared-memory descriptor, 128-byte swizzle: start>>4 [0,14), leading byte offset>>4 [16,30), stride byte
// offset>>4 [32,46), version 1 [46,48), layout type SWIZZLE_128B = 2 [61,64).  Tile bases are 1024-byte aligned.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                               uint32_t layout_type = 2u) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void sts4(uint8_t* base, uint32_t off, float4 v) {
  *reinterpret_cast<float4*>(base + off) = v;
}
// byte offset of 16-byte chunk `chunk` of 128-byte line `line` inside a swizzled region (region base 1024-aligned)
__device__ __forceinline__ uint32_t swz(int line, int chunk) {
  return (uint32_t)line * 128u + (uint32_t)((chunk ^ (line & 7)) << 4);
}
// MN-major fp32/tf32 operands use the SWIZZLE_128B_BASE32B layout (UMMA layout type 1): atoms of 4 K-rows x 128
// bytes (32 fp32 along M/N); inside an atom the 32-BYTE chunk index is XORed with the K-row index.  Returns the
// byte offset of 16-byte chunk `chunk` (0..7) of K-row `row4` (0..3) inside a 512-byte atom.
__device__ __forceinline__ uint32_t swz32(int row4, int chunk) {
  return (uint32_t)row4 * 128u + (uint32_t)(((chunk >> 1) ^ row4) << 5) + (uint32_t)((chunk & 1) << 4);
}

// ---------------------------------------------------------------------------------------------------
// tile plan: the type-sorted edge list cut into single-type tiles of <= 128 edges (device side, no host read),
// plus the per-position gather indices and weights (psrc/pdst/palpha) so the kernels need one index load per row.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_tc_plan_scan(const int* __restrict__ type_ptr, int ntypes, int max_tiles,
                                                       int* __restrict__ tile_off /*[ntypes+1]*/,
                                                       int* __restrict__ head /*{n_tiles, unit_alpha, 0, 0}*/) {
  __shared__ int buf[1024];
  __shared__ int carry;
  const int tid = threadIdx.x;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int u0 = 0; u0 < ntypes; u0 += 1024) {
    const int u = u0 + tid;
    int nt = 0;
    if (u < ntypes) {
      const int cnt = type_ptr[u + 1] - type_ptr[u];
      nt = cnt > 0 ? (cnt + TILE - 1) / TILE : 0;
    }
    buf[tid] = nt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = tid >= o ? buf[tid - o] : 0;
      __syncthreads();
      buf[tid] += t;
      __syncthreads();
    }
    if (u < ntypes) tile_off[u] = carry + buf[tid] - nt;
    __syncthreads();
    if (tid == 0) carry += buf[1023];
    __syncthreads();
  }
  if (tid == 0) {
    tile_off[ntypes] = carry;
    head[0] = min(carry, max_tiles);
    head[1] = 1;  // cleared by k_tc_plan_gather if any edge weight differs from 1
  }
}

// tile t -> (type, first sorted position, edge count): binary search of t in tile_off
__global__ void __launch_bounds__(256) k_tc_plan_fill(const int* __restrict__ type_ptr, int ntypes,
                                                      const int* __restrict__ tile_off,
                                                      const int* __restrict__ n_tiles, int* __restrict__ tile_type,
                                                      int* __restrict__ tile_pos, int* __restrict__ tile_cnt) {
  const int total = *n_tiles;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < total; t += gridDim.x * 256) {
    int lo = 0, hi = ntypes - 1;   // last u with tile_off[u] <= t (types without tiles share their successor's offset)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int u = lo;
    const int p0 = type_ptr[u], cnt = type_ptr[u + 1] - p0;
    const int k = t - tile_off[u];
    tile_type[t] = u;
    tile_pos[t] = p0 + k * TILE;
    tile_cnt[t] = min(TILE, cnt - k * TILE);
  }
}

// psrc[p] = edge_src[type_eid[p]], pdst likewise, palpha[p] = edge_w[type_eid[p]] (1 if edge_w is null)
__global__ void __launch_bounds__(256) k_tc_plan_gather(const int* __restrict__ type_ptr, int ntypes, int cap,
                                                        const int* __restrict__ type_eid,
                                                        const int* __restrict__ edge_src,
                                                        const int* __restrict__ edge_dst,
                                                        const float* __restrict__ edge_w, int* __restrict__ psrc,
                                                        int* __restrict__ pdst, float* __restrict__ palpha,
                                                        int* __restrict__ head) {
  const int E = min(type_ptr[ntypes], cap);
  bool unit = true;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < E; p += gridDim.x * 256) {
    const int e = type_eid[p];
    psrc[p] = edge_src[e];
    pdst[p] = edge_dst[e];
    const float w = edge_w ? edge_w[e] : 1.f;
    palpha[p] = w;
    unit = unit && (w == 1.f);
  }
  if (!unit) head[1] = 0;
}

struct TcPlan {
  const int* head;      // {n_tiles, unit_alpha}
  const int* tile_off;
  const int* tile_type;
  const int* tile_pos;
  const int* tile_cnt;
  const int* psrc;
  const int* pdst;
  const float* palpha;
};

// Bimg[(u * NKB + kb)][n][chunk ^ (n & 7)][4] = Bm[u][n][kb*32 + chunk*4 + 0..3]: the exact shared-memory image of
// one K-block of one type's matrix (K-major, 128-byte swizzle), so a stage's B operand is ONE contiguous bulk copy
__global__ void __launch_bounds__(256) k_tc_swizzle_table(const float* __restrict__ Bm, int ntypes, int DP,
                                                          float* __restrict__ Bimg) {
  const int nkb = DP / KB;
  const long long total = (long long)ntypes * DP * (DP / 4);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c4 = (int)(i % (DP / 4));
    const int n = (int)((i / (DP / 4)) % DP);
    const int u = (int)(i / ((long long)DP * (DP / 4)));
    const int kb = c4 >> 3, chunk = c4 & 7;
    const float4 v = *reinterpret_cast<const float4*>(Bm + ((size_t)u * DP + n) * DP + c4 * 4);
    float* dst = Bimg + (((size_t)u * nkb + kb) * DP + n) * KB + ((chunk ^ (n & 7)) << 2);
    *reinterpret_cast<float4*>(dst) = v;
  }
}

struct TcGemm {
  TcPlan plan;
  const int* type_eid;  // sorted position -> edge id (output row)
  const int* prow;      // sorted position -> row of A (plan.psrc or plan.pdst)
  const float* A;       // [*, lda]
  const float* Bimg;    // pre-swizzled image of the B matrices
  int use_alpha;
  float* Y;             // [E, ldy]
  int lda, K, ldy, N;
  // dense mode (plain GEMM on contiguous rows, no plan): tile t -> row block t / G, N-block ("type") t % G
  int dense, G;
  long long rows;
  int kseg;             // K segments accumulated per tile: segment s reads A columns s*acol.. and B matrix u*kseg+s
  int acol;
  long long ycol;       // output offset (floats) between N-blocks
  int accumulate;       // Y += result
  int nsplit;           // > 0: output columns [j*nsplit, (j+1)*nsplit) go to buffer j (Y + j*ycol), G must be 1
  const float* bias;    // [G*N] or null
};

// tile t -> (type u, first row / sorted position, rows in the tile)
__device__ __forceinline__ void gemm_tile(const TcGemm& a, int t, int& u, int& pos, int& cnt) {
  if (a.dense) {
    const int rt = t / a.G;
    u = t - rt * a.G;
    pos = rt * TILE;
    const long long rem = a.rows - (long long)pos;
    cnt = rem < TILE ? (int)rem : TILE;
  } else {
    u = a.plan.tile_type[t];
    pos = a.plan.tile_pos[t];
    cnt = a.plan.tile_cnt[t];
  }
}

template <int DP>
struct GemmCfg {
  static constexpr int A_BYTES = TILE * 128;
  static constexpr int B_BYTES = DP * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = 4;
  static constexpr int EPI_BYTES = 4 * 32 * 33 * 4;  // per epilogue warp: 32 rows x 32 columns, row stride 33
  static constexpr int SMEM = NSTAGE * STAGE + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static constexpr int TCOLS = 2 * DP;  // double-buffered accumulator
};

// ---------------------------------------------------------------------------------------------------
// Y[e, 0:N] = alpha_e * sum_k A[row(e), k] * B[type(e)][n][k]       (edges in type-sorted tiles)
// ---------------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(THREADS, 1) k_tc_edge_gemm(TcGemm a) {
  using C = GemmCfg<DP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(smem + C::NSTAGE * C::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE + C::EPI_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::NSTAGE + s); };
  auto accfull_bar = [&](int s) { return bar_base + 8u * (2 * C::NSTAGE + s); };
  auto accempty_bar = [&](int s) { return bar_base + 8u * (2 * C::NSTAGE + 2 + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) {
      mbar_init(full_bar(s), PRODUCERS + 1);  // 128 cp.async completions + the bulk copy's expect_tx arrival
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accfull_bar(s), 1);
      mbar_init(accempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = a.dense ? (int)((a.rows + TILE - 1) / TILE) * a.G : a.plan.head[0];
  const int per = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = blockIdx.x * per;
  const int t1 = min(t0 + per, n_tiles);
  constexpr int NKB = DP / KB;
  const int nkbu = (a.K + KB - 1) / KB;   // K-blocks that hold data (K < DP: the all-zero ones are skipped)

  if (warp < 4) {
    // ===================== producers =====================
    const int sub = tid >> 3, chunk = tid & 7;
    int stage = 0, phase = 0;
    int rows_next[8];   // A rows of the NEXT tile (prefetched one tile ahead)
    auto load_rows = [&](int t, int* rows) {
      if (t < t1) {
        int u_, pos, cnt;
        gemm_tile(a, t, u_, pos, cnt);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 16 + sub;
          rows[i] = r < cnt ? (a.dense ? pos + r : __ldg(a.prow + pos + r)) : -1;
        }
      }
    };
    load_rows(t0, rows_next);
    for (int t = t0; t < t1; ++t) {
      int u, pos_, cnt_;
      gemm_tile(a, t, u, pos_, cnt_);
      int rows[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rows[i] = rows_next[i];
      load_rows(t + 1, rows_next);
      for (int sk = 0, seg = 0, kb = 0; sk < a.kseg * nkbu; ++sk, kb = (kb + 1 == nkbu ? 0 : kb + 1), seg += (kb == 0)) {
        const float* bimg = a.Bimg + (size_t)(u * a.kseg + seg) * NKB * (DP * KB);
        const int kk = kb * KB + chunk * 4;
        const int acol = seg * a.acol + kk;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t As = smem_base + stage * C::STAGE;
        if (tid == 0) {
          mbar_arrive_expect_tx(full_bar(stage), C::B_BYTES);
          bulk_copy(As + C::A_BYTES, bimg + (size_t)kb * (DP * KB), C::B_BYTES, full_bar(stage));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = rows[i] >= 0 && kk < a.K;
          const float* src = ok ? a.A + (size_t)rows[i] * a.lda + acol : a.A;
          cp_async16(As + swz(i * 16 + sub, chunk), src, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(full_bar(stage));
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE, DP, 0, 0);
      int stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int i = t - t0, acc = i & 1, use = i >> 1;
        mbar_wait(accempty_bar(acc), (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * DP);
        for (int kb = 0; kb < a.kseg * nkbu; ++kb) {
          mbar_wait(full_bar(stage), phase);
          fence_proxy_async();   // cp.async wrote through the generic proxy; the MMA reads through the async proxy
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE;
          const uint64_t ad = make_sdesc(sa, 16, 1024);
          const uint64_t bd = make_sdesc(sa + C::A_BYTES, 16, 1024);
#pragma unroll
          for (int j = 0; j < KB / 8; ++j)  // K = 8 per instruction: +32 bytes inside the swizzle line
            umma_tf32(d, ad + 2u * j, bd + 2u * j, idesc, (kb | j) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(accfull_bar(acc));
      }
    }
  } else {
    // ===================== epilogue =====================
    // One TMEM lane (= one edge row) per thread comes out of tcgen05.ld; storing it as is would touch 32 different
    // 128-byte lines per warp instruction.  Each warp transposes its 32 x 32 block through shared memory so that
    // 8 lanes write one row's 128 contiguous bytes (4 rows per instruction).
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int r = q * 32 + lane;
    float* tb = epi + q * (32 * 33);
    const int orow = lane >> 3, ocol = (lane & 7) * 4;
    for (int t = t0; t < t1; ++t) {
      const int i = t - t0, acc = i & 1, use = i >> 1;
      int u, pos, cnt;
      gemm_tile(a, t, u, pos, cnt);
      const int e = r < cnt ? (a.dense ? pos + r : __ldg(a.type_eid + pos + r)) : -1;
      const float al = (e >= 0 && a.use_alpha) ? __ldg(a.plan.palpha + pos + r) : 1.f;
      const long long ybase = a.dense ? (long long)u * a.ycol : 0;
      const float* bias = a.bias ? a.bias + (size_t)u * a.N : nullptr;
      int erow[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) erow[it] = __shfl_sync(0xffffffffu, e, it * 4 + orow);
      mbar_wait(accfull_bar(acc), use & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < DP; c0 += 32) {
        if (c0 >= a.N) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * DP + c0), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) tb[lane * 33 + c] = al * v[c];
        __syncwarp();
        if (c0 + ocol < a.N) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + c0 + ocol));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const float* p = tb + (it * 4 + orow) * 33 + ocol;
            if (erow[it] >= 0) {
              const int cc = c0 + ocol;
              const long long yoff = a.nsplit > 0 ? (long long)(cc / a.nsplit) * a.ycol + (cc % a.nsplit) : ybase + cc;
              float4* y = reinterpret_cast<float4*>(a.Y + (size_t)erow[it] * a.ldy + yoff);
              float4 o = make_float4(p[0] + bv.x, p[1] + bv.y, p[2] + bv.z, p[3] + bv.w);
              if (a.accumulate) {
                const float4 old = *y;
                o.x += old.x;
                o.y += old.y;
                o.z += old.z;
                o.w += old.w;
              }
              *y = o;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(accempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// dT[u][l][k] = sum_{e of type u} alpha_e H[src_e, l] dM[dst_e, k]: K = the edges.  Both operands are gathered
// rows, i.e. MN-major: a stage holds 32 edges, every 128-byte row piece is one K-row of a [4 x 32] atom of the
// SWIZZLE_128B_BASE32B layout (the only shared-memory layout tcgen05 accepts for MN-major 32-bit operands).
// A CTA walks a contiguous range of tiles and keeps accumulating in TMEM while the type does not change; each
// (CTA, type) segment is flushed to partial slot (cta + type) -- strictly increasing along the sorted list --
// and k_tc_table_reduce sums the slots of a type in fixed order.
// Unit edge weights (0/1 adjacency, every reference dataset): 16-byte cp.async gathers as in the edge GEMM.
// General weights: the H rows are scaled by alpha_e in registers on their way to shared memory.
// ---------------------------------------------------------------------------------------------------
struct TcGrad {
  TcPlan plan;
  const float* H;    // [*, nf]
  const float* dM;   // [*, mf]
  float* partial;    // [slots][DP][DP]
  int nf, mf;
  int use_alpha;     // 0: all weights are 1 (HEAD form); 1: the plan's edge weights
  // dense mode (out[g] = H^T dM[:, g*bcol ...] on contiguous rows): tile t -> type t / RT, row block t % RT
  int dense, G, RT;
  long long rows;
  int ldh, ldm, bcol;
};

__device__ __forceinline__ void grad_tile(const TcGrad& a, int t, int& u, int& pos, int& cnt) {
  if (a.dense) {
    u = t / a.RT;
    pos = (t - u * a.RT) * TILE;
    const long long rem = a.rows - (long long)pos;
    cnt = rem < TILE ? (int)rem : TILE;
  } else {
    u = a.plan.tile_type[t];
    pos = a.plan.tile_pos[t];
    cnt = a.plan.tile_cnt[t];
  }
}

template <int DP>
struct GradCfg {
  static constexpr int MP = DP < 128 ? 128 : DP;  // UMMA M is 128: narrower tables are zero padded
  static constexpr int MB = MP / 128;
  static constexpr int KST = DP == 256 ? 32 : 64;  // edges per stage
  static constexpr int EPT = KST / 16;             // edges per producer thread and stage
  static constexpr int A_BYTES = KST * MP * 4;
  static constexpr int B_BYTES = KST * DP * 4;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = DP == 64 ? 4 : 3;
  static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;
  static constexpr int TCOLS = MB * DP < 32 ? 32 : MB * DP;
  static constexpr uint32_t SBO = 512;               // between [4 x 32] atoms along K (groups of 4 edges)
  static constexpr uint32_t LBO = (KST / 4) * 512;   // between atoms along M/N (32 columns)
};

template <int DP>
__global__ void __launch_bounds__(THREADS, 1) k_tc_table_grad(TcGrad a) {
  using C = GradCfg<DP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 2);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::NSTAGE + s); };
  const uint32_t accfull_bar = bar_base + 8u * (2 * C::NSTAGE);
  const uint32_t accempty_bar = bar_base + 8u * (2 * C::NSTAGE + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) {
      mbar_init(full_bar(s), PRODUCERS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accfull_bar, 1);
    mbar_init(accempty_bar, 128);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = a.dense ? a.G * a.RT : a.plan.head[0];
  const bool unit_alpha = a.dense || !a.use_alpha || a.plan.head[1] != 0;
  const int per = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = blockIdx.x * per;
  const int t1 = min(t0 + per, n_tiles);

  if (warp < 4) {
    // ===================== producers =====================
    const int sub = tid >> 3, chunk = tid & 7;   // sub: 0..15 -> edges sub, sub+16, ... of the stage
    int stage = 0, phase = 0;
    if (unit_alpha) {
      // 32-column blocks of the A operand that lie entirely behind nf (a 64-wide table in the M = 128 operand) are never
      // written again: clear them in every stage once instead of zero-filling them with cp.async per stage
      for (int st = 0; st < C::NSTAGE; ++st)
        for (int mb = 0; mb < C::MP / 32; ++mb)
          if (mb * 32 >= a.nf)
            for (int i = tid; i < C::KST * 8; i += PRODUCERS)
              sts4(smem + st * C::STAGE, (uint32_t)mb * C::LBO + (uint32_t)i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
      fence_proxy_async();
    }
    // gather indices of the NEXT stage are loaded while the current one is being issued (a dependent L2 / DRAM round
    // trip per stage otherwise: 84 stages per CTA at B = 16 384)
    int isrc[C::EPT], idst[C::EPT];
    auto load_idx = [&](int t, int s0) {
      if (a.dense || t >= t1) return;
      int u2, pos2, cnt2;
      grad_tile(a, t, u2, pos2, cnt2);
#pragma unroll
      for (int h = 0; h < C::EPT; ++h) {
        const int r = s0 + sub + 16 * h;
        isrc[h] = r < cnt2 ? __ldg(a.plan.psrc + pos2 + r) : -1;
        idst[h] = r < cnt2 ? __ldg(a.plan.pdst + pos2 + r) : -1;
      }
    };
    load_idx(t0, 0);
    for (int t = t0; t < t1; ++t) {
      int u_, pos, cnt;
      grad_tile(a, t, u_, pos, cnt);
      for (int s0 = 0; s0 < cnt; s0 += C::KST) {
        const float* hrow[C::EPT];
        const float* mrow[C::EPT];
        float al[C::EPT];
#pragma unroll
        for (int h = 0; h < C::EPT; ++h) {
          const int r = s0 + sub + 16 * h;
          const bool ok = r < cnt;
          if (a.dense) {
            hrow[h] = ok ? a.H + (size_t)(pos + r) * a.ldh : nullptr;
            mrow[h] = ok ? a.dM + (size_t)(pos + r) * a.ldm + (size_t)u_ * a.bcol : nullptr;
          } else {
            hrow[h] = ok ? a.H + (size_t)isrc[h] * a.nf : nullptr;
            mrow[h] = ok ? a.dM + (size_t)idst[h] * a.mf : nullptr;
          }
          al[h] = (ok && !unit_alpha) ? __ldg(a.plan.palpha + pos + r) : 1.f;
        }
        if (s0 + C::KST < cnt) load_idx(t, s0 + C::KST);
        else load_idx(t + 1, 0);
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t As = smem_base + stage * C::STAGE;
        const uint32_t Bs = As + C::A_BYTES;
        if (unit_alpha) {
#pragma unroll
          for (int mb = 0; mb < C::MP / 32; ++mb) {
            if (mb * 32 >= a.nf) continue;   // padding columns of the M = 128 operand: zeroed once, below
#pragma unroll
            for (int h = 0; h < C::EPT; ++h) {
              const int r = sub + 16 * h, col = mb * 32 + chunk * 4;
              const bool ok = hrow[h] != nullptr && col < a.nf;
              cp_async16(As + (uint32_t)mb * C::LBO + (uint32_t)(r >> 2) * C::SBO + swz32(r & 3, chunk),
                         ok ? hrow[h] + col : a.H, ok ? 16u : 0u);
            }
          }
#pragma unroll
          for (int nb = 0; nb < DP / 32; ++nb)
#pragma unroll
            for (int h = 0; h < C::EPT; ++h) {
              const int r = sub + 16 * h, col = nb * 32 + chunk * 4;
              const bool ok = mrow[h] != nullptr && col < a.mf;
              cp_async16(Bs + (uint32_t)nb * C::LBO + (uint32_t)(r >> 2) * C::SBO + swz32(r & 3, chunk),
                         ok ? mrow[h] + col : a.dM, ok ? 16u : 0u);
            }
          cp_async_arrive_noinc(full_bar(stage));
        } else {
          uint8_t* Ag = smem + stage * C::STAGE;
          uint8_t* Bg = Ag + C::A_BYTES;
          {
            float4 v[C::EPT * C::MP / 32];
#pragma unroll
            for (int mb = 0; mb < C::MP / 32; ++mb)
#pragma unroll
              for (int h = 0; h < C::EPT; ++h) {
                const int col = mb * 32 + chunk * 4;
                float4 x = (hrow[h] != nullptr && col < a.nf) ? ldg4(hrow[h] + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[mb * C::EPT + h] = make_float4(al[h] * x.x, al[h] * x.y, al[h] * x.z, al[h] * x.w);
              }
#pragma unroll
            for (int mb = 0; mb < C::MP / 32; ++mb)
#pragma unroll
              for (int h = 0; h < C::EPT; ++h) {
                const int r = sub + 16 * h;
                sts4(Ag, (uint32_t)mb * C::LBO + (uint32_t)(r >> 2) * C::SBO + swz32(r & 3, chunk), v[mb * C::EPT + h]);
              }
          }
          {
            float4 v[C::EPT * DP / 32];
#pragma unroll
            for (int nb = 0; nb < DP / 32; ++nb)
#pragma unroll
              for (int h = 0; h < C::EPT; ++h) {
                const int col = nb * 32 + chunk * 4;
                v[nb * C::EPT + h] = (mrow[h] != nullptr && col < a.mf) ? ldg4(mrow[h] + col) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
            for (int nb = 0; nb < DP / 32; ++nb)
#pragma unroll
              for (int h = 0; h < C::EPT; ++h) {
                const int r = sub + 16 * h;
                sts4(Bg, (uint32_t)nb * C::LBO + (uint32_t)(r >> 2) * C::SBO + swz32(r & 3, chunk), v[nb * C::EPT + h]);
              }
          }
          fence_proxy_async();
          mbar_arrive(full_bar(stage));
        }
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, DP, 1, 1);
      int stage = 0, phase = 0, seg = 0, cur = -1;
      bool fresh = true;
      for (int t = t0; t < t1; ++t) {
        int u, pos_, cnt;
        grad_tile(a, t, u, pos_, cnt);
        if (u != cur) {
          if (cur >= 0) {
            umma_commit(accfull_bar);
            mbar_wait(accempty_bar, seg & 1);  // the epilogue has drained the accumulator of segment `seg`
            tc_fence_after();
            ++seg;
          }
          cur = u;
          fresh = true;
        }
        for (int s0 = 0; s0 < cnt; s0 += C::KST) {
          mbar_wait(full_bar(stage), phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE;
#pragma unroll
          for (int j = 0; j < C::KST / 8; ++j) {
            // K = 8 per instruction = two 4-edge atoms (SBO apart); operands span M/N in atoms LBO apart
            const uint64_t bd = make_sdesc(sa + C::A_BYTES + j * 1024, C::LBO, C::SBO, 1u);
#pragma unroll
            for (int m = 0; m < C::MB; ++m) {
              const uint64_t ad = make_sdesc(sa + m * 4 * C::LBO + j * 1024, C::LBO, C::SBO, 1u);
              umma_tf32(tmem_base + (uint32_t)(m * DP), ad, bd, idesc, fresh ? 0u : 1u);
            }
            fresh = false;
          }
          umma_commit(empty_bar(stage));
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (cur >= 0) umma_commit(accfull_bar);
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    int seg = 0;
    for (int t = t0; t < t1; ++t) {
      int u, pos_, cnt_, un = -1;
      grad_tile(a, t, u, pos_, cnt_);
      if (t + 1 < t1) grad_tile(a, t + 1, un, pos_, cnt_);
      const bool last = (t + 1 == t1) || (un != u);
      if (!last) continue;
      mbar_wait(accfull_bar, seg & 1);
      tc_fence_after();
      float* out = a.partial + (size_t)(blockIdx.x + u) * DP * DP;
#pragma unroll 1
      for (int m = 0; m < C::MB; ++m) {
        const int l = m * 128 + q * 32 + lane;
        if (m * 128 + q * 32 < DP) {  // warp-uniform: quadrants beyond a narrow table hold only padding
#pragma unroll 1
          for (int c0 = 0; c0 < DP; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * DP + c0), v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(out + (size_t)l * DP + c0 + 4 * j) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(accempty_bar);
      ++seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// Fused masked GRU forward (gru_update.py:26-35, 66-68) for widths 33..128: both gate products and the gate
// arithmetic in ONE kernel, so the [rows, 3d] pre-activations never travel through HBM.
//   K segment 0: A = messages m, B = W_ih  -> accumulators NI, R, Z      (TMEM columns [0,DP) [DP,2DP) [2DP,3DP))
//   K segment 1: A = states   h, B = W_hh  -> accumulators     R, Z, NH  (R, Z accumulate on top of segment 0)
//   epilogue   : r = sigmoid(R + b)mu, z = sigmoid(Z + b)mu, n = tanh(NI + b + r (NH + b))mu, h' = ((1-z)n + z h)mu;
//                each 32 x 32 accumulator block is transposed through shared memory so that every global access of
//                the epilogue (h in, h' and the four saved gate planes out) is a full 128-byte line per 8 lanes.
// ---------------------------------------------------------------------------------------------------
struct TcGru {
  const float* m;      // [rows, d]
  const float* h;      // [rows, d]
  const float* mask;   // [rows]
  const float* Bimg;   // [column blocks][2 segments][NKB][3 gate blocks][DP][32] swizzled
  const float* b_ih;   // [3d]
  const float* b_hh;   // [3d]
  float* hout;         // [rows, d]
  float* gates;        // [rows, 4d]: sigmoid r | sigmoid z | tanh n | nh
  long long rows;
  int d;
  int ncb;             // output-column blocks of DP (d = 256: two blocks of 128, the A tiles are read once per block)
  // aggregation folded into the A-tile producer (adjacent_message_agg.py:18): when Y is given, row i of the message
  // operand is sum_{e in [row_ptr[i], row_ptr[i+1])} Y[e, :] (the per-edge messages of the grouped GEMM, CSR order),
  // summed in edge order while staging; m is ignored and m_out (if given) receives the sums for the backward
  const float* Y;      // [E, d]
  const int* row_ptr;  // [rows + 1]
  float* m_out;        // [rows, d]
};

// MUFU.TANH (max relative error 2^-11, the same order as the TF32 operands feeding it): the gate arithmetic of the
// fused GRU epilogue is issue-bound on 8 warps, so the 1-instruction forms matter
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

template <int DP, int KP>   // DP: output columns per tile (N of the MMAs), KP: padded contraction width
struct GruCfg {
  static constexpr int A_BYTES = TILE * 128;
  static constexpr int B_BYTES = 3 * DP * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = DP == 64 ? 4 : 3;
  static constexpr int EPW = 8;                          // epilogue warps: two per TMEM lane quadrant, split by column chunk
  static constexpr int GRU_THREADS = PRODUCERS + 32 + 32 * EPW;
  static constexpr int EPI_BYTES = EPW * 32 * 33 * 4;
  static constexpr int SMEM = NSTAGE * STAGE + EPI_BYTES + 1024 + 256;
  static constexpr int NACC = (8 * DP <= 512) ? 2 : 1;   // accumulator sets (4*DP columns each) that fit in TMEM
  static constexpr int TCOLS = 512;
};

// image[(((cb*2 + s)*NKB + kb)*3 + j)][n][chunk ^ (n & 7)][4]: column block cb, segment s, K-block kb, gate block j;
// W_* are [d, 3d] input-major
__global__ void __launch_bounds__(256) k_tc_gru_pack(const float* __restrict__ W_ih, const float* __restrict__ W_hh,
                                                     int d, int DP, int KP, int ncb, float* __restrict__ img) {
  const int nkb = KP / KB;
  const long long total = (long long)ncb * 2 * 3 * DP * KP;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int k = (int)(i % KP);
    const int nl = (int)((i / KP) % DP);
    const int j = (int)((i / ((long long)DP * KP)) % 3);
    const int sgm = (int)((i / (3LL * DP * KP)) % 2);
    const int cb = (int)(i / (6LL * DP * KP));
    const int n = cb * DP + nl;
    // segment 0 blocks: NI, R, Z = gates 2, 0, 1 of W_ih; segment 1 blocks: R, Z, NH = gates 0, 1, 2 of W_hh
    const int gate = sgm == 0 ? (j == 0 ? 2 : j - 1) : j;
    const float* W = sgm == 0 ? W_ih : W_hh;
    const float v = (n < d && k < d) ? __ldg(W + (size_t)k * 3 * d + (size_t)gate * d + n) : 0.f;
    const int kb = k >> 5, chunk = (k & 31) >> 2, jj = k & 3;
    img[(((((size_t)cb * 2 + sgm) * nkb + kb) * 3 + j) * DP + nl) * KB + ((chunk ^ (nl & 7)) << 2) + jj] = v;
  }
}

template <int DP, int KP>
__global__ void __launch_bounds__(GruCfg<DP, KP>::GRU_THREADS, 1) k_tc_gru_fwd(TcGru a) {
  using C = GruCfg<DP, KP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(smem + C::NSTAGE * C::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE + C::EPI_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::NSTAGE + s); };
  auto accfull_bar = [&](int s) { return bar_base + 8u * (2 * C::NSTAGE + s); };
  auto accempty_bar = [&](int s) { return bar_base + 8u * (2 * C::NSTAGE + 2 + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) {
      mbar_init(full_bar(s), PRODUCERS + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accfull_bar(s), 1);
      mbar_init(accempty_bar(s), 32 * C::EPW);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ncb = a.ncb;
  const int n_tiles = (int)((a.rows + TILE - 1) / TILE) * ncb;   // tile t -> row block t / ncb, column block t % ncb
  const int per = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = blockIdx.x * per;
  const int t1 = min(t0 + per, n_tiles);
  constexpr int NKB = KP / KB;
  const int d = a.d;

  if (warp < 4) {
    // ===================== producers =====================
    const int sub = tid >> 3, chunk = tid & 7;
    int stage = 0, phase = 0;
    for (int t = t0; t < t1; ++t) {
      const long long pos = (long long)(t / ncb) * TILE;
      const int cb = t % ncb;
      for (int sk = 0; sk < 2 * NKB; ++sk) {
        const int seg = sk / NKB, kb = sk - seg * NKB;
        const float* A = seg == 0 ? a.m : a.h;
        const int kk = kb * KB + chunk * 4;
        // the CSR ranges of this thread's eight rows: loaded before the wait for the stage (two dependent round trips)
        int e0[8], e1[8];
        const bool agg = seg == 0 && a.Y != nullptr;
        if (agg) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long row = pos + i * 16 + sub;
            const bool ok = row < a.rows && kk < d;
            e0[i] = ok ? __ldg(a.row_ptr + row) : 0;
            e1[i] = ok ? __ldg(a.row_ptr + row + 1) : 0;
          }
        }
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t As = smem_base + stage * C::STAGE;
        if (tid == 0) {
          mbar_arrive_expect_tx(full_bar(stage), C::B_BYTES);
          bulk_copy(As + C::A_BYTES, a.Bimg + (size_t)((cb * 2 + seg) * NKB + kb) * 3 * (DP * KB), C::B_BYTES,
                    full_bar(stage));
        }
        if (agg) {
          uint8_t* Ag = smem + stage * C::STAGE;
          // four rows at a time, the first four edges of each with all 16 loads in flight (molecular graphs: degree <= 4;
          // a sequential loop per row costs one DRAM latency per ROW); absent edges add +0, the order stays e ascending
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float4 y[4][4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int i = half * 4 + ii;
              const float* yb = a.Y + (size_t)e0[i] * d + kk;
              const int ne = e1[i] - e0[i];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                y[ii][j] = j < ne ? ldg4(yb + j * d) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int i = half * 4 + ii;
              const long long row = pos + i * 16 + sub;
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc.x += y[ii][j].x;
                acc.y += y[ii][j].y;
                acc.z += y[ii][j].z;
                acc.w += y[ii][j].w;
              }
              for (int e = e0[i] + 4; e < e1[i]; ++e) {
                const float4 z = ldg4(a.Y + (size_t)e * d + kk);
                acc.x += z.x;
                acc.y += z.y;
                acc.z += z.z;
                acc.w += z.w;
              }
              sts4(Ag, swz(i * 16 + sub, chunk), acc);
              if (a.m_out && cb == 0 && row < a.rows && kk < d)
                *reinterpret_cast<float4*>(a.m_out + (size_t)row * d + kk) = acc;
            }
          }
          fence_proxy_async();
          mbar_arrive(full_bar(stage));
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long row = pos + i * 16 + sub;
          const bool ok = row < a.rows && kk < d;
          cp_async16(As + swz(i * 16 + sub, chunk), ok ? A + (size_t)row * d + kk : A, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(full_bar(stage));
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE, DP, 0, 0);
      int stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int i = t - t0, acc = i % C::NACC, use = i / C::NACC;
        mbar_wait(accempty_bar(acc), (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t dbase = tmem_base + (uint32_t)(acc * 4 * DP);
        for (int sk = 0; sk < 2 * NKB; ++sk) {
          const int seg = sk / NKB, kb = sk - seg * NKB;
          mbar_wait(full_bar(stage), phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE;
          const uint64_t ad = make_sdesc(sa, 16, 1024);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            // segment 0 blocks -> NI, R, Z (column blocks 0, 1, 2); segment 1 blocks -> R, Z, NH (1, 2, 3)
            const uint32_t dcol = dbase + (uint32_t)((seg + j) * DP);
            const uint64_t bd = make_sdesc(sa + C::A_BYTES + j * DP * 128, 16, 1024);
            const bool carried = seg == 1 && j < 2;   // R and Z continue the sums started by segment 0
#pragma unroll
            for (int ks = 0; ks < KB / 8; ++ks)
              umma_tf32(dcol, ad + 2u * ks, bd + 2u * ks, idesc, (carried || (kb | ks) != 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(accfull_bar(acc));
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quadrant of this warp
    const int ew = warp - (MMA_WARP + 1);   // 0 .. EPW-1
    const int half = ew >> 2;               // which 32-column chunks this warp takes
    float* tb = epi + ew * (32 * 33);
    const int orow = lane >> 3, ocol = (lane & 7) * 4;
    for (int t = t0; t < t1; ++t) {
      const int i = t - t0, acc = i % C::NACC, use = i / C::NACC;
      const long long pos = (long long)(t / ncb) * TILE + q * 32;
      const int cbase = (t % ncb) * DP;   // first output column of this tile
      mbar_wait(accfull_bar(acc), use & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 4 * DP);
      float mu[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const long long row = pos + it * 4 + orow;
        mu[it] = row < a.rows ? __ldg(a.mask + row) : 0.f;
      }
#pragma unroll 1
      for (int c0 = 32 * half; c0 < DP; c0 += 32 * (C::EPW / 4)) {
        if (cbase + c0 >= d) break;
        const int col = cbase + c0 + ocol;
        const bool cok = col < d;
        float4 sr[8], sz[8], nh[8];
        float v[32];
        // ---- R ----
        tmem_ld32(tacc + (uint32_t)(DP + c0), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) tb[lane * 33 + c] = v[c];
        __syncwarp();
        {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cok) {
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.b_ih + col));
            const float4 b2 = __ldg(reinterpret_cast<const float4*>(a.b_hh + col));
            b = make_float4(b1.x + b2.x, b1.y + b2.y, b1.z + b2.z, b1.w + b2.w);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const float* p = tb + (it * 4 + orow) * 33 + ocol;
            sr[it] = make_float4(fast_sigmoid(p[0] + b.x), fast_sigmoid(p[1] + b.y), fast_sigmoid(p[2] + b.z), fast_sigmoid(p[3] + b.w));
          }
        }
        __syncwarp();
        // ---- Z ----
        tmem_ld32(tacc + (uint32_t)(2 * DP + c0), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) tb[lane * 33 + c] = v[c];
        __syncwarp();
        {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cok) {
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.b_ih + d + col));
            const float4 b2 = __ldg(reinterpret_cast<const float4*>(a.b_hh + d + col));
            b = make_float4(b1.x + b2.x, b1.y + b2.y, b1.z + b2.z, b1.w + b2.w);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const float* p = tb + (it * 4 + orow) * 33 + ocol;
            sz[it] = make_float4(fast_sigmoid(p[0] + b.x), fast_sigmoid(p[1] + b.y), fast_sigmoid(p[2] + b.z), fast_sigmoid(p[3] + b.w));
          }
        }
        __syncwarp();
        // ---- NH ----
        tmem_ld32(tacc + (uint32_t)(3 * DP + c0), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) tb[lane * 33 + c] = v[c];
        __syncwarp();
        {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cok) b = __ldg(reinterpret_cast<const float4*>(a.b_hh + 2 * d + col));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const float* p = tb + (it * 4 + orow) * 33 + ocol;
            nh[it] = make_float4(p[0] + b.x, p[1] + b.y, p[2] + b.z, p[3] + b.w);
          }
        }
        __syncwarp();
        // ---- NI, then the gate arithmetic and the coalesced stores ----
        tmem_ld32(tacc + (uint32_t)c0, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) tb[lane * 33 + c] = v[c];
        __syncwarp();
        if (cok) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.b_ih + 2 * d + col));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const long long row = pos + it * 4 + orow;
            if (row < a.rows) {
              const float* p = tb + (it * 4 + orow) * 33 + ocol;
              const float m_ = mu[it];
              const float4 hv = __ldg(reinterpret_cast<const float4*>(a.h + (size_t)row * d + col));
              float4 tn, ho;
              tn.x = fast_tanh(p[0] + b.x + sr[it].x * m_ * nh[it].x);
              tn.y = fast_tanh(p[1] + b.y + sr[it].y * m_ * nh[it].y);
              tn.z = fast_tanh(p[2] + b.z + sr[it].z * m_ * nh[it].z);
              tn.w = fast_tanh(p[3] + b.w + sr[it].w * m_ * nh[it].w);
              ho.x = ((1.f - sz[it].x * m_) * tn.x * m_ + sz[it].x * m_ * hv.x) * m_;
              ho.y = ((1.f - sz[it].y * m_) * tn.y * m_ + sz[it].y * m_ * hv.y) * m_;
              ho.z = ((1.f - sz[it].z * m_) * tn.z * m_ + sz[it].z * m_ * hv.z) * m_;
              ho.w = ((1.f - sz[it].w * m_) * tn.w * m_ + sz[it].w * m_ * hv.w) * m_;
              *reinterpret_cast<float4*>(a.hout + (size_t)row * d + col) = ho;
              float* g = a.gates + (size_t)row * 4 * d + col;
              *reinterpret_cast<float4*>(g) = sr[it];
              *reinterpret_cast<float4*>(g + d) = sz[it];
              *reinterpret_cast<float4*>(g + 2 * d) = tn;
              *reinterpret_cast<float4*>(g + 3 * d) = nh[it];
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(accempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// GRU weight gradients for widths <= 64 in ONE pass over the gate gradients (gru_update.py:27-28 backward):
//     dW_ih [d, 3d] = m^T (dar | daz | dan),     dW_hh [d, 3d] = h^T (dar | daz | dnh)
// Both products contract over the rows, so every operand is MN-major (the table-gradient machinery above).  m and h
// are stacked into one M = 128 operand ([m row | h row], 64 columns each) and the four gate-gradient blocks into one
// N = 256 operand: one tcgen05.mma per 8 rows gives D[128][256] = [m^T; h^T] (dar|daz|dan|dnh), of which the reduction
// keeps m^T (dar|daz|dan) and h^T (dar|daz|dnh).  Every CTA accumulates its rows in TMEM and writes ONE partial.
// (The per-product kernels read m / h three times and the gate gradients twice: 1.04 GB instead of 0.58 GB at
// B = 16 384, d = 64.)
// ---------------------------------------------------------------------------------------------------
struct TcGruParam {
  const float* m;     // [rows, d]
  const float* h;     // [rows, d]
  const float* dg;    // [rows, ldg]: blocks dar | daz | dan | dnh, d columns each
  float* partial;     // [grid][128][256]
  long long rows;
  int d, ldg;
  // k_tc_gru_param_point only: the saved gates [rows, 4d], dh' [rows, d], mask [rows]; outputs dg [rows, 6d] and the
  // bias partials [grid * 16][4d]
  const float* gates;
  const float* dh_out;
  const float* mask;
  float* dg_out;
  float* bias_part;
};
struct GpCfg {
  static constexpr int KST = 32;                 // rows per stage
  static constexpr int EPT = KST / 16;
  static constexpr int A_BYTES = KST * 128 * 4;
  static constexpr int B_BYTES = KST * 256 * 4;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = 4;
  static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;
  static constexpr int TCOLS = 256;
  static constexpr uint32_t SBO = 512;
  static constexpr uint32_t LBO = (KST / 4) * 512;
};

__global__ void __launch_bounds__(THREADS, 1) k_tc_gru_param_grad(TcGruParam a) {
  using C = GpCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 2);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::NSTAGE + s); };
  const uint32_t accfull_bar = bar_base + 8u * (2 * C::NSTAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) {
      mbar_init(full_bar(s), PRODUCERS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_chunks = (a.rows + C::KST - 1) / C::KST;
  const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const long long c0 = (long long)blockIdx.x * per;
  const long long c1 = c0 + per < n_chunks ? c0 + per : n_chunks;
  const int d = a.d;

  if (warp < 4) {
    // ===================== producers =====================
    const int sub = tid >> 3, chunk = tid & 7;
    int stage = 0, phase = 0;
    for (long long c = c0; c < c1; ++c) {
      const long long pos = c * C::KST;
      mbar_wait(empty_bar(stage), phase ^ 1);
      const uint32_t As = smem_base + stage * C::STAGE;
      const uint32_t Bs = As + C::A_BYTES;
#pragma unroll
      for (int hh = 0; hh < C::EPT; ++hh) {
        const int r = sub + 16 * hh;
        const long long row = pos + r;
        const bool rok = row < a.rows;
        const uint32_t roff = (uint32_t)(r >> 2) * C::SBO + swz32(r & 3, chunk);
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          const int col = (mb & 1) * 32 + chunk * 4;
          const float* base = mb < 2 ? a.m : a.h;
          const bool ok = rok && col < d;
          cp_async16(As + (uint32_t)mb * C::LBO + roff, ok ? base + (size_t)row * d + col : base, ok ? 16u : 0u);
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const int col = (nb & 1) * 32 + chunk * 4;
          const bool ok = rok && col < d;
          cp_async16(Bs + (uint32_t)nb * C::LBO + roff,
                     ok ? a.dg + (size_t)row * a.ldg + (size_t)(nb >> 1) * d + col : a.dg, ok ? 16u : 0u);
        }
      }
      cp_async_arrive_noinc(full_bar(stage));
      if (++stage == C::NSTAGE) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0 && c0 < c1) {
      constexpr uint32_t idesc = make_idesc(128, 256, 1, 1);
      int stage = 0, phase = 0;
      bool fresh = true;
      for (long long c = c0; c < c1; ++c) {
        mbar_wait(full_bar(stage), phase);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE;
#pragma unroll
        for (int j = 0; j < C::KST / 8; ++j) {
          const uint64_t ad = make_sdesc(sa + j * 1024, C::LBO, C::SBO, 1u);
          const uint64_t bd = make_sdesc(sa + C::A_BYTES + j * 1024, C::LBO, C::SBO, 1u);
          umma_tf32(tmem_base, ad, bd, idesc, fresh ? 0u : 1u);
          fresh = false;
        }
        umma_commit(empty_bar(stage));
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(accfull_bar);
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int l = q * 32 + lane;
    float* out = a.partial + (size_t)blockIdx.x * 128 * 256 + (size_t)l * 256;
    if (c0 < c1) {
      mbar_wait(accfull_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 256; cc += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(out + cc + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    } else {
      for (int cc = 0; cc < 256; cc += 4) *reinterpret_cast<float4*>(out + cc) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// The same with the pointwise GRU backward folded into the producers (no separate pass that writes the gate gradients and no
// re-read of them here): reads gates / h / dh' / m once, writes the bias partials and -- when asked -- dg [rows, 6d].
// The raw operands (four gate planes, h, dh', mask of a 32-row stage) travel through a two-deep cp.async ring in shared
// memory, so the loads of stages c+1 and c+2 are in flight while stage c is converted into the MMA operand (a first
// version loaded them into registers: 49 KB in flight per SM with a bubble per stage, 47 % of the DRAM peak).  A thread
// converts exactly the elements it copied, so no barrier is needed between the copy and the conversion.
struct GppCfg {
  static constexpr int KST = 32;
  static constexpr int EPT = 2;
  static constexpr int A_BYTES = KST * 128 * 4;
  static constexpr int B_BYTES = KST * 256 * 4;
  static constexpr int STAGE = A_BYTES + B_BYTES;       // operand stage
  static constexpr int NSTAGE = 2;
  static constexpr int RAW_ITEMS = 24;                  // float4 per thread and stage: (row 2) x (column half 2) x (plane 6)
  static constexpr int RAW_BYTES = RAW_ITEMS * PRODUCERS * 16 + 2 * PRODUCERS * 4;   // + the rows' mask values
  static constexpr int RAW_STAGE = (RAW_BYTES + 1023) / 1024 * 1024;
  static constexpr int NRAW = 2;
  static constexpr int SMEM = NSTAGE * STAGE + NRAW * RAW_STAGE + 1024 + 256;
  static constexpr int TCOLS = 256;
  static constexpr uint32_t SBO = 512;
  static constexpr uint32_t LBO = (KST / 4) * 512;
};

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) k_tc_gru_param_point(TcGruParam a) {
  using C = GppCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* rawbuf = smem + C::NSTAGE * C::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(rawbuf + C::NRAW * C::RAW_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 2);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t raw_base = smem_u32(rawbuf);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::NSTAGE + s); };
  const uint32_t accfull_bar = bar_base + 8u * (2 * C::NSTAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) {
      mbar_init(full_bar(s), 2 * PRODUCERS);   // per producer: one arrival for its stores, one for its async copies
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_chunks = (a.rows + C::KST - 1) / C::KST;
  const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const long long c0 = (long long)blockIdx.x * per;
  const long long c1 = c0 + per < n_chunks ? c0 + per : n_chunks;
  const int d = a.d;

  if (warp < 4) {
    // ===================== producers: pointwise GRU backward + operand staging =====================
    // A thread owns the 4-column groups (half hb, chunk) of rows sub and sub + 16 of every stage.
    const int sub = tid >> 3, chunk = tid & 7;
    float4 bsum[2][4];
#pragma unroll
    for (int hb = 0; hb < 2; ++hb)
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) bsum[hb][q4] = make_float4(0.f, 0.f, 0.f, 0.f);
    // raw item k = (hh * 2 + hb) * 6 + plane lives at raw[(k * PRODUCERS + tid) * 16]: conflict-free both ways
    auto issue_raw = [&](long long c) {
      if (c < c1) {
        const long long pos = c * C::KST;
        const uint32_t rb = raw_base + (uint32_t)((c - c0) & 1) * C::RAW_STAGE;
#pragma unroll
        for (int hh = 0; hh < C::EPT; ++hh) {
          const long long row = pos + sub + 16 * hh;
          const bool rok = row < a.rows;
          cp_async4(rb + C::RAW_ITEMS * PRODUCERS * 16 + (uint32_t)(hh * PRODUCERS + tid) * 4, rok ? a.mask + row : a.mask,
                    rok ? 4u : 0u);
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const int col = hb * 32 + chunk * 4;
            const bool ok = rok && col < d;
            const uint32_t n = ok ? 16u : 0u;
            const float* g = a.gates + (ok ? (size_t)row * 4 * d + col : 0);
            const uint32_t k0 = (uint32_t)((hh * 2 + hb) * 6);
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) cp_async16(rb + ((k0 + pl) * PRODUCERS + tid) * 16, ok ? g + pl * d : a.gates, n);
            cp_async16(rb + ((k0 + 4) * PRODUCERS + tid) * 16, ok ? a.h + (size_t)row * d + col : a.h, n);
            cp_async16(rb + ((k0 + 5) * PRODUCERS + tid) * 16, ok ? a.dh_out + (size_t)row * d + col : a.dh_out, n);
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_raw(c0);
    issue_raw(c0 + 1);
    int stage = 0, phase = 0;
    for (long long c = c0; c < c1; ++c) {
      const long long pos = c * C::KST;
      asm volatile("cp.async.wait_group 1;" ::: "memory");     // this thread's copies of stage c have landed
      const uint8_t* rw = rawbuf + ((c - c0) & 1) * C::RAW_STAGE;
      mbar_wait(empty_bar(stage), phase ^ 1);
      const uint32_t As = smem_base + stage * C::STAGE;
      uint8_t* Bg = smem + stage * C::STAGE + C::A_BYTES;
#pragma unroll
      for (int hh = 0; hh < C::EPT; ++hh) {
        const int r = sub + 16 * hh;
        const long long row = pos + r;
        const bool rok = row < a.rows;
        const uint32_t roff = (uint32_t)(r >> 2) * C::SBO + swz32(r & 3, chunk);
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          const int col = (mb & 1) * 32 + chunk * 4;
          const float* base = mb < 2 ? a.m : a.h;
          const bool ok = rok && col < d;
          cp_async16(As + (uint32_t)mb * C::LBO + roff, ok ? base + (size_t)row * d + col : base, ok ? 16u : 0u);
        }
        const float m_ = *reinterpret_cast<const float*>(rw + C::RAW_ITEMS * PRODUCERS * 16 + (hh * PRODUCERS + tid) * 4);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          const int col = hb * 32 + chunk * 4;
          const bool ok = rok && col < d;
          const int k0 = (hh * 2 + hb) * 6;
          const float4 sr = *reinterpret_cast<const float4*>(rw + ((k0 + 0) * PRODUCERS + tid) * 16);
          const float4 sz = *reinterpret_cast<const float4*>(rw + ((k0 + 1) * PRODUCERS + tid) * 16);
          const float4 tn = *reinterpret_cast<const float4*>(rw + ((k0 + 2) * PRODUCERS + tid) * 16);
          const float4 nh = *reinterpret_cast<const float4*>(rw + ((k0 + 3) * PRODUCERS + tid) * 16);
          const float4 hv = *reinterpret_cast<const float4*>(rw + ((k0 + 4) * PRODUCERS + tid) * 16);
          const float4 dv = *reinterpret_cast<const float4*>(rw + ((k0 + 5) * PRODUCERS + tid) * 16);
          float4 o[6];
#define MPNN_GP(X)                                              \
  {                                                             \
    const float r_ = sr.X * m_, z_ = sz.X * m_, n_ = tn.X * m_; \
    const float go = dv.X * m_;                                 \
    const float dn = go * (1.f - z_);                           \
    const float dz = go * (hv.X - n_);                          \
    const float dan = dn * m_ * (1.f - tn.X * tn.X);            \
    const float dr = dan * nh.X;                                \
    o[0].X = dr * m_ * sr.X * (1.f - sr.X);                     \
    o[1].X = dz * m_ * sz.X * (1.f - sz.X);                     \
    o[2].X = dan;                                               \
    o[3].X = dan * r_;                                          \
    const float gz = go * z_;                                   \
    o[4].X = __uint_as_float(__float_as_uint(gz) & 0xffffe000u); \
    o[5].X = gz - o[4].X;                                       \
  }
          MPNN_GP(x) MPNN_GP(y) MPNN_GP(z) MPNN_GP(w)
#undef MPNN_GP
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            sts4(Bg, (uint32_t)(2 * q4 + hb) * C::LBO + roff, o[q4]);   // (zero where the raw copy was zero-filled)
            bsum[hb][q4].x += o[q4].x;
            bsum[hb][q4].y += o[q4].y;
            bsum[hb][q4].z += o[q4].z;
            bsum[hb][q4].w += o[q4].w;
          }
          if (ok && a.dg_out) {
            float* og = a.dg_out + (size_t)row * 6 * d + col;
#pragma unroll
            for (int q6 = 0; q6 < 6; ++q6) *reinterpret_cast<float4*>(og + q6 * d) = o[q6];
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(full_bar(stage));              // the B image (generic stores)
      cp_async_arrive_noinc(full_bar(stage));    // the A image (and every older copy of this thread)
      issue_raw(c + 2);                          // refill the raw stage just consumed
      if (++stage == C::NSTAGE) {
        stage = 0;
        phase ^= 1;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    // bias column sums of this (CTA, sub): one [4d] row of the partial array (k_gru_bias_final sums the rows in order)
    {
      float* bp = a.bias_part + ((size_t)blockIdx.x * 16 + sub) * 4 * d;
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int col = hb * 32 + chunk * 4;
        if (col < d) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) *reinterpret_cast<float4*>(bp + q4 * d + col) = bsum[hb][q4];
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0 && c0 < c1) {
      constexpr uint32_t idesc = make_idesc(128, 256, 1, 1);
      int stage = 0, phase = 0;
      bool fresh = true;
      for (long long c = c0; c < c1; ++c) {
        mbar_wait(full_bar(stage), phase);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE;
#pragma unroll
        for (int j = 0; j < C::KST / 8; ++j) {
          const uint64_t ad = make_sdesc(sa + j * 1024, C::LBO, C::SBO, 1u);
          const uint64_t bd = make_sdesc(sa + C::A_BYTES + j * 1024, C::LBO, C::SBO, 1u);
          umma_tf32(tmem_base, ad, bd, idesc, fresh ? 0u : 1u);
          fresh = false;
        }
        umma_commit(empty_bar(stage));
        if (++stage == C::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(accfull_bar);
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int l = q * 32 + lane;
    float* out = a.partial + (size_t)blockIdx.x * 128 * 256 + (size_t)l * 256;
    if (c0 < c1) {
      mbar_wait(accfull_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 256; cc += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(out + cc + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    } else {
      for (int cc = 0; cc < 256; cc += 4) *reinterpret_cast<float4*>(out + cc) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// GRU data gradients for widths <= 64 WITHOUT a gate-gradient array in HBM (gru_update.py:27-34 backward):
//     (dm | dh) [rows, 2d] = (dar | daz | dan | dnh | hi(go z) | lo(go z)) [rows, 6d]  x  Wc [6d, 2d]
// The producers read the saved gates, h and dh' of a 128-row x 32-column block ONCE, compute the six gate-gradient
// blocks and store each into its own stage as the K-major A operand (six stages = the six K segments of one
// (row tile, 32-column block) group; the B operand of a stage is the matching 32 x 2d slice of Wc, one bulk copy).
// The MMA warp walks the six stages in order; the accumulator (N = 128 columns: dm | dh) is double-buffered in TMEM;
// the epilogue sends columns [0, d) to dm and [d, 2d) to dh.  With mpnn_tc_gru_param_point the gate gradients of
// the wide GRU backward never reach HBM.
// ---------------------------------------------------------------------------------------------------
struct TcGruData {
  const float *gates, *h, *dh_out, *mask;
  const float* Bimg;   // [6][DP / 32][DP][32] swizzled (k_tc_pack_image, DP = 128)
  float *dm, *dh;
  long long rows;
  int d;
  unsigned long long* dbg;   // profiling aid (mpnn_tc_debug): %globaltimer stamps of CTA 0's producer thread 0
};
struct GdCfg {
  static constexpr int DP = 128;
  static constexpr int NSEG = 6;
  static constexpr int NST = 3;                          // operand stages: three gate-gradient blocks per pass, two passes
  static constexpr int A_BYTES = TILE * 128;
  static constexpr int B_BYTES = DP * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int RAW_ITEMS = 24;                   // float4 per thread and half tile: (row 4) x (plane 6)
  static constexpr int RAW_HALF = (RAW_ITEMS * PRODUCERS * 16 + 4 * PRODUCERS * 4 + 1023) / 1024 * 1024;
  static constexpr int EPI_BYTES = 4 * 32 * 33 * 4;
  static constexpr int SMEM = NST * STAGE + 2 * RAW_HALF + EPI_BYTES + 1024 + 256;
  static constexpr int TCOLS = 2 * DP;
};

// Producer schedule per (row tile, 32-column block) group: the raw operands of its two half tiles (64 rows each) arrive
// through two cp.async buffers.  Pass 0 converts both halves: blocks dar, daz, dan go to the three operand stages, the
// other three (dnh, hi, lo) overwrite the raw slots they were computed from.  Pass 1 moves those into the same stages
// once the MMAs of pass 0 have drained them, and refills each raw buffer with the NEXT group's half as soon as it has been
// read: up to 96 KB of loads are in flight per SM while a group is converted (a first version loaded 24 float4 per
// thread into registers, converted, loaded the next 24: 40 % of the DRAM peak).
__global__ void __launch_bounds__(THREADS, 1) k_tc_gru_data_grad(TcGruData a) {
  using C = GdCfg;
  constexpr int DP = C::DP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* rawbuf = smem + C::NST * C::STAGE;
  float* epi = reinterpret_cast<float*>(rawbuf + 2 * C::RAW_HALF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(rawbuf + 2 * C::RAW_HALF + C::EPI_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::NST + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t raw_base = smem_u32(rawbuf);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::NST + s); };
  auto accfull_bar = [&](int s) { return bar_base + 8u * (2 * C::NST + s); };
  auto accempty_bar = [&](int s) { return bar_base + 8u * (2 * C::NST + 2 + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::NST; ++s) {
      mbar_init(full_bar(s), PRODUCERS + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accfull_bar(s), 1);
      mbar_init(accempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int d = a.d;
  const int nkb = (d + KB - 1) / KB;
  constexpr int NKB = DP / KB;
  const int n_tiles = (int)((a.rows + TILE - 1) / TILE);
  const int per = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = blockIdx.x * per;
  const int t1 = min(t0 + per, n_tiles);
  const int n_groups = (t1 - t0) * nkb;   // group g -> tile t0 + g / nkb, column block g % nkb

  if (warp < 4) {
    // ===================== producers =====================
    const int sub = tid >> 3, chunk = tid & 7;
    // raw item k = ii * 6 + plane of half tile hf lives at raw[hf][(k * PRODUCERS + tid) * 16]
    auto issue_raw = [&](int g, int hf) {
      if (g < n_groups) {
        const long long pos = (long long)(t0 + g / nkb) * TILE;
        const int col = (g % nkb) * KB + chunk * 4;
        const uint32_t rb = raw_base + (uint32_t)hf * C::RAW_HALF;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const long long row = pos + (hf * 4 + ii) * 16 + sub;
          const bool ok = col < d && row < a.rows;
          const uint32_t n = ok ? 16u : 0u;
          cp_async4(rb + C::RAW_ITEMS * PRODUCERS * 16 + (uint32_t)(ii * PRODUCERS + tid) * 4, ok ? a.mask + row : a.mask,
                    ok ? 4u : 0u);
          const float* gp = a.gates + (ok ? (size_t)row * 4 * d + col : 0);
#pragma unroll
          for (int pl = 0; pl < 4; ++pl) cp_async16(rb + ((ii * 6 + pl) * PRODUCERS + tid) * 16, ok ? gp + pl * d : a.gates, n);
          cp_async16(rb + ((ii * 6 + 4) * PRODUCERS + tid) * 16, ok ? a.h + (size_t)row * d + col : a.h, n);
          cp_async16(rb + ((ii * 6 + 5) * PRODUCERS + tid) * 16, ok ? a.dh_out + (size_t)row * d + col : a.dh_out, n);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_raw(0, 0);
    issue_raw(0, 1);
    int phase = 0;
    int ds = 0;
    auto stamp = [&]() {
      if (a.dbg && blockIdx.x == 0 && tid == 0 && ds < 64) {
        unsigned long long tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        a.dbg[ds] = tt;
      }
      ++ds;
    };
    for (int g = 0; g < n_groups; ++g) {
      const int kb = g % nkb;
      stamp();
      // ---- pass 0: dar, daz, dan -> stages 0..2; dnh, hi, lo -> back into the raw slots ----
#pragma unroll 1
      for (int sg = 0; sg < C::NST; ++sg) {
        mbar_wait(empty_bar(sg), phase ^ 1);
        if (tid == 0) {
          mbar_arrive_expect_tx(full_bar(sg), C::B_BYTES);
          bulk_copy(smem_base + sg * C::STAGE + C::A_BYTES, a.Bimg + (size_t)(sg * NKB + kb) * (DP * KB), C::B_BYTES,
                    full_bar(sg));
        }
      }
      stamp();
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        if (hf == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        stamp();
        uint8_t* rw = rawbuf + hf * C::RAW_HALF;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int r = (hf * 4 + ii) * 16 + sub;
          float4* slot = reinterpret_cast<float4*>(rw + ((ii * 6) * PRODUCERS + tid) * 16);
          const float4 sr = slot[0 * PRODUCERS], sz = slot[1 * PRODUCERS], tn = slot[2 * PRODUCERS], nh = slot[3 * PRODUCERS],
                       hv = slot[4 * PRODUCERS], dv = slot[5 * PRODUCERS];
          const float m_ = *reinterpret_cast<const float*>(rw + C::RAW_ITEMS * PRODUCERS * 16 + (ii * PRODUCERS + tid) * 4);
          float4 o[6];
#define MPNN_GD(X)                                              \
  {                                                             \
    const float r_ = sr.X * m_, z_ = sz.X * m_, n_ = tn.X * m_; \
    const float go = dv.X * m_;                                 \
    const float dn = go * (1.f - z_);                           \
    const float dz = go * (hv.X - n_);                          \
    const float dan = dn * m_ * (1.f - tn.X * tn.X);            \
    const float dr = dan * nh.X;                                \
    o[0].X = dr * m_ * sr.X * (1.f - sr.X);                     \
    o[1].X = dz * m_ * sz.X * (1.f - sz.X);                     \
    o[2].X = dan;                                               \
    o[3].X = dan * r_;                                          \
    const float gz = go * z_;                                   \
    o[4].X = __uint_as_float(__float_as_uint(gz) & 0xffffe000u); \
    o[5].X = gz - o[4].X;                                       \
  }
          MPNN_GD(x) MPNN_GD(y) MPNN_GD(z) MPNN_GD(w)
#undef MPNN_GD
#pragma unroll
          for (int sg = 0; sg < C::NST; ++sg) {
            sts4(smem + sg * C::STAGE, swz(r, chunk), o[sg]);
            slot[sg * PRODUCERS] = o[3 + sg];
          }
        }
      }
      stamp();
      fence_proxy_async();
#pragma unroll 1
      for (int sg = 0; sg < C::NST; ++sg) mbar_arrive(full_bar(sg));
      phase ^= 1;
      // ---- pass 1: dnh, hi, lo -> stages 0..2; each raw buffer is refilled with the next group's half once read ----
#pragma unroll 1
      for (int sg = 0; sg < C::NST; ++sg) {
        mbar_wait(empty_bar(sg), phase ^ 1);
        if (tid == 0) {
          mbar_arrive_expect_tx(full_bar(sg), C::B_BYTES);
          bulk_copy(smem_base + sg * C::STAGE + C::A_BYTES, a.Bimg + (size_t)((3 + sg) * NKB + kb) * (DP * KB), C::B_BYTES,
                    full_bar(sg));
        }
      }
      stamp();
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        uint8_t* rw = rawbuf + hf * C::RAW_HALF;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int r = (hf * 4 + ii) * 16 + sub;
          const float4* slot = reinterpret_cast<const float4*>(rw + ((ii * 6) * PRODUCERS + tid) * 16);
#pragma unroll
          for (int sg = 0; sg < C::NST; ++sg) sts4(smem + sg * C::STAGE, swz(r, chunk), slot[sg * PRODUCERS]);
        }
        issue_raw(g + 1, hf);
      }
      fence_proxy_async();
#pragma unroll 1
      for (int sg = 0; sg < C::NST; ++sg) mbar_arrive(full_bar(sg));
      phase ^= 1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE, DP, 0, 0);
      int phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int i = t - t0, acc = i & 1, use = i >> 1;
        mbar_wait(accempty_bar(acc), (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t dtm = tmem_base + (uint32_t)(acc * DP);
        for (int kp = 0; kp < 2 * nkb; ++kp) {      // (column block, pass)
          for (int sg = 0; sg < C::NST; ++sg) {
            mbar_wait(full_bar(sg), phase);
            fence_proxy_async();
            tc_fence_after();
            const uint32_t sa = smem_base + sg * C::STAGE;
            const uint64_t ad = make_sdesc(sa, 16, 1024);
            const uint64_t bd = make_sdesc(sa + C::A_BYTES, 16, 1024);
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) umma_tf32(dtm, ad + 2u * j, bd + 2u * j, idesc, (kp | sg | j) != 0 ? 1u : 0u);
            umma_commit(empty_bar(sg));
          }
          phase ^= 1;
        }
        umma_commit(accfull_bar(acc));
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    float* tb = epi + q * (32 * 33);
    const int orow = lane >> 3, ocol = (lane & 7) * 4;
    for (int t = t0; t < t1; ++t) {
      const int i = t - t0, acc = i & 1, use = i >> 1;
      const long long pos = (long long)t * TILE + q * 32;
      mbar_wait(accfull_bar(acc), use & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < DP; c0 += 32) {
        if (c0 >= 2 * d) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * DP + c0), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) tb[lane * 33 + c] = v[c];
        __syncwarp();
        const int cc = c0 + ocol;
        if (cc < 2 * d) {
          float* out = cc < d ? a.dm + cc : a.dh + (cc - d);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const long long row = pos + it * 4 + orow;
            if (row < a.rows) {
              const float* p = tb + (it * 4 + orow) * 33 + ocol;
              *reinterpret_cast<float4*>(out + (size_t)row * d) = make_float4(p[0], p[1], p[2], p[3]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(accempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// dW_ih[k][g*d + c] = sum_cta partial[cta][k][g*64 + c];  dW_hh[k][g*d + c] = sum_cta partial[cta][64 + k][blk(g)*64 + c],
// blk = (0, 1, 3): fixed order over the CTAs
__global__ void __launch_bounds__(256) k_tc_gru_param_reduce(const float* __restrict__ partial, int n_cta, int d,
                                                             float* __restrict__ dW_ih, float* __restrict__ dW_hh) {
  const int per = 3 * d * d;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= 2 * per) return;
  const int which = e / per, rem = e - which * per;
  const int k = rem / (3 * d), j = rem - k * 3 * d;
  const int g = j / d, c = j - g * d;
  const int blk = (which == 1 && g == 2) ? 3 : g;
  const float* p = partial + (size_t)(which * 64 + k) * 256 + blk * 64 + c;
  float s = 0.f;
  int i = 0;
  for (; i + 4 <= n_cta; i += 4) {
    const float v0 = p[(size_t)i * 32768], v1 = p[(size_t)(i + 1) * 32768], v2 = p[(size_t)(i + 2) * 32768],
                v3 = p[(size_t)(i + 3) * 32768];
    s += v0;
    s += v1;
    s += v2;
    s += v3;
  }
  for (; i < n_cta; ++i) s += p[(size_t)i * 32768];
  (which ? dW_hh : dW_ih)[rem] = s;
}

// out[u*su + l*sl + k] (l < M, k < N) = sum over the CTAs whose tile range touches type u of partial[cta + u][l][k]
// (fixed order).  Plan mode: the tile range of a type comes from tile_off; dense mode: type u owns tiles [u*RT, (u+1)*RT).
__global__ void __launch_bounds__(256) k_tc_table_reduce(TcPlan plan, int dense, int RT, int grid_ctas, int ntypes,
                                                         int DP, int M, int N, long long su, long long sl,
                                                         const float* __restrict__ partial, float* __restrict__ out) {
  const int u = blockIdx.y;
  int n_tiles, f, l;
  if (dense) {
    n_tiles = ntypes * RT;
    f = u * RT;
    l = f + RT;
  } else {
    n_tiles = plan.head[0];
    f = plan.tile_off[u];
    l = (u < ntypes ? plan.tile_off[u + 1] : f);
  }
  const int per = (n_tiles + grid_ctas - 1) / grid_ctas;
  int c0 = 0, c1 = -1;
  if (l > f && per > 0 && f < n_tiles) {
    c0 = f / per;
    c1 = (min(l, n_tiles) - 1) / per;
  }
  const int elems = DP * DP;
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < elems; idx += gridDim.x * 256) {
    const int li = idx / DP, ki = idx - li * DP;
    if (li >= M || ki >= N) continue;
    float s = 0.f;
    for (int c = c0; c <= c1; ++c) s += partial[(size_t)(c + u) * elems + idx];
    out[(size_t)u * su + (size_t)li * sl + ki] = s;
  }
}

// image[(b*NKB + kb)][n][chunk ^ (n & 7)][j] = W[n*sn + k*sk + g*sg + s*ss], b = g*kseg + s, k = kb*32 + chunk*4 + j
// (zero outside N x K):
// packs nb strided weight blocks straight into the shared-memory image the dense GEMM bulk-copies per stage
__global__ void __launch_bounds__(256) k_tc_pack_image(const float* __restrict__ W, long long sn, long long sk,
                                                       long long sg, long long ss, int kseg, int nb, int N, int K,
                                                       int DP, float* __restrict__ img) {
  const int nkb = DP / KB;
  const long long total = (long long)nb * DP * DP;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int k = (int)(i % DP);
    const int n = (int)((i / DP) % DP);
    const int b = (int)(i / ((long long)DP * DP));
    const int g = b / kseg, sgm = b - g * kseg;
    const float v = (n < N && k < K) ? __ldg(W + (size_t)n * sn + (size_t)k * sk + (size_t)g * sg + (size_t)sgm * ss) : 0.f;
    const int kb = k >> 5, chunk = (k & 31) >> 2, j = k & 3;
    img[(((size_t)b * nkb + kb) * DP + n) * KB + ((chunk ^ (n & 7)) << 2) + j] = v;
  }
}

// Process-wide precision switch (mpnn_set_tensor_cores): with the tensor cores off every width is served by the fp32
// kernels (per-edge contraction, tile GEMMs), at fp32 accuracy and a fraction of the speed.
int g_tc_enabled = 1;
unsigned long long* g_tc_dbg = nullptr;

int tc_dp(int nf, int mf) {
  int d = nf > mf ? nf : mf;
  if (!g_tc_enabled || d <= 32 || d > 256 || (nf & 3) || (mf & 3)) return -1;
  return pow2_at_least(d, 64);
}
int tc_grid() { return mpnn_num_sms(); }
int tc_max_tiles(int edge_capacity, int ntypes) { return edge_capacity / TILE + ntypes + 1; }

struct PlanBuf {
  int* head;
  int* tile_off;
  int* tile_type;
  int* tile_pos;
  int* tile_cnt;
  int* psrc;
  int* pdst;
  float* palpha;
};
size_t plan_words(int edge_capacity, int ntypes) {
  const size_t mt = (size_t)tc_max_tiles(edge_capacity, ntypes);
  const size_t cap = (size_t)(edge_capacity > 0 ? edge_capacity : 1);
  return 4 + ((size_t)ntypes + 2) + 3 * mt + 3 * cap;
}
PlanBuf carve_plan(void* ws, int edge_capacity, int ntypes) {
  const size_t mt = (size_t)tc_max_tiles(edge_capacity, ntypes);
  const size_t cap = (size_t)(edge_capacity > 0 ? edge_capacity : 1);
  int* p = (int*)ws;
  PlanBuf b;
  b.head = p;
  b.tile_off = p + 4;
  b.tile_type = b.tile_off + ntypes + 2;
  b.tile_pos = b.tile_type + mt;
  b.tile_cnt = b.tile_pos + mt;
  b.psrc = b.tile_cnt + mt;
  b.pdst = b.psrc + cap;
  b.palpha = reinterpret_cast<float*>(b.pdst + cap);
  return b;
}
TcPlan as_plan(const PlanBuf& b) {
  return TcPlan{b.head, b.tile_off, b.tile_type, b.tile_pos, b.tile_cnt, b.psrc, b.pdst, b.palpha};
}

template <typename K>
int set_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? 0 : -1;
}

}  // namespace

extern "C" {

// padded feature width served by the tensor-core typed path (64, 128 or 256); -1 = not served
int mpnn_tc_dp(int nf, int mf) { return tc_dp(nf, mf); }

// 1 (default): widths 33..256 run on the tcgen05 kernels with TF32 operands; 0: fp32 kernels everywhere.  Returns the
// previous setting.
int mpnn_set_tensor_cores(int enabled) {
  const int prev = g_tc_enabled;
  g_tc_enabled = enabled ? 1 : 0;
  return prev;
}
int mpnn_tensor_cores_enabled(void) { return g_tc_enabled; }

// plan buffer: the type-sorted edge list cut into single-type tiles; built once per edge list
size_t mpnn_tc_plan_bytes(int edge_capacity, int unique_capacity) {
  return align_up(plan_words(edge_capacity, unique_capacity + 1) * sizeof(int), 256);
}

int mpnn_tc_plan(const int* type_ptr, const int* type_eid, const int* edge_src, const int* edge_dst,
                 const float* edge_w, int edge_capacity, int unique_capacity, void* plan, size_t plan_bytes_,
                 cudaStream_t stream) {
  const int ntypes = unique_capacity + 1;   // type_ptr has unique_capacity+1 entries: the last "type" is empty
  MPNN_REQUIRE(type_ptr && type_eid && edge_src && edge_dst && plan, MPNN_ERR_ARG, "tc_plan: null argument");
  MPNN_REQUIRE(plan_bytes_ >= mpnn_tc_plan_bytes(edge_capacity, unique_capacity), MPNN_ERR_WORKSPACE,
               "tc_plan: plan buffer too small");
  PlanBuf b = carve_plan(plan, edge_capacity, ntypes);
  // types 0..unique_capacity-1 own [type_ptr[u], type_ptr[u+1]); entry `unique_capacity` closes the last one
  const int max_tiles = tc_max_tiles(edge_capacity, ntypes);
  k_tc_plan_scan<<<1, 1024, 0, stream>>>(type_ptr, unique_capacity, max_tiles, b.tile_off, b.head);
  MPNN_CHECK_LAUNCH("k_tc_plan_scan");
  int fgrid = ceil_div(max_tiles, 256);
  if (fgrid > 4 * mpnn_num_sms()) fgrid = 4 * mpnn_num_sms();
  k_tc_plan_fill<<<fgrid, 256, 0, stream>>>(type_ptr, unique_capacity, b.tile_off, b.head, b.tile_type, b.tile_pos,
                                            b.tile_cnt);
  MPNN_CHECK_LAUNCH("k_tc_plan_fill");
  int ggrid = ceil_div(edge_capacity > 0 ? edge_capacity : 1, 256);
  if (ggrid > 8 * mpnn_num_sms()) ggrid = 8 * mpnn_num_sms();
  k_tc_plan_gather<<<ggrid, 256, 0, stream>>>(type_ptr, unique_capacity, edge_capacity, type_eid, edge_src, edge_dst,
                                              edge_w, b.psrc, b.pdst, b.palpha, b.head);
  MPNN_CHECK_LAUNCH("k_tc_plan_gather");
  return MPNN_OK;
}

size_t mpnn_tc_edge_gemm_workspace_bytes(int unique_capacity, int DP) {
  return (size_t)(unique_capacity + 1) * DP * DP * sizeof(float);
}

// Y[e, 0:N] = alpha_e * Bm[uid_e] (N x K, K contiguous, padded to DP x DP) . A[row_e, 0:K]
// forward (use_dst = 0): A = H, row_e = src_e, Bm = tableT, K = nf, N = mf;
// backward (use_dst = 1): A = dM, row_e = dst_e, Bm = table, K = mf, N = nf.
int mpnn_tc_edge_gemm(const void* plan, int edge_capacity, int unique_capacity, const int* type_eid, int use_dst,
                      const float* A, int lda, int K, const float* Bm, int DP, int use_alpha, float* Y, int ldy,
                      int N, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(plan && type_eid && A && Bm && Y && workspace, MPNN_ERR_ARG, "tc_edge_gemm: null argument");
  MPNN_REQUIRE((K & 3) == 0 && (N & 3) == 0 && (lda & 3) == 0 && (ldy & 3) == 0 && K <= DP && N <= DP,
               MPNN_ERR_UNSUPPORTED, "tc_edge_gemm: widths must be multiples of 4 and <= DP");
  MPNN_REQUIRE(DP == 64 || DP == 128 || DP == 256, MPNN_ERR_UNSUPPORTED,
               "tc_edge_gemm: DP must be 64, 128 or 256 (got %d)", DP);
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_edge_gemm_workspace_bytes(unique_capacity, DP), MPNN_ERR_WORKSPACE,
               "tc_edge_gemm: workspace too small");
  PlanBuf b = carve_plan(const_cast<void*>(plan), edge_capacity, unique_capacity + 1);
  float* img = (float*)workspace;
  {
    const long long total = (long long)(unique_capacity + 1) * DP * (DP / 4);
    int g = ceil_div(total, 256);
    if (g > 8 * mpnn_num_sms()) g = 8 * mpnn_num_sms();
    k_tc_swizzle_table<<<g, 256, 0, stream>>>(Bm, unique_capacity + 1, DP, img);
    MPNN_CHECK_LAUNCH("k_tc_swizzle_table");
  }
  TcGemm a;
  a.plan = as_plan(b);
  a.type_eid = type_eid;
  a.prow = use_dst ? b.pdst : b.psrc;
  a.A = A;
  a.Bimg = img;
  a.use_alpha = use_alpha;
  a.Y = Y;
  a.lda = lda;
  a.K = K;
  a.ldy = ldy;
  a.N = N;
  a.dense = 0;
  a.G = 1;
  a.rows = 0;
  a.kseg = 1;
  a.acol = 0;
  a.ycol = 0;
  a.accumulate = 0;
  a.nsplit = 0;
  a.bias = nullptr;
  const int grid = tc_grid();
  switch (DP) {
    case 64:
      MPNN_REQUIRE(set_smem(k_tc_edge_gemm<64>, GemmCfg<64>::SMEM) == 0, MPNN_ERR_CUDA, "tc_edge_gemm: smem attribute");
      k_tc_edge_gemm<64><<<grid, THREADS, GemmCfg<64>::SMEM, stream>>>(a);
      break;
    case 128:
      MPNN_REQUIRE(set_smem(k_tc_edge_gemm<128>, GemmCfg<128>::SMEM) == 0, MPNN_ERR_CUDA, "tc_edge_gemm: smem attribute");
      k_tc_edge_gemm<128><<<grid, THREADS, GemmCfg<128>::SMEM, stream>>>(a);
      break;
    default:
      MPNN_REQUIRE(set_smem(k_tc_edge_gemm<256>, GemmCfg<256>::SMEM) == 0, MPNN_ERR_CUDA, "tc_edge_gemm: smem attribute");
      k_tc_edge_gemm<256><<<grid, THREADS, GemmCfg<256>::SMEM, stream>>>(a);
      break;
  }
  MPNN_CHECK_LAUNCH("k_tc_edge_gemm");
  return MPNN_OK;
}

size_t mpnn_tc_table_grad_workspace_bytes(int unique_capacity, int DP) {
  return (size_t)(tc_grid() + unique_capacity + 2) * DP * DP * sizeof(float);
}

// dT [(unique_capacity+1)][DP][DP]:  dT[u][l][k] = sum_{e of type u} alpha_e H[src_e, l] dM[dst_e, k]
// (alpha = the edge weights the plan was built with if use_alpha, else 1)
int mpnn_tc_table_grad(const void* plan, int edge_capacity, int unique_capacity, const float* H, int nf,
                       const float* dM, int mf, int DP, int use_alpha, float* dT, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(plan && H && dM && dT && workspace, MPNN_ERR_ARG, "tc_table_grad: null argument");
  MPNN_REQUIRE((nf & 3) == 0 && (mf & 3) == 0 && nf <= DP && mf <= DP, MPNN_ERR_UNSUPPORTED,
               "tc_table_grad: widths must be multiples of 4 and <= DP");
  MPNN_REQUIRE(DP == 64 || DP == 128 || DP == 256, MPNN_ERR_UNSUPPORTED,
               "tc_table_grad: DP must be 64, 128 or 256 (got %d)", DP);
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_table_grad_workspace_bytes(unique_capacity, DP), MPNN_ERR_WORKSPACE,
               "tc_table_grad: workspace too small");
  PlanBuf b = carve_plan(const_cast<void*>(plan), edge_capacity, unique_capacity + 1);
  TcGrad a;
  a.plan = as_plan(b);
  a.H = H;
  a.dM = dM;
  a.partial = (float*)workspace;
  a.nf = nf;
  a.mf = mf;
  a.use_alpha = use_alpha;
  a.dense = 0;
  a.G = a.RT = 0;
  a.rows = 0;
  a.ldh = a.ldm = a.bcol = 0;
  const int grid = tc_grid();
  switch (DP) {
    case 64:
      MPNN_REQUIRE(set_smem(k_tc_table_grad<64>, GradCfg<64>::SMEM) == 0, MPNN_ERR_CUDA, "tc_table_grad: smem attribute");
      k_tc_table_grad<64><<<grid, THREADS, GradCfg<64>::SMEM, stream>>>(a);
      break;
    case 128:
      MPNN_REQUIRE(set_smem(k_tc_table_grad<128>, GradCfg<128>::SMEM) == 0, MPNN_ERR_CUDA, "tc_table_grad: smem attribute");
      k_tc_table_grad<128><<<grid, THREADS, GradCfg<128>::SMEM, stream>>>(a);
      break;
    default:
      MPNN_REQUIRE(set_smem(k_tc_table_grad<256>, GradCfg<256>::SMEM) == 0, MPNN_ERR_CUDA, "tc_table_grad: smem attribute");
      k_tc_table_grad<256><<<grid, THREADS, GradCfg<256>::SMEM, stream>>>(a);
      break;
  }
  MPNN_CHECK_LAUNCH("k_tc_table_grad");
  dim3 rgrid(ceil_div(DP * DP, 256 * 4), unique_capacity + 1);
  k_tc_table_reduce<<<rgrid, 256, 0, stream>>>(a.plan, 0, 0, grid, unique_capacity, DP, DP, DP, (long long)DP * DP, DP,
                                               a.partial, dT);
  MPNN_CHECK_LAUNCH("k_tc_table_reduce");
  return MPNN_OK;
}

// ---- dense GEMMs on the same kernels (GRU gates, readout projections at widths 33..256) ----------------------
// The GRU update (gru_update.py:27-28) and the readout projections (graph_level_output.py:36) are plain
// [rows, K] x [K, N] products on contiguous rows: the grouped-GEMM kernel with an identity plan.
size_t mpnn_tc_dense_workspace_bytes(int n_blocks, int DP) { return (size_t)n_blocks * DP * DP * sizeof(float); }

int mpnn_tc_dense_gemm_ll(const float* A, long long rows, int lda, int K, int kseg, int acol, const float* W,
                          long long w_sn, long long w_sk, long long w_sg, long long w_ss, int G, int N, const float* bias,
                          float* Y, int ldy, long long ycol, int nsplit, int accumulate, int DP, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream);

// Y[r, g*ycol + n] (+)= sum_{s < kseg} sum_{k < K} A[r, s*acol + k] * W[n*w_sn + k*w_sk + g*w_sg + s*w_ss] + bias[g*N + n]
// for g < G, n < N.  K, N <= DP in {64, 128, 256}; widths, strides of A / Y multiples of 4 floats.
int mpnn_tc_dense_gemm(const float* A, long long rows, int lda, int K, int kseg, int acol, const float* W,
                       long long w_sn, long long w_sk, long long w_sg, long long w_ss, int G, int N, const float* bias,
                       float* Y, int ldy, int ycol, int accumulate, int DP, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  return mpnn_tc_dense_gemm_ll(A, rows, lda, K, kseg, acol, W, w_sn, w_sk, w_sg, w_ss, G, N, bias, Y, ldy, (long long)ycol,
                               0, accumulate, DP, workspace, workspace_bytes, stream);
}

// the same with a 64-bit offset between output blocks that may be different buffers (mpnn_gru_bwd's dm / dh): either
// the G blocks, or -- nsplit > 0, G = 1 -- every nsplit columns of the one N-wide product
int mpnn_tc_dense_gemm_ll(const float* A, long long rows, int lda, int K, int kseg, int acol, const float* W,
                          long long w_sn, long long w_sk, long long w_sg, long long w_ss, int G, int N, const float* bias,
                          float* Y, int ldy, long long ycol, int nsplit, int accumulate, int DP, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(nsplit == 0 || (G == 1 && nsplit > 0 && (nsplit & 3) == 0), MPNN_ERR_ARG, "tc_dense_gemm: bad nsplit");
  MPNN_REQUIRE(A && W && Y && workspace && rows > 0 && G > 0 && kseg > 0, MPNN_ERR_ARG, "tc_dense_gemm: bad argument");
  MPNN_REQUIRE(DP == 64 || DP == 128 || DP == 256, MPNN_ERR_UNSUPPORTED, "tc_dense_gemm: DP must be 64, 128 or 256");
  MPNN_REQUIRE((K & 3) == 0 && (N & 3) == 0 && (lda & 3) == 0 && (ldy & 3) == 0 && (acol & 3) == 0 && (ycol & 3) == 0 &&
                   K <= DP && N <= DP,
               MPNN_ERR_UNSUPPORTED, "tc_dense_gemm: widths must be multiples of 4 and <= DP");
  MPNN_REQUIRE(rows < (1ll << 31) - TILE, MPNN_ERR_UNSUPPORTED, "tc_dense_gemm: too many rows");
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_dense_workspace_bytes(G * kseg, DP), MPNN_ERR_WORKSPACE,
               "tc_dense_gemm: workspace too small");
  float* img = (float*)workspace;
  {
    const long long total = (long long)G * kseg * DP * DP;
    int g = ceil_div(total, 256);
    if (g > 8 * mpnn_num_sms()) g = 8 * mpnn_num_sms();
    k_tc_pack_image<<<g, 256, 0, stream>>>(W, w_sn, w_sk, w_sg, w_ss, kseg, G * kseg, N, K, DP, img);
    MPNN_CHECK_LAUNCH("k_tc_pack_image");
  }
  TcGemm a;
  memset(&a, 0, sizeof(a));
  a.A = A;
  a.Bimg = img;
  a.Y = Y;
  a.lda = lda;
  a.K = K;
  a.ldy = ldy;
  a.N = N;
  a.dense = 1;
  a.G = G;
  a.rows = rows;
  a.kseg = kseg;
  a.acol = acol;
  a.ycol = ycol;
  a.accumulate = accumulate;
  a.nsplit = nsplit;
  a.bias = bias;
  const long long tiles = ((rows + TILE - 1) / TILE) * G;
  const int grid = (int)(tiles < tc_grid() ? tiles : tc_grid());
  switch (DP) {
    case 64:
      MPNN_REQUIRE(set_smem(k_tc_edge_gemm<64>, GemmCfg<64>::SMEM) == 0, MPNN_ERR_CUDA, "tc_dense_gemm: smem attribute");
      k_tc_edge_gemm<64><<<grid, THREADS, GemmCfg<64>::SMEM, stream>>>(a);
      break;
    case 128:
      MPNN_REQUIRE(set_smem(k_tc_edge_gemm<128>, GemmCfg<128>::SMEM) == 0, MPNN_ERR_CUDA, "tc_dense_gemm: smem attribute");
      k_tc_edge_gemm<128><<<grid, THREADS, GemmCfg<128>::SMEM, stream>>>(a);
      break;
    default:
      MPNN_REQUIRE(set_smem(k_tc_edge_gemm<256>, GemmCfg<256>::SMEM) == 0, MPNN_ERR_CUDA, "tc_dense_gemm: smem attribute");
      k_tc_edge_gemm<256><<<grid, THREADS, GemmCfg<256>::SMEM, stream>>>(a);
      break;
  }
  MPNN_CHECK_LAUNCH("k_tc_edge_gemm (dense)");
  return MPNN_OK;
}

size_t mpnn_tc_dense_grad_workspace_bytes(int G, int DP) {
  return (size_t)(tc_grid() + G + 2) * DP * DP * sizeof(float);
}

// out[g*o_sg + l*o_sl + k] = sum_r X[r, l] * D[r, g*dcol + k]   (l < M, k < N, g < G): the weight gradients X^T D
int mpnn_tc_dense_gemm_tn(const float* X, long long rows, int ldx, int M, const float* D, int ldd, int dcol, int G,
                          int N, int DP, float* out, long long o_sg, long long o_sl, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(X && D && out && workspace && rows > 0 && G > 0, MPNN_ERR_ARG, "tc_dense_gemm_tn: bad argument");
  MPNN_REQUIRE(DP == 64 || DP == 128 || DP == 256, MPNN_ERR_UNSUPPORTED, "tc_dense_gemm_tn: DP must be 64, 128 or 256");
  MPNN_REQUIRE((M & 3) == 0 && (N & 3) == 0 && (ldx & 3) == 0 && (ldd & 3) == 0 && (dcol & 3) == 0 && M <= DP && N <= DP,
               MPNN_ERR_UNSUPPORTED, "tc_dense_gemm_tn: widths must be multiples of 4 and <= DP");
  MPNN_REQUIRE(rows < (1ll << 31) - TILE, MPNN_ERR_UNSUPPORTED, "tc_dense_gemm_tn: too many rows");
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_dense_grad_workspace_bytes(G, DP), MPNN_ERR_WORKSPACE,
               "tc_dense_gemm_tn: workspace too small");
  TcGrad a;
  memset(&a, 0, sizeof(a));
  a.H = X;
  a.dM = D;
  a.partial = (float*)workspace;
  a.nf = M;
  a.mf = N;
  a.use_alpha = 0;
  a.dense = 1;
  a.G = G;
  a.RT = (int)((rows + TILE - 1) / TILE);
  a.rows = rows;
  a.ldh = ldx;
  a.ldm = ldd;
  a.bcol = dcol;
  const int grid = tc_grid();
  switch (DP) {
    case 64:
      MPNN_REQUIRE(set_smem(k_tc_table_grad<64>, GradCfg<64>::SMEM) == 0, MPNN_ERR_CUDA, "tc_dense_gemm_tn: smem attribute");
      k_tc_table_grad<64><<<grid, THREADS, GradCfg<64>::SMEM, stream>>>(a);
      break;
    case 128:
      MPNN_REQUIRE(set_smem(k_tc_table_grad<128>, GradCfg<128>::SMEM) == 0, MPNN_ERR_CUDA, "tc_dense_gemm_tn: smem attribute");
      k_tc_table_grad<128><<<grid, THREADS, GradCfg<128>::SMEM, stream>>>(a);
      break;
    default:
      MPNN_REQUIRE(set_smem(k_tc_table_grad<256>, GradCfg<256>::SMEM) == 0, MPNN_ERR_CUDA, "tc_dense_gemm_tn: smem attribute");
      k_tc_table_grad<256><<<grid, THREADS, GradCfg<256>::SMEM, stream>>>(a);
      break;
  }
  MPNN_CHECK_LAUNCH("k_tc_table_grad (dense)");
  dim3 rgrid(ceil_div(DP * DP, 256 * 4), G);
  k_tc_table_reduce<<<rgrid, 256, 0, stream>>>(a.plan, 1, a.RT, grid, G, DP, M, N, o_sg, o_sl, a.partial, out);
  MPNN_CHECK_LAUNCH("k_tc_table_reduce (dense)");
  return MPNN_OK;
}

// ---- nn.Linear-shaped helpers on the dense mode: W [N, K] row-major (out-features major), K, N up to 1024 -------
// The contraction and output widths are cut into equal blocks of <= 256 (multiples of 4).
static bool tc_split(int W, int* parts, int* width) {
  for (int p = 1; p <= 8; ++p) {
    if (W % p) continue;
    const int w = W / p;
    if (w <= 256 && (w & 3) == 0) {
      *parts = p;
      *width = w;
      return true;
    }
  }
  return false;
}
static int tc_linear_plan(int K, int N, int* kseg, int* Ks, int* G, int* Nb) {
  if (!g_tc_enabled || (K <= 32 && N <= 32) || !tc_split(K, kseg, Ks) || !tc_split(N, G, Nb)) return -1;
  const int m = *Ks > *Nb ? *Ks : *Nb;
  return pow2_at_least(m, 64);
}

// 1 if Y = X W^T (+ b) with these widths is served by the tensor-core dense mode
int mpnn_tc_linear_supported(int K, int N) {
  int a, b, c, d;
  return tc_linear_plan(K, N, &a, &b, &c, &d) > 0 ? 1 : 0;
}

size_t mpnn_tc_linear_workspace_bytes(int K, int N) {
  int kseg, Ks, G, Nb;
  const int DP = tc_linear_plan(K, N, &kseg, &Ks, &G, &Nb);
  if (DP < 0) return 0;
  const size_t img = align_up(mpnn_tc_dense_workspace_bytes(G * kseg, DP), 256);
  const int Gmax = G > kseg ? G : kseg;
  return img + mpnn_tc_dense_grad_workspace_bytes(Gmax, DP);
}

// Y[rows, N] (+)= X[rows, K] W^T + bias          (graph_level_output.py:36: the i / j projections)
int mpnn_tc_linear_fwd(const float* X, long long rows, int ldx, int K, const float* W, int N, const float* bias,
                       float* Y, int ldy, int accumulate, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  int kseg, Ks, G, Nb;
  const int DP = tc_linear_plan(K, N, &kseg, &Ks, &G, &Nb);
  MPNN_REQUIRE(DP > 0, MPNN_ERR_UNSUPPORTED, "tc_linear_fwd: widths K=%d N=%d not served", K, N);
  return mpnn_tc_dense_gemm(X, rows, ldx, Ks, kseg, Ks, W, K, 1, (long long)Nb * K, Ks, G, Nb, bias, Y, ldy, Nb, accumulate,
                            DP, workspace, workspace_bytes, stream);
}

// dX[rows, K] (+)= dY[rows, N] W
int mpnn_tc_linear_bwd_data(const float* dY, long long rows, int ldd, int N, const float* W, int K, float* dX, int ldx,
                            int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int kseg, Ks, G, Nb;
  const int DP = tc_linear_plan(K, N, &kseg, &Ks, &G, &Nb);
  MPNN_REQUIRE(DP > 0, MPNN_ERR_UNSUPPORTED, "tc_linear_bwd_data: widths K=%d N=%d not served", K, N);
  // contraction over the N out-features (G segments of Nb), output over the K in-features (kseg blocks of Ks)
  return mpnn_tc_dense_gemm(dY, rows, ldd, Nb, G, Nb, W, 1, K, Ks, (long long)Nb * K, kseg, Ks, nullptr, dX, ldx, Ks,
                            accumulate, DP, workspace, workspace_bytes, stream);
}

// dW[N, K] = dY^T X
int mpnn_tc_linear_bwd_weight(const float* dY, long long rows, int ldd, int N, const float* X, int ldx, int K,
                              float* dW, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int kseg, Ks, G, Nb;
  const int DP = tc_linear_plan(K, N, &kseg, &Ks, &G, &Nb);
  MPNN_REQUIRE(DP > 0, MPNN_ERR_UNSUPPORTED, "tc_linear_bwd_weight: widths K=%d N=%d not served", K, N);
  for (int mb = 0; mb < G; ++mb) {  // out-feature blocks: rows of dW
    int rc = mpnn_tc_dense_gemm_tn(dY + (size_t)mb * Nb, rows, ldd, Nb, X, ldx, Ks, kseg, Ks, DP,
                                   dW + (size_t)mb * Nb * K, Ks, K, workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  return MPNN_OK;
}

// ---- GRU weight gradients for widths <= 64 in one pass (k_tc_gru_param_grad) --------------------------------------
size_t mpnn_tc_gru_param_workspace_bytes(void) { return (size_t)tc_grid() * 128 * 256 * sizeof(float); }

// dg [rows, ldg]: the gate-gradient blocks dar | daz | dan | dnh (d columns each, as mpnn_gru_bwd lays them out)
int mpnn_tc_gru_param_grad(const float* m, const float* h, const float* dg, int ldg, long long rows, int d,
                           float* dW_ih, float* dW_hh, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(m && h && dg && dW_ih && dW_hh && workspace && rows > 0, MPNN_ERR_ARG, "tc_gru_param_grad: bad argument");
  MPNN_REQUIRE(d > 0 && d <= 64 && (d & 3) == 0 && (ldg & 3) == 0 && ldg >= 4 * d, MPNN_ERR_UNSUPPORTED,
               "tc_gru_param_grad: width %d not served", d);
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_gru_param_workspace_bytes(), MPNN_ERR_WORKSPACE,
               "tc_gru_param_grad: workspace too small");
  TcGruParam a = {m, h, dg, (float*)workspace, rows, d, ldg, nullptr, nullptr, nullptr, nullptr, nullptr};
  const int grid = tc_grid();
  MPNN_REQUIRE(set_smem(k_tc_gru_param_grad, GpCfg::SMEM) == 0, MPNN_ERR_CUDA, "tc_gru_param_grad: smem attribute");
  k_tc_gru_param_grad<<<grid, THREADS, GpCfg::SMEM, stream>>>(a);
  MPNN_CHECK_LAUNCH("k_tc_gru_param_grad");
  k_tc_gru_param_reduce<<<ceil_div(6 * d * d, 256), 256, 0, stream>>>(a.partial, grid, d, dW_ih, dW_hh);
  MPNN_CHECK_LAUNCH("k_tc_gru_param_reduce");
  return MPNN_OK;
}

// Pointwise GRU backward + weight gradients in one pass (widths <= 64): reads the saved gates [rows, 4d], m, h, dh', mask;
// writes dg [rows, 6d] (dar | daz | dan | dnh | hi(go z) | lo(go z): the operand of the data product), bias partials
// [mpnn_tc_gru_param_bias_parts()][4d] (sum them in order for db_ih / db_hh) and dW_ih, dW_hh [d, 3d].
int mpnn_tc_gru_param_bias_parts(void) { return tc_grid() * 16; }

int mpnn_tc_gru_param_point(const float* m, const float* h, const float* mask, const float* gates, const float* dh_out,
                            long long rows, int d, float* dg, float* bias_part, float* dW_ih, float* dW_hh,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MPNN_REQUIRE(m && h && mask && gates && dh_out && bias_part && dW_ih && dW_hh && workspace && rows > 0,
               MPNN_ERR_ARG, "tc_gru_param_point: bad argument");
  MPNN_REQUIRE(d > 0 && d <= 64 && (d & 3) == 0, MPNN_ERR_UNSUPPORTED, "tc_gru_param_point: width %d not served", d);
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_gru_param_workspace_bytes(), MPNN_ERR_WORKSPACE,
               "tc_gru_param_point: workspace too small");
  TcGruParam a = {m, h, nullptr, (float*)workspace, rows, d, 6 * d, gates, dh_out, mask, dg, bias_part};
  const int grid = tc_grid();
  MPNN_REQUIRE(set_smem(k_tc_gru_param_point, GppCfg::SMEM) == 0, MPNN_ERR_CUDA, "tc_gru_param_point: smem attribute");
  k_tc_gru_param_point<<<grid, THREADS, GppCfg::SMEM, stream>>>(a);
  MPNN_CHECK_LAUNCH("k_tc_gru_param_point");
  k_tc_gru_param_reduce<<<ceil_div(6 * d * d, 256), 256, 0, stream>>>(a.partial, grid, d, dW_ih, dW_hh);
  MPNN_CHECK_LAUNCH("k_tc_gru_param_reduce");
  return MPNN_OK;
}

// profiling aid: DEVICE buffer of 64 x uint64 that receives %globaltimer stamps of k_tc_gru_data_grad's producer (CTA 0,
// thread 0; seven per group: start, stages free, half 0 landed, half 1 landed, converted, pass-1 stages free, ...); NULL = off
void mpnn_tc_debug(unsigned long long* buf) { g_tc_dbg = buf; }

// GRU data gradients for widths <= 64 straight from the saved gates (k_tc_gru_data_grad): dm, dh [rows, d].
// Wc: the combined weights [6][2d][d] (rows [0,d) of a block -> dm, [d,2d) -> dh; blocks = dar | daz | dan | dnh | I | I).
size_t mpnn_tc_gru_data_workspace_bytes(void) { return (size_t)6 * 128 * 128 * sizeof(float); }

int mpnn_tc_gru_data_grad(const float* gates, const float* h, const float* dh_out, const float* mask, const float* Wc,
                          long long rows, int d, float* dm, float* dh, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream) {
  MPNN_REQUIRE(gates && h && dh_out && mask && Wc && dm && dh && workspace && rows > 0, MPNN_ERR_ARG,
               "tc_gru_data_grad: bad argument");
  MPNN_REQUIRE(d > 0 && d <= 64 && (d & 3) == 0, MPNN_ERR_UNSUPPORTED, "tc_gru_data_grad: width %d not served", d);
  MPNN_REQUIRE(rows < (1ll << 31) - TILE, MPNN_ERR_UNSUPPORTED, "tc_gru_data_grad: too many rows");
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_gru_data_workspace_bytes(), MPNN_ERR_WORKSPACE,
               "tc_gru_data_grad: workspace too small");
  float* img = (float*)workspace;
  k_tc_pack_image<<<ceil_div(6LL * 128 * 128, 256), 256, 0, stream>>>(Wc, d, 1, 0, (long long)2 * d * d, 6, 6, 2 * d, d, 128,
                                                                     img);
  MPNN_CHECK_LAUNCH("k_tc_pack_image");
  TcGruData a = {gates, h, dh_out, mask, img, dm, dh, rows, d, g_tc_dbg};
  const long long tiles = (rows + TILE - 1) / TILE;
  const int grid = (int)(tiles < tc_grid() ? tiles : tc_grid());
  MPNN_REQUIRE(set_smem(k_tc_gru_data_grad, GdCfg::SMEM) == 0, MPNN_ERR_CUDA, "tc_gru_data_grad: smem attribute");
  k_tc_gru_data_grad<<<grid, THREADS, GdCfg::SMEM, stream>>>(a);
  MPNN_CHECK_LAUNCH("k_tc_gru_data_grad");
  return MPNN_OK;
}

// ---- fused masked GRU forward on the tensor cores (widths 33..256, multiples of 4) -------------------------------
int mpnn_tc_gru_supported(int d) { return (g_tc_enabled && d > 32 && d <= 256 && (d & 3) == 0) ? 1 : 0; }

static void tc_gru_dims(int d, int* DP, int* KP, int* ncb) {
  *KP = pow2_at_least(d, 64);
  *DP = *KP > 128 ? 128 : *KP;      // four accumulators of DP columns must fit in the 512 TMEM columns
  *ncb = *KP / *DP;
}

size_t mpnn_tc_gru_workspace_bytes(int d) {
  if (!mpnn_tc_gru_supported(d)) return 0;
  int DP, KP, ncb;
  tc_gru_dims(d, &DP, &KP, &ncb);
  return (size_t)ncb * 2 * 3 * DP * KP * sizeof(float);
}

int mpnn_tc_gru_fwd_agg(const float* Y, const int* row_ptr, const float* m, const float* h, const float* mask,
                        const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, long long rows,
                        int d, float* m_out, float* h_out, float* gates, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream);

// h_out [rows, d], gates [rows, 4d] (sigmoid r | sigmoid z | tanh n | nh: what mpnn_gru_bwd reads)
int mpnn_tc_gru_fwd(const float* m, const float* h, const float* mask, const float* W_ih, const float* W_hh,
                    const float* b_ih, const float* b_hh, long long rows, int d, float* h_out, float* gates,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return mpnn_tc_gru_fwd_agg(nullptr, nullptr, m, h, mask, W_ih, W_hh, b_ih, b_hh, rows, d, nullptr, h_out, gates,
                             workspace, workspace_bytes, stream);
}

// The same with the aggregation folded in (Y, row_ptr non-NULL): message row i = sum of Y[row_ptr[i] : row_ptr[i+1]]
// in edge order (the fixed-order CSR sum of mpnn_segment_sum), computed by the A-tile producer while staging; the sums
// are also written to m_out [rows, d] (what the backward reads as `m`).
int mpnn_tc_gru_fwd_agg(const float* Y, const int* row_ptr, const float* m, const float* h, const float* mask,
                        const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, long long rows,
                        int d, float* m_out, float* h_out, float* gates, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream) {
  MPNN_REQUIRE((Y != nullptr) == (row_ptr != nullptr) && (Y || m), MPNN_ERR_ARG, "tc_gru_fwd: message operand missing");
  MPNN_REQUIRE(mpnn_tc_gru_supported(d), MPNN_ERR_UNSUPPORTED, "tc_gru_fwd: width %d not served", d);
  MPNN_REQUIRE(rows > 0 && rows < (1ll << 30), MPNN_ERR_ARG, "tc_gru_fwd: bad row count");
  MPNN_REQUIRE(workspace_bytes >= mpnn_tc_gru_workspace_bytes(d), MPNN_ERR_WORKSPACE, "tc_gru_fwd: workspace too small");
  int DP, KP, ncb;
  tc_gru_dims(d, &DP, &KP, &ncb);
  float* img = (float*)workspace;
  k_tc_gru_pack<<<ceil_div((long long)ncb * 6 * DP * KP, 256), 256, 0, stream>>>(W_ih, W_hh, d, DP, KP, ncb, img);
  MPNN_CHECK_LAUNCH("k_tc_gru_pack");
  TcGru a = {m, h, mask, img, b_ih, b_hh, h_out, gates, rows, d, ncb, Y, row_ptr, m_out};
  const long long tiles = ((rows + TILE - 1) / TILE) * ncb;
  const int grid = (int)(tiles < tc_grid() ? tiles : tc_grid());
  if (KP == 64) {
    MPNN_REQUIRE(set_smem(k_tc_gru_fwd<64, 64>, GruCfg<64, 64>::SMEM) == 0, MPNN_ERR_CUDA, "tc_gru_fwd: smem attribute");
    k_tc_gru_fwd<64, 64><<<grid, GruCfg<64, 64>::GRU_THREADS, GruCfg<64, 64>::SMEM, stream>>>(a);
  } else if (KP == 128) {
    MPNN_REQUIRE(set_smem(k_tc_gru_fwd<128, 128>, GruCfg<128, 128>::SMEM) == 0, MPNN_ERR_CUDA, "tc_gru_fwd: smem attribute");
    k_tc_gru_fwd<128, 128><<<grid, GruCfg<128, 128>::GRU_THREADS, GruCfg<128, 128>::SMEM, stream>>>(a);
  } else {
    MPNN_REQUIRE(set_smem(k_tc_gru_fwd<128, 256>, GruCfg<128, 256>::SMEM) == 0, MPNN_ERR_CUDA, "tc_gru_fwd: smem attribute");
    k_tc_gru_fwd<128, 256><<<grid, GruCfg<128, 256>::GRU_THREADS, GruCfg<128, 256>::SMEM, stream>>>(a);
  }
  MPNN_CHECK_LAUNCH("k_tc_gru_fwd");
  return MPNN_OK;
}

}  // extern "C"
