// Adam over a list of small parameter tensors as ONE launch ("fused multi-tensor Adam").
// The reference drivers train with torch.optim.Adam (test_lipo.py:138-139); its fused multi-tensor kernel gives one
// CTA a whole 64 K chunk of a tensor, so the ~100 K parameters of an edge-network model take 17 us on 27 CTAs.  Here
// the tensors are laid end to end in an index space cut into 1 K-element blocks: ~100 CTAs, 4 us.
// Arithmetic = torch.optim.Adam (amsgrad off, maximize off, L2-style weight decay):
//   g += wd p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// The step count lives on the device (CUDA-graph replay): every CTA reads it on entry, the last CTA to finish writes
// t+1 back (ticket counter, self-resetting) -- no separate increment launch.
#include "common.cuh"

namespace {

constexpr int ADAM_MAXT = 72;        // tensors per launch (pointer table travels as a kernel argument)
constexpr int ADAM_BLOCK = 1024;     // elements per CTA (256 threads x 4)

struct AdamPack {
  float* p[ADAM_MAXT];
  const float* g[ADAM_MAXT];
  float* m[ADAM_MAXT];
  float* v[ADAM_MAXT];
  int first_block[ADAM_MAXT + 1];    // CTA range of tensor i
  int numel[ADAM_MAXT];
  int n;
  const int* guard[8];               // overflow flags of the step's edge lists (counts[2] of each): any set -> no update
  int n_guard;
};

__global__ void __launch_bounds__(256) k_adam(AdamPack pk, float* __restrict__ step, unsigned int* __restrict__ ticket,
                                              int bump, float lr, float b1, float b2, float eps, float wd) {
  for (int q = 0; q < pk.n_guard; ++q)
    if (pk.guard[q][2] != 0) return;   // a batch overflowed the captured capacities: its gradients are not applied
  const float t = step[0] + 1.f;
  int i = 0;
  while (i + 1 < pk.n && (int)blockIdx.x >= pk.first_block[i + 1]) ++i;
  const int base = ((int)blockIdx.x - pk.first_block[i]) * ADAM_BLOCK;
  const float bc1 = 1.f - powf(b1, t);
  const float bc2s = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  float* __restrict__ P = pk.p[i];
  const float* __restrict__ G = pk.g[i];
  float* __restrict__ M = pk.m[i];
  float* __restrict__ V = pk.v[i];
  const int n = pk.numel[i];
  float pv[4], gv[4], mv[4], vv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = base + k * 256 + threadIdx.x;
    const bool in = e < n;
    pv[k] = in ? P[e] : 0.f;
    gv[k] = in ? G[e] : 0.f;
    mv[k] = in ? M[e] : 0.f;
    vv[k] = in ? V[e] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = base + k * 256 + threadIdx.x;
    if (e < n) {
      const float g = wd != 0.f ? fmaf(wd, pv[k], gv[k]) : gv[k];
      const float m = b1 * mv[k] + (1.f - b1) * g;
      const float v = b2 * vv[k] + (1.f - b2) * g * g;
      const float denom = sqrtf(v) / bc2s + eps;
      M[e] = m;
      V[e] = v;
      P[e] = pv[k] - step_size * (m / denom);
    }
  }
  if (!bump) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {   // every CTA has read step[0] before it took a ticket
      step[0] = t;
      *ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Data-parallel form: gradient all-reduce FUSED into the optimizer step over NVLink peer memory (the path's
// only collective is the parameter-gradient sum).  Every rank owns a symmetric buffer (torch symmetric memory: the same
// allocation mapped into every peer) holding two gradient regions (alternating by step parity) and a flag array.
// CTA c of rank r, for its 1 K-element chunk:
//   1. packs its slice of the local gradients into region (t & 1) of the local buffer,
//   2. publishes flag[r][c] = t in EVERY peer's buffer (system-scope release),
//   3. waits until every peer's flag[.][c] in the LOCAL buffer has reached t,
//   4. sums the chunk over the ranks in rank order (identical on every rank, bit-reproducible), scales by 1/W, applies Adam.
// Chunks are independent: no grid barrier, no separate all-reduce / scale / unpack launches, and the wire traffic of a
// chunk overlaps the arithmetic of the others.  Two regions make the write-after-read hazard impossible: a peer can only
// start overwriting region (t & 1) at step t + 2, which needs this rank's flags of step t + 1, i.e. this step finished.
struct AdamDdp {
  float* flat[8];        // every rank's symmetric buffer (peer-mapped addresses), [2][region_floats] gradients
  unsigned* flags[8];    // every rank's flag array inside its buffer: [world][n_chunks]
  long long goff[ADAM_MAXT];   // offset of tensor i inside a region
  long long region_floats;
  int world, rank, n_chunks;
  float scale;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) k_adam_ddp(AdamPack pk, AdamDdp dd, float* __restrict__ step,
                                                  unsigned int* __restrict__ ticket, float lr, float b1, float b2,
                                                  float eps, float wd) {
  const float t = step[0] + 1.f;
  const unsigned epoch = (unsigned)t;
  int i = 0;
  while (i + 1 < pk.n && (int)blockIdx.x >= pk.first_block[i + 1]) ++i;
  const int base = ((int)blockIdx.x - pk.first_block[i]) * ADAM_BLOCK;
  const int n = pk.numel[i];
  const long long roff = (long long)(epoch & 1u) * dd.region_floats + dd.goff[i];
  const float* __restrict__ G = pk.g[i];
  // 1. pack
  float* mine = dd.flat[dd.rank] + roff;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = base + k * 256 + threadIdx.x;
    if (e < n) mine[e] = G[e];
  }
  __syncthreads();
  // 2. publish, 3. wait
  if (threadIdx.x < dd.world && threadIdx.x != dd.rank) {
    __threadfence_system();
    st_release_sys(dd.flags[threadIdx.x] + (size_t)dd.rank * dd.n_chunks + blockIdx.x, epoch);
  }
  if (threadIdx.x < dd.world && threadIdx.x != dd.rank) {
    const unsigned* f = dd.flags[dd.rank] + (size_t)threadIdx.x * dd.n_chunks + blockIdx.x;
    while (ld_acquire_sys(f) < epoch) __nanosleep(100);
  }
  __syncthreads();
  // 4. reduce in rank order + Adam
  const float bc1 = 1.f - powf(b1, t);
  const float bc2s = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  float* __restrict__ P = pk.p[i];
  float* __restrict__ M = pk.m[i];
  float* __restrict__ V = pk.v[i];
  float gv[4], pv[4], mv[4], vv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) gv[k] = 0.f;
  for (int r = 0; r < dd.world; ++r) {
    const float* src = dd.flat[r] + roff;
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = base + k * 256 + threadIdx.x;
      x[k] = e < n ? __ldcv(src + e) : 0.f;   // peer memory: never from a stale cache line
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) gv[k] += x[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = base + k * 256 + threadIdx.x;
    const bool in = e < n;
    pv[k] = in ? P[e] : 0.f;
    mv[k] = in ? M[e] : 0.f;
    vv[k] = in ? V[e] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = base + k * 256 + threadIdx.x;
    if (e < n) {
      float g = gv[k] * dd.scale;
      if (wd != 0.f) g = fmaf(wd, pv[k], g);
      const float m = b1 * mv[k] + (1.f - b1) * g;
      const float v = b2 * vv[k] + (1.f - b2) * g * g;
      const float denom = sqrtf(v) / bc2s + eps;
      M[e] = m;
      V[e] = v;
      P[e] = pv[k] - step_size * (m / denom);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      step[0] = t;
      *ticket = 0u;
    }
  }
}

}  // namespace

extern "C" {

// n tensors: params[i], grads[i], exp_avg[i], exp_avg_sq[i] (device pointers, numel[i] floats each; host arrays).
// step: device float (number of steps taken so far; incremented by this call); ticket: device uint32, zero on first use.
// guards: HOST array of n_guards (<= 8) device pointers to `counts` arrays of the step's edge lists: when any counts[2]
// (capacity overflow of a captured step) is set, the whole update is skipped on the device.
int mpnn_adam_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const long long* numel, float* step, unsigned int* ticket, float lr,
                   float beta1, float beta2, float eps, float weight_decay, const int* const* guards, int n_guards,
                   cudaStream_t stream) {
  MPNN_REQUIRE(n >= 0 && step && ticket && n_guards >= 0 && n_guards <= 8, MPNN_ERR_ARG, "adam_step: bad arguments");
  int done = 0;
  while (done < n) {
    AdamPack pk;
    memset(&pk, 0, sizeof(pk));
    int blocks = 0, k = 0;
    while (done + k < n && k < ADAM_MAXT) {
      const long long ne = numel[done + k];
      MPNN_REQUIRE(ne > 0 && ne < (1ll << 31), MPNN_ERR_ARG, "adam_step: tensor %d has %lld elements", done + k, ne);
      pk.p[k] = params[done + k];
      pk.g[k] = grads[done + k];
      pk.m[k] = exp_avg[done + k];
      pk.v[k] = exp_avg_sq[done + k];
      pk.numel[k] = (int)ne;
      pk.first_block[k] = blocks;
      blocks += ceil_div(ne, ADAM_BLOCK);
      ++k;
    }
    pk.first_block[k] = blocks;
    pk.n = k;
    pk.n_guard = n_guards;
    for (int q = 0; q < n_guards; ++q) pk.guard[q] = guards[q];
    done += k;
    k_adam<<<blocks, 256, 0, stream>>>(pk, step, ticket, done == n ? 1 : 0, lr, beta1, beta2, eps, weight_decay);
    MPNN_CHECK_LAUNCH("k_adam");
  }
  return MPNN_OK;
}

// Data-parallel Adam: gradient sum over `world` ranks fused into the step (see k_adam_ddp).  flat / flags: HOST arrays of
// `world` peer-mapped device pointers (this rank's own buffer at index `rank`); every rank's buffer holds
// [2][region_floats] floats followed by its flag array [world][n_chunks] (zero on first use); goff[i] = offset of tensor i
// inside a region (the same on every rank).  All n tensors must fit one launch (n <= 72).  Returns the number of chunks
// (CTAs) when called with params == NULL (sizing query).
int mpnn_adam_step_ddp(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const long long* numel, const long long* goff, float* step,
                       unsigned int* ticket, float lr, float beta1, float beta2, float eps, float weight_decay,
                       float* const* flat, unsigned* const* flags, long long region_floats, int world, int rank,
                       cudaStream_t stream) {
  MPNN_REQUIRE(n >= 1 && n <= ADAM_MAXT && world >= 1 && world <= 8 && rank >= 0 && rank < world, MPNN_ERR_ARG,
               "adam_step_ddp: bad arguments (n %d, world %d)", n, world);
  AdamPack pk;
  AdamDdp dd;
  memset(&pk, 0, sizeof(pk));
  memset(&dd, 0, sizeof(dd));
  int blocks = 0;
  for (int k = 0; k < n; ++k) {
    const long long ne = numel[k];
    MPNN_REQUIRE(ne > 0 && ne < (1ll << 31), MPNN_ERR_ARG, "adam_step_ddp: tensor %d has %lld elements", k, ne);
    pk.numel[k] = (int)ne;
    pk.first_block[k] = blocks;
    blocks += ceil_div(ne, ADAM_BLOCK);
    if (params) {
      pk.p[k] = params[k];
      pk.g[k] = grads[k];
      pk.m[k] = exp_avg[k];
      pk.v[k] = exp_avg_sq[k];
      dd.goff[k] = goff[k];
    }
  }
  pk.first_block[n] = blocks;
  pk.n = n;
  if (!params) return blocks;
  MPNN_REQUIRE(step && ticket && flat && flags, MPNN_ERR_ARG, "adam_step_ddp: null buffers");
  for (int r = 0; r < world; ++r) {
    dd.flat[r] = flat[r];
    dd.flags[r] = flags[r];
  }
  dd.region_floats = region_floats;
  dd.world = world;
  dd.rank = rank;
  dd.n_chunks = blocks;
  dd.scale = 1.f / (float)world;
  k_adam_ddp<<<blocks, 256, 0, stream>>>(pk, dd, step, ticket, lr, beta1, beta2, eps, weight_decay);
  MPNN_CHECK_LAUNCH("k_adam_ddp");
  return MPNN_OK;
}

}  // extern "C"
