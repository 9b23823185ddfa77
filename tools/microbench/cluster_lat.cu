// Launch + barrier cost of thread-block clusters (decides whether small batch-coupled ops belong in one cluster
// launch or in two grid launches).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_lat cluster_lat.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_plain(float* out, int nsync) {
  extern __shared__ float sm[];
  sm[threadIdx.x] = threadIdx.x;
  for (int i = 0; i < nsync; ++i) __syncthreads();
  if (threadIdx.x == 0 && sm[1] < 0) out[blockIdx.x] = sm[0];
}
__global__ void k_cluster(float* out, int nsync, int gather) {
  extern __shared__ float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  sm[threadIdx.x] = threadIdx.x;
  float acc = 0.f;
  for (int i = 0; i < nsync; ++i) {
    cluster.sync();
    if (gather && threadIdx.x < 64)
      for (unsigned k = 0; k < cluster.num_blocks(); ++k) acc += cluster.map_shared_rank(sm, k)[threadIdx.x];
  }
  if (nsync) cluster.sync();
  if (threadIdx.x == 0 && acc < -1.f) out[blockIdx.x] = acc;
}

template <typename F>
float time_us(F launch, int iters, cudaStream_t st) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 20; ++i) launch();
  cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int i = 0; i < iters; ++i) launch();
  cudaEventRecord(b, st);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms * 1000.f / iters;
}

int main() {
  float* out;
  cudaMalloc(&out, 4096);
  cudaStream_t st;
  cudaStreamCreate(&st);
  const int iters = 2000;
  for (int threads : {256, 1024}) {
    for (size_t smem : {(size_t)8 * 1024, (size_t)100 * 1024}) {
      cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      printf("threads=%d smem=%zuKB\n", threads, smem / 1024);
      printf("  plain   8 CTAs                     %.2f us\n", time_us([&] { k_plain<<<8, threads, smem, st>>>(out, 3); }, iters, st));
      for (int nc : {2, 8}) {
        for (int nsync : {0, 1, 3}) {
          for (int gather : {0, 1}) {
            if (gather && !nsync) continue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nc);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = nc;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            float t = time_us([&] { cudaLaunchKernelEx(&cfg, k_cluster, out, nsync, gather); }, iters, st);
            printf("  cluster %d CTAs syncs=%d gather=%d     %.2f us\n", nc, nsync, gather, t);
          }
        }
      }
    }
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
