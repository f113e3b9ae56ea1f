// What do shared::cta and shared::cluster (mapa) addresses look like inside a cluster of 6?  (Is the CTA rank encoded in the
// 32-bit shared address, as CUTLASS' Sm100MmaPeerBitMask = 0xFEFFFFFF implies?)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k() {
  __shared__ uint64_t bar[4];
  uint32_t rank, local = static_cast<uint32_t>(__cvta_generic_to_shared(&bar[1]));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0 && blockIdx.x < 6) {
    uint32_t m[6];
    for (uint32_t r = 0; r < 6; ++r) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(m[r]) : "r"(local), "r"(r));
    printf("block %d rank %u local %08x mapa: %08x %08x %08x %08x %08x %08x\n", blockIdx.x, rank, local, m[0], m[1], m[2], m[3],
           m[4], m[5]);
  }
}
int main() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(12);
  cfg.blockDim = dim3(32);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 6; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k);
  printf("launch: %s\n", cudaGetErrorString(e));
  printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
