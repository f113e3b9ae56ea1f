// Microbenchmark (development aid): TMA box-load throughput per SM when EVERY SM is loading the same L2-resident matrix
// (the situation of K1B / K11 / K3F operand streams), as a function of the box row width.  One CTA per SM; thread 0
// keeps `inflight` boxes outstanding.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate_all tma_rate_all.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../flexdiffuse_b200/csrc/fd_common.cuh"
using namespace fd;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int box_bytes, int rows, int inner_elems,
                                            int k_extent, int inflight, int iters, int issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4][6];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 24; ++i) mbar_init(&bars[0][0] + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // `issuers` warps (<= 4) each keep their own stream of boxes; total smem slots shared: 6 slots split between them
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    uint64_t* bar = bars[w];
    const int nslot = 6 / issuers;   // slots of 32 KB per issuer
    long long t0 = clock64();
    for (int it = 0; it < iters + inflight; ++it) {
      if (it >= inflight) mbar_wait(&bar[(it - inflight) % nslot], ((it - inflight) / nslot) & 1);
      if (it < iters) {
        mbar_expect_tx(&bar[it % nslot], box_bytes * rows);
        tma_load_2d(smem + (w * nslot + it % nslot) * 32768, &tm, &bar[it % nslot], (it * inner_elems) % k_extent, 0);
      }
    }
    if (w == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  void* g;
  cudaMalloc(&g, 257 * 768 * 4);
  cudaMemset(g, 0, 257 * 768 * 4);
  long long* d;
  cudaMalloc(&d, 8 * 148);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768 + 1024);
  struct Cfg { int box_bytes, rows, elem; CUtensorMapSwizzle sw; const char* name; };
  Cfg cfgs[] = {{64, 256, 2, CU_TENSOR_MAP_SWIZZLE_64B, "fp16 64B x256 SW64"},   {128, 256, 2, CU_TENSOR_MAP_SWIZZLE_128B, "fp16 128B x256 SW128"},
                {128, 128, 2, CU_TENSOR_MAP_SWIZZLE_128B, "fp16 128B x128 SW128"}, {128, 80, 4, CU_TENSOR_MAP_SWIZZLE_NONE, "fp32 128B x80 none"},
                {256, 80, 4, CU_TENSOR_MAP_SWIZZLE_NONE, "fp32 256B x80 none"},   {32, 256, 2, CU_TENSOR_MAP_SWIZZLE_32B, "fp16 32B x256 SW32"}};
  for (auto& c : cfgs) {
    CUtensorMap tm;
    const int kext = 768;
    uint64_t dims[2] = {(uint64_t)kext, 256};
    uint64_t strides[1] = {(uint64_t)kext * c.elem};
    uint32_t box[2] = {(uint32_t)(c.box_bytes / c.elem), (uint32_t)c.rows};
    if (encode_tmap(&tm, c.elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, c.sw) != 0) {
      printf("encode failed: %s\n", fd_last_error_string());
      return 1;
    }
    for (int grid : {148})
      for (int issuers : {1, 2, 3})
      for (int inflight : {2}) {
        const int iters = 192;
        for (int rep = 0; rep < 2; ++rep) { k<<<grid, 128, 6 * 32768 + 1024>>>(tm, c.box_bytes, c.rows, c.box_bytes / c.elem, kext, inflight, iters, issuers, d); cudaDeviceSynchronize(); }
        long long h[148]; cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-22s grid %3d issuers %d inflight %d each: %7.1f cycles per box per issuer, %5.1f B/clk per SM  (%s)\n", c.name, grid, issuers, inflight,
               double(mx) / iters, double(c.box_bytes) * c.rows * iters * issuers / mx, cudaGetErrorString(cudaGetLastError()));
      }
  }
  return 0;
}
