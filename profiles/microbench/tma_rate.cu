// Microbenchmark (development aid): cycles for one TMA 2-D box load {inner bytes, 256 rows} from an
// L2-resident bf16 matrix into shared memory, as a function of the inner box width and the number
// of loads kept in flight.  One CTA; thread 0 issues and waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../flexdiffuse_b200/csrc/fd_common.cuh"
using namespace fd;

__global__ void __launch_bounds__(128, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int box_bytes,
                                                          int inflight, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[6];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int inner_elems = box_bytes / 2;
    long long t0 = clock64();
    for (int it = 0; it < iters + inflight; ++it) {
      if (it >= inflight) mbar_wait(&bar[(it - inflight) % 6], ((it - inflight) / 6) & 1);
      if (it < iters) {
        mbar_expect_tx(&bar[it % 6], box_bytes * 256);
        tma_load_2d(smem + (it % 6) * 32768, &tm, &bar[it % 6], (it * inner_elems) % 768, 0);
      }
    }
    out[0] = clock64() - t0;
  }
}

int main() {
  __nv_bfloat16* g;
  cudaMalloc(&g, 257 * 768 * 2);
  cudaMemset(g, 0, 257 * 768 * 2);
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768 + 1024);
  for (int box_bytes : {32, 64, 128}) {
    CUtensorMap tm;
    uint64_t dims[2] = {768, 256};
    uint64_t strides[1] = {768 * 2};
    uint32_t box[2] = {static_cast<uint32_t>(box_bytes / 2), 256};
    CUtensorMapSwizzle sw = box_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (encode_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, strides, box, sw) != 0) { printf("encode failed: %s\n", fd_last_error_string()); return 1; }
    for (int inflight : {1, 2, 4, 6}) {
      const int iters = 96;
      for (int rep = 0; rep < 2; ++rep) { tma_rate_kernel<<<1, 128, 6 * 32768 + 1024>>>(tm, box_bytes, inflight, iters, d); cudaDeviceSynchronize(); }
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("box {%3d B x 256 rows} = %5d B, %d in flight: %7.1f cycles per box, %5.1f B/clk  (%s)\n", box_bytes, box_bytes * 256,
             inflight, double(h) / iters, box_bytes * 256.0 * iters / h, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
