// Microbenchmark (development aid): cycles per tcgen05.mma (cta_group::1, SS operands, SWIZZLE_128B
// K-major, accumulating into one TMEM tile) as a function of N and kind, with nothing else running
// on the SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../flexdiffuse_b200/csrc/fd_common.cuh"

using namespace fd;

template <bool TF32>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int distinct_k, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  // zero operands: A 128 rows x 128 B x 4 k-blocks, B 256 rows x 128 B x 4 k-blocks
  for (int i = threadIdx.x; i < (128 + 256) * 128 * 4 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(TF32 ? UMMA_TF32 : UMMA_BF16, 128, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 4 * 128 * 128);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int kb = it % distinct_k;  // which 128 B-wide k-block
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_desc_sw128(a0 + kb * 128 * 128, 16, 1024) + 2 * k;
        const uint64_t bd = umma_desc_sw128(b0 + kb * 256 * 128, 16, 1024) + 2 * k;
        if (TF32) mma_tf32_ss(tmem, ad, bd, idesc, 1);
        else mma_f16_ss(tmem, ad, bd, idesc, 1);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  // TMEM read-back cost: 256 columns per lane, (a) wait after every x16 load, (b) 4 loads per wait
  {
    const uint32_t lane_addr = static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16;
    float acc = 0.f;
    long long ta = clock64();
    for (int c = 0; c < 256; c += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem + lane_addr + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) acc += __uint_as_float(v[q]);
    }
    long long tb = clock64();
    for (int c = 0; c < 256; c += 64) {
      uint32_t v[4][16];
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_ld_x16(tmem + lane_addr + c + 16 * g, v[g]);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int q = 0; q < 16; ++q) acc += __uint_as_float(v[g][q]);
    }
    long long tc = clock64();
    if (threadIdx.x == 0) {
      out[148 + blockIdx.x] = tb - ta;
      out[296 + blockIdx.x] = tc - tb;
    }
    if (acc == 123.456f) out[0] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 3 * 148 * sizeof(long long));
  const int smem = (128 + 256) * 128 * 4 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 256;
  for (int grid : {1, 148})
    for (int tf32 = 0; tf32 < 2; ++tf32)
      for (int N : {64, 80, 128, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (tf32) mma_rate_kernel<true><<<grid, 128, smem>>>(N, iters, 4, d);
          else mma_rate_kernel<false><<<grid, 128, smem>>>(N, iters, 4, d);
          cudaDeviceSynchronize();
        }
        long long h[444];
        cudaMemcpy(h, d, 444 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid=%3d kind=%s M=128 N=%3d K=%2d : %7.1f cycles per MMA (%d MMAs)  err=%s\n", grid,
               tf32 ? "tf32" : "bf16", N, tf32 ? 8 : 16, double(mx) / (iters * 4), iters * 4,
               cudaGetErrorString(cudaGetLastError()));
        printf("      tmem read-back of 256 cols: wait-each %lld cycles, 4-per-wait %lld cycles\n", h[148], h[296]);
      }
  return 0;
}
