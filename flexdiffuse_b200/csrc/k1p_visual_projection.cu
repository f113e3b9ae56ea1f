// K1P -- CLIP `visual_projection` as the prologue of K1: the last Linear of the vision tower
// (encode/clip.py:100, [257 x 1024] . [1024 x 768]^T per image, no bias) on tcgen05, producing the
// fp32 guide embeddings K1 consumes.  SURVEY 8a: "line 100 projection is fused into K1".
//
// The reference computes it in fp32 and everything downstream is decided on near-ties of 100 * cos, so
// the product must be fp32-accurate: same two-term fp16 split as K1 (x 2^k = h1 + h2, three exact
// products h2.h1 + h1.h2 + h1.h1, fp32 accumulation in TMEM).  The weight planes are split once per
// weight version by `k1p_split_kernel`; the activations are split on the fly by the feed warps.
//   CTA = 128 rows (tokens) x 256 output features; K chunks of 64; warp 0 TMA (weight planes),
//   warp 1 MMA issuer, warps 2-7 activation feed, then all warps drain TMEM -> fp32 global.
#include <cuda_fp16.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int P_THREADS = 256;
constexpr int P_FEED_THREADS = 192;
constexpr int P_KC = 64;
constexpr int P_BN = 256;
constexpr int P_A_PLANE = 128 * 128;          // activations: 128 rows x 128 B
constexpr int P_W_PLANE = P_BN * 128;         // weights: 256 rows x 128 B
constexpr int P_STAGE = 2 * P_A_PLANE + 2 * P_W_PLANE;   // 98304
constexpr int P_SMEM = 1024 + P_STAGE + 256;
constexpr float P_ACT_SCALE = 64.0f;          // |activation| < 1023
constexpr float P_W_SCALE = 1024.0f;          // |weight| < 63
constexpr int P_ITEMS = (128 * 8 + P_FEED_THREADS - 1) / P_FEED_THREADS;  // 6

__device__ int g_k1p_range_flag;

// fp32 [rows, K] -> two fp16 planes of x * scale
__global__ void __launch_bounds__(256) k1p_split_kernel(const float* __restrict__ x, __half* __restrict__ h1p,
                                                        __half* __restrict__ h2p, int64_t n, float scale) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 v = *reinterpret_cast<const float4*>(x + i);
  const float xs[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
  unsigned short a[4], b[4];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bad |= !(fabsf(xs[k]) < 65504.f);
    const __half h = __float2half_rn(xs[k]);
    a[k] = __half_as_ushort(h);
    b[k] = __half_as_ushort(__float2half_rn(xs[k] - __half2float(h)));
  }
  if (bad) g_k1p_range_flag = 1;
  *reinterpret_cast<uint2*>(h1p + i) = make_uint2(a[0] | (static_cast<uint32_t>(a[1]) << 16), a[2] | (static_cast<uint32_t>(a[3]) << 16));
  *reinterpret_cast<uint2*>(h2p + i) = make_uint2(b[0] | (static_cast<uint32_t>(b[1]) << 16), b[2] | (static_cast<uint32_t>(b[3]) << 16));
}

__global__ void __launch_bounds__(P_THREADS, 1)
k1p_project_kernel(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                   const float* __restrict__ act, float* __restrict__ out, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage + P_STAGE);
  uint64_t* empty_bar = full_bar + 1;
  uint64_t* done_bar = empty_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * P_BN;
  const int num_kc = K / P_KC;
  if (tid == 0) {
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    mbar_init(full_bar, 1 + P_FEED_THREADS / 32);
    mbar_init(empty_bar, 1);
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait_backoff(empty_bar, (kc & 1) ^ 1);
        mbar_expect_tx(full_bar, 2u * P_W_PLANE);
        tma_load_2d(stage + 2 * P_A_PLANE, &tm_w1, full_bar, kc * P_KC, n0);
        tma_load_2d(stage + 2 * P_A_PLANE + P_W_PLANE, &tm_w2, full_bar, kc * P_KC, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(UMMA_F16, 128, P_BN, 0, 0);
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait_backoff(full_bar, kc & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(stage);
        const uint64_t a1 = umma_desc_sw128(base, 16, 1024), a2 = umma_desc_sw128(base + P_A_PLANE, 16, 1024);
        const uint64_t w1 = umma_desc_sw128(base + 2 * P_A_PLANE, 16, 1024);
        const uint64_t w2 = umma_desc_sw128(base + 2 * P_A_PLANE + P_W_PLANE, 16, 1024);
#pragma unroll
        for (int ks = 0; ks < P_KC / 16; ++ks) {  // small terms first
          mma_f16_ss(tmem_base, a2 + 2 * ks, w1 + 2 * ks, idesc, (kc | ks) != 0);
          mma_f16_ss(tmem_base, a1 + 2 * ks, w2 + 2 * ks, idesc, 1);
          mma_f16_ss(tmem_base, a1 + 2 * ks, w1 + 2 * ks, idesc, 1);
        }
        tc_commit(empty_bar);
      }
      tc_commit(done_bar);
    }
  } else {
    // activation feed: 128 rows x 64 elements per chunk, 8 elements per item
    const int tt = tid - 64;
    float4 rb[P_ITEMS][2];
    auto load_chunk = [&](int kc) {
#pragma unroll
      for (int j = 0; j < P_ITEMS; ++j) {
        const int f = tt + j * P_FEED_THREADS;
        const int row = m0 + (f >> 3);
        if (f < 128 * 8 && row < M) {
          const float4* src = reinterpret_cast<const float4*>(act + static_cast<size_t>(row) * K + kc * P_KC) + 2 * (f & 7);
          rb[j][0] = __ldg(src);
          rb[j][1] = __ldg(src + 1);
        } else {
          rb[j][0] = rb[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    load_chunk(0);
    bool bad = false;
    for (int kc = 0; kc < num_kc; ++kc) {
      mbar_wait_backoff(empty_bar, (kc & 1) ^ 1);
#pragma unroll
      for (int j = 0; j < P_ITEMS; ++j) {
        const int f = tt + j * P_FEED_THREADS;
        if (f < 128 * 8) {
          const float x[8] = {rb[j][0].x, rb[j][0].y, rb[j][0].z, rb[j][0].w,
                              rb[j][1].x, rb[j][1].y, rb[j][1].z, rb[j][1].w};
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float x0 = x[2 * q] * P_ACT_SCALE, x1 = x[2 * q + 1] * P_ACT_SCALE;
            bad |= !(fabsf(x0) < 65504.f) | !(fabsf(x1) < 65504.f);
            const __half a0 = __float2half_rn(x0), a1 = __float2half_rn(x1);
            const __half b0 = __float2half_rn(x0 - __half2float(a0)), b1 = __float2half_rn(x1 - __half2float(a1));
            ph[q] = static_cast<uint32_t>(__half_as_ushort(a0)) | (static_cast<uint32_t>(__half_as_ushort(a1)) << 16);
            pl[q] = static_cast<uint32_t>(__half_as_ushort(b0)) | (static_cast<uint32_t>(__half_as_ushort(b1)) << 16);
          }
          const uint32_t off = sw128_offset(f >> 3, f & 7);
          *reinterpret_cast<uint4*>(stage + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(stage + P_A_PLANE + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar);
      if (kc + 1 < num_kc) load_chunk(kc + 1);
    }
    if (bad) g_k1p_range_flag = 1;
  }
  mbar_wait(done_bar, 0);
  tc_fence_after();
  // drain: warp w reads TMEM lanes 32 (w % 4) .., columns split between the two warps of a quarter
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int row = m0 + quarter * 32 + lane;
    const float unscale = 1.0f / (P_ACT_SCALE * P_W_SCALE);
    for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
      uint32_t v[2][16];
      tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v[0]);
      tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0 + 16, v[1]);
      tmem_ld_wait();
      if (row < M) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const int n = n0 + c0 + 16 * g + q;
            if (n + 3 < N) {
              *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * N + n) =
                  make_float4(__uint_as_float(v[g][q]) * unscale, __uint_as_float(v[g][q + 1]) * unscale,
                              __uint_as_float(v[g][q + 2]) * unscale, __uint_as_float(v[g][q + 3]) * unscale);
            } else {
              for (int e = 0; e < 4; ++e)
                if (n + e < N) out[static_cast<size_t>(row) * N + n + e] = __uint_as_float(v[g][q + e]) * unscale;
            }
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace
}  // namespace fd

extern "C" int64_t fd_visual_projection_workspace_bytes(int N, int K) {
  return 2 * static_cast<int64_t>(N) * K * 2 + 64;
}

extern "C" int fd_visual_projection(const float* hidden_dev, const float* weight_dev, float* out_dev, int M, int N,
                                    int K, void* workspace_dev, int64_t workspace_bytes, int weights_changed,
                                    void* stream) {
  using namespace fd;
  FD_REQUIRE(hidden_dev && weight_dev && out_dev && workspace_dev, "fd_visual_projection: NULL pointer");
  FD_REQUIRE(M > 0 && N > 0 && K > 0, "fd_visual_projection: non-positive shape");
  FD_REQUIRE(K % P_KC == 0 && N % 4 == 0, "fd_visual_projection: need K %% 64 == 0 and N %% 4 == 0 (K=%d, N=%d)", K, N);
  FD_REQUIRE(workspace_bytes >= fd_visual_projection_workspace_bytes(N, K), "fd_visual_projection: workspace too small");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(hidden_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(weight_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(out_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(workspace_dev) % 16 == 0,
             "fd_visual_projection: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* w1 = static_cast<__half*>(workspace_dev);
  __half* w2 = w1 + static_cast<int64_t>(N) * K;
  if (weights_changed) {
    const int64_t n = static_cast<int64_t>(N) * K;
    k1p_split_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(weight_dev, w1, w2, n, P_W_SCALE);
    FD_CUDA_OK(cudaGetLastError());
  }
  CUtensorMap t1, t2;
  for (int which = 0; which < 2; ++which) {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {P_KC, P_BN};
    rc = encode_tmap(which ? &t2 : &t1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, which ? w2 : w1, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  FD_CUDA_OK(cudaFuncSetAttribute(k1p_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
  dim3 grid((M + 127) / 128, (N + P_BN - 1) / P_BN);
  k1p_project_kernel<<<grid, P_THREADS, P_SMEM, st>>>(t1, t2, hidden_dev, out_dev, M, N, K);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

// 1 when an operand of a projection since the last call was outside the fp16-split range (|x| >= 1023 for
// activations, >= 63 for weights, or non-finite); clears the flag.  Synchronises the device.
extern "C" int fd_visual_projection_range_flag(void) {
  int v = 0, z = 0;
  if (cudaMemcpyFromSymbol(&v, fd::g_k1p_range_flag, sizeof(int)) != cudaSuccess) return FD_ERR_CUDA;
  if (v) cudaMemcpyToSymbol(fd::g_k1p_range_flag, &z, sizeof(int));
  return v;
}
