// K11 -- fp32-accurate Linear on tcgen05 for the CLIP towers:
//   out[M, N] = act( x[M, K] . W[N, K]^T + bias[N] )
// Replaces the fp32 `nn.Linear`s (q / k / v / out_proj, fc1, fc2) of the ViT-L/14 vision tower and the
// text tower that `CLIPEncoder.image` / `.prompt` run (reference encode/clip.py:57-65, 86-100 through
// transformers' CLIPModel).  The reference computes them in fp32 and K1's decisions hang on near-ties
// of the resulting embeddings, so a single bf16 / tf32 pass is not acceptable (3e-4 relative error on the
// tower output, measured); in fp32 they are CUDA-core GEMMs: 155 GFLOP per image, 7.6 ms on a B200.
//
// Same arithmetic as K1 / K1P: every operand row is scaled by a power of two and split into two fp16
// terms, x 2^e = h1 + h2 (11 + 11 mantissa bits); the three exact products h2.w1 + h1.w2 + h1.w1 are
// accumulated in fp32 in TMEM, which carries the result to fp32 accuracy (error <= ~4e-7 |x| |w| per row
// pair; tests/test_linear_x3_gpu.py).  Unlike K1P the scale is chosen PER ROW from the row's maximum
// (max |x| 2^e in [2^13, 2^14)), so no activation range has to be assumed inside the towers.
//   k11_split_rows_kernel   fp32 rows -> h1 / h2 planes + the row's inverse scale (weights once per
//                           version, activations once per call; one warp per row)
//   k11_gemm_kernel<BN>     both operands by TMA (SWIZZLE_128B, K chunks of 64, multi-stage ring),
//                           warp 0 producer, warp 1 MMA issuer (M128 N=BN K16, three products per k-step),
//                           warps 2-5 epilogue: TMEM -> x inv_a[row] x inv_w[col] (+ bias, activation) ->
//                           fp32 global, or split-K partials
//   k11_reduce_kernel       split-K partials in a fixed order (+ bias, activation)
#include <cuda_fp16.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int L_THREADS = 192;
constexpr int L_KC = 64;
constexpr int L_A_PLANE = 128 * 128;   // 128 rows x 128 B

__device__ int g_k11_flag;  // 1: non-finite operand; >= 16: a bounded wait gave up (development aid)

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == FD_LINEAR_ACT_QUICK_GELU) return v / (1.0f + __expf(-1.702f * v));
  if (act == FD_LINEAR_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  return v;
}

// one warp per row: the row maximum picks the power of two, then the two-term split
__global__ void __launch_bounds__(256) k11_split_rows_kernel(const float* __restrict__ x, __half* __restrict__ h1p,
                                                             __half* __restrict__ h2p, float* __restrict__ inv_scale,
                                                             int rows, int K) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * K);
  const int k4 = K >> 2;
  float mx = 0.f;
  bool bad = false;
  for (int c = lane; c < k4; c += 32) {
    const float4 v = __ldg(xr + c);
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    bad |= !(fabsf(v.x) <= 3.0e38f) | !(fabsf(v.y) <= 3.0e38f) | !(fabsf(v.z) <= 3.0e38f) | !(fabsf(v.w) <= 3.0e38f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  bad = __any_sync(0xffffffffu, bad);
  int ex = 0;
  if (mx > 0.f) frexpf(mx, &ex);            // mx = m 2^ex, m in [0.5, 1)
  ex = max(-100, min(100, ex));
  const float scale = exp2f(static_cast<float>(14 - ex));   // mx * scale in [2^13, 2^14)
  if (bad && lane == 0) g_k11_flag = 1;
  if (lane == 0) inv_scale[row] = exp2f(static_cast<float>(ex - 14));
  uint2* o1 = reinterpret_cast<uint2*>(h1p + static_cast<size_t>(row) * K);
  uint2* o2 = reinterpret_cast<uint2*>(h2p + static_cast<size_t>(row) * K);
  for (int c = lane; c < k4; c += 32) {
    const float4 v = __ldg(xr + c);
    const float a = v.x * scale, b = v.y * scale, cc = v.z * scale, d = v.w * scale;
    const __half2 p0 = __floats2half2_rn(a, b), p1 = __floats2half2_rn(cc, d);
    const float2 f0 = __half22float2(p0), f1 = __half22float2(p1);
    const __half2 q0 = __floats2half2_rn(a - f0.x, b - f0.y), q1 = __floats2half2_rn(cc - f1.x, d - f1.y);
    o1[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
    o2[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&q0), *reinterpret_cast<const uint32_t*>(&q1));
  }
}

struct LArgs {
  const float* inv_a;   // [M]
  const float* inv_w;   // [N]
  const float* bias;    // [N] or nullptr
  float* out;           // [M, N] (split_k == 1) or partials [split_k, M, N]
  int M, N, K, act, split_k, kc_per_split;
};

template <int BN>
__global__ void __launch_bounds__(L_THREADS, 1)
k11_gemm_kernel(const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_a2,
                const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2, const LArgs a) {
  constexpr int W_PLANE = BN * 128;
  constexpr int STAGE = 2 * L_A_PLANE + 2 * W_PLANE;
  constexpr int STAGES = (200 * 1024) / STAGE;   // BN 128: 3 stages of 64 KB, BN 64: 4 of 48 KB
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(stage + STAGES * STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN, z = blockIdx.z;
  const int kc0 = z * a.kc_per_split;
  const int num_kc = min(a.kc_per_split, a.K / L_KC - kc0);
  if (tid == 0) {
    tma_prefetch_desc(&tm_a1);
    tma_prefetch_desc(&tm_a2);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < num_kc; ++i) {
        const int s = i % STAGES;
        mbar_wait_bounded(&empty[s], ((i / STAGES) & 1) ^ 1, &g_k11_flag, 16);
        mbar_expect_tx(&full[s], STAGE);
        uint8_t* st = stage + s * STAGE;
        const int k = (kc0 + i) * L_KC;
        tma_load_2d(st, &tm_a1, &full[s], k, m0);
        tma_load_2d(st + L_A_PLANE, &tm_a2, &full[s], k, m0);
        tma_load_2d(st + 2 * L_A_PLANE, &tm_w1, &full[s], k, n0);
        tma_load_2d(st + 2 * L_A_PLANE + W_PLANE, &tm_w2, &full[s], k, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(UMMA_F16, 128, BN, 0, 0);
      for (int i = 0; i < num_kc; ++i) {
        const int s = i % STAGES;
        mbar_wait_bounded(&full[s], (i / STAGES) & 1, &g_k11_flag, 17);
        tc_fence_after();
        const uint32_t base = smem_u32(stage + s * STAGE);
        const uint64_t a1 = umma_desc_sw128(base, 16, 1024), a2 = umma_desc_sw128(base + L_A_PLANE, 16, 1024);
        const uint64_t w1 = umma_desc_sw128(base + 2 * L_A_PLANE, 16, 1024);
        const uint64_t w2 = umma_desc_sw128(base + 2 * L_A_PLANE + W_PLANE, 16, 1024);
#pragma unroll
        for (int ks = 0; ks < L_KC / 16; ++ks) {  // small terms first
          mma_f16_ss(tmem_base, a2 + 2 * ks, w1 + 2 * ks, idesc, (i | ks) != 0);
          mma_f16_ss(tmem_base, a1 + 2 * ks, w2 + 2 * ks, idesc, 1);
          mma_f16_ss(tmem_base, a1 + 2 * ks, w1 + 2 * ks, idesc, 1);
        }
        tc_commit(&empty[s]);
      }
      tc_commit(done);
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 = rows m0 + 32 (w % 4) + lane
    const int quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    const float ia = row < a.M ? a.inv_a[row] : 0.f;
    mbar_wait_bounded(done, 0, &g_k11_flag, 18);
    tc_fence_after();
    float* orow = a.out + (static_cast<size_t>(z) * a.M + (row < a.M ? row : 0)) * a.N;
    const bool final_out = a.split_k == 1;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[2][16];
      tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v[0]);
      tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0 + 16, v[1]);
      tmem_ld_wait();
      if (row < a.M) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const int n = n0 + c0 + 16 * g + q;
            if (n < a.N) {  // N % 4 == 0
              const float4 iw = __ldg(reinterpret_cast<const float4*>(a.inv_w + n));
              float4 o = make_float4(__uint_as_float(v[g][q]) * (ia * iw.x), __uint_as_float(v[g][q + 1]) * (ia * iw.y),
                                     __uint_as_float(v[g][q + 2]) * (ia * iw.z), __uint_as_float(v[g][q + 3]) * (ia * iw.w));
              if (final_out) {
                if (a.bias) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n));
                  o.x += b.x;
                  o.y += b.y;
                  o.z += b.z;
                  o.w += b.w;
                }
                o.x = act_apply(o.x, a.act);
                o.y = act_apply(o.y, a.act);
                o.z = act_apply(o.z, a.act);
                o.w = act_apply(o.w, a.act);
              }
              *reinterpret_cast<float4*>(orow + n) = o;
            }
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// out = act( sum_z partial[z] + bias ), z in ascending order
__global__ void __launch_bounds__(256) k11_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                                         float* __restrict__ out, int64_t mn4, int n4, int split_k, int act) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= mn4) return;
  float4 s = __ldg(reinterpret_cast<const float4*>(partial) + i);
  for (int z = 1; z < split_k; ++z) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(partial) + static_cast<int64_t>(z) * mn4 + i);
    s.x += p.x;
    s.y += p.y;
    s.z += p.z;
    s.w += p.w;
  }
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (i % n4));
    s.x += b.x;
    s.y += b.y;
    s.z += b.z;
    s.w += b.w;
  }
  s.x = act_apply(s.x, act);
  s.y = act_apply(s.y, act);
  s.z = act_apply(s.z, act);
  s.w = act_apply(s.w, act);
  reinterpret_cast<float4*>(out)[i] = s;
}

template <int BN>
int launch_k11(const CUtensorMap* tm, const LArgs& a, cudaStream_t st) {
  constexpr int STAGE = 2 * L_A_PLANE + 2 * BN * 128;
  constexpr int STAGES = (200 * 1024) / STAGE;
  constexpr int SMEM = 1024 + STAGES * STAGE + 256;
  static_assert(SMEM <= 227 * 1024, "K11 shared memory");
  FD_CUDA_OK(cudaFuncSetAttribute(k11_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  dim3 grid((a.M + 127) / 128, (a.N + BN - 1) / BN, a.split_k);
  k11_gemm_kernel<BN><<<grid, L_THREADS, SMEM, st>>>(tm[0], tm[1], tm[2], tm[3], a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int64_t fd_linear_x3_operand_bytes(int rows, int K) {
  // [h1 plane][h2 plane][inverse row scales], each 256-byte aligned
  const int64_t plane = (static_cast<int64_t>(rows) * K * 2 + 255) / 256 * 256;
  return 2 * plane + (static_cast<int64_t>(rows) * 4 + 255) / 256 * 256;
}

extern "C" int fd_linear_x3_split(const float* x_dev, int rows, int K, void* operand_dev, int64_t operand_bytes,
                                  void* stream) {
  using namespace fd;
  FD_REQUIRE(x_dev && operand_dev, "fd_linear_x3_split: NULL pointer");
  FD_REQUIRE(rows > 0 && K > 0 && K % L_KC == 0, "fd_linear_x3_split: need rows > 0 and K %% 64 == 0 (rows=%d, K=%d)", rows, K);
  FD_REQUIRE(operand_bytes >= fd_linear_x3_operand_bytes(rows, K), "fd_linear_x3_split: operand buffer too small");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(x_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(operand_dev) % 256 == 0,
             "fd_linear_x3_split: x must be 16-byte and the operand buffer 256-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int64_t plane = (static_cast<int64_t>(rows) * K * 2 + 255) / 256 * 256;
  uint8_t* base = static_cast<uint8_t*>(operand_dev);
  k11_split_rows_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_dev, reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + plane),
      reinterpret_cast<float*>(base + 2 * plane), rows, K);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_linear_x3(const void* act_operand_dev, int M, const void* weight_operand_dev, int N, int K,
                            const float* bias_dev, int act, float* out_dev, float* partial_dev, int split_k,
                            void* stream) {
  using namespace fd;
  FD_REQUIRE(act_operand_dev && weight_operand_dev && out_dev, "fd_linear_x3: NULL pointer");
  FD_REQUIRE(M > 0 && N > 0 && K > 0, "fd_linear_x3: non-positive shape");
  FD_REQUIRE(K % L_KC == 0 && N % 4 == 0, "fd_linear_x3: need K %% 64 == 0 and N %% 4 == 0 (K=%d, N=%d)", K, N);
  FD_REQUIRE(act >= 0 && act <= FD_LINEAR_ACT_GELU, "fd_linear_x3: unknown activation %d", act);
  FD_REQUIRE(split_k >= 1 && split_k <= 16 && (split_k == 1 || partial_dev), "fd_linear_x3: split_k=%d needs a partial buffer",
             split_k);
  FD_REQUIRE(reinterpret_cast<uintptr_t>(out_dev) % 16 == 0 && (!bias_dev || reinterpret_cast<uintptr_t>(bias_dev) % 16 == 0) &&
                 (!partial_dev || reinterpret_cast<uintptr_t>(partial_dev) % 16 == 0),
             "fd_linear_x3: out / bias / partial must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int num_kc = K / L_KC;
  if (split_k > num_kc) split_k = num_kc;
  const int bn = N >= 2048 || (N % 128 == 0 && static_cast<int64_t>((M + 127) / 128) * (N / 128) * split_k >= 96) ? 128 : 64;
  const uint8_t* ab = static_cast<const uint8_t*>(act_operand_dev);
  const uint8_t* wb = static_cast<const uint8_t*>(weight_operand_dev);
  const int64_t a_plane = (static_cast<int64_t>(M) * K * 2 + 255) / 256 * 256;
  const int64_t w_plane = (static_cast<int64_t>(N) * K * 2 + 255) / 256 * 256;
  CUtensorMap tm[4];
  for (int which = 0; which < 4; ++which) {
    const bool is_w = which >= 2;
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(is_w ? N : M)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {L_KC, static_cast<uint32_t>(is_w ? bn : 128)};
    const void* p = is_w ? static_cast<const void*>(wb + (which & 1) * w_plane) : static_cast<const void*>(ab + (which & 1) * a_plane);
    rc = encode_tmap(&tm[which], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  LArgs a;
  a.inv_a = reinterpret_cast<const float*>(ab + 2 * a_plane);
  a.inv_w = reinterpret_cast<const float*>(wb + 2 * w_plane);
  a.bias = bias_dev;
  a.out = split_k == 1 ? out_dev : partial_dev;
  a.M = M;
  a.N = N;
  a.K = K;
  a.act = act;
  a.split_k = split_k;
  a.kc_per_split = (num_kc + split_k - 1) / split_k;
  rc = bn == 128 ? launch_k11<128>(tm, a, st) : launch_k11<64>(tm, a, st);
  if (rc != FD_OK) return rc;
  if (split_k > 1) {
    const int64_t mn4 = static_cast<int64_t>(M) * N / 4;
    k11_reduce_kernel<<<static_cast<unsigned>((mn4 + 255) / 256), 256, 0, st>>>(partial_dev, bias_dev, out_dev, mn4, N / 4,
                                                                               split_k, act);
    FD_CUDA_OK(cudaGetLastError());
  }
  return FD_OK;
}

// development aid / health check: 1 = a non-finite operand was split since the last call, >= 16 = a pipeline wait
// gave up; clears the flag.  Synchronises the device.
extern "C" int fd_linear_x3_flag(void) {
  int v = 0, z = 0;
  if (cudaMemcpyFromSymbol(&v, fd::g_k11_flag, sizeof(int)) != cudaSuccess) return FD_ERR_CUDA;
  if (v) cudaMemcpyToSymbol(fd::g_k11_flag, &z, sizeof(int));
  return v;
}
