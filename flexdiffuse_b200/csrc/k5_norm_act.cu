// K5 / K6 -- normalisation + activation glue of the UNet forward that the denoising loop
// (pipeline/guide.py:56-58 -> unet(...)) spends most of its non-GEMM time in.
//
// The round-1 ncu launch list (profiles/r01/SUMMARY.md) showed that at B=1 the ATen GroupNorm path
// (NHWC->NCHW copy, RowwiseMoments, ComputeFusedParams, apply, separate SiLU, separate broadcast add
// of the time embedding) and the GEGLU gelu/mul pair on strided views are ~45 % of a denoising step.
// These are pure streaming kernels over L2-resident activations:
//
//   K5  y = act( GroupNorm( x + bias[n,c] ) * gamma[c] + beta[c] )        NHWC bf16 in / out
//       diffusers ResnetBlock2D: norm1 -> SiLU, (+ time_emb_proj) -> norm2 -> SiLU;
//       SpatialTransformer.norm, conv_norm_out.  Two launches: per-slab partial (sum, sumsq) per
//       group, then normalise + affine + SiLU.  No atomics, deterministic.
//   K6  out[m, f] = in[m, f] * gelu(in[m, F + f])                          diffusers GEGLU
#include <math.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_GROUPS = 32;
constexpr int GN_MAX_SLABS = 64;

struct GnArgs {
  const __nv_bfloat16* x;      // [N, HW, C] (channels_last view of [N, C, H, W])
  const __nv_bfloat16* bias;   // [N, C] or nullptr
  const __nv_bfloat16* gamma;  // [C]
  const __nv_bfloat16* beta;   // [C]
  float* partial;              // [N, slabs, G, 2]
  __nv_bfloat16* y;            // [N, HW, C]
  int HW, C, G, slabs, rows_per_slab;
  float eps;
  int act_silu;
};

__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

// phase 1: partial sums.  grid (slabs, N).  Thread t owns channel pairs t, t+256, ... (coalesced
// across the warp); a channel pair never straddles a group because C/G is even.
__global__ void __launch_bounds__(GN_THREADS) k5_gn_stats_kernel(const GnArgs a) {
  __shared__ float s_sum[GN_MAX_GROUPS], s_sq[GN_MAX_GROUPS];
  const int n = blockIdx.y, slab = blockIdx.x;
  const int pairs = a.C >> 1;
  const int cg2 = (a.C / a.G) >> 1;  // channel pairs per group
  if (threadIdx.x < GN_MAX_GROUPS) {
    s_sum[threadIdx.x] = 0.f;
    s_sq[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int r0 = slab * a.rows_per_slab;
  const int r1 = min(a.HW, r0 + a.rows_per_slab);
  const uint32_t* xb = reinterpret_cast<const uint32_t*>(a.x) + static_cast<size_t>(n) * a.HW * pairs;
  const uint32_t* bb = a.bias ? reinterpret_cast<const uint32_t*>(a.bias) + static_cast<size_t>(n) * pairs : nullptr;
  for (int p = threadIdx.x; p < pairs; p += GN_THREADS) {
    float2 b = bb ? bf2_to_f2(bb[p]) : make_float2(0.f, 0.f);
    float s = 0.f, q = 0.f;
    int r = r0;
    for (; r + 4 <= r1; r += 4) {
      uint32_t v0 = xb[static_cast<size_t>(r) * pairs + p];
      uint32_t v1 = xb[static_cast<size_t>(r + 1) * pairs + p];
      uint32_t v2 = xb[static_cast<size_t>(r + 2) * pairs + p];
      uint32_t v3 = xb[static_cast<size_t>(r + 3) * pairs + p];
      float2 f0 = bf2_to_f2(v0), f1 = bf2_to_f2(v1), f2 = bf2_to_f2(v2), f3 = bf2_to_f2(v3);
      f0.x += b.x; f0.y += b.y; f1.x += b.x; f1.y += b.y;
      f2.x += b.x; f2.y += b.y; f3.x += b.x; f3.y += b.y;
      s += (f0.x + f0.y) + (f1.x + f1.y) + (f2.x + f2.y) + (f3.x + f3.y);
      q += (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y) + (f2.x * f2.x + f2.y * f2.y) +
           (f3.x * f3.x + f3.y * f3.y);
    }
    for (; r < r1; ++r) {
      float2 f = bf2_to_f2(xb[static_cast<size_t>(r) * pairs + p]);
      f.x += b.x;
      f.y += b.y;
      s += f.x + f.y;
      q += f.x * f.x + f.y * f.y;
    }
    const int g = p / cg2;
    atomicAdd(&s_sum[g], s);
    atomicAdd(&s_sq[g], q);
  }
  __syncthreads();
  if (threadIdx.x < a.G) {
    float* dst = a.partial + ((static_cast<size_t>(n) * a.slabs + slab) * a.G + threadIdx.x) * 2;
    dst[0] = s_sum[threadIdx.x];
    dst[1] = s_sq[threadIdx.x];
  }
}

// phase 2: normalise + affine + optional SiLU.  Same grid / mapping.
__global__ void __launch_bounds__(GN_THREADS) k5_gn_apply_kernel(const GnArgs a) {
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  const int n = blockIdx.y, slab = blockIdx.x;
  const int pairs = a.C >> 1;
  const int cg = a.C / a.G, cg2 = cg >> 1;
  if (threadIdx.x < a.G) {
    float s = 0.f, q = 0.f;
    const float* src = a.partial + (static_cast<size_t>(n) * a.slabs * a.G + threadIdx.x) * 2;
    for (int k = 0; k < a.slabs; ++k) {
      s += src[static_cast<size_t>(k) * a.G * 2];
      q += src[static_cast<size_t>(k) * a.G * 2 + 1];
    }
    const float cnt = static_cast<float>(a.HW) * cg;
    const float mean = s / cnt;
    const float var = fmaxf(q / cnt - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + a.eps);
  }
  __syncthreads();
  const int r0 = slab * a.rows_per_slab;
  const int r1 = min(a.HW, r0 + a.rows_per_slab);
  const size_t base = static_cast<size_t>(n) * a.HW * pairs;
  const uint32_t* xb = reinterpret_cast<const uint32_t*>(a.x) + base;
  uint32_t* yb = reinterpret_cast<uint32_t*>(a.y) + base;
  const uint32_t* bb = a.bias ? reinterpret_cast<const uint32_t*>(a.bias) + static_cast<size_t>(n) * pairs : nullptr;
  const uint32_t* gb = reinterpret_cast<const uint32_t*>(a.gamma);
  const uint32_t* tb = reinterpret_cast<const uint32_t*>(a.beta);
  for (int p = threadIdx.x; p < pairs; p += GN_THREADS) {
    const int g = p / cg2;
    const float mean = s_mean[g], rstd = s_rstd[g];
    const float2 gam = bf2_to_f2(gb[p]), bet = bf2_to_f2(tb[p]);
    const float2 b = bb ? bf2_to_f2(bb[p]) : make_float2(0.f, 0.f);
    // y = (x + b - mean) * rstd * gamma + beta  =  x * sc + sh
    const float scx = rstd * gam.x, scy = rstd * gam.y;
    const float shx = (b.x - mean) * scx + bet.x, shy = (b.y - mean) * scy + bet.y;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const float2 f = bf2_to_f2(xb[static_cast<size_t>(r) * pairs + p]);
      float ox = f.x * scx + shx, oy = f.y * scy + shy;
      if (a.act_silu) {
        ox = ox / (1.0f + __expf(-ox));
        oy = oy / (1.0f + __expf(-oy));
      }
      __nv_bfloat162 o = __floats2bfloat162_rn(ox, oy);
      yb[static_cast<size_t>(r) * pairs + p] = *reinterpret_cast<uint32_t*>(&o);
    }
  }
}

// K6: GEGLU.  in [M, 2F] bf16 -> out [M, F] bf16; 8 elements (16 B) per thread.
__global__ void __launch_bounds__(256) k6_geglu_kernel(const __nv_bfloat16* __restrict__ in,
                                                       __nv_bfloat16* __restrict__ out, int64_t M, int F) {
  const int f8 = F >> 3;
  const int64_t total = M * f8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = i / f8;
    const int c = static_cast<int>(i - m * f8) << 3;
    const uint4 xv = *reinterpret_cast<const uint4*>(in + m * 2 * F + c);
    const uint4 gv = *reinterpret_cast<const uint4*>(in + m * 2 * F + F + c);
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const uint32_t gs[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = bf2_to_f2(xs[k]), g = bf2_to_f2(gs[k]);
      // exact (erf) GELU, as torch.nn.functional.gelu's default
      const float gx = 0.5f * g.x * (1.0f + erff(g.x * 0.70710678118654752f));
      const float gy = 0.5f * g.y * (1.0f + erff(g.y * 0.70710678118654752f));
      __nv_bfloat162 o = __floats2bfloat162_rn(x.x * gx, x.y * gy);
      os[k] = *reinterpret_cast<uint32_t*>(&o);
    }
    *reinterpret_cast<uint4*>(out + m * F + c) = make_uint4(os[0], os[1], os[2], os[3]);
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_groupnorm_act_workspace_bytes(int N, int G) {
  return N * fd::GN_MAX_SLABS * G * 2 * static_cast<int>(sizeof(float));
}

extern "C" int fd_groupnorm_act(const void* x_bf16_dev, const void* bias_bf16_dev, const void* gamma_bf16_dev,
                                const void* beta_bf16_dev, void* workspace_dev, void* y_bf16_dev, int N, int HW,
                                int C, int G, float eps, int act_silu, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && gamma_bf16_dev && beta_bf16_dev && workspace_dev && y_bf16_dev,
             "fd_groupnorm_act: NULL pointer");
  FD_REQUIRE(N > 0 && HW > 0 && C > 0, "fd_groupnorm_act: non-positive shape");
  FD_REQUIRE(G > 0 && G <= GN_MAX_GROUPS && C % G == 0 && (C / G) % 2 == 0,
             "fd_groupnorm_act: need G <= %d, C %% G == 0 and an even number of channels per group (C=%d G=%d)",
             GN_MAX_GROUPS, C, G);
  FD_REQUIRE(N <= 65535, "fd_groupnorm_act: N exceeds grid limits");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_groupnorm_act: cannot query SM count");
  GnArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x_bf16_dev);
  a.bias = static_cast<const __nv_bfloat16*>(bias_bf16_dev);
  a.gamma = static_cast<const __nv_bfloat16*>(gamma_bf16_dev);
  a.beta = static_cast<const __nv_bfloat16*>(beta_bf16_dev);
  a.partial = static_cast<float*>(workspace_dev);
  a.y = static_cast<__nv_bfloat16*>(y_bf16_dev);
  a.HW = HW;
  a.C = C;
  a.G = G;
  a.eps = eps;
  a.act_silu = act_silu;
  // about two CTAs per SM over the whole batch, at least 8 rows per slab
  int slabs = (2 * sms + N - 1) / N;
  if (slabs > GN_MAX_SLABS) slabs = GN_MAX_SLABS;
  if (slabs > (HW + 7) / 8) slabs = (HW + 7) / 8;
  if (slabs < 1) slabs = 1;
  a.rows_per_slab = (HW + slabs - 1) / slabs;
  a.slabs = (HW + a.rows_per_slab - 1) / a.rows_per_slab;
  dim3 grid(a.slabs, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k5_gn_stats_kernel<<<grid, GN_THREADS, 0, st>>>(a);
  k5_gn_apply_kernel<<<grid, GN_THREADS, 0, st>>>(a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_geglu(const void* in_bf16_dev, void* out_bf16_dev, int64_t M, int F, void* stream) {
  using namespace fd;
  FD_REQUIRE(in_bf16_dev && out_bf16_dev, "fd_geglu: NULL pointer");
  FD_REQUIRE(M > 0 && F > 0 && F % 8 == 0, "fd_geglu: need M > 0 and F a positive multiple of 8");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(in_bf16_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(out_bf16_dev) % 16 == 0,
             "fd_geglu: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_geglu: cannot query SM count");
  const int64_t total = M * (F >> 3);
  int64_t want = (total + 255) / 256, cap = static_cast<int64_t>(sms) * 8;
  k6_geglu_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_bf16_dev), static_cast<__nv_bfloat16*>(out_bf16_dev), M, F);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
