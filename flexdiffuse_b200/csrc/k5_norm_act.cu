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
//       channel pair, a tiny per-group finalise, then normalise + affine + SiLU: three launches, no
//       atomics, fixed summation order (bit-reproducible).
//   K7  y = x + h + bias[c]: the resnet residual add with conv2's bias folded in (cuDNN would
//       otherwise add every convolution bias with a separate broadcast kernel).
//   K6  out[m, f] = in[m, f] * gelu(in[m, F + f])                          diffusers GEGLU
#include <math.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int GN_THREADS = 256;   // 8 warps: warp w takes rows w, w+8, ... of the slab
constexpr int GN_WARPS = GN_THREADS / 32;
constexpr int GN_ROWS = 32;        // rows (pixels) per CTA
constexpr int GN_MAX_GROUPS = 32;

struct GnArgs {
  const __nv_bfloat16* x;      // [N, HW, C] (channels_last view of [N, C, H, W])
  const __nv_bfloat16* bias;   // [N, C] or nullptr
  const __nv_bfloat16* gamma;  // [C]
  const __nv_bfloat16* beta;   // [C]
  float2* partial;             // [N, slabs, C/2]  per channel-pair (sum, sumsq) of one slab
  float2* stats;               // [N, G]           (mean, rstd)
  __nv_bfloat16* y;            // [N, HW, C]
  int HW, C, G, slabs;
  float eps;
  int act_silu;
};

__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

// phase 1.  grid (C/64, slabs, N): one CTA = 32 channel pairs x 32 rows; a warp reads 128
// contiguous bytes per row, 4 rows in flight per thread.  Fixed reduction order (no atomics).
__global__ void __launch_bounds__(GN_THREADS) k5_gn_stats_kernel(const GnArgs a) {
  __shared__ float2 red[GN_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pairs = a.C >> 1;
  const int p = blockIdx.x * 32 + lane;
  const int slab = blockIdx.y, n = blockIdx.z;
  const int r0 = slab * GN_ROWS;
  const uint32_t* xb = reinterpret_cast<const uint32_t*>(a.x) + static_cast<size_t>(n) * a.HW * pairs + p;
  float2 b = make_float2(0.f, 0.f);
  if (a.bias) b = bf2_to_f2(reinterpret_cast<const uint32_t*>(a.bias)[static_cast<size_t>(n) * pairs + p]);
  uint32_t v[GN_ROWS / GN_WARPS];
#pragma unroll
  for (int i = 0; i < GN_ROWS / GN_WARPS; ++i) {
    const int r = r0 + warp + i * GN_WARPS;
    v[i] = (r < a.HW) ? xb[static_cast<size_t>(r) * pairs] : 0u;
  }
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < GN_ROWS / GN_WARPS; ++i) {
    const int r = r0 + warp + i * GN_WARPS;
    if (r < a.HW) {
      float2 f = bf2_to_f2(v[i]);
      f.x += b.x;
      f.y += b.y;
      s += f.x + f.y;
      q += f.x * f.x + f.y * f.y;
    }
  }
  red[warp][lane] = make_float2(s, q);
  __syncthreads();
  if (warp == 0) {
    float2 t = red[0][lane];
#pragma unroll
    for (int w = 1; w < GN_WARPS; ++w) {
      t.x += red[w][lane].x;
      t.y += red[w][lane].y;
    }
    a.partial[(static_cast<size_t>(n) * a.slabs + slab) * pairs + p] = t;
  }
}

// phase 2.  grid (G, N): reduce the partials of one group to (mean, rstd), fixed order.
__global__ void __launch_bounds__(256) k5_gn_finalize_kernel(const GnArgs a) {
  __shared__ float2 red[256];
  const int g = blockIdx.x, n = blockIdx.y;
  const int pairs = a.C >> 1, cg2 = (a.C / a.G) >> 1;
  const int total = a.slabs * cg2;
  float s = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int slab = i / cg2, k = i - slab * cg2;
    const float2 t = a.partial[(static_cast<size_t>(n) * a.slabs + slab) * pairs + g * cg2 + k];
    s += t.x;
    q += t.y;
  }
  red[threadIdx.x] = make_float2(s, q);
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[threadIdx.x].x += red[threadIdx.x + o].x;
      red[threadIdx.x].y += red[threadIdx.x + o].y;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float cnt = static_cast<float>(a.HW) * (a.C / a.G);
    const float mean = red[0].x / cnt;
    const float var = fmaxf(red[0].y / cnt - mean * mean, 0.f);
    a.stats[static_cast<size_t>(n) * a.G + g] = make_float2(mean, rsqrtf(var + a.eps));
  }
}

// phase 3: normalise + affine + optional SiLU, same tiling as phase 1.
__global__ void __launch_bounds__(GN_THREADS) k5_gn_apply_kernel(const GnArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pairs = a.C >> 1, cg2 = (a.C / a.G) >> 1;
  const int p = blockIdx.x * 32 + lane;
  const int slab = blockIdx.y, n = blockIdx.z;
  const int r0 = slab * GN_ROWS;
  const size_t base = static_cast<size_t>(n) * a.HW * pairs + p;
  const uint32_t* xb = reinterpret_cast<const uint32_t*>(a.x) + base;
  uint32_t* yb = reinterpret_cast<uint32_t*>(a.y) + base;
  uint32_t v[GN_ROWS / GN_WARPS];
#pragma unroll
  for (int i = 0; i < GN_ROWS / GN_WARPS; ++i) {
    const int r = r0 + warp + i * GN_WARPS;
    v[i] = (r < a.HW) ? xb[static_cast<size_t>(r) * pairs] : 0u;
  }
  const float2 st = a.stats[static_cast<size_t>(n) * a.G + p / cg2];
  const float2 gam = bf2_to_f2(reinterpret_cast<const uint32_t*>(a.gamma)[p]);
  const float2 bet = bf2_to_f2(reinterpret_cast<const uint32_t*>(a.beta)[p]);
  float2 b = make_float2(0.f, 0.f);
  if (a.bias) b = bf2_to_f2(reinterpret_cast<const uint32_t*>(a.bias)[static_cast<size_t>(n) * pairs + p]);
  // y = (x + b - mean) * rstd * gamma + beta  =  x * sc + sh
  const float scx = st.y * gam.x, scy = st.y * gam.y;
  const float shx = (b.x - st.x) * scx + bet.x, shy = (b.y - st.x) * scy + bet.y;
#pragma unroll
  for (int i = 0; i < GN_ROWS / GN_WARPS; ++i) {
    const int r = r0 + warp + i * GN_WARPS;
    if (r < a.HW) {
      const float2 f = bf2_to_f2(v[i]);
      float ox = f.x * scx + shx, oy = f.y * scy + shy;
      if (a.act_silu) {
        ox = ox / (1.0f + __expf(-ox));
        oy = oy / (1.0f + __expf(-oy));
      }
      __nv_bfloat162 o = __floats2bfloat162_rn(ox, oy);
      yb[static_cast<size_t>(r) * pairs] = *reinterpret_cast<uint32_t*>(&o);
    }
  }
}

// K7: y = x + h + bias[c]   (residual add with the convolution bias folded in), NHWC bf16
__global__ void __launch_bounds__(256) k7_add_bias_residual_kernel(const uint4* __restrict__ x,
                                                                   const uint4* __restrict__ h,
                                                                   const uint4* __restrict__ bias, uint4* __restrict__ y,
                                                                   int64_t total8, int c8) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 xv = x[i], hv = h[i], bv = bias[i % c8];
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, hs[4] = {hv.x, hv.y, hv.z, hv.w},
                   bs[4] = {bv.x, bv.y, bv.z, bv.w};
    uint32_t os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = bf2_to_f2(xs[k]), b = bf2_to_f2(hs[k]), c = bf2_to_f2(bs[k]);
      __nv_bfloat162 o = __floats2bfloat162_rn(a.x + (b.x + c.x), a.y + (b.y + c.y));
      os[k] = *reinterpret_cast<uint32_t*>(&o);
    }
    y[i] = make_uint4(os[0], os[1], os[2], os[3]);
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

extern "C" int64_t fd_groupnorm_act_workspace_bytes(int N, int HW, int C, int G) {
  const int64_t slabs = (HW + fd::GN_ROWS - 1) / fd::GN_ROWS;
  return (static_cast<int64_t>(N) * slabs * (C / 2) + static_cast<int64_t>(N) * G) * 8;
}

extern "C" int fd_groupnorm_act(const void* x_bf16_dev, const void* bias_bf16_dev, const void* gamma_bf16_dev,
                                const void* beta_bf16_dev, void* workspace_dev, void* y_bf16_dev, int N, int HW,
                                int C, int G, float eps, int act_silu, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && gamma_bf16_dev && beta_bf16_dev && workspace_dev && y_bf16_dev,
             "fd_groupnorm_act: NULL pointer");
  FD_REQUIRE(N > 0 && HW > 0 && C > 0, "fd_groupnorm_act: non-positive shape");
  FD_REQUIRE(G > 0 && G <= GN_MAX_GROUPS && C % G == 0 && (C / G) % 2 == 0 && C % 64 == 0,
             "fd_groupnorm_act: need G <= %d, C %% G == 0, even channels per group, C %% 64 == 0 (C=%d G=%d)",
             GN_MAX_GROUPS, C, G);
  const int slabs = (HW + GN_ROWS - 1) / GN_ROWS;
  FD_REQUIRE(N <= 65535 && slabs <= 65535, "fd_groupnorm_act: shape exceeds grid limits");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(workspace_dev) % 8 == 0, "fd_groupnorm_act: workspace must be 8-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  GnArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x_bf16_dev);
  a.bias = static_cast<const __nv_bfloat16*>(bias_bf16_dev);
  a.gamma = static_cast<const __nv_bfloat16*>(gamma_bf16_dev);
  a.beta = static_cast<const __nv_bfloat16*>(beta_bf16_dev);
  a.partial = static_cast<float2*>(workspace_dev);
  a.stats = a.partial + static_cast<size_t>(N) * slabs * (C / 2);
  a.y = static_cast<__nv_bfloat16*>(y_bf16_dev);
  a.HW = HW;
  a.C = C;
  a.G = G;
  a.slabs = slabs;
  a.eps = eps;
  a.act_silu = act_silu;
  dim3 grid(C / 64, slabs, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k5_gn_stats_kernel<<<grid, GN_THREADS, 0, st>>>(a);
  k5_gn_finalize_kernel<<<dim3(G, N), 256, 0, st>>>(a);
  k5_gn_apply_kernel<<<grid, GN_THREADS, 0, st>>>(a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_add_bias_residual(const void* x_bf16_dev, const void* h_bf16_dev, const void* bias_bf16_dev,
                                    void* y_bf16_dev, int64_t n_elem, int C, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && h_bf16_dev && bias_bf16_dev && y_bf16_dev, "fd_add_bias_residual: NULL pointer");
  FD_REQUIRE(C > 0 && C % 8 == 0 && n_elem > 0 && n_elem % C == 0,
             "fd_add_bias_residual: need C %% 8 == 0 and n_elem a multiple of C");
  auto mis = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 != 0; };
  FD_REQUIRE(!mis(x_bf16_dev) && !mis(h_bf16_dev) && !mis(bias_bf16_dev) && !mis(y_bf16_dev),
             "fd_add_bias_residual: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_add_bias_residual: cannot query SM count");
  const int64_t total8 = n_elem / 8;
  int64_t want = (total8 + 255) / 256, cap = static_cast<int64_t>(sms) * 8;
  k7_add_bias_residual_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0,
                                static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x_bf16_dev), static_cast<const uint4*>(h_bf16_dev),
      static_cast<const uint4*>(bias_bf16_dev), static_cast<uint4*>(y_bf16_dev), total8, C / 8);
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
