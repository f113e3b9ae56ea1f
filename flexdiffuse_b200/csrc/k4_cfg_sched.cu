// K4 -- classifier-free-guidance combine + DDIM / PLMS / LMS scheduler update, one pass.
//
// Replaces (reference, pure Python over diffusers 0.3.0):
//   pipeline/guide.py:61-63    noise_pred = u + guidance * (c - u)
//   pipeline/flex.py:280-285   latents = scheduler.step(noise_pred, t, latents).prev_sample
//   pipeline/flex.py:270-274   LMS model-input pre-scale  latents / sqrt(sigma^2 + 1)
// which in the reference is 3 + (5..15) tiny elementwise launches per step.
//
// Pure HBM streaming: per element it reads u, c, x (+ up to 3 history tensors, + noise) and
// writes x' (+ eps for the history, + the scaled model input).  DDIM: 4 tensors = 16 B/elem
// fp32 (262,144 B per 4x64x64 sample-step).  128-bit loads/stores, streaming cache
// hints (nothing is re-read inside a launch), grid = multiple of the SM count.
#include "fd_common.cuh"

namespace fd {
namespace {

struct K4Args {
  const void* u;
  const void* c;
  const float* x;
  const float* h1;
  const float* h2;
  const float* h3;
  const float* noise;
  float* x_out;
  float* eps_out;
  void* scaled_out;
  fd_sched_coeffs k;
  int64_t n4;  // number of 4-element groups
};

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 ld_stream_bf4(const __nv_bfloat16* p) {
  uint2 raw = __ldcs(reinterpret_cast<const uint2*>(p));
  float4 r;
  r.x = __uint_as_float(raw.x << 16);
  r.y = __uint_as_float(raw.x & 0xFFFF0000u);
  r.z = __uint_as_float(raw.y << 16);
  r.w = __uint_as_float(raw.y & 0xFFFF0000u);
  return r;
}

template <bool EPS_BF16>
__device__ __forceinline__ float4 ld_eps(const void* base, int64_t i4) {
  if constexpr (EPS_BF16)
    return ld_stream_bf4(reinterpret_cast<const __nv_bfloat16*>(base) + 4 * i4);
  else
    return ld_stream_f4(reinterpret_cast<const float*>(base) + 4 * i4);
}

#define FD_F4_OP(dst, expr)            \
  do {                                 \
    { const int _q = 0; (dst).x = (expr); } \
    { const int _q = 1; (dst).y = (expr); } \
    { const int _q = 2; (dst).z = (expr); } \
    { const int _q = 3; (dst).w = (expr); } \
  } while (0)

__device__ __forceinline__ float f4_get(const float4& v, int q) {
  return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w;
}

// NH = number of history tensors (0..3).  All coefficient branches are grid-uniform.
template <bool EPS_BF16, int NH, bool SCALED_BF16>
__global__ void __launch_bounds__(256) k4_cfg_sched_kernel(const K4Args a) {
  const fd_sched_coeffs k = a.k;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n4;
       i += stride) {
    // issue every load before the first use (memory-level parallelism)
    float4 c = ld_eps<EPS_BF16>(a.c, i);
    float4 u = k.use_cfg ? ld_eps<EPS_BF16>(a.u, i) : c;
    float4 x = ld_stream_f4(a.x + 4 * i);
    float4 h1, h2, h3, nz;
    if constexpr (NH >= 1) h1 = ld_stream_f4(a.h1 + 4 * i);
    if constexpr (NH >= 2) h2 = ld_stream_f4(a.h2 + 4 * i);
    if constexpr (NH >= 3) h3 = ld_stream_f4(a.h3 + 4 * i);
    const bool has_noise = a.noise != nullptr;
    if (has_noise) nz = ld_stream_f4(a.noise + 4 * i);

    float4 eps;
    if (k.use_cfg) {
      // u + g * (c - u), same association as pipeline/guide.py:62-63
      FD_F4_OP(eps, f4_get(u, _q) + k.guidance * (f4_get(c, _q) - f4_get(u, _q)));
    } else {
      eps = c;
    }
    float4 e = eps;
    if constexpr (NH == 0) {
      if (k.w[0] != 1.0f) FD_F4_OP(e, k.w[0] * f4_get(eps, _q));
    }
    if constexpr (NH == 1)
      FD_F4_OP(e, k.w[0] * f4_get(eps, _q) + k.w[1] * f4_get(h1, _q));
    if constexpr (NH == 2)
      FD_F4_OP(e, k.w[0] * f4_get(eps, _q) + k.w[1] * f4_get(h1, _q) + k.w[2] * f4_get(h2, _q));
    if constexpr (NH == 3)
      FD_F4_OP(e, k.w[0] * f4_get(eps, _q) + k.w[1] * f4_get(h1, _q) + k.w[2] * f4_get(h2, _q) +
                      k.w[3] * f4_get(h3, _q));
    float4 xn;
    FD_F4_OP(xn, k.a * f4_get(x, _q) + k.b * f4_get(e, _q));
    if (has_noise) FD_F4_OP(xn, f4_get(xn, _q) + k.c_noise * f4_get(nz, _q));

    __stcs(reinterpret_cast<float4*>(a.x_out + 4 * i), xn);
    if (a.eps_out) __stcs(reinterpret_cast<float4*>(a.eps_out + 4 * i), eps);
    if (a.scaled_out) {
      float4 s;
      FD_F4_OP(s, f4_get(xn, _q) * k.in_scale);
      if constexpr (SCALED_BF16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(s.x, s.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(s.z, s.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        // kept cache-resident: the UNet's conv_in reads it next
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.scaled_out) + 4 * i) = pk;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.scaled_out) + 4 * i) = s;
      }
    }
  }
}

template <bool EPS_BF16, bool SCALED_BF16>
void launch_nh(int nh, dim3 grid, cudaStream_t st, const K4Args& a) {
  switch (nh) {
    case 0: k4_cfg_sched_kernel<EPS_BF16, 0, SCALED_BF16><<<grid, 256, 0, st>>>(a); break;
    case 1: k4_cfg_sched_kernel<EPS_BF16, 1, SCALED_BF16><<<grid, 256, 0, st>>>(a); break;
    case 2: k4_cfg_sched_kernel<EPS_BF16, 2, SCALED_BF16><<<grid, 256, 0, st>>>(a); break;
    default: k4_cfg_sched_kernel<EPS_BF16, 3, SCALED_BF16><<<grid, 256, 0, st>>>(a); break;
  }
}

}  // namespace
}  // namespace fd

namespace fd {
namespace {

constexpr int MAX_ENTITIES = 16;
struct K9Args {
  const void* eps;  // [2 + E, C, H, W]: uncond, background, entities
  float* out_u;     // [C, H, W] fp32 copy of the uncond prediction
  float* out_c;     // [C, H, W] fp32 composite conditional prediction
  int C, H, W, E;
  fd_entity_box box[MAX_ENTITIES];
};

// K9: rectangular per-entity lerp of the entity noise predictions into the background prediction
// (composition/guide.py:66-87), entities applied in declaration order.
template <bool BF16>
__global__ void __launch_bounds__(256) k9_composite_eps_kernel(const K9Args a) {
  const int plane = a.H * a.W, chw = a.C * plane;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chw; i += gridDim.x * blockDim.x) {
    const int pix = i % plane, y = pix / a.W, x = pix - y * a.W;
    auto ld = [&](int sample) -> float {
      if constexpr (BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.eps)[sample * chw + i]);
      else return reinterpret_cast<const float*>(a.eps)[sample * chw + i];
    };
    float v = ld(1);
    for (int e = 0; e < a.E; ++e) {
      const fd_entity_box b = a.box[e];
      if (x >= b.ox && x < b.ox + b.sx && y >= b.oy && y < b.oy + b.sy) v = v + b.blend * (ld(2 + e) - v);
    }
    a.out_u[i] = ld(0);
    a.out_c[i] = v;
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_composite_eps(const void* eps_dev, int eps_dtype, const fd_entity_box* boxes, int n_entities,
                                int C, int H, int W, float* out_uncond_dev, float* out_cond_dev, void* stream) {
  using namespace fd;
  FD_REQUIRE(eps_dev && out_uncond_dev && out_cond_dev, "fd_composite_eps: NULL pointer");
  FD_REQUIRE(n_entities >= 0 && n_entities <= MAX_ENTITIES, "fd_composite_eps: at most %d entities", MAX_ENTITIES);
  FD_REQUIRE(n_entities == 0 || boxes, "fd_composite_eps: boxes is NULL");
  FD_REQUIRE(C > 0 && H > 0 && W > 0, "fd_composite_eps: non-positive shape");
  FD_REQUIRE(eps_dtype == FD_DTYPE_F32 || eps_dtype == FD_DTYPE_BF16, "fd_composite_eps: bad dtype");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  K9Args a;
  a.eps = eps_dev;
  a.out_u = out_uncond_dev;
  a.out_c = out_cond_dev;
  a.C = C;
  a.H = H;
  a.W = W;
  a.E = n_entities;
  for (int e = 0; e < n_entities; ++e) a.box[e] = boxes[e];
  const int chw = C * H * W;
  const int grid = (chw + 255) / 256;
  if (eps_dtype == FD_DTYPE_BF16)
    k9_composite_eps_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  else
    k9_composite_eps_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_cfg_sched_step(const void* eps_uncond_dev, const void* eps_cond_dev,
                                 int eps_dtype, const float* x_dev, const float* h1_dev,
                                 const float* h2_dev, const float* h3_dev,
                                 const float* noise_dev, const fd_sched_coeffs* coeffs,
                                 int64_t n_elem, float* x_out_dev, float* eps_out_dev,
                                 void* scaled_out_dev, int scaled_dtype, void* stream) {
  using namespace fd;
  FD_REQUIRE(coeffs != nullptr, "fd_cfg_sched_step: coeffs is NULL");
  FD_REQUIRE(eps_cond_dev && x_dev && x_out_dev, "fd_cfg_sched_step: NULL eps_cond/x/x_out");
  FD_REQUIRE(!coeffs->use_cfg || eps_uncond_dev, "fd_cfg_sched_step: use_cfg without eps_uncond");
  FD_REQUIRE(n_elem > 0 && n_elem % 4 == 0, "fd_cfg_sched_step: n_elem=%lld must be a positive multiple of 4",
             (long long)n_elem);
  FD_REQUIRE(eps_dtype == FD_DTYPE_F32 || eps_dtype == FD_DTYPE_BF16, "fd_cfg_sched_step: bad eps_dtype %d", eps_dtype);
  FD_REQUIRE(scaled_dtype == FD_DTYPE_F32 || scaled_dtype == FD_DTYPE_BF16, "fd_cfg_sched_step: bad scaled_dtype %d",
             scaled_dtype);
  // history pointers must be a prefix: h2 needs h1, h3 needs h2
  FD_REQUIRE(!(h2_dev && !h1_dev) && !(h3_dev && !h2_dev), "fd_cfg_sched_step: history pointers must be a prefix");
  auto misaligned = [](const void* p, size_t al) { return p && (reinterpret_cast<uintptr_t>(p) % al) != 0; };
  const size_t eps_al = eps_dtype == FD_DTYPE_BF16 ? 8 : 16;
  FD_REQUIRE(!misaligned(eps_uncond_dev, eps_al) && !misaligned(eps_cond_dev, eps_al) && !misaligned(x_dev, 16) &&
                 !misaligned(h1_dev, 16) && !misaligned(h2_dev, 16) && !misaligned(h3_dev, 16) &&
                 !misaligned(noise_dev, 16) && !misaligned(x_out_dev, 16) && !misaligned(eps_out_dev, 16) &&
                 !misaligned(scaled_out_dev, scaled_dtype == FD_DTYPE_BF16 ? 8 : 16),
             "fd_cfg_sched_step: pointers must be 16-byte aligned (8 for bf16)");
  int rc = check_device();
  if (rc != FD_OK) return rc;

  K4Args a;
  a.u = eps_uncond_dev;
  a.c = eps_cond_dev;
  a.x = x_dev;
  a.h1 = h1_dev;
  a.h2 = h2_dev;
  a.h3 = h3_dev;
  a.noise = noise_dev;
  a.x_out = x_out_dev;
  a.eps_out = eps_out_dev;
  a.scaled_out = scaled_out_dev;
  a.k = *coeffs;
  a.n4 = n_elem / 4;
  const int nh = h3_dev ? 3 : h2_dev ? 2 : h1_dev ? 1 : 0;

  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_cfg_sched_step: cannot query SM count");
  // 8 resident CTAs of 256 threads per SM; grid-stride beyond that.
  int64_t want = (a.n4 + 255) / 256;
  int64_t cap = static_cast<int64_t>(sms) * 8;
  dim3 grid(static_cast<unsigned>(want < cap ? want : cap));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool eb = eps_dtype == FD_DTYPE_BF16, sb = scaled_dtype == FD_DTYPE_BF16;
  if (eb && sb) launch_nh<true, true>(nh, grid, st, a);
  else if (eb) launch_nh<true, false>(nh, grid, st, a);
  else if (sb) launch_nh<false, true>(nh, grid, st, a);
  else launch_nh<false, false>(nh, grid, st, a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
