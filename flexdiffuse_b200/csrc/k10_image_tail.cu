// K10 -- image post-processing tail of the VAE decode, one pass:
//   (x / 2 + 0.5).clamp(0, 1) -> * 255 -> round -> uint8, NHWC
// Replaces pipeline/flex.py:119-124 (`(image / 2 + 0.5).clamp(0, 1)`, `.cpu().permute(0, 2, 3, 1)`,
// then numpy_to_pil's `(images * 255).round().astype('uint8')`): three elementwise kernels, a fp32
// upcast and a 12-byte-per-pixel device-to-host copy become one kernel and 3 bytes per pixel.
// The decoder output is channels-last, i.e. its memory already IS [B, H, W, 3]: the kernel is a flat
// elementwise map.  Arithmetic follows the reference's dtype order exactly (x / 2 and + 0.5 rounded
// in the tensor's own dtype, the * 255 product in fp32, round-half-even as numpy.round), so the
// bytes equal what the reference path produces from the same decoder output.
#include "fd_common.cuh"

namespace fd {
namespace {

__device__ __forceinline__ uint32_t to_u8(float y01) {
  return static_cast<uint32_t>(__float2int_rn(__fmul_rn(y01, 255.0f)));  // y01 in [0, 1]: no overflow
}
__device__ __forceinline__ float tail_bf16(__nv_bfloat16 x) {
  // bf16 tensor ops: each result rounded to bf16 (x / 2 is exact; + 0.5 rounds)
  const float h = __bfloat162float(x) * 0.5f;
  const float s = __bfloat162float(__float2bfloat16_rn(h + 0.5f));
  return fminf(fmaxf(s, 0.0f), 1.0f);
}
__device__ __forceinline__ float tail_f32(float x) {
  return fminf(fmaxf(__fadd_rn(__fmul_rn(x, 0.5f), 0.5f), 0.0f), 1.0f);
}

template <bool BF16>
__global__ void __launch_bounds__(256) k10_image_tail_kernel(const void* __restrict__ xin, uint8_t* __restrict__ out,
                                                             int64_t n) {
  const int64_t n8 = n / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float y[8];
    if (BF16) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(xin) + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        y[2 * k] = tail_bf16(__ushort_as_bfloat16(static_cast<unsigned short>(w[k] & 0xffffu)));
        y[2 * k + 1] = tail_bf16(__ushort_as_bfloat16(static_cast<unsigned short>(w[k] >> 16)));
      }
    } else {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(xin) + 2 * i);
      const float4 b = __ldcs(reinterpret_cast<const float4*>(xin) + 2 * i + 1);
      const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = tail_f32(x[k]);
    }
    uint2 o;
    o.x = to_u8(y[0]) | (to_u8(y[1]) << 8) | (to_u8(y[2]) << 16) | (to_u8(y[3]) << 24);
    o.y = to_u8(y[4]) | (to_u8(y[5]) << 8) | (to_u8(y[6]) << 16) | (to_u8(y[7]) << 24);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
  // tail (n % 8 elements): first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < n - n8 * 8) {
    const int64_t i = n8 * 8 + threadIdx.x;
    const float y = BF16 ? tail_bf16(reinterpret_cast<const __nv_bfloat16*>(xin)[i])
                         : tail_f32(reinterpret_cast<const float*>(xin)[i]);
    out[i] = static_cast<uint8_t>(to_u8(y));
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_image_tail_u8(const void* x_dev, int dtype, int64_t n_elem, void* out_u8_dev, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_dev && out_u8_dev, "fd_image_tail_u8: NULL pointer");
  FD_REQUIRE(n_elem > 0, "fd_image_tail_u8: n_elem must be positive");
  FD_REQUIRE(dtype == FD_DTYPE_F32 || dtype == FD_DTYPE_BF16, "fd_image_tail_u8: dtype must be f32 or bf16");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(x_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(out_u8_dev) % 8 == 0,
             "fd_image_tail_u8: input must be 16-byte and output 8-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_image_tail_u8: cannot query SM count");
  int64_t blocks = (n_elem / 8 + 255) / 256;
  if (blocks > 8LL * sms) blocks = 8LL * sms;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == FD_DTYPE_BF16)
    k10_image_tail_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x_dev, static_cast<uint8_t*>(out_u8_dev), n_elem);
  else
    k10_image_tail_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x_dev, static_cast<uint8_t*>(out_u8_dev), n_elem);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
