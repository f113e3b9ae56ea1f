// K15 -- data-movement glue of the UNet's up path (fd_concat_channels, fd_upsample_nearest2x).
// The skip-connection concat of the UNet's up blocks, `torch.cat([x, skip], dim=1)` inside diffusers'
// CrossAttnUpBlock2D / UpBlock2D (reached from the UNet call at pipeline/guide.py:56-58), on channels-last bf16 tensors:
//   y[p, 0:Ca] = a[p, :],  y[p, Ca:Ca+Cb] = b[p, :]        p = pixel (n, h, w)
// ATen's CatArrayBatchedCopy moves it at ~2 TB/s (15 us for the 960-channel 64 x 64 case at two samples, 12 launches = 2.3 % of
// a B = 1 step, 3 % at B = 8).  Here one thread moves one 16-byte vector of the OUTPUT row, so stores are fully coalesced
// and each load is a coalesced run of one source row; grid-stride, 4 vectors in flight per thread.
#include "fd_common.cuh"

namespace fd {
namespace {

__global__ void __launch_bounds__(256) k15_concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                         uint4* __restrict__ y, uint32_t total, uint32_t va, uint32_t vb) {
  const uint32_t vy = va + vb;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  auto src = [&](uint32_t k) -> const uint4* {   // 32-bit index arithmetic (the host checks total < 2^31)
    const uint32_t p = k / vy, c = k - p * vy;
    return c < va ? a + static_cast<size_t>(p) * va + c : b + static_cast<size_t>(p) * vb + (c - va);
  };
  for (; static_cast<uint64_t>(i) + 3ull * stride < total; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *src(i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) y[i + u * stride] = v[u];
  }
  for (; i < total; i += stride) y[i] = *src(i);
}

// Nearest-neighbour 2x upsample of a channels-last bf16 activation (diffusers Upsample2D: F.interpolate(scale_factor=2,
// mode="nearest") before its convolution).  One thread per 16-byte vector of the INPUT: read once, written to the four
// output pixels it covers (each store coalesced along the channels).  ATen's upsample_nearest2d_nhwc ran at 0.8 TB/s
// (132 us for the 640-channel 32 x 32 -> 64 x 64 case at 16 samples).
__global__ void __launch_bounds__(256) k15_upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint32_t total,
                                                             uint32_t vc, uint32_t H, uint32_t W) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint32_t p = i / vc, c = i - p * vc;          // input pixel (n, h, w) flattened, vector column
    const uint32_t w = p % W, nh = p / W;               // nh = n * H + h
    const uint4 v = x[i];
    const size_t o = (static_cast<size_t>(2 * nh) * (2 * W) + 2 * w) * vc + c;   // output pixel (n, 2h, 2w)
    const size_t row = static_cast<size_t>(2 * W) * vc;
    y[o] = v;
    y[o + vc] = v;
    y[o + row] = v;
    y[o + row + vc] = v;
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_upsample_nearest2x(const void* x_bf16_dev, void* y_bf16_dev, int N, int H, int W, int C, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && y_bf16_dev, "fd_upsample_nearest2x: NULL pointer");
  FD_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "fd_upsample_nearest2x: need positive sizes and C %% 8 == 0 (C=%d)", C);
  FD_REQUIRE(reinterpret_cast<uintptr_t>(x_bf16_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(y_bf16_dev) % 16 == 0,
             "fd_upsample_nearest2x: pointers must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(N) * H * W * (C / 8);
  FD_REQUIRE(total < (1ll << 29), "fd_upsample_nearest2x: tensor too large (%lld 16-byte vectors)", static_cast<long long>(total));
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_upsample_nearest2x: cannot query SM count");
  const int64_t want = (total + 255) / 256, cap = static_cast<int64_t>(sms) * 16;
  k15_upsample2x_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x_bf16_dev), static_cast<uint4*>(y_bf16_dev), static_cast<uint32_t>(total),
      static_cast<uint32_t>(C / 8), static_cast<uint32_t>(H), static_cast<uint32_t>(W));
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_concat_channels(const void* a_bf16_dev, const void* b_bf16_dev, void* y_bf16_dev, int64_t pixels, int Ca,
                                  int Cb, void* stream) {
  using namespace fd;
  FD_REQUIRE(a_bf16_dev && b_bf16_dev && y_bf16_dev, "fd_concat_channels: NULL pointer");
  FD_REQUIRE(pixels > 0 && Ca > 0 && Cb > 0 && Ca % 8 == 0 && Cb % 8 == 0,
             "fd_concat_channels: need pixels > 0 and channel counts that are multiples of 8 (Ca=%d Cb=%d)", Ca, Cb);
  auto mis = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 != 0; };
  FD_REQUIRE(!mis(a_bf16_dev) && !mis(b_bf16_dev) && !mis(y_bf16_dev), "fd_concat_channels: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_concat_channels: cannot query SM count");
  const int64_t total = pixels * ((Ca + Cb) / 8);
  FD_REQUIRE(total < (1ll << 31), "fd_concat_channels: tensor too large (%lld 16-byte vectors)", static_cast<long long>(total));
  const int64_t want = (total + 1023) / 1024, cap = static_cast<int64_t>(sms) * 8;
  k15_concat_kernel<<<static_cast<unsigned>(want < cap ? (want < 1 ? 1 : want) : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a_bf16_dev), static_cast<const uint4*>(b_bf16_dev), static_cast<uint4*>(y_bf16_dev), static_cast<uint32_t>(total),
      static_cast<uint32_t>(Ca / 8), static_cast<uint32_t>(Cb / 8));
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
