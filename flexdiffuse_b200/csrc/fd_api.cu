// Library-level entry points of the C ABI (include/flexdiffuse_b200.h): version, error string,
// architecture gate, tensor-map encoding helper.  No compute lives here.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "fd_common.cuh"

namespace fd {

static thread_local char g_err[512] = "no error";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_device() {
  static thread_local int ok_device = -1;  // last device that passed the gate on this thread
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess && dev == ok_device) return FD_OK;
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(FD_ERR_ARCH, "no CUDA device: %s (flexdiffuse_b200 has no CPU fallback)",
                     cudaGetErrorString(e));
  }
  const int rc = fd_arch_check(dev);
  if (rc == FD_OK) ok_device = dev;
  return rc;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached_n = 0;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev == cached_dev) return cached_n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached_dev = dev;
  cached_n = n;
  return n;
}

namespace {
// Encoding a tensor map costs a few microseconds of host time; the same (pointer, shape) pairs
// recur every denoising step (the caching allocator hands back the same blocks), so keep a small
// thread-local direct-mapped cache.  A tensor map only describes addresses and strides, so a
// stale entry for a re-used pointer with identical geometry is still correct.
struct TmapKey {
  const void* base;
  uint64_t dims[5], strides[4];
  uint32_t box[5], rank, dtype, swizzle;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapEntry {
  bool valid = false;
  TmapKey key;
  CUtensorMap map;
};
constexpr int TMAP_CACHE = 256;
thread_local TmapEntry g_tmap_cache[TMAP_CACHE];
}  // namespace

static int encode_tmap_uncached(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                                const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                CUtensorMapSwizzle swizzle);

int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                CUtensorMapSwizzle swizzle) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.rank = rank;
  key.dtype = static_cast<uint32_t>(dtype);
  key.swizzle = static_cast<uint32_t>(swizzle);
  uint64_t h = reinterpret_cast<uint64_t>(base) * 0x9E3779B97F4A7C15ull;
  for (uint32_t i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i + 1 < rank) key.strides[i] = strides_bytes[i];
    h = (h ^ (dims[i] * 31 + box[i])) * 0x9E3779B97F4A7C15ull;
  }
  TmapEntry& e = g_tmap_cache[(h >> 40) % TMAP_CACHE];
  if (e.valid && e.key == key) {
    *map = e.map;
    return FD_OK;
  }
  int rc = encode_tmap_uncached(map, dtype, rank, base, dims, strides_bytes, box, swizzle);
  if (rc == FD_OK) {
    e.valid = true;
    e.key = key;
    e.map = *map;
  }
  return rc;
}

static int encode_tmap_uncached(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                                const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                CUtensorMapSwizzle swizzle) {
  static thread_local PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e =
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return set_error(FD_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  uint32_t elem_strides[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(map, dtype, rank, const_cast<void*>(base), dims, strides_bytes, box,
                      elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return FD_OK;
}

}  // namespace fd

extern "C" {

int fd_version(void) { return FD_ABI_VERSION; }

const char* fd_last_error_string(void) { return fd::g_err; }

int fd_arch_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fd::set_error(FD_ERR_ARCH,
                         "no CUDA device visible (%s); flexdiffuse_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= n)
    return fd::set_error(FD_ERR_ARCH, "device %d out of range (%d visible)", device, n);
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10)
    return fd::set_error(FD_ERR_ARCH,
                         "device %d is sm_%d%d; flexdiffuse_b200 is built for sm_100a only",
                         device, major, minor);
  return FD_OK;
}

int fd_sm_count(void) { return fd::sm_count(); }

}  // extern "C"
