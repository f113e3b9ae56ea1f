// Library-level entry points of the C ABI (include/flexdiffuse_b200.h): version, error string,
// architecture gate, tensor-map encoding helper.  No compute lives here.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

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
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(FD_ERR_ARCH, "no CUDA device: %s (flexdiffuse_b200 has no CPU fallback)",
                     cudaGetErrorString(e));
  }
  return fd_arch_check(dev);
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
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
