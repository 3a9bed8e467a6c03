// Error plumbing, version and device gate of the csmae_b200 C-ABI (include/csmae_b200.h).
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace {
thread_local char g_err[512] = "";
}

void csm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* csm_last_error(void) { return g_err; }

extern "C" int csm_version(void) { return 100; }  // 0.1.0

// Refuses anything that is not a Blackwell sm_100 part: there is no fallback path.
extern "C" int csm_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    csm_set_error("csm_device_check: cudaGetDeviceProperties(%d) failed: %s", device, cudaGetErrorString(e));
    return CSM_ERR_CUDA;
  }
  if (prop.major != 10 || prop.minor != 0) {
    csm_set_error("csm_device_check: device %d is sm_%d%d; this library contains sm_100a code only", device,
                  prop.major, prop.minor);
    return CSM_ERR_DEVICE;
  }
  return prop.multiProcessorCount;
}

// Zeroes device memory with a memset node (no kernel): loss accumulators, the flat gradient buffer.
extern "C" int csm_zero_async(void* ptr, long long nbytes, cudaStream_t stream) {
  CSM_CHECK_ARG(ptr != nullptr && nbytes >= 0, "csm_zero_async: bad arguments");
  cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(nbytes), stream);
  if (e != cudaSuccess) {
    csm_set_error("csm_zero_async: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return CSM_ERR_CUDA;
  }
  return CSM_OK;
}
