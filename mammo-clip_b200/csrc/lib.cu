// Library plumbing: error text, version, device check, SM-count cache.
#include "common.cuh"
#include "mclip_internal.h"
#include <stdarg.h>

static thread_local char g_err[1024] = "";

void mclip_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mclip_last_error(void) { return g_err; }
extern "C" int mclip_version(void) { return MCLIP_ABI_VERSION; }

int mclip_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" int mclip_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  MCLIP_CHECK_CUDA(cudaGetDevice(&dev));
  MCLIP_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MCLIP_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    mclip_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, major, minor);
    return MCLIP_ERR_DEVICE;
  }
  return MCLIP_OK;
}
