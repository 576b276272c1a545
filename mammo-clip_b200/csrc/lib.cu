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

#include <cudaTypedefs.h>
int mclip_tmap_encode_bf16(CUtensorMap* m, const void* ptr, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                           const unsigned* box, int swizzle128) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = (PFN_cuTensorMapEncodeTiled_v12000)fp;
  }
  if (!enc) { mclip_set_error("cuTensorMapEncodeTiled not available from the driver"); return MCLIP_ERR_CUDA; }
  if ((uintptr_t)ptr & 15) { mclip_set_error("TMA operand %p is not 16-byte aligned", ptr); return MCLIP_ERR_INVALID; }
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    st[i] = strides_bytes[i];
    if (st[i] % 16) { mclip_set_error("TMA stride %llu is not a multiple of 16 bytes", strides_bytes[i]); return MCLIP_ERR_INVALID; }
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (swizzle128 & 1) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   (swizzle128 & 2) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mclip_set_error("cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank); return MCLIP_ERR_CUDA; }
  return MCLIP_OK;
}
