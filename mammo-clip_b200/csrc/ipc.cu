// Peer-mapped ("symmetric") device buffers for the fused all-gather: plain cudaMalloc + CUDA IPC handles, exchanged by the
// host through torch.distributed (object all-gather).  Replaces the NCCL all_gather / reduce_scatter transport of
// util/dist_autograd.py:10,22 with direct NVLink stores from inside the loss kernel.
#include "common.cuh"
#include "mclip_internal.h"

extern "C" int mclip_ipc_alloc(long long bytes, void** out) {
  MCLIP_REQUIRE(bytes > 0 && out, "mclip_ipc_alloc: bad arguments");
  void* p = nullptr;
  MCLIP_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  MCLIP_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  MCLIP_CHECK_CUDA(cudaDeviceSynchronize());
  *out = p;
  return MCLIP_OK;
}
extern "C" int mclip_ipc_free(void* p) {
  if (p) MCLIP_CHECK_CUDA(cudaFree(p));
  return MCLIP_OK;
}
extern "C" int mclip_ipc_get_handle(void* p, void* handle64) {
  MCLIP_REQUIRE(p && handle64, "mclip_ipc_get_handle: null");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  MCLIP_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  return MCLIP_OK;
}
extern "C" int mclip_ipc_open_handle(const void* handle64, void** out) {
  MCLIP_REQUIRE(handle64 && out, "mclip_ipc_open_handle: null");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  MCLIP_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return MCLIP_OK;
}
extern "C" int mclip_ipc_close_handle(void* p) {
  if (p) MCLIP_CHECK_CUDA(cudaIpcCloseMemHandle(p));
  return MCLIP_OK;
}
