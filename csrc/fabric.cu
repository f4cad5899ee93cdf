// NVLink fabric runtime: raw device allocations (outside any caching allocator so they can be
// exported), CUDA-IPC export / import of the center variable + control block, peer-access setup,
// pinned host staging and async copies.  This replaces the reference's socket layer
// (distkeras/networking.py:18-99) on the GPU path: after `dk_ipc_open` a worker process holds a
// plain device pointer into the PS GPU's HBM and every "message" is a load / store / atomic
// issued from inside a kernel.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

extern "C" {

int dk_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int dk_set_device(int dev) {
  DK_HOST_CHECK(cudaSetDevice(dev));
  return 0;
}

int dk_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor, long* total_mem) {
  cudaDeviceProp p;
  DK_HOST_CHECK(cudaGetDeviceProperties(&p, dev));
  *sm_count = p.multiProcessorCount;
  *cc_major = p.major;
  *cc_minor = p.minor;
  *total_mem = static_cast<long>(p.totalGlobalMem);
  return 0;
}

// Raw allocation (cudaMalloc granularity, IPC-exportable), zero-initialised.
int dk_fabric_alloc(long bytes, void** out) {
  void* p = nullptr;
  DK_HOST_CHECK(cudaMalloc(&p, static_cast<size_t>(bytes)));
  DK_HOST_CHECK(cudaMemset(p, 0, static_cast<size_t>(bytes)));
  DK_HOST_CHECK(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

int dk_fabric_free(void* p) {
  DK_HOST_CHECK(cudaFree(p));
  return 0;
}

int dk_ipc_export(void* p, void* handle_out64) {
  cudaIpcMemHandle_t h;
  DK_HOST_CHECK(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "unexpected IPC handle size");
  memcpy(handle_out64, &h, sizeof(h));
  return 0;
}

int dk_ipc_open(const void* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  DK_HOST_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out = p;
  return 0;
}

int dk_ipc_close(void* p) {
  DK_HOST_CHECK(cudaIpcCloseMemHandle(p));
  return 0;
}

int dk_can_access_peer(int dev, int peer) {
  int ok = 0;
  if (cudaDeviceCanAccessPeer(&ok, dev, peer) != cudaSuccess) return 0;
  return ok;
}

// Single-process multi-GPU mode: make `peer`'s memory addressable from kernels on `dev`.
int dk_enable_peer_access(int dev, int peer) {
  int cur = 0;
  DK_HOST_CHECK(cudaGetDevice(&cur));
  DK_HOST_CHECK(cudaSetDevice(dev));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    e = cudaSuccess;
  }
  cudaSetDevice(cur);
  DK_HOST_CHECK(e);
  return 0;
}

int dk_host_alloc_pinned(long bytes, void** out) {
  void* p = nullptr;
  DK_HOST_CHECK(cudaHostAlloc(&p, static_cast<size_t>(bytes), cudaHostAllocPortable));
  *out = p;
  return 0;
}

int dk_host_free_pinned(void* p) {
  DK_HOST_CHECK(cudaFreeHost(p));
  return 0;
}

int dk_host_register(void* p, long bytes) {
  DK_HOST_CHECK(cudaHostRegister(p, static_cast<size_t>(bytes), cudaHostRegisterPortable));
  return 0;
}

int dk_host_unregister(void* p) {
  DK_HOST_CHECK(cudaHostUnregister(p));
  return 0;
}

// kind: 1 = H2D, 2 = D2H, 3 = D2D
int dk_memcpy_async(void* dst, const void* src, long bytes, int kind, void* stream) {
  cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice
                               : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  DK_HOST_CHECK(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), k, (cudaStream_t)stream));
  return 0;
}

int dk_memset_async(void* dst, int value, long bytes, void* stream) {
  DK_HOST_CHECK(cudaMemsetAsync(dst, value, static_cast<size_t>(bytes), (cudaStream_t)stream));
  return 0;
}

int dk_memcpy2d_async(void* dst, long dpitch, const void* src, long spitch, long width, long height,
                      void* stream) {
  DK_HOST_CHECK(cudaMemcpy2DAsync(dst, static_cast<size_t>(dpitch), src, static_cast<size_t>(spitch),
                                  static_cast<size_t>(width), static_cast<size_t>(height),
                                  cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int dk_stream_sync(void* stream) {
  DK_HOST_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int dk_device_sync() {
  DK_HOST_CHECK(cudaDeviceSynchronize());
  return 0;
}

const char* dk_build_info() { return "distkeras_b200 native runtime: sm_100a, tcgen05/TMEM/TMA"; }

}  // extern "C"
