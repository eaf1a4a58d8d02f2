// Shared host plumbing of the C ABI translation units (mpcb_api.cu, nmpc_api.cu): the thread-local error string behind
// mpcb_last_error() and RAII-less device / pinned buffers that grow on demand.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "../../include/mpcb200.h"

namespace mpcb {

int api_fail(int code, const std::string& msg);   // records msg for mpcb_last_error() and returns code

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return mpcb::api_fail(MPCB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

inline bool is_pinned_or_device(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

inline cudaError_t upload(DevBuf<double>& b, const double* src, size_t n) {
  cudaError_t e = b.ensure(n);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(b.p, src, n * sizeof(double), cudaMemcpyHostToDevice);
}

}  // namespace mpcb
