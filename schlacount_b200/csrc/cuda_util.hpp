// CUDA helpers (product code).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "common.hpp"

namespace hlac {

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess)
    throw Error(HLAC_ERR_CUDA, std::string(what) + " failed: " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
}
#define HLAC_CUDA(x) ::hlac::cuda_check((x), #x, __FILE__, __LINE__)

// growable device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;   // elements
  ~DevBuf() { if (p) cudaFree(p); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void reserve(size_t n, cudaStream_t s = 0, bool keep = false, size_t used = 0) {
    if (n <= cap) return;
    size_t ncap = cap ? cap : 1;
    while (ncap < n) ncap = ncap + ncap / 2 + 1024;
    T* np = nullptr;
    HLAC_CUDA(cudaMalloc(&np, ncap * sizeof(T)));
    if (keep && p && used) HLAC_CUDA(cudaMemcpyAsync(np, p, used * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) { HLAC_CUDA(cudaStreamSynchronize(s)); cudaFree(p); }
    p = np; cap = ncap;
  }
  void upload(const T* h, size_t n, cudaStream_t s = 0) {
    reserve(n, s);
    if (n) HLAC_CUDA(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) { HLAC_CUDA(cudaGetDevice(&prev)); HLAC_CUDA(cudaSetDevice(dev)); }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

}  // namespace hlac
