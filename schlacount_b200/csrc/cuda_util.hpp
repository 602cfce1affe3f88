// CUDA helpers (product code).
#pragma once
#include <cuda_runtime.h>

#include <time.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>

#include "common.hpp"

namespace hlac {

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess)
    throw Error(HLAC_ERR_CUDA, std::string(what) + " failed: " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
}
#define HLAC_CUDA(x) ::hlac::cuda_check((x), #x, __FILE__, __LINE__)

// Device memory comes from a process-wide cache of freed blocks (per device, best fit within 25 % slack): a genotyping pass
// makes ~80 allocations of scratch and table memory and cudaMalloc + cudaFree cost 0.15 ms each on a B200 box (measured:
// 25 ms of a 140 ms pass), with occasional multi-millisecond outliers. A block returns to the cache only after its device
// has drained (what cudaFree would have waited for as well), so a cached block is never still in use when it is handed out
// again. The cache holds at most HLAC_ALLOC_CACHE_MB (default 8192) per process, is emptied when an allocation fails, and
// is switched off by HLAC_ALLOC_CACHE_MB=0.
struct AllocStats { unsigned long long n_malloc = 0, n_free = 0, n_reused = 0; double t_malloc = 0, t_free = 0, max_malloc = 0, max_free = 0; };
struct AllocCache {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;   // (device, bytes) -> block
  std::unordered_map<void*, std::pair<int, size_t>> live;     // block -> (device, bytes)
  size_t cached_bytes = 0, limit = (size_t)8192 << 20;
  bool limit_read = false, poison = false;   // HLAC_ALLOC_POISON=1: every block is handed out filled with 0xA5
  AllocStats st;
};
// never destroyed: buffers may still be freed while the process is shutting down (static destructors of other objects)
inline AllocCache& alloc_cache() { static AllocCache* const c = new AllocCache; return *c; }
inline double alloc_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
inline void alloc_cache_trim_locked(AllocCache& c, size_t keep_bytes) {   // frees cached blocks (largest first per key order is not needed: any)
  while (c.cached_bytes > keep_bytes && !c.free_blocks.empty()) {
    auto it = std::prev(c.free_blocks.end());
    int prev = 0; cudaGetDevice(&prev);
    if (prev != it->first.first) cudaSetDevice(it->first.first);
    cudaFree(it->second);
    if (prev != it->first.first) cudaSetDevice(prev);
    c.cached_bytes -= it->first.second;
    c.free_blocks.erase(it);
  }
}
inline cudaError_t dev_malloc(void** p, size_t bytes) {
  AllocCache& c = alloc_cache();
  const double t0 = alloc_now();
  std::lock_guard<std::mutex> lk(c.mu);
  if (!c.limit_read) { c.limit_read = true; if (const char* e = getenv("HLAC_ALLOC_CACHE_MB")) c.limit = (size_t)strtoull(e, nullptr, 10) << 20; c.poison = getenv("HLAC_ALLOC_POISON") != nullptr; }
  int dev = 0; cudaGetDevice(&dev);
  bytes = bytes < ((size_t)1 << 20) ? ((bytes + 255) & ~(size_t)255) : ((bytes + 65535) & ~(size_t)65535);
  if (bytes == 0) bytes = 256;
  cudaError_t e = cudaSuccess;
  auto it = c.free_blocks.lower_bound({dev, bytes});
  if (it != c.free_blocks.end() && it->first.first == dev && it->first.second <= bytes + bytes / 4) {
    *p = it->second; bytes = it->first.second;
    c.cached_bytes -= bytes; c.free_blocks.erase(it); c.st.n_reused++;
  } else {
    e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) { (void)cudaGetLastError(); alloc_cache_trim_locked(c, 0); e = cudaMalloc(p, bytes); }
  }
  if (e == cudaSuccess) c.live[*p] = {dev, bytes};
  if (e == cudaSuccess && c.poison) { cudaDeviceSynchronize(); cudaMemset(*p, 0xA5, bytes); cudaDeviceSynchronize(); }   // debugging aid: nothing may rely on zeroed memory
  const double dt = alloc_now() - t0;
  c.st.n_malloc++; c.st.t_malloc += dt; if (dt > c.st.max_malloc) c.st.max_malloc = dt;
  return e;
}
inline void dev_free(void* p) {
  if (!p) return;
  AllocCache& c = alloc_cache();
  const double t0 = alloc_now();
  std::unique_lock<std::mutex> lk(c.mu);
  auto it = c.live.find(p);
  if (it == c.live.end()) { lk.unlock(); cudaFree(p); return; }
  const int dev = it->second.first; const size_t bytes = it->second.second;
  c.live.erase(it);
  lk.unlock();
  int prev = 0; cudaGetDevice(&prev);
  if (prev != dev) cudaSetDevice(dev);
  if (c.limit == 0 || bytes > c.limit) cudaFree(p);
  else {
    cudaDeviceSynchronize();   // nothing on this device still uses the block (cudaFree implies the same wait)
    lk.lock();
    c.free_blocks.insert({{dev, bytes}, p}); c.cached_bytes += bytes;
    if (c.cached_bytes > c.limit) alloc_cache_trim_locked(c, c.limit / 2);
    lk.unlock();
  }
  if (prev != dev) cudaSetDevice(prev);
  const double dt = alloc_now() - t0;
  lk.lock();
  c.st.n_free++; c.st.t_free += dt; if (dt > c.st.max_free) c.st.max_free = dt;
}

// growable device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;   // elements
  ~DevBuf() { if (p) dev_free(p); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void reserve(size_t n, cudaStream_t s = 0, bool keep = false, size_t used = 0) {
    if (n <= cap) return;
    size_t ncap = cap ? cap : 1;
    while (ncap < n) ncap = ncap + ncap / 2 + 1024;
    T* np = nullptr;
    HLAC_CUDA(dev_malloc((void**)&np, ncap * sizeof(T)));
    if (keep && p && used) HLAC_CUDA(cudaMemcpyAsync(np, p, used * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) { HLAC_CUDA(cudaStreamSynchronize(s)); dev_free(p); }
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
