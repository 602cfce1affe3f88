// Device-wide exclusive prefix sum of u32 values into u64 offsets (out[n] = total), three small kernels. Used wherever a
// CSR offset array is built on the device (path colour lists, class vectors). Product code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "cuda_util.hpp"

namespace hlac {
namespace scan_detail {
constexpr int T = 256, PER = 8, TILE = T * PER;
static __global__ void __launch_bounds__(T) k_tile_sums(const uint32_t* __restrict__ in, uint64_t n, uint64_t* tile_sum) {
  __shared__ uint64_t w[T / 32];
  const uint64_t base = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * PER;
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < PER; k++) if (base + k < n) s += in[base + k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint64_t t = 0; for (int k = 0; k < T / 32; k++) t += w[k]; tile_sum[blockIdx.x] = t; }
}
// one block: exclusive scan of the tile sums in place; total to tile_sum[n_tiles]
static __global__ void __launch_bounds__(1024) k_scan_tiles(uint64_t* tile_sum, uint32_t n_tiles) {
  __shared__ uint64_t part[1024];
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < n_tiles; b0 += 1024) {
    const uint32_t i = b0 + threadIdx.x;
    const uint64_t v = i < n_tiles ? tile_sum[i] : 0;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const uint64_t add = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n_tiles) tile_sum[i] = carry + part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sum[n_tiles] = carry;
}
static __global__ void __launch_bounds__(T) k_tile_scan(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ tile_off, uint64_t* out) {
  __shared__ uint64_t w[T / 32];
  const uint64_t base = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * PER;
  uint32_t v[PER]; uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < PER; k++) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
  uint64_t incl = s;
  for (int o = 1; o < 32; o <<= 1) { const uint64_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((threadIdx.x & 31) >= (unsigned)o) incl += u; }
  if ((threadIdx.x & 31) == 31) w[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint64_t off = tile_off[blockIdx.x] + incl - s;
  for (unsigned k = 0; k < (threadIdx.x >> 5); k++) off += w[k];
#pragma unroll
  for (int k = 0; k < PER; k++) { if (base + k < n) out[base + k] = off; off += v[k]; }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == T - 1) out[n] = tile_off[gridDim.x];   // total (tile_off[n_tiles])
}
}  // namespace scan_detail

// out must hold n + 1 values; tmp is scratch the caller keeps alive until the stream has run the kernels
inline void exclusive_scan_u32(const uint32_t* in, uint64_t* out, uint64_t n, DevBuf<uint64_t>& tmp, cudaStream_t st) {
  using namespace scan_detail;
  if (n == 0) { HLAC_CUDA(cudaMemsetAsync(out, 0, 8, st)); return; }
  const uint32_t tiles = (uint32_t)((n + TILE - 1) / TILE);
  tmp.reserve(tiles + 1, st);
  k_tile_sums<<<tiles, T, 0, st>>>(in, n, tmp.p);
  k_scan_tiles<<<1, 1024, 0, st>>>(tmp.p, tiles);
  k_tile_scan<<<tiles, T, 0, st>>>(in, n, tmp.p, out);
  HLAC_CUDA(cudaGetLastError());
}
}  // namespace hlac
