// Stable LSD radix sort of (key, value) pairs on the device, hand-written (8-bit digits). Used by every grouping step of the
// library that needs an order (index build: k-mer occurrences; genotyping finish: (barcode, UMI) records; EM: the transposed
// class table; the bincode counts file; the counting fallback), so that no library kernel runs anywhere in the product.
//
// One pass = three kernels over tiles of 256 threads x 16 elements:
//   k_rs_hist     per-tile digit histogram (warp-private shared-memory counters, match_any to merge equal digits of a warp)
//   exclusive scan of the (digit-major, tile-minor) counts  (scan.cuh)
//   k_rs_scatter  every warp owns a contiguous 512-element sub-tile and walks it 32 elements at a time: an element's position is
//                 global offset of (digit, tile) + elements of that digit in earlier warps of the tile + in earlier rounds of the
//                 warp + among lower lanes of this round (match_any ballot) - which is exactly the stable order.
// The last pass of a bit range that is not a multiple of 8 wide masks its digit, so bits outside [begin_bit, end_bit) never order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "cuda_util.hpp"
#include "scan.cuh"

namespace hlac {
namespace rs_detail {
constexpr int T = 256, PER = 16, TILE = T * PER, WARPS = T / 32, WTILE = 32 * PER;

template <typename K>
static __global__ void __launch_bounds__(T) k_rs_hist(const K* __restrict__ keys, uint64_t n, int shift, unsigned dmask, uint32_t n_tiles, uint32_t* hist /* [256][n_tiles] */) {
  __shared__ uint32_t h[WARPS][256];
  const unsigned t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (unsigned d = lane; d < 256; d += 32) h[w][d] = 0;
  __syncwarp();
  const uint64_t base = (uint64_t)blockIdx.x * TILE + (uint64_t)w * WTILE;
  for (int r = 0; r < PER; r++) {
    const uint64_t i = base + (uint64_t)r * 32 + lane;
    const bool ok = i < n;
    const unsigned d = ok ? (unsigned)(keys[i] >> shift) & dmask : 0x100u;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
    if (ok && (peers & ((1u << lane) - 1)) == 0) h[w][d] += __popc(peers);   // the lowest lane of every digit group adds the group
    __syncwarp();
  }
  __syncthreads();
  uint32_t c = 0;
  for (int k = 0; k < WARPS; k++) c += h[k][t];
  hist[(size_t)t * n_tiles + blockIdx.x] = c;
}

template <typename K, typename V>
static __global__ void __launch_bounds__(T) k_rs_scatter(const K* __restrict__ keys, const V* __restrict__ vals, uint64_t n, int shift, unsigned dmask, uint32_t n_tiles,
                                                          const uint64_t* __restrict__ offs /* [256][n_tiles] exclusive */, K* keys_out, V* vals_out) {
  __shared__ uint32_t h[WARPS][256];     // phase 1: per-warp digit counts; phase 2: exclusive over the warps
  __shared__ uint64_t gbase[256];
  const unsigned t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (unsigned d = lane; d < 256; d += 32) h[w][d] = 0;
  gbase[t] = offs[(size_t)t * n_tiles + blockIdx.x];
  __syncwarp();
  const uint64_t base = (uint64_t)blockIdx.x * TILE + (uint64_t)w * WTILE;
  K key[PER]; uint32_t rank[PER];
#pragma unroll
  for (int r = 0; r < PER; r++) {
    const uint64_t i = base + (uint64_t)r * 32 + lane;
    const bool ok = i < n;
    key[r] = ok ? keys[i] : (K)0;
    const unsigned d = ok ? (unsigned)(key[r] >> shift) & dmask : 0x100u;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
    const unsigned lower = peers & ((1u << lane) - 1);
    uint32_t before = 0;
    if (ok) before = h[w][d];                       // elements of this digit in the warp's earlier rounds
    __syncwarp();
    if (ok && lower == 0) h[w][d] = before + __popc(peers);
    __syncwarp();
    rank[r] = before + __popc(lower);
  }
  __syncthreads();
  {   // exclusive prefix over the warps, per digit (thread t = digit t)
    uint32_t run = 0;
    for (int k = 0; k < WARPS; k++) { const uint32_t c = h[k][t]; h[k][t] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < PER; r++) {
    const uint64_t i = base + (uint64_t)r * 32 + lane;
    if (i < n) {
      const unsigned d = (unsigned)(key[r] >> shift) & dmask;
      const uint64_t pos = gbase[d] + h[w][d] + rank[r];
      keys_out[pos] = key[r];
      if (vals) vals_out[pos] = vals[i];
    }
  }
}
}  // namespace rs_detail

// scratch a sorter keeps between calls
struct RadixSortScratch {
  DevBuf<uint32_t> hist; DevBuf<uint64_t> offs, scan_tmp;
};

// Sorts n pairs by bits [begin_bit, end_bit) of the key, stable. Ping-pongs between (k0, v0) and (k1, v1); returns true if the
// result is in (k1, v1), false if in (k0, v0). vals may be null (keys only: pass V = uint32_t and null pointers).
template <typename K, typename V>
bool radix_sort_pairs(K* k0, K* k1, V* v0, V* v1, uint64_t n, int begin_bit, int end_bit, RadixSortScratch& sc, cudaStream_t st) {
  using namespace rs_detail;
  if (n == 0) return false;
  const uint32_t n_tiles = (uint32_t)((n + TILE - 1) / TILE);
  sc.hist.reserve((size_t)256 * n_tiles, st); sc.offs.reserve((size_t)256 * n_tiles + 1, st);
  bool in_second = false;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    K* ki = in_second ? k1 : k0; K* ko = in_second ? k0 : k1;
    V* vi = in_second ? v1 : v0; V* vo = in_second ? v0 : v1;
    const unsigned dmask = end_bit - shift >= 8 ? 0xFFu : (1u << (end_bit - shift)) - 1u;
    k_rs_hist<K><<<n_tiles, T, 0, st>>>(ki, n, shift, dmask, n_tiles, sc.hist.p);
    exclusive_scan_u32(sc.hist.p, sc.offs.p, (uint64_t)256 * n_tiles, sc.scan_tmp, st);
    k_rs_scatter<K, V><<<n_tiles, T, 0, st>>>(ki, vi, n, shift, dmask, n_tiles, sc.offs.p, ko, vo);
    HLAC_CUDA(cudaGetLastError());
    in_second = !in_second;
  }
  return in_second;
}
}  // namespace hlac
