// build_index on the device ([EXT] debruijn_mapping build_index; call sites src/mapping.rs:652-662, src/hla.rs:258).
//
// The stranded k = 24 coloured compacted de Bruijn graph from the allele sequences, as kernels:
//   1  k_enum            one (k-mer, tx << 16 | left-ext mask << 8 | right-ext mask) pair per k-mer occurrence
//   2  radix sort        stable, by k-mer (radix_sort.cuh; occurrences are generated in ascending tx order, so every k-mer's tx list
//                        comes out sorted)
//   3  k_heads + scan    distinct k-mers; k_uniq_summary (one warp per k-mer): ext masks, size and hash of its de-duplicated tx list
//   4  colours           k_col_insert: hash table keyed by (list hash, list size), per slot the smallest k-mer rank that uses it;
//                        the k-mers that define a colour, scanned in k-mer order, give the colour ids in first-appearance order
//                        (the reference's numbering); k_col_fill materialises every colour's tx list from its defining k-mer,
//                        k_col_verify compares EVERY k-mer's list with its colour element by element (a hash collision aborts)
//   5  compaction        k_links: a -> b is an interior link iff a has one right extension, b one left extension and both one
//                        colour; pointer jumping (k_jump) gives every k-mer its node head and its distance from it; heads scanned in
//                        k-mer order give the node ids (the reference's order); k_node_fill writes the node sequences 2-bit packed
//   6  finalize          k_adj (edge tables: binary search of the extended terminal k-mers), k_dict (exact dictionary, CAS insert),
//                        k_lmers (+ popcount) for the presence bitmap of the skip-scan seed filter
// Everything lands in a HostIndex (the .idx writer and the inspection calls work on it) that hlac_index::upload then makes resident.
// Declines (returns false, the caller builds on the host) for tiny inputs and for graphs with closed cycles of interior links.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "cuda_util.hpp"
#include "device_index.hpp"
#include "host_index.hpp"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace hlac {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
  return x;
}

// ---- 1: occurrences ---------------------------------------------------------------------------------------------------
__global__ void k_enum(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ seq_off, const uint64_t* __restrict__ first,
                       uint32_t n_tx, uint64_t* keys, uint64_t* vals) {
  for (uint32_t t = blockIdx.x; t < n_tx; t += gridDim.x) {
    const uint64_t len = seq_off[t + 1] - seq_off[t];
    if (len < (uint64_t)K) continue;
    const uint8_t* s = bases + seq_off[t];
    const uint64_t nk = len - K + 1;
    for (uint64_t i = threadIdx.x; i < nk; i += blockDim.x) {
      uint64_t v = 0;
#pragma unroll
      for (int k = 0; k < K; k++) v = (v << 2) | s[i + k];
      keys[first[t] + i] = v;
      vals[first[t] + i] = ((uint64_t)t << 16) | (uint64_t)(i > 0 ? 1u << s[i - 1] : 0u) << 8 | (uint64_t)(i + K < len ? 1u << s[i + K] : 0u);
    }
  }
}

// ---- 3: distinct k-mers -------------------------------------------------------------------------------------------------
__global__ void k_heads(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* flag) {
  const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) flag[o] = (o == 0 || keys[o] != keys[o - 1]) ? 1u : 0u;
}
__global__ void k_run_begin(const uint32_t* __restrict__ flag, const uint64_t* __restrict__ excl, uint64_t n, uint64_t nu, uint64_t* ub) {
  const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n && flag[o]) ub[excl[o]] = o;
  if (o == 0) ub[nu] = n;
}
// one warp per distinct k-mer: ext masks, de-duplicated tx list (its size, a hash of the set), and per occurrence the
// "first occurrence of this tx within the k-mer" flag used to lay out the colour lists
__global__ void k_uniq_summary(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, const uint64_t* __restrict__ ub, uint64_t nu,
                               uint64_t* uk, uint8_t* ul, uint8_t* ur, uint32_t* untx, uint64_t* uhash, uint32_t* dflag) {
  const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (q >= nu) return;
  const uint64_t b = ub[q], e = ub[q + 1];
  uint32_t l = 0, r = 0, cnt = 0;
  uint64_t h = 0;
  for (uint64_t o = b + lane; o < e; o += 32) {
    const uint64_t v = vals[o];
    l |= (uint32_t)(v >> 8) & 0xFFu; r |= (uint32_t)v & 0xFFu;
    const uint32_t tx = (uint32_t)(v >> 16);
    const bool d = o == b || (uint32_t)(vals[o - 1] >> 16) != tx;
    dflag[o] = d ? 1u : 0u;
    if (d) { cnt++; h += mix64((uint64_t)tx + 0x9E3779B97F4A7C15ULL); }
  }
  for (int s = 16; s > 0; s >>= 1) {
    l |= __shfl_down_sync(0xFFFFFFFFu, l, s); r |= __shfl_down_sync(0xFFFFFFFFu, r, s);
    cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, s); h += __shfl_down_sync(0xFFFFFFFFu, h, s);
  }
  if (lane == 0) { uk[q] = keys[b]; ul[q] = (uint8_t)l; ur[q] = (uint8_t)r; untx[q] = cnt; uhash[q] = h; }
}

// ---- 4: colours -------------------------------------------------------------------------------------------------------
__global__ void k_col_insert(const uint64_t* __restrict__ uhash, const uint32_t* __restrict__ untx, uint64_t nu, unsigned long long* tkeys, uint32_t* tmin,
                             uint64_t mask, uint32_t* slot) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nu) return;
  const unsigned long long key = mix64(uhash[q] ^ ((uint64_t)untx[q] * 0xD1B54A32D192ED03ULL)) | 1ULL;   // 0 marks an empty slot
  uint64_t h = mix64(key) & mask;
  for (;;) {
    unsigned long long cur = tkeys[h];
    if (cur == 0) { cur = atomicCAS(tkeys + h, 0ULL, key); if (cur == 0) cur = key; }
    if (cur == key) break;
    h = (h + 1) & mask;
  }
  atomicMin(tmin + h, (uint32_t)q);
  slot[q] = (uint32_t)h;
}
__global__ void k_col_definers(const uint32_t* __restrict__ tmin, const uint32_t* __restrict__ slot, uint64_t nu, uint32_t* isdef) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nu) isdef[q] = tmin[slot[q]] == (uint32_t)q ? 1u : 0u;
}
__global__ void k_col_assign(const uint32_t* __restrict__ tmin, const uint32_t* __restrict__ slot, const uint32_t* __restrict__ isdef,
                             const uint64_t* __restrict__ def_excl, const uint32_t* __restrict__ untx, uint64_t nu, uint32_t* ucol, uint32_t* col_len) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nu) return;
  const uint32_t c = (uint32_t)def_excl[tmin[slot[q]]];
  ucol[q] = c;
  if (isdef[q]) col_len[c] = untx[q];
}
// occurrence o of k-mer q is the pos-th distinct tx of that k-mer
__global__ void k_col_fill(const uint64_t* __restrict__ vals, const uint32_t* __restrict__ dflag, const uint64_t* __restrict__ dscan,
                           const uint32_t* __restrict__ hflag, const uint64_t* __restrict__ hexcl, const uint64_t* __restrict__ ub,
                           const uint32_t* __restrict__ isdef, const uint32_t* __restrict__ ucol, const uint64_t* __restrict__ col_off, uint64_t n,
                           uint32_t* col_tx) {
  const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n || !dflag[o]) return;
  const uint64_t q = hexcl[o] + hflag[o] - 1;
  if (!isdef[q]) return;
  col_tx[col_off[ucol[q]] + (dscan[o] - dscan[ub[q]])] = (uint32_t)(vals[o] >> 16);
}
__global__ void k_col_verify(const uint64_t* __restrict__ vals, const uint32_t* __restrict__ dflag, const uint64_t* __restrict__ dscan,
                             const uint32_t* __restrict__ hflag, const uint64_t* __restrict__ hexcl, const uint64_t* __restrict__ ub,
                             const uint32_t* __restrict__ isdef, const uint32_t* __restrict__ ucol, const uint64_t* __restrict__ col_off,
                             const uint32_t* __restrict__ untx, const uint32_t* __restrict__ col_tx, uint64_t n, unsigned long long* bad) {
  const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n || !dflag[o]) return;
  const uint64_t q = hexcl[o] + hflag[o] - 1;
  if (isdef[q]) return;
  const uint32_t c = ucol[q];
  const uint64_t pos = dscan[o] - dscan[ub[q]], len = col_off[c + 1] - col_off[c];
  if (untx[q] != len || pos >= len || col_tx[col_off[c] + pos] != (uint32_t)(vals[o] >> 16)) atomicAdd(bad, 1ULL);
}

// ---- 5: compaction ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t find_kmer(const uint64_t* __restrict__ uk, uint64_t nu, uint64_t kmer) {
  uint64_t lo = 0, hi = nu;
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (uk[mid] < kmer) lo = mid + 1; else hi = mid; }
  return (lo < nu && uk[lo] == kmer) ? (int64_t)lo : -1;
}
__device__ __forceinline__ bool one_bit(uint32_t m) { return m != 0 && (m & (m - 1)) == 0; }
__global__ void k_links(const uint64_t* __restrict__ uk, const uint8_t* __restrict__ ul, const uint8_t* __restrict__ ur, const uint32_t* __restrict__ ucol,
                        uint64_t nu, uint32_t* link, uint32_t* prev) {
  const uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nu || !one_bit(ur[a])) return;
  const int64_t b = find_kmer(uk, nu, ((uk[a] << 2) | (uint64_t)(__ffs(ur[a]) - 1)) & KMER_MASK);
  if (b < 0 || (uint64_t)b == a || !one_bit(ul[b]) || ucol[b] != ucol[a]) return;
  link[a] = (uint32_t)b; prev[b] = (uint32_t)a;   // b has ONE left extension: a is its only possible predecessor
}
__global__ void k_jump_init(const uint32_t* __restrict__ prev, uint64_t nu, uint32_t* p, uint32_t* d) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nu) return;
  const bool has = prev[x] != NONE32;
  p[x] = has ? prev[x] : (uint32_t)x; d[x] = has ? 1u : 0u;
}
__global__ void k_jump(const uint32_t* __restrict__ p, const uint32_t* __restrict__ d, uint64_t nu, uint32_t* p2, uint32_t* d2, unsigned int* changed) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nu) return;
  const uint32_t px = p[x], ppx = p[px];
  p2[x] = ppx; d2[x] = d[x] + d[px];
  if (ppx != px) *changed = 1u;
}
__global__ void k_head_flags(const uint32_t* __restrict__ prev, uint64_t nu, uint32_t* hf) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x < nu) hf[x] = prev[x] == NONE32 ? 1u : 0u;
}
// every k-mer: its node (= rank of its head among the heads), the node's k-mer count and tail; a k-mer whose chain never
// reaches a head sits on a closed cycle
__global__ void k_node_of(const uint32_t* __restrict__ p, const uint32_t* __restrict__ d, const uint32_t* __restrict__ prev, const uint32_t* __restrict__ link,
                          const uint64_t* __restrict__ head_excl, uint64_t nu, uint32_t* node_of, uint32_t* nkm, uint32_t* tail, unsigned long long* cycles) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nu) return;
  const uint32_t h = p[x];
  if (prev[h] != NONE32) { atomicAdd(cycles, 1ULL); return; }
  const uint32_t nd = (uint32_t)head_excl[h];
  node_of[x] = nd;
  if (link[x] == NONE32) { tail[nd] = (uint32_t)x; nkm[nd] = d[x] + 1; }   // the tail is the farthest k-mer of its node
}
__global__ void k_node_len(const uint32_t* __restrict__ nkm, uint32_t nn, uint32_t* len) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nn) len[i] = nkm[i] + K - 1;
}
__global__ void k_node_fill(const uint64_t* __restrict__ uk, const uint8_t* __restrict__ ul, const uint8_t* __restrict__ ur, const uint32_t* __restrict__ ucol,
                            const uint32_t* __restrict__ d, const uint32_t* __restrict__ node_of, const uint32_t* __restrict__ tail,
                            const uint64_t* __restrict__ start64, uint64_t nu, uint32_t* seq32, uint32_t* start, uint32_t* exts, uint32_t* colour) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nu) return;
  const uint32_t nd = node_of[x];
  const uint64_t s0 = start64[nd];
  auto put = [&](uint64_t g, uint32_t base) { atomicOr(seq32 + (g >> 4), base << (30 - 2 * (g & 15))); };
  if (d[x] == 0) {   // the head writes its whole k-mer and the node's record
    for (int i = 0; i < K; i++) put(s0 + i, (uint32_t)(uk[x] >> (2 * (K - 1 - i))) & 3u);
    start[nd] = (uint32_t)s0; colour[nd] = ucol[x];
    exts[nd] = ((uint32_t)ur[tail[nd]] << 4) | ul[x];
  } else put(s0 + K - 1 + d[x], (uint32_t)uk[x] & 3u);
}

// ---- 6: finalize ----------------------------------------------------------------------------------------------------------
__global__ void k_adj(const uint64_t* __restrict__ uk, const uint32_t* __restrict__ d, const uint32_t* __restrict__ node_of, const uint32_t* __restrict__ nkm,
                      const uint32_t* __restrict__ tail, const uint32_t* __restrict__ exts, const uint64_t* __restrict__ head_of_node, uint64_t nu, uint32_t nn,
                      uint32_t* radj, uint32_t* ladj) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, b = threadIdx.x & 3;
  if (i >= nn) return;
  const uint64_t fk = uk[head_of_node[i]], lk = uk[tail[i]];
  uint32_t r = NONE32, l = NONE32;
  if ((exts[i] >> 4) >> b & 1) {   // successor: the node whose FIRST k-mer is the extended last k-mer
    const int64_t x = find_kmer(uk, nu, ((lk << 2) | b) & KMER_MASK);
    if (x >= 0 && d[x] == 0) r = node_of[x];
  }
  if ((exts[i] & 0xF) >> b & 1) {  // predecessor: the node whose LAST k-mer is the extended first k-mer
    const int64_t x = find_kmer(uk, nu, (fk >> 2) | ((uint64_t)b << (2 * K - 2)));
    if (x >= 0 && d[x] + 1 == nkm[node_of[x]]) l = node_of[x];
  }
  radj[(size_t)i * 4 + b] = r; ladj[(size_t)i * 4 + b] = l;
}
__global__ void k_head_of_node(const uint32_t* __restrict__ d, const uint32_t* __restrict__ node_of, uint64_t nu, uint64_t* head_of_node) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x < nu && d[x] == 0) head_of_node[node_of[x]] = x;
}
__global__ void k_dict(const uint64_t* __restrict__ uk, const uint32_t* __restrict__ d, const uint32_t* __restrict__ node_of, uint64_t nu, uint32_t* dict,
                       uint32_t bits) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nu) return;
  const uint64_t v = uk[x];
  const uint32_t mask = (1u << bits) - 1;
  uint32_t h = (uint32_t)((v * 0x9E3779B97F4A7C15ULL) >> (64 - bits));   // dict_hash
  unsigned long long* slots = reinterpret_cast<unsigned long long*>(dict);   // slot = {lo, hi, node, offset}; (lo, hi) all ones = empty
  for (;;) {
    if (atomicCAS(slots + (size_t)h * 2, ~0ULL, (unsigned long long)v) == ~0ULL) break;
    h = (h + 1) & mask;
  }
  dict[(size_t)h * 4 + 2] = node_of[x]; dict[(size_t)h * 4 + 3] = d[x];
}
// l-mers ending at every node base (l-mers across node boundaries never occur in a read that matches: the filter tests windows
// that lie inside a dictionary k-mer, and every k-mer lies inside one node)
__global__ void k_lmers(const uint32_t* __restrict__ seq32, const uint32_t* __restrict__ start, const uint32_t* __restrict__ len, uint32_t nn, uint32_t l, uint32_t* bitmap) {
  for (uint32_t i = blockIdx.x; i < nn; i += gridDim.x) {
    const uint32_t s = start[i], n = len[i];
    for (uint32_t o = l - 1 + threadIdx.x; o < n; o += blockDim.x) {
      uint32_t v = 0;
      for (uint32_t k = 0; k < l; k++) { const uint64_t g = (uint64_t)s + o + 1 - l + k; v = (v << 2) | ((seq32[g >> 4] >> (30 - 2 * (g & 15))) & 3u); }
      atomicOr(bitmap + (v >> 5), 1u << (v & 31));
    }
  }
}
__global__ void k_popcount(const uint32_t* __restrict__ w, uint64_t n, unsigned long long* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned c = i < n ? __popc(w[i]) : 0;
  for (int s = 16; s > 0; s >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, s);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

template <typename T>
void fetch(std::vector<T>& dst, const T* src, size_t n) {
  dst.resize(n);
  if (n) HLAC_CUDA(cudaMemcpy(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost));
}
inline unsigned blocks(uint64_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace

bool device_build_index(const std::vector<Bases>& seqs, const std::vector<std::string>& names, HostIndex& ix, int device, DeviceIndexParts* keep) {
  const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto tick = [&](const char* what) {
    if (!dbg) return;
    cudaDeviceSynchronize();
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[hlac timing] device build: %-26s %.3f s\n", what, std::chrono::duration<double>(t - t_last).count());
    t_last = t;
  };
  const size_t n_tx = seqs.size();
  std::vector<uint64_t> seq_off(n_tx + 1, 0), first(n_tx + 1, 0);
  for (size_t t = 0; t < n_tx; t++) {
    seq_off[t + 1] = seq_off[t] + seqs[t].size();
    first[t + 1] = first[t] + (seqs[t].size() >= (size_t)K ? seqs[t].size() - K + 1 : 0);
  }
  const uint64_t total = first[n_tx];
  if (total < (1u << 16) || total >= (1ull << 32) || n_tx >= (1u << 24)) return false;   // small graphs: the host build is instant
  DeviceGuard guard(device);
  cudaStream_t st = 0;
  // ---- 1: occurrences ----
  std::vector<uint8_t> flat(seq_off[n_tx]);
  for (size_t t = 0; t < n_tx; t++) if (!seqs[t].empty()) memcpy(flat.data() + seq_off[t], seqs[t].data(), seqs[t].size());
  DevBuf<uint8_t> d_bases; DevBuf<uint64_t> d_seq_off, d_first, k0, k1, v0, v1, scan_tmp;
  d_bases.upload(flat.data(), std::max<size_t>(flat.size(), 1), st); d_seq_off.upload(seq_off.data(), n_tx + 1, st); d_first.upload(first.data(), n_tx + 1, st);
  k0.reserve(total, st); k1.reserve(total, st); v0.reserve(total, st); v1.reserve(total, st);
  k_enum<<<(unsigned)std::min<size_t>(n_tx, 65535), 256, 0, st>>>(d_bases.p, d_seq_off.p, d_first.p, (uint32_t)n_tx, k0.p, v0.p);
  HLAC_CUDA(cudaGetLastError());
  tick("upload + enumerate");
  // ---- 2: stable sort by k-mer ----
  RadixSortScratch rs;
  const bool second = radix_sort_pairs(k0.p, k1.p, v0.p, v1.p, total, 0, 2 * K, rs, st);
  const uint64_t* keys = second ? k1.p : k0.p; const uint64_t* vals = second ? v1.p : v0.p;
  tick("sort occurrences");
  // ---- 3: distinct k-mers ----
  DevBuf<uint32_t> hflag, dflag; DevBuf<uint64_t> hexcl, dscan;
  hflag.reserve(total, st); dflag.reserve(total, st); hexcl.reserve(total + 1, st); dscan.reserve(total + 1, st);
  k_heads<<<blocks(total), 256, 0, st>>>(keys, total, hflag.p);
  exclusive_scan_u32(hflag.p, hexcl.p, total, scan_tmp, st);
  uint64_t nu = 0;
  HLAC_CUDA(cudaMemcpyAsync(&nu, hexcl.p + total, 8, cudaMemcpyDeviceToHost, st));
  HLAC_CUDA(cudaStreamSynchronize(st));
  if (nu >= 0xFFFFFFF0ull) return false;
  DevBuf<uint64_t> ub, uk, uhash; DevBuf<uint8_t> ul, ur; DevBuf<uint32_t> untx;
  ub.reserve(nu + 1, st); uk.reserve(nu, st); uhash.reserve(nu, st); ul.reserve(nu, st); ur.reserve(nu, st); untx.reserve(nu, st);
  k_run_begin<<<blocks(total), 256, 0, st>>>(hflag.p, hexcl.p, total, nu, ub.p);
  k_uniq_summary<<<blocks(nu * 32), 256, 0, st>>>(keys, vals, ub.p, nu, uk.p, ul.p, ur.p, untx.p, uhash.p, dflag.p);
  exclusive_scan_u32(dflag.p, dscan.p, total, scan_tmp, st);
  HLAC_CUDA(cudaGetLastError());
  tick("distinct k-mers");
  // ---- 4: colours ----
  uint64_t tcap = 1 << 10;
  while (tcap < 2 * nu) tcap <<= 1;
  DevBuf<unsigned long long> tkeys, ctr; DevBuf<uint32_t> tmin, slot, isdef, ucol, col_len; DevBuf<uint64_t> def_excl, col_off;
  tkeys.reserve(tcap, st); tmin.reserve(tcap, st); slot.reserve(nu, st); isdef.reserve(nu, st); ucol.reserve(nu, st); def_excl.reserve(nu + 1, st); ctr.reserve(4, st);
  HLAC_CUDA(cudaMemsetAsync(tkeys.p, 0, tcap * 8, st)); HLAC_CUDA(cudaMemsetAsync(tmin.p, 0xFF, tcap * 4, st)); HLAC_CUDA(cudaMemsetAsync(ctr.p, 0, 32, st));
  k_col_insert<<<blocks(nu), 256, 0, st>>>(uhash.p, untx.p, nu, tkeys.p, tmin.p, tcap - 1, slot.p);
  k_col_definers<<<blocks(nu), 256, 0, st>>>(tmin.p, slot.p, nu, isdef.p);
  exclusive_scan_u32(isdef.p, def_excl.p, nu, scan_tmp, st);
  uint64_t ncol = 0;
  HLAC_CUDA(cudaMemcpyAsync(&ncol, def_excl.p + nu, 8, cudaMemcpyDeviceToHost, st));
  HLAC_CUDA(cudaStreamSynchronize(st));
  col_len.reserve(ncol, st); col_off.reserve(ncol + 1, st);
  k_col_assign<<<blocks(nu), 256, 0, st>>>(tmin.p, slot.p, isdef.p, def_excl.p, untx.p, nu, ucol.p, col_len.p);
  exclusive_scan_u32(col_len.p, col_off.p, ncol, scan_tmp, st);
  uint64_t ncoltx = 0;
  HLAC_CUDA(cudaMemcpyAsync(&ncoltx, col_off.p + ncol, 8, cudaMemcpyDeviceToHost, st));
  HLAC_CUDA(cudaStreamSynchronize(st));
  DevBuf<uint32_t> col_tx; col_tx.reserve(std::max<uint64_t>(ncoltx, 1), st);
  k_col_fill<<<blocks(total), 256, 0, st>>>(vals, dflag.p, dscan.p, hflag.p, hexcl.p, ub.p, isdef.p, ucol.p, col_off.p, total, col_tx.p);
  k_col_verify<<<blocks(total), 256, 0, st>>>(vals, dflag.p, dscan.p, hflag.p, hexcl.p, ub.p, isdef.p, ucol.p, col_off.p, untx.p, col_tx.p, total, ctr.p);
  HLAC_CUDA(cudaGetLastError());
  tick("colours");
  // ---- 5: compaction ----
  DevBuf<uint32_t> link, prev, pa, pb, da, db, hf, node_of, nkm, tail; DevBuf<uint64_t> head_excl; DevBuf<unsigned int> changed;
  link.reserve(nu, st); prev.reserve(nu, st); pa.reserve(nu, st); pb.reserve(nu, st); da.reserve(nu, st); db.reserve(nu, st); hf.reserve(nu, st);
  node_of.reserve(nu, st); head_excl.reserve(nu + 1, st); changed.reserve(1, st);
  HLAC_CUDA(cudaMemsetAsync(link.p, 0xFF, nu * 4, st)); HLAC_CUDA(cudaMemsetAsync(prev.p, 0xFF, nu * 4, st));
  k_links<<<blocks(nu), 256, 0, st>>>(uk.p, ul.p, ur.p, ucol.p, nu, link.p, prev.p);
  k_jump_init<<<blocks(nu), 256, 0, st>>>(prev.p, nu, pa.p, da.p);
  uint32_t *p = pa.p, *d = da.p, *p2 = pb.p, *d2 = db.p;
  for (int round = 0; round < 40; round++) {
    HLAC_CUDA(cudaMemsetAsync(changed.p, 0, 4, st));
    k_jump<<<blocks(nu), 256, 0, st>>>(p, d, nu, p2, d2, changed.p);
    std::swap(p, p2); std::swap(d, d2);
    unsigned int ch = 0;
    HLAC_CUDA(cudaMemcpyAsync(&ch, changed.p, 4, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    if (!ch) break;
  }
  k_head_flags<<<blocks(nu), 256, 0, st>>>(prev.p, nu, hf.p);
  exclusive_scan_u32(hf.p, head_excl.p, nu, scan_tmp, st);
  uint64_t nn = 0;
  HLAC_CUDA(cudaMemcpyAsync(&nn, head_excl.p + nu, 8, cudaMemcpyDeviceToHost, st));
  HLAC_CUDA(cudaStreamSynchronize(st));
  if (nn == 0 || nn >= 0xFFFFFFF0ull) return false;
  nkm.reserve(nn, st); tail.reserve(nn, st);
  HLAC_CUDA(cudaMemsetAsync(nkm.p, 0, nn * 4, st));
  k_node_of<<<blocks(nu), 256, 0, st>>>(p, d, prev.p, link.p, head_excl.p, nu, node_of.p, nkm.p, tail.p, ctr.p + 1);
  unsigned long long c2[2] = {0, 0};
  HLAC_CUDA(cudaMemcpyAsync(c2, ctr.p, 16, cudaMemcpyDeviceToHost, st));
  HLAC_CUDA(cudaStreamSynchronize(st));
  if (c2[0]) throw Error(HLAC_ERR_STATE, "tx-list hash collision while interning colours");
  if (c2[1]) return false;   // closed cycles of interior links: the host build numbers them after the open paths
  DevBuf<uint32_t> nlen, nstart, nexts, ncolour, seq32; DevBuf<uint64_t> start64, head_of_node;
  nlen.reserve(nn, st); nstart.reserve(nn, st); nexts.reserve(nn, st); ncolour.reserve(nn, st); start64.reserve(nn + 1, st); head_of_node.reserve(nn, st);
  k_node_len<<<blocks(nn), 256, 0, st>>>(nkm.p, (uint32_t)nn, nlen.p);
  exclusive_scan_u32(nlen.p, start64.p, nn, scan_tmp, st);
  uint64_t n_bases = 0;
  HLAC_CUDA(cudaMemcpyAsync(&n_bases, start64.p + nn, 8, cudaMemcpyDeviceToHost, st));
  HLAC_CUDA(cudaStreamSynchronize(st));
  if (n_bases >= (1ULL << 32)) throw Error(HLAC_ERR_LIMIT, "index exceeds 2^32 bases");
  const size_t seq_words = (n_bases + 15) / 16 + 4;   // padding so that 3-word window reads never run off the end
  seq32.reserve(seq_words, st);
  HLAC_CUDA(cudaMemsetAsync(seq32.p, 0, seq_words * 4, st));
  k_node_fill<<<blocks(nu), 256, 0, st>>>(uk.p, ul.p, ur.p, ucol.p, d, node_of.p, tail.p, start64.p, nu, seq32.p, nstart.p, nexts.p, ncolour.p);
  HLAC_CUDA(cudaGetLastError());
  tick("compact nodes");
  // ---- 6: finalize ----
  DevBuf<uint32_t> radj, ladj, dict, bitmap;
  radj.reserve(nn * 4, st); ladj.reserve(nn * 4, st);
  k_head_of_node<<<blocks(nu), 256, 0, st>>>(d, node_of.p, nu, head_of_node.p);
  k_adj<<<blocks(nn * 4), 256, 0, st>>>(uk.p, d, node_of.p, nkm.p, tail.p, nexts.p, head_of_node.p, nu, (uint32_t)nn, radj.p, ladj.p);
  uint32_t dict_bits = 4;
  while ((1ULL << dict_bits) < nu * 2 + 2) dict_bits++;
  if (dict_bits > 31) throw Error(HLAC_ERR_LIMIT, "too many k-mers for the dictionary");
  dict.reserve((size_t)4 << dict_bits, st);
  HLAC_CUDA(cudaMemsetAsync(dict.p, 0xFF, ((size_t)16 << dict_bits), st));
  k_dict<<<blocks(nu), 256, 0, st>>>(uk.p, d, node_of.p, nu, dict.p, dict_bits);
  uint32_t lmer = 8;
  for (;; lmer++) {   // smallest l in 8..14 with density <= 1/16
    const uint64_t bits = 1ULL << (2 * lmer);
    bitmap.reserve(bits / 32, st);
    HLAC_CUDA(cudaMemsetAsync(bitmap.p, 0, bits / 8, st));
    HLAC_CUDA(cudaMemsetAsync(ctr.p + 2, 0, 8, st));
    k_lmers<<<(unsigned)std::min<uint64_t>(nn, 65535), 128, 0, st>>>(seq32.p, nstart.p, nlen.p, (uint32_t)nn, lmer, bitmap.p);
    k_popcount<<<blocks(bits / 32), 256, 0, st>>>(bitmap.p, bits / 32, ctr.p + 2);
    unsigned long long distinct = 0;
    HLAC_CUDA(cudaMemcpyAsync(&distinct, ctr.p + 2, 8, cudaMemcpyDeviceToHost, st));
    HLAC_CUDA(cudaStreamSynchronize(st));
    if (distinct * 16 <= bits || lmer == 14) break;
  }
  HLAC_CUDA(cudaGetLastError());
  tick("finalize (adj/dict/bitmap)");
  // ---- into the HostIndex ----
  ix = HostIndex();
  ix.tx_names = names;
  ix.n_bases = n_bases; ix.n_kmers = nu; ix.dict_bits = dict_bits; ix.lmer = lmer;
  fetch(ix.seq32, seq32.p, seq_words); fetch(ix.start, nstart.p, nn); fetch(ix.len, nlen.p, nn); fetch(ix.colour, ncolour.p, nn);
  { std::vector<uint32_t> e; fetch(e, nexts.p, nn); ix.exts.resize(nn); for (size_t i = 0; i < nn; i++) ix.exts[i] = (uint8_t)e[i]; }
  fetch(ix.col_off, col_off.p, ncol + 1); fetch(ix.col_tx, col_tx.p, ncoltx);
  fetch(ix.radj, radj.p, nn * 4); fetch(ix.ladj, ladj.p, nn * 4);
  fetch(ix.bitmap, bitmap.p, (size_t)((1ULL << (2 * lmer)) / 32));
  if (keep) {   // the resident copies stay where they are; the dictionary (the largest part, read by kernels only) is not downloaded
    auto give = [](auto& dst, auto& src) { std::swap(dst.p, src.p); std::swap(dst.cap, src.cap); };
    give(keep->seq32, seq32); give(keep->dict, dict); give(keep->bitmap, bitmap); give(keep->col_tx, col_tx); give(keep->col_off, col_off);
    keep->device = device; keep->valid = true;
  } else fetch(ix.dict, dict.p, (size_t)4 << dict_bits);
  tick("download");
  return true;
}

}  // namespace hlac
