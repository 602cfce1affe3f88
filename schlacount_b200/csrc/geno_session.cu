// Genotyping session: mapping_wrapper's hot loops (src/mapping.rs:337-402) and EqClassDb::{count, eq_class_counts}
// (src/mapping.rs:128-228) on the GPU.
//
// The reference's per-read class is NOT a set: it is the colour list of the last visited node, permuted by the
// no-truncate merge (mapping.rs:283-308 semantics inside the pinned map_read; SURVEY.md finding #2) against every
// other visited node's colour list, in visit order. It is therefore a pure function of the read's ordered colour
// PATH. Lists can be thousands of alleles long, so the kernels never materialise a class per read; they intern paths:
//
//   per push
//   K1g k_geno_map        thread per read: fwd, else revcomp (unless forward_only) walk; emits coverage + two 64-bit
//                         order-sensitive hashes of the colour path
//   K2a k_intern_insert   counted reads (cov > MIN_SCORE_CALL) insert hash A into a global open-addressing table; the
//                         CAS winner allocates a dense path id and registers itself as the path's representative
//   K2b k_geno_paths      one thread per NEW path re-walks its representative read and writes the colour path
//   K2c k_intern_lookup   counted reads look their path id up, atomicMin the path's first-seen ordinal, atomicAdd its
//                         read count, and append a (barcode, umi, path id) record (EqClassDb::count, mapping.rs:128-163)
//   at finish
//   K3a k_resolve_paths   thread per distinct path: exact emulation of the merges -> class vector (+ its hash)
//       host              interns equal vectors, orders classes by first-seen ordinal (= eq_class_id order,
//                         mapping.rs:147-154), sums counts_reads, builds path -> class-rank table
//   K3b radix sort (radix_sort.cuh) + k_umi_min   group records by (barcode, umi); the UMI's class is the lowest class id among its
//                         reads (mapping.rs:165-167,184-193), then sort+dedup on the host per distinct class
//                         (mapping.rs:205-207); counts_umi histogram by atomicAdd
#include <algorithm>
#include <cstdio>
#include <ctime>
#include <map>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

#include "caller.hpp"
#include "comm.hpp"
#include "device_index.hpp"
#include "radix_sort.cuh"
#include "scan.cuh"

using namespace hlac;

namespace {

constexpr int GENO_THREADS = 256;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
  return x;
}

struct HashVis {   // order-sensitive hashes of the colour path: four 32-bit multiplicative accumulators (one IMAD each
                   // per visited node; the 64-bit mixes this replaced were 25 % of k_geno_map's instructions), avalanched
                   // once at the end into the 2 x 64-bit identity of the path
  uint32_t a0 = 0x811c9dc5u, a1 = 0x9e3779b9u, b0 = 0x85ebca6bu, b1 = 0xc2b2ae35u;
  uint64_t a = 0, b = 0;
  uint32_t n = 0;
  __device__ __forceinline__ void visit(uint32_t colour, uint32_t) {
    a0 = (a0 ^ colour) * 0x01000193u;
    a1 = (a1 + colour) * 0xcc9e2d51u + 0x1b873593u;
    b0 = (b0 ^ (colour + n)) * 0x9e3779b1u;
    b1 = (b1 + __funnelshift_l(colour, colour, 13)) * 0x85ebca77u + n;
    n++;
  }
  __device__ __forceinline__ void finish() {
    a = mix64((((uint64_t)a0 << 32) | a1) ^ n); if (a == 0) a = 1;   // 0 marks an empty table slot
    b = mix64((((uint64_t)b0 << 32) | b1) ^ ((uint64_t)n << 32));
  }
};
struct CountVis2 { uint32_t n = 0; __device__ __forceinline__ void visit(uint32_t, uint32_t) { n++; } };
struct WriteVis { uint32_t* out; uint32_t n = 0; __device__ __forceinline__ void visit(uint32_t colour, uint32_t) { out[n++] = colour; } };
struct CompareVis {   // the walk's colour path against a stored one
  const uint32_t* want; uint32_t n_want; uint32_t k = 0; bool ok = true;
  __device__ __forceinline__ void visit(uint32_t colour, uint32_t) { if (k >= n_want || want[k] != colour) ok = false; k++; }
};

struct SharedWords {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator[](int i) const { return p[i]; }
};
struct BitmapGlobal {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return __ldg(p + i); }
};

struct GenoMapArgs {
  DevIndexView ix;
  const uint64_t* seq; const uint16_t* len;
  uint64_t n; uint32_t wpr, nw, row_stride;
  int forward_only;
  uint32_t* cov;            // per read (0 = None)
  uint64_t* hash_a; uint64_t* hash_b;
  uint8_t* strand;          // 0 fwd, 1 revcomp (for the representative re-walk)
  unsigned long long* stats;  // [0] some_aln, [1] long_aln, [2] bitmap tests, [3] dict probes
};

// loads read i of the batch into this thread's shared-memory row (forward strand, zero padded)
__device__ __forceinline__ void load_read_row(const uint64_t* seq, uint64_t i, uint32_t wpr, uint32_t* fw) {
  for (uint32_t w = 0; w < wpr; w++) {
    const uint64_t v = __ldg(seq + i * wpr + w);
    fw[2 * w] = (uint32_t)(v >> 32); fw[2 * w + 1] = (uint32_t)v;
  }
  fw[2 * wpr] = 0; fw[2 * wpr + 1] = 0;
}

// Warp-staged like k_count_map (count_session.cu): each warp owns a tile of 32*GENO_R reads; load -> forward seed scan
// -> (survivors) reverse-complement seed scan -> dense walk queue. mapping.rs:341-363: forward, else reverse
// complement; the unmapped pass (forward_only) never tries the reverse complement (mapping.rs:383).
constexpr int GENO_R = 2;
template <int STRIDE>
__global__ void __launch_bounds__(GENO_THREADS) k_geno_map(const GenoMapArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int TILE = 32 * GENO_R;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per_warp = TILE * a.row_stride + TILE * 2 + TILE / 2 + 3 * TILE / 4;   // u32 words
  uint32_t* rows = smem + warp * per_warp;
  uint32_t* seed_node = rows + TILE * a.row_stride;
  uint32_t* seed_info = seed_node + TILE;                                   // off(20) | pos(10) | strand(1)
  uint16_t* lens = reinterpret_cast<uint16_t*>(seed_info + TILE);
  uint8_t* q = reinterpret_cast<uint8_t*>(lens + TILE);                     // [0] scan fwd, [1] scan rev, [2] walk
  unsigned long long some = 0, lng = 0;
  uint32_t n_tests = 0, n_probes = 0;
  BitmapGlobal bm{a.ix.bitmap};
  const uint64_t n_warps = (uint64_t)gridDim.x * (GENO_THREADS / 32);
  for (uint64_t tile = (uint64_t)blockIdx.x * (GENO_THREADS / 32) + warp; tile * TILE < a.n; tile += n_warps) {
    const uint64_t base = tile * TILE;
    unsigned cnt0 = 0, cnt1 = 0, cntw = 0;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < GENO_R; r++) {
      const unsigned slot = r * 32 + lane;
      const uint64_t i = base + slot;
      const bool valid = i < a.n;
      if (valid) {
        load_read_row(a.seq, i, a.wpr, rows + slot * a.row_stride);
        lens[slot] = (uint16_t)min((uint32_t)a.len[i], 32u * a.wpr);   // never read past the packed words
        a.cov[i] = 0; a.hash_a[i] = 0; a.hash_b[i] = 0; a.strand[i] = 0;
      }
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, valid);
      if (valid) q[cnt0 + __popc(ballot & ((1u << lane) - 1))] = (uint8_t)slot;
      cnt0 += __popc(ballot);
    }
    __syncwarp();
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
      if (p == 1 && a.forward_only) break;
      const unsigned nt = p ? cnt1 : cnt0;
      const uint8_t* qc = q + (p ? TILE : 0);
      for (unsigned t0 = 0; t0 < nt; t0 += 32) {
        const bool task = t0 + lane < nt;
        bool found = false;
        unsigned slot = 0;
        if (task) {
          slot = qc[t0 + lane];
          uint32_t* fw = rows + slot * a.row_stride;
          uint32_t* rc = fw + a.nw;
          const int L = lens[slot];
          if (p == 1) make_revcomp(SharedWords{fw}, rc, L, (int)a.nw);
          int pos = 0; uint32_t node = 0, off = 0;
          if (L >= K) found = find_seed<STRIDE>(a.ix, SharedWords{p ? rc : fw}, bm, pos, L - K, node, off, n_tests, n_probes);
          if (found) { seed_node[slot] = node; seed_info[slot] = (off << 12) | ((uint32_t)pos << 2) | (uint32_t)p; }
        }
        __syncwarp();
        const unsigned bf = __ballot_sync(0xFFFFFFFFu, task && found), bn = __ballot_sync(0xFFFFFFFFu, task && !found);
        const unsigned below = (1u << lane) - 1;
        if (task && found) q[2 * TILE + cntw + __popc(bf & below)] = (uint8_t)slot;
        if (p == 0 && task && !found) q[TILE + cnt1 + __popc(bn & below)] = (uint8_t)slot;
        cntw += __popc(bf);
        if (p == 0) cnt1 += __popc(bn);
      }
      __syncwarp();
    }
    for (unsigned t0 = 0; t0 < cntw; t0 += 32) {
      if (t0 + lane < cntw) {
        const unsigned slot = q[2 * TILE + t0 + lane];
        const uint64_t i = base + slot;
        const uint32_t info = seed_info[slot];
        const int p = (int)(info & 3u), pos = (int)((info >> 2) & 1023u);
        uint32_t* fw = rows + slot * a.row_stride;
        HashVis vis;
        const uint32_t cov = walk_from_seed<STRIDE>(a.ix, GlobalGraph{a.ix.nodes, a.ix.seq32}, SharedWords{p ? fw + a.nw : fw}, bm, (int)lens[slot], pos, seed_node[slot], info >> 12, vis, n_tests, n_probes);
        vis.finish();
        a.cov[i] = cov; a.hash_a[i] = vis.a; a.hash_b[i] = vis.b; a.strand[i] = (uint8_t)p;
        some++;
        if (cov > MIN_SCORE_CALL) lng++;
      }
    }
  }
  unsigned long long m[4] = {some, lng, n_tests, n_probes};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    unsigned long long v = m[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(a.stats + k, v);
  }
}

struct PathTable {
  unsigned long long* keys;   // hash A, 0 = empty
  uint32_t* ids;              // dense path id
  uint64_t mask;
};
struct PathEntries {          // dense, indexed by path id
  uint64_t* hash_a;
  uint64_t* hash_b;
  unsigned long long* first_ord;
  uint32_t* count;
  uint64_t* path_off;
  uint32_t* path_len;
  uint32_t* rep_read;         // batch-local index of the representative read (valid for paths new in this push)
};

__global__ void k_intern_insert(PathTable t, PathEntries e, const uint32_t* __restrict__ cov, const uint64_t* __restrict__ ha,
                                const uint64_t* __restrict__ hb, uint64_t n, unsigned long long* n_entries) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || cov[i] <= MIN_SCORE_CALL) return;
  const unsigned long long key = ha[i];
  uint64_t h = mix64(key) & t.mask;
  for (;;) {
    const unsigned long long cur = t.keys[h];
    if (cur == key) return;
    if (cur == 0) {
      const unsigned long long old = atomicCAS(t.keys + h, 0ULL, key);
      if (old == 0) {   // winner: allocate the dense id and register as representative
        const uint32_t id = (uint32_t)atomicAdd(n_entries, 1ULL);
        t.ids[h] = id;
        e.hash_a[id] = key; e.hash_b[id] = hb[i]; e.first_ord[id] = ~0ULL; e.count[id] = 0; e.rep_read[id] = (uint32_t)i;
        e.path_off[id] = 0; e.path_len[id] = 0;
        return;
      }
      if (old == key) return;
    }
    h = (h + 1) & t.mask;
  }
}

struct GenoPathArgs {
  DevIndexView ix;
  const uint64_t* seq; const uint16_t* len; const uint8_t* strand;
  uint32_t wpr, nw, row_stride;
  PathEntries e;
  uint32_t first_new, n_new;
  uint32_t* arena; unsigned long long* arena_cursor;
};

template <int STRIDE>
__global__ void __launch_bounds__(GENO_THREADS) k_geno_paths(const GenoPathArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* fw = smem + threadIdx.x * a.row_stride;
  uint32_t* rc = fw + a.nw;
  const uint32_t k = blockIdx.x * GENO_THREADS + threadIdx.x;
  if (k >= a.n_new) return;
  const uint32_t id = a.first_new + k;
  const uint64_t i = a.e.rep_read[id];
  load_read_row(a.seq, i, a.wpr, fw);
  const int L = min((int)a.len[i], (int)(32 * a.wpr));
  const bool rev = a.strand[i] != 0;
  if (rev) make_revcomp(SharedWords{fw}, rc, L, (int)a.nw);
  SharedWords rd{rev ? rc : fw};
  uint32_t cov, t0 = 0, t1 = 0;
  CountVis2 cv;
  map_read_walk<STRIDE>(a.ix, rd, BitmapGlobal{a.ix.bitmap}, L, cv, cov, t0, t1);
  const unsigned long long off = atomicAdd(a.arena_cursor, (unsigned long long)cv.n);
  WriteVis wv{a.arena + off};
  map_read_walk<STRIDE>(a.ix, rd, BitmapGlobal{a.ix.bitmap}, L, wv, cov, t0, t1);
  a.e.path_off[id] = off; a.e.path_len[id] = cv.n;
}

// Exactness of the interning: a read joins an interned path when its two 64-bit path hashes match. This kernel removes the
// residual doubt: every counted read is walked again and its colour path compared, id by id, with the stored path it was
// assigned to; a single mismatch fails the push. (One extra walk per counted read; HLAC_GENO_VERIFY=0 skips it.)
template <int STRIDE>
__global__ void __launch_bounds__(GENO_THREADS) k_geno_verify(const GenoPathArgs a, const uint32_t* __restrict__ cov, const uint32_t* __restrict__ pid, uint64_t n,
                                                               unsigned long long* mismatches) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* fw = smem + threadIdx.x * a.row_stride;
  uint32_t* rc = fw + a.nw;
  const uint64_t i = (uint64_t)blockIdx.x * GENO_THREADS + threadIdx.x;
  if (i >= n || cov[i] <= MIN_SCORE_CALL) return;
  const uint32_t id = pid[i];
  load_read_row(a.seq, i, a.wpr, fw);
  const int L = min((int)a.len[i], (int)(32 * a.wpr));
  const bool rev = a.strand[i] != 0;
  if (rev) make_revcomp(SharedWords{fw}, rc, L, (int)a.nw);
  SharedWords rd{rev ? rc : fw};
  uint32_t c2, t0 = 0, t1 = 0;
  CompareVis cv{a.arena + a.e.path_off[id], a.e.path_len[id]};
  const bool some = map_read_walk<STRIDE>(a.ix, rd, BitmapGlobal{a.ix.bitmap}, L, cv, c2, t0, t1);
  if (!some || !cv.ok || cv.k != cv.n_want || c2 != cov[i]) atomicAdd(mismatches, 1ULL);
}

struct Record { uint32_t bc, umi, path; };

__global__ void k_intern_lookup(PathTable t, PathEntries e, const uint32_t* __restrict__ cov, const uint64_t* __restrict__ ha,
                                const uint64_t* __restrict__ hb, const uint32_t* __restrict__ cell, const uint32_t* __restrict__ umi,
                                uint64_t n, uint64_t ord0, const uint64_t* __restrict__ ordinal, uint64_t* rec_key, uint32_t* rec_path, unsigned long long* rec_cursor,
                                uint32_t* read_path, unsigned long long* collisions, uint32_t* batch_pid) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  bool counted = false; uint32_t id = NONE32;
  if (i < n && cov[i] > MIN_SCORE_CALL) {
    counted = true;
    const unsigned long long key = ha[i];
    uint64_t h = mix64(key) & t.mask;
    while (t.keys[h] != key) h = (h + 1) & t.mask;
    id = t.ids[h];
    if (e.hash_b[id] != hb[i]) atomicAdd(collisions, 1ULL);
    atomicMin(e.first_ord + id, (unsigned long long)(ordinal ? ordinal[i] : ord0 + i));
    atomicAdd(e.count + id, 1u);
  }
  if (i < n && read_path) read_path[i] = id;
  if (i < n && batch_pid) batch_pid[i] = id;
  const unsigned ballot = __ballot_sync(0xFFFFFFFFu, counted);
  if (ballot) {
    unsigned long long at = 0;
    if (lane == 0) at = atomicAdd(rec_cursor, (unsigned long long)__popc(ballot));
    at = __shfl_sync(0xFFFFFFFFu, at, 0);
    if (counted) {
      const unsigned long long w = at + __popc(ballot & ((1u << lane) - 1));
      rec_key[w] = ((uint64_t)cell[i] << 32) | umi[i];
      rec_path[w] = id;
    }
  }
}

__global__ void k_rehash(const unsigned long long* __restrict__ okeys, const uint32_t* __restrict__ oids, uint64_t ocap, PathTable t) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ocap || okeys[i] == 0) return;
  const unsigned long long key = okeys[i];
  uint64_t h = mix64(key) & t.mask;
  for (;;) {
    if (atomicCAS(t.keys + h, 0ULL, key) == 0) { t.ids[h] = oids[i]; return; }
    h = (h + 1) & t.mask;
  }
}

// rebuild the table from dense entries (after hlac_geno_import_paths)
__global__ void k_table_build(PathTable t, const uint64_t* __restrict__ hash_a, uint32_t n) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  const unsigned long long key = hash_a[id];
  uint64_t h = mix64(key) & t.mask;
  for (;;) {
    if (atomicCAS(t.keys + h, 0ULL, key) == 0) { t.ids[h] = id; return; }
    h = (h + 1) & t.mask;
  }
}
__global__ void k_remap(uint32_t* ids, uint64_t n, const uint32_t* __restrict__ map) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ids[i] != NONE32) ids[i] = map[ids[i]];
}

// per path: length of its class vector = |colours[last colour of the path]|
__global__ void k_class_len(const uint32_t* __restrict__ arena, const uint64_t* __restrict__ path_off, const uint32_t* __restrict__ path_len,
                            const uint64_t* __restrict__ col_off, uint32_t n, uint32_t* out, uint32_t* last_colour) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  const uint32_t c = arena[path_off[id] + path_len[id] - 1];
  out[id] = (uint32_t)(col_off[c + 1] - col_off[c]);
  last_colour[id] = c;
}

// order-sensitive but additive hashes of a class vector (position-keyed, so a block can reduce them in parallel);
// only used to find interning candidates — equality is verified element by element by k_verify_classes
__device__ __forceinline__ uint64_t vec_h1(uint32_t x, uint32_t i) { return mix64(((uint64_t)x << 32) | i); }
__device__ __forceinline__ uint64_t vec_h2(uint32_t x, uint32_t i) { return mix64((((uint64_t)i << 32) | x) ^ 0x9e3779b97f4a7c15ULL); }

// Exact emulation of `v1 = colours[last]; for n in others (visit order): merge_no_truncate(v1, colours[n])`.
__global__ void k_resolve_paths(const uint32_t* __restrict__ arena, const uint64_t* __restrict__ path_off, const uint32_t* __restrict__ path_len,
                                const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ col_tx, uint32_t n,
                                const uint64_t* __restrict__ vec_off, uint32_t* vec, uint64_t* vec_hash, uint64_t* vec_hash2) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  const uint32_t* path = arena + path_off[id];
  const uint32_t plen = path_len[id];
  const uint32_t last = path[plen - 1];
  uint32_t* v1 = vec + vec_off[id];
  const uint32_t n1 = (uint32_t)(col_off[last + 1] - col_off[last]);
  for (uint32_t k = 0; k < n1; k++) v1[k] = col_tx[col_off[last] + k];
  for (uint32_t s = 0; s + 1 < plen; s++) {
    const uint32_t c = path[s];
    const uint32_t* v2 = col_tx + col_off[c];
    const uint32_t n2 = (uint32_t)(col_off[c + 1] - col_off[c]);
    uint32_t f = 0, i = 0, j = 0;
    while (i < n1 && j < n2) {
      const uint32_t x = v1[i], y = v2[j];
      if (x < y) i++;
      else if (x > y) j++;
      else { v1[i] = v1[f]; v1[f] = x; i++; j++; f++; }
    }
  }
  uint64_t h = 0, h2 = 0;
  for (uint32_t k = 0; k < n1; k++) { h += vec_h1(v1[k], k); h2 += vec_h2(v1[k], k); }
  vec_hash[id] = h;
  vec_hash2[id] = h2;
}

// The same emulation, data-parallel: one thread block per path, the vector in shared memory. It rests on a property of
// the reference's no-truncate merge (verified against brute force): walking v1 against a sorted v2,
// element i is a match IFF v1[i] is in v2 AND v1[i] is a left-to-right maximum of v1 (the second pointer is a running
// maximum of ranks, so any earlier larger element has already carried it past v1[i]). The r-th match moves to
// position r; the element it displaces ends at the match position whose chain (m -> rank of the match sitting at m)
// leads back to it. So one merge = bitset membership + prefix-maximum scan + prefix sum + a gather.
// per colour: membership bitset over tx ids, one warp per colour (built once per index)
__global__ void k_colour_bitsets(const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ col_tx, uint32_t n_colours,
                                 uint32_t words, uint32_t* bits) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (c >= n_colours) return;
  uint32_t* b = bits + (size_t)c * words;
  for (uint64_t q = col_off[c] + lane; q < col_off[c + 1]; q += 32) { const uint32_t x = col_tx[q]; atomicOr(&b[x >> 5], 1u << (x & 31)); }
}

constexpr int RES_ITEMS = 8;
// T = uint16_t when every tx id fits in 16 bits (halves the shared memory), else uint32_t. 256 threads per block: 1024-thread
// blocks for the long vectors (which only fit one or two blocks per SM) were measured slower (0.44 s vs 0.33 s).
//
// r02: where the displaced elements go. The r-th match (at position p_r) swaps with position r; what sat at position r is
// either an original non-match or, if position r is itself an earlier match position, whatever that earlier swap left
// there: C(r) = C(rank of the match at r). Following that chain one link at a time (r01) is a serial pointer chase whose
// length is ~ r / (number of non-matches before r) - thousands of dependent shared-memory loads when a single allele drops
// out at the front of a 2 000-element vector; the ncu capture of r01 showed 17 % of the kernel's instructions there at
// 2.9 active lanes and 53 % of the stall samples on the barrier behind it. Inside a stretch between two consecutive
// non-matches the link length is constant (= the number of non-matches so far, j), so the whole stretch is crossed with one
// division: from position m, with u the last non-match before m, the chain next touches m - (floor((m - u - 1) / j) + 1) j.
// A chain now costs at most one hop per non-match it passes. The positions of the non-matches (needed for u) are written,
// during the scan, into the half of the output vector the matches do not use (matches fill B from the front, the list of
// non-match positions fills it from the back); the gather then runs in two phases (source index into rk, then move).
// Membership of an element in the colour being merged is only ever needed for left-to-right maxima, and is read straight
// from the per-index colour bitsets (global memory, L1-resident rows) when they exist, so the per-merge copy of the bitset
// into shared memory (11 % of r01's instructions) is gone as well.
template <typename T, int RES_THREADS>
__global__ void __launch_bounds__(RES_THREADS) k_resolve_paths_par(const uint32_t* __restrict__ arena, const uint64_t* __restrict__ path_off,
    const uint32_t* __restrict__ path_len, const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ col_tx, uint32_t n_paths,
    const uint64_t* __restrict__ vec_off, uint32_t* vec, uint64_t* vec_hash, uint64_t* vec_hash2, uint32_t n_max, uint32_t bitset_words,
    unsigned int* next_path, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ col_bits) {
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* bits = sm;                                // membership bitset of the current v2 over tx ids (only without col_bits)
  const uint32_t n_pad = (n_max + 7u) & ~7u;          // 16-byte aligned arrays, readable in whole 8-element groups
  T* A = reinterpret_cast<T*>(bits + (col_bits ? 0u : ((bitset_words + 3u) & ~3u)));   // current vector
  T* B = A + n_pad;                                   // next vector
  uint16_t* rk = reinterpret_cast<uint16_t*>(B + n_pad);   // per position: rank + 1 of the match sitting there, 0 if none
  __shared__ uint32_t w_max[2][RES_THREADS / 32], w_cnt[2][RES_THREADS / 32], w_last[2][RES_THREADS / 32];
  __shared__ unsigned int s_id;
  __shared__ unsigned long long s_h[2][RES_THREADS / 32];
  const unsigned t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (;;) {
    __syncthreads();
    if (t == 0) s_id = atomicAdd(next_path, 1u);
    __syncthreads();
    if (s_id >= n_paths) return;
    const uint32_t id = ids[s_id];   // n_paths = length of this launch's id list
    const uint32_t* path = arena + path_off[id];
    const uint32_t plen = path_len[id];
    const uint32_t last = path[plen - 1];
    const uint32_t n = (uint32_t)(col_off[last + 1] - col_off[last]);
    for (uint32_t k = t; k < n; k += RES_THREADS) A[k] = (T)col_tx[col_off[last] + k];
    for (uint32_t sidx = 0; sidx + 1 < plen; sidx++) {
      const uint32_t c = path[sidx];
      const uint32_t* mem_row = col_bits ? col_bits + (size_t)c * bitset_words : bits;
      __syncthreads();   // the previous merge has finished with `bits` / B / rk; orders the A writes of the previous merge / the load
      if (!col_bits) {   // no precomputed colour bitsets: build this colour's in shared memory
        const uint32_t* v2 = col_tx + col_off[c];
        const uint32_t n2 = (uint32_t)(col_off[c + 1] - col_off[c]);
        for (uint32_t k = t; k < bitset_words; k += RES_THREADS) bits[k] = 0;
        __syncthreads();
        for (uint32_t k = t; k < n2; k += RES_THREADS) { const uint32_t x = v2[k]; atomicOr(&bits[x >> 5], 1u << (x & 31)); }
        __syncthreads();
      }
      // matches: left-to-right maxima of A that are members of v2. The r-th match is scattered to B[r] right away, the
      // position of the j-th non-match to B[n - 1 - j]. Each thread owns RES_ITEMS consecutive elements of a tile
      // (thread-local scan, then warp and block scans).
      uint32_t carry = 0, base = 0, last_pos = 0;   // values are compared as x + 1 so that 0 means "nothing yet"
      int par = 0;
      for (uint32_t t0 = 0; t0 < n; t0 += RES_THREADS * RES_ITEMS, par ^= 1) {
        const uint32_t i0 = t0 + t * RES_ITEMS;
        uint32_t x[RES_ITEMS], pm[RES_ITEMS];   // values, thread-local inclusive prefix maxima of (value + 1)
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) x[k] = 0u;
        if (i0 < n) {   // RES_ITEMS consecutive elements as 128-bit shared-memory loads (conflict-free; the tail past n is padding)
          if (sizeof(T) == 2) {
            const uint4 q = *reinterpret_cast<const uint4*>(A + i0);
            x[0] = q.x & 0xFFFFu; x[1] = q.x >> 16; x[2] = q.y & 0xFFFFu; x[3] = q.y >> 16; x[4] = q.z & 0xFFFFu; x[5] = q.z >> 16; x[6] = q.w & 0xFFFFu; x[7] = q.w >> 16;
          } else {
            const uint4 q0 = *reinterpret_cast<const uint4*>(A + i0), q1 = *reinterpret_cast<const uint4*>(A + i0 + 4);
            x[0] = q0.x; x[1] = q0.y; x[2] = q0.z; x[3] = q0.w; x[4] = q1.x; x[5] = q1.y; x[6] = q1.z; x[7] = q1.w;
          }
        }
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) { run = max(run, i0 + k < n ? x[k] + 1 : 0u); pm[k] = run; }
        uint32_t v = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= (unsigned)o) v = max(v, u); }
        uint32_t before = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        if (lane == 0) before = 0;
        if (lane == 31) w_max[par][w] = v;
        __syncthreads();
        uint32_t excl = max(carry, before), tile_max = carry;
#pragma unroll
        for (unsigned k = 0; k < RES_THREADS / 32; k++) { if (k < w) excl = max(excl, w_max[par][k]); tile_max = max(tile_max, w_max[par][k]); }
        unsigned mflags = 0, cnt = 0;
        uint32_t my_last = 0;
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) {
          const uint32_t prev = k ? max(excl, pm[k - 1]) : excl;
          bool match = i0 + k < n && x[k] + 1 > prev;                                        // a left-to-right maximum ...
          if (match) match = (mem_row[x[k] >> 5] >> (x[k] & 31)) & 1u;                        // ... that is a member of v2
          if (match) { mflags |= 1u << k; cnt++; my_last = i0 + k + 1; }
        }
        uint32_t incl = cnt;   // warp inclusive scan of the per-thread match counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (unsigned)o) incl += u; }
        uint32_t wl = my_last;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_down_sync(0xFFFFFFFFu, wl, o));
        if (lane == 31) w_cnt[par][w] = incl;
        if (lane == 0) w_last[par][w] = wl;
        __syncthreads();
        uint32_t rank = base + incl - cnt, tile_cnt = 0;
#pragma unroll
        for (unsigned k = 0; k < RES_THREADS / 32; k++) { if (k < w) rank += w_cnt[par][k]; tile_cnt += w_cnt[par][k]; last_pos = max(last_pos, w_last[par][k]); }
        uint32_t rkv[RES_ITEMS];
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) {
          const bool match = (mflags >> k) & 1u;
          rkv[k] = match ? rank + 1 : 0u;
          if (match) { B[rank] = (T)x[k]; rank++; }
          else if (i0 + k < n) B[n - 1 - (i0 + k - rank)] = (T)(i0 + k);   // the (i0 + k - rank)-th non-match sits here
        }
        if (i0 < n) *reinterpret_cast<uint4*>(rk + i0) = make_uint4(rkv[0] | (rkv[1] << 16), rkv[2] | (rkv[3] << 16), rkv[4] | (rkv[5] << 16), rkv[6] | (rkv[7] << 16));
        carry = tile_max; base += tile_cnt;
      }
      __syncthreads();
      const uint32_t M = base;
      if (M == 0 || last_pos == M) continue;   // nothing moves: the matches already occupy 0..M-1 (last_pos = last match position + 1)
      // phase 1: for every position behind the matches, where its new element comes from
      for (uint32_t i = M + t; i < n; i += RES_THREADS) {
        uint32_t src = i;
        if (rk[i]) {                                       // a match left from here: a displaced element arrives
          uint32_t m = rk[i] - 1u;                         // ... the content of position m when the match got there
          for (;;) {
            const uint32_t r1 = rk[m];
            if (r1 == 0) break;                            // an original non-match: this is the element
            const uint32_t j = m - (r1 - 1u);              // non-matches before m = link length of this stretch
            if (j == 0) break;
            const uint32_t u = B[n - j];                   // position of the last non-match before m (the (j-1)-th, 0-based)
            m -= ((m - u - 1u) / j + 1u) * j;
          }
          src = m;
        }
        rk[i] = (uint16_t)src;
      }
      __syncthreads();
      // phase 2: move
      for (uint32_t i = M + t; i < n; i += RES_THREADS) B[i] = A[rk[i]];
      __syncthreads();
      T* tmp = A; A = B; B = tmp;
    }
    __syncthreads();
    // write the vector out and hash it
    uint32_t* dst = vec + vec_off[id];
    unsigned long long h1 = 0, h2 = 0;
    for (uint32_t k = t; k < n; k += RES_THREADS) { const uint32_t x = A[k]; dst[k] = x; h1 += vec_h1(x, k); h2 += vec_h2(x, k); }
    for (int o = 16; o > 0; o >>= 1) { h1 += __shfl_down_sync(0xFFFFFFFFu, h1, o); h2 += __shfl_down_sync(0xFFFFFFFFu, h2, o); }
    if (lane == 0) { s_h[0][w] = h1; s_h[1][w] = h2; }
    __syncthreads();
    if (t == 0) {
      unsigned long long a = 0, b = 0;
      for (unsigned k = 0; k < RES_THREADS / 32; k++) { a += s_h[0][k]; b += s_h[1][k]; }
      vec_hash[id] = a; vec_hash2[id] = b;
    }
  }
}

// r02b: the same merge with ONE WARP per path, for the vectors of up to RES_WARP_MAX elements (most of them: the ncu capture
// of the block kernel showed its shortest-vector launch taking half the time at ~5 600 warp-instructions per merge whatever
// the length, since all 8 warps run the whole tile code and meet at two block barriers for a vector that fills a fraction of
// one 2 048-element tile). A warp walks its vector in tiles of 32 lanes x 8 elements with the running maximum / match count
// carried in registers: no block barrier, no cross-warp partials, instruction count proportional to the length. The warps
// of a block are independent (own slice of shared memory, own path from the shared work counter). Needs the per-index
// colour bitsets in global memory.
constexpr uint32_t RES_WARP_MAX = 2048;
template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_resolve_paths_warp(const uint32_t* __restrict__ arena, const uint64_t* __restrict__ path_off,
    const uint32_t* __restrict__ path_len, const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ col_tx, uint32_t n_paths,
    const uint64_t* __restrict__ vec_off, uint32_t* vec, uint64_t* vec_hash, uint64_t* vec_hash2, uint32_t cap /* multiple of 8 */,
    uint32_t bitset_words, unsigned int* next_path, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ col_bits) {
  extern __shared__ __align__(16) uint32_t sm[];
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char* mine = reinterpret_cast<unsigned char*>(sm) + (size_t)w * cap * (2 * sizeof(T) + 2);
  T* A = reinterpret_cast<T*>(mine);
  T* B = A + cap;
  uint16_t* rk = reinterpret_cast<uint16_t*>(B + cap);
  for (;;) {
    unsigned slot = 0;
    if (lane == 0) slot = atomicAdd(next_path, 1u);
    slot = __shfl_sync(FULL, slot, 0);
    if (slot >= n_paths) return;
    const uint32_t id = ids[slot];
    const uint32_t* path = arena + path_off[id];
    const uint32_t plen = path_len[id];
    const uint32_t last = path[plen - 1];
    const uint64_t cb = col_off[last];
    const uint32_t n = (uint32_t)(col_off[last + 1] - cb);
    __syncwarp();   // the previous path's output loop has finished reading A
    for (uint32_t k = lane; k < n; k += 32) A[k] = (T)col_tx[cb + k];
    for (uint32_t sidx = 0; sidx + 1 < plen; sidx++) {
      const uint32_t* mem_row = col_bits + (size_t)path[sidx] * bitset_words;
      __syncwarp();
      uint32_t carry = 0, base = 0, last_pos = 0;   // values are compared as x + 1 so that 0 means "nothing yet"
      for (uint32_t t0 = 0; t0 < n; t0 += 256) {
        const uint32_t i0 = t0 + lane * RES_ITEMS;
        uint32_t x[RES_ITEMS], pm[RES_ITEMS];
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) x[k] = 0u;
        if (i0 < n) {
          if (sizeof(T) == 2) {
            const uint4 q = *reinterpret_cast<const uint4*>(A + i0);
            x[0] = q.x & 0xFFFFu; x[1] = q.x >> 16; x[2] = q.y & 0xFFFFu; x[3] = q.y >> 16; x[4] = q.z & 0xFFFFu; x[5] = q.z >> 16; x[6] = q.w & 0xFFFFu; x[7] = q.w >> 16;
          } else {
            const uint4 q0 = *reinterpret_cast<const uint4*>(A + i0), q1 = *reinterpret_cast<const uint4*>(A + i0 + 4);
            x[0] = q0.x; x[1] = q0.y; x[2] = q0.z; x[3] = q0.w; x[4] = q1.x; x[5] = q1.y; x[6] = q1.z; x[7] = q1.w;
          }
        }
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) { run = max(run, i0 + k < n ? x[k] + 1 : 0u); pm[k] = run; }
        uint32_t v = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, v, o); if (lane >= (unsigned)o) v = max(v, u); }
        uint32_t before = __shfl_up_sync(FULL, v, 1);
        if (lane == 0) before = 0;
        const uint32_t excl = max(carry, before);
        carry = max(carry, __shfl_sync(FULL, v, 31));
        unsigned mflags = 0, cnt = 0;
        uint32_t my_last = 0;
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) {
          const uint32_t prev = k ? max(excl, pm[k - 1]) : excl;
          bool match = i0 + k < n && x[k] + 1 > prev;
          if (match) match = (__ldg(mem_row + (x[k] >> 5)) >> (x[k] & 31)) & 1u;
          if (match) { mflags |= 1u << k; cnt++; my_last = i0 + k + 1; }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if (lane >= (unsigned)o) incl += u; }
        last_pos = max(last_pos, __reduce_max_sync(FULL, my_last));
        uint32_t rank = base + incl - cnt;
        base += __shfl_sync(FULL, incl, 31);
        uint32_t rkv[RES_ITEMS];
#pragma unroll
        for (int k = 0; k < RES_ITEMS; k++) {
          const bool match = (mflags >> k) & 1u;
          rkv[k] = match ? rank + 1 : 0u;
          if (match) { B[rank] = (T)x[k]; rank++; }
          else if (i0 + k < n) B[n - 1 - (i0 + k - rank)] = (T)(i0 + k);
        }
        if (i0 < n) *reinterpret_cast<uint4*>(rk + i0) = make_uint4(rkv[0] | (rkv[1] << 16), rkv[2] | (rkv[3] << 16), rkv[4] | (rkv[5] << 16), rkv[6] | (rkv[7] << 16));
      }
      const uint32_t M = base;
      if (M == 0 || last_pos == M) continue;   // nothing moves (uniform over the warp)
      __syncwarp();
      for (uint32_t i = M + lane; i < n; i += 32) {
        uint32_t src = i;
        const uint32_t r0 = rk[i];
        if (r0) {
          uint32_t m = r0 - 1u;
          for (;;) {
            const uint32_t r1 = rk[m];
            if (r1 == 0) break;
            const uint32_t j = m - (r1 - 1u);
            if (j == 0) break;
            const uint32_t u = B[n - j];
            m -= ((m - u - 1u) / j + 1u) * j;
          }
          src = m;
        }
        rk[i] = (uint16_t)src;   // i >= M, while the chains only read rk below M
      }
      __syncwarp();
      for (uint32_t i = M + lane; i < n; i += 32) B[i] = A[rk[i]];   // only now: the back of B held the non-match positions
      T* tmp = A; A = B; B = tmp;
    }
    __syncwarp();
    uint32_t* dst = vec + vec_off[id];
    unsigned long long h1 = 0, h2 = 0;
    for (uint32_t k = lane; k < n; k += 32) { const uint32_t x = A[k]; dst[k] = x; h1 += vec_h1(x, k); h2 += vec_h2(x, k); }
    for (int o = 16; o > 0; o >>= 1) { h1 += __shfl_down_sync(FULL, h1, o); h2 += __shfl_down_sync(FULL, h2, o); }
    if (lane == 0) { vec_hash[id] = h1; vec_hash2[id] = h2; }
  }
}

// every path's vector must equal its class representative's vector element by element (interning is by hash)
__global__ void k_verify_classes(const uint32_t* __restrict__ vec, const uint64_t* __restrict__ vec_off, const uint32_t* __restrict__ rep_of_path,
                                 uint32_t n, unsigned long long* mismatches) {
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (p >= n) return;
  const uint32_t r = rep_of_path[p];
  if (r == p) return;
  const uint64_t a = vec_off[p], b = vec_off[r], len = vec_off[p + 1] - a;
  bool bad = (vec_off[r + 1] - b) != len;
  for (uint64_t k = lane; k < len && !bad; k += 32) bad = vec[a + k] != vec[b + k];
  if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicAdd(mismatches, 1ULL);
}

// counts_reads in class order: one block per class copies its representative's vector to its final place
__global__ void k_gather_classes(const uint32_t* __restrict__ vec, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ dst_off,
                                 uint32_t* out) {
  const uint32_t c = blockIdx.x;
  const uint64_t a = src_off[c], b = dst_off[c], len = dst_off[c + 1] - b;
  for (uint64_t k = threadIdx.x; k < len; k += blockDim.x) out[b + k] = vec[a + k];
}

// reads_explained[t] += reads of every class whose (sorted) member list is colour list `col` (mapping.rs:196-198)
__global__ void k_reads_explained(const uint32_t* __restrict__ cols, const unsigned long long* __restrict__ totals, uint32_t n,
                                  const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ col_tx, unsigned long long* expl) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (i >= n) return;
  const uint32_t c = cols[i];
  const unsigned long long t = totals[i];
  for (uint64_t q = col_off[c] + lane; q < col_off[c + 1]; q += 32) atomicAdd(expl + col_tx[q], t);
}

__global__ void k_map_rank(const uint32_t* __restrict__ rec_path, const uint32_t* __restrict__ rank_of_path, uint64_t n, uint32_t* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rank_of_path[rec_path[i]];
}

// one thread per (barcode, umi) run of the sorted records: UMI class = lowest class rank among its reads
__global__ void k_umi_min(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ranks, uint64_t n, uint32_t* umi_hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  if (i > 0 && keys[i - 1] == k) return;
  uint32_t best = ranks[i];
  for (uint64_t j = i + 1; j < n && keys[j] == k; j++) best = min(best, ranks[j]);
  atomicAdd(umi_hist + best, 1u);
}

// ---- multi-rank union of the interned paths, on the device (hlac_geno_attach_comm) -------------------------
// one warp per path: its colour ids from wherever the arena holds them to their place in a compact array
__global__ void k_pack_colours(const uint32_t* __restrict__ arena, const uint64_t* __restrict__ path_off, const uint32_t* __restrict__ path_len,
                               const uint64_t* __restrict__ dst_off, uint32_t n, uint32_t* out) {
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (p >= n) return;
  const uint32_t* src = arena + path_off[p];
  uint32_t* dst = out + dst_off[p];
  for (uint32_t k = lane; k < path_len[p]; k += 32) dst[k] = src[k];
}
// every gathered entry inserts its hash A; the winner of a slot leaves its entry index there (the representative)
__global__ void k_union_insert(PathTable t, const uint64_t* __restrict__ g_hash_a, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = g_hash_a[i];
  uint64_t h = mix64(key) & t.mask;
  for (;;) {
    const unsigned long long cur = t.keys[h];
    if (cur == key) return;
    if (cur == 0) {
      const unsigned long long old = atomicCAS(t.keys + h, 0ULL, key);
      if (old == 0) { t.ids[h] = (uint32_t)i; return; }
      if (old == key) return;
    }
    h = (h + 1) & t.mask;
  }
}
// occupied slots get dense union ids (the numbering is local to a rank: nothing global depends on it)
__global__ void k_union_number(PathTable t, uint64_t cap, unsigned long long* counter, uint32_t* rep_of_uid) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= cap || t.keys[h] == 0) return;
  const uint32_t uid = (uint32_t)atomicAdd(counter, 1ULL);
  rep_of_uid[uid] = t.ids[h];
  t.ids[h] = uid;
}
struct UnionOut { uint64_t* hash_a; uint64_t* hash_b; unsigned long long* first_ord; uint32_t* count; uint64_t* path_off; uint32_t* path_len; };
// every gathered entry folds into its union path: min of first-seen ordinals, sum of read counts; the entry must be the same
// path as the representative (hash B, length and the colour ids themselves), else two different paths collided on hash A
__global__ void k_union_fill(PathTable t, const uint64_t* __restrict__ g_hash_a, const uint64_t* __restrict__ g_hash_b,
                             const unsigned long long* __restrict__ g_ord, const uint32_t* __restrict__ g_cnt, const uint32_t* __restrict__ g_len,
                             const uint64_t* __restrict__ g_coff, const uint32_t* __restrict__ g_col, uint64_t n, const uint32_t* __restrict__ rep_of_uid,
                             UnionOut u, unsigned long long* collisions, unsigned long long* overflow) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = g_hash_a[i];
  uint64_t h = mix64(key) & t.mask;
  while (t.keys[h] != key) h = (h + 1) & t.mask;
  const uint32_t uid = t.ids[h], rep = rep_of_uid[uid];
  atomicMin(u.first_ord + uid, g_ord[i]);
  const uint32_t old = atomicAdd(u.count + uid, g_cnt[i]);
  if (old + g_cnt[i] < old) atomicAdd(overflow, 1ULL);
  if (i == rep) { u.hash_a[uid] = key; u.hash_b[uid] = g_hash_b[i]; u.path_off[uid] = g_coff[i]; u.path_len[uid] = g_len[i]; }
  else {
    bool bad = g_hash_b[i] != g_hash_b[rep] || g_len[i] != g_len[rep];
    for (uint32_t k = 0; k < g_len[i] && !bad; k++) bad = g_col[g_coff[i] + k] != g_col[g_coff[rep] + k];
    if (bad) atomicAdd(collisions, 1ULL);
  }
}
// local path id -> union id
__global__ void k_local_to_union(PathTable t, const uint64_t* __restrict__ local_hash_a, uint32_t n, uint32_t* map) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = local_hash_a[i];
  uint64_t h = mix64(key) & t.mask;
  while (t.keys[h] != key) h = (h + 1) & t.mask;
  map[i] = t.ids[h];
}
// paths this rank resolves: hash A mod n_ranks == rank (every rank derives the same partition from the union alone)
__global__ void k_owned_mask(const uint64_t* __restrict__ hash_a, uint32_t n, uint32_t n_ranks, uint32_t rank, const uint32_t* __restrict__ len_in, uint32_t* len_owned,
                             uint32_t* owner_count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t o = (uint32_t)(hash_a[i] % n_ranks);
  len_owned[i] = o == rank ? len_in[i] : 0u;
  atomicAdd(owner_count + o, 1u);
}
// an owner's resolved fingerprints, packed for the exchange: (hash A, vector hash 1, vector hash 2) per owned path
__global__ void k_pack_fingerprints(const uint64_t* __restrict__ hash_a, const uint64_t* __restrict__ vh1, const uint64_t* __restrict__ vh2, uint32_t n, uint32_t n_ranks,
                                    uint32_t rank, uint64_t* out, unsigned long long* cursor) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (uint32_t)(hash_a[i] % n_ranks) != rank) return;
  const unsigned long long at = atomicAdd(cursor, 1ULL);
  out[3 * at] = hash_a[i]; out[3 * at + 1] = vh1[i]; out[3 * at + 2] = vh2[i];
}
__global__ void k_unpack_fingerprints(PathTable t, const uint64_t* __restrict__ in, uint64_t n, uint64_t* vh1, uint64_t* vh2) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = in[3 * i];
  uint64_t h = mix64(key) & t.mask;
  while (t.keys[h] != key) h = (h + 1) & t.mask;
  const uint32_t uid = t.ids[h];
  vh1[uid] = in[3 * i + 1]; vh2[uid] = in[3 * i + 2];
}
// per class: which of the (<= 128) candidate alleles it contains. A class vector is a permutation of the colour list of its
// path's last node, so membership is read off that (sorted) colour list - the resolved vector itself is not needed, which
// matters when the vector lives on another rank.
__global__ void k_class_cand_mask_colour(const uint32_t* __restrict__ cls_colour, uint32_t n, const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ col_tx,
                                         const uint8_t* __restrict__ cand_idx, uint64_t* mask) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (c >= n) return;
  const uint32_t col = cls_colour[c];
  uint64_t lo = 0, hi = 0;
  for (uint64_t q = col_off[col] + lane; q < col_off[col + 1]; q += 32) {
    const uint8_t ci = cand_idx[col_tx[q]];
    if (ci < 64) lo |= 1ull << ci; else if (ci < 128) hi |= 1ull << (ci - 64);
  }
  for (int o = 16; o > 0; o >>= 1) { lo |= __shfl_down_sync(0xFFFFFFFFu, lo, o); hi |= __shfl_down_sync(0xFFFFFFFFu, hi, o); }
  if (lane == 0) { mask[2 * (size_t)c] = lo; mask[2 * (size_t)c + 1] = hi; }
}

// pair_reads_explained (em.rs:36-44) for every pair of candidate alleles at once: a class with read count r and candidate
// set M adds r to single[a] for a in M and to both[a][b] for a < b in M; reads in classes containing a OR b are then
// single[a] + single[b] - both[a][b]. One thread per class (a class holds a handful of candidates).
__global__ void k_pair_matrix(const uint64_t* __restrict__ mask, const uint32_t* __restrict__ cls_reads, uint32_t n, unsigned long long* single, unsigned long long* both) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const uint64_t m[2] = {mask[2 * (size_t)c], mask[2 * (size_t)c + 1]};
  if ((m[0] | m[1]) == 0) return;
  const unsigned long long r = cls_reads[c];
  for (int wa = 0; wa < 2; wa++)
    for (uint64_t ba = m[wa]; ba; ba &= ba - 1) {
      const int a = wa * 64 + __ffsll((long long)ba) - 1;
      atomicAdd(single + a, r);
      for (int wb = wa; wb < 2; wb++)
        for (uint64_t bb = wb == wa ? (ba & (ba - 1)) : m[wb]; bb; bb &= bb - 1) {
          const int b = wb * 64 + __ffsll((long long)bb) - 1;
          atomicAdd(both + a * 128 + b, r);
        }
    }
}

}  // namespace

static inline double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

struct hlac_geno {
  hlac_index* idx = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_sm = 1;
  int stride = 1;   // seed-scan stride (hlac_set_seed_stride) the session was created with
  // staging + per-batch temporaries
  DevBuf<uint64_t> st_seq, hash_a, hash_b;
  DevBuf<uint16_t> st_len;
  DevBuf<uint32_t> st_cell, st_umi, cov, batch_pid;
  bool verify = true;   // re-walk every counted read against its interned path (HLAC_GENO_VERIFY=0 turns it off)
  DevBuf<uint64_t> st_ord;
  DevBuf<uint8_t> strand;
  // path table
  DevBuf<unsigned long long> t_keys; DevBuf<uint32_t> t_ids; uint64_t t_cap = 0;
  DevBuf<uint64_t> e_hash_a, e_hash_b, e_path_off; DevBuf<unsigned long long> e_first_ord; DevBuf<uint32_t> e_count, e_path_len, e_rep;
  DevBuf<uint32_t> arena;
  DevBuf<unsigned long long> ctr;   // [0] n_entries [1] arena cursor [2] rec cursor [3] collisions [4..7] stats
  // records of counted reads
  DevBuf<uint64_t> rec_key; DevBuf<uint32_t> rec_path;
  DevBuf<uint32_t> read_path; bool record_reads = false;
  bool lean = false;   // finish does not materialise counts_reads on the host (hlac_geno_set_lean)   // optional: path id of every read pushed
  uint64_t reads_pushed = 0, last_n = 0, n_entries = 0, arena_used = 0, n_records_ub = 0;
  // finish results kept for hlac_geno_read_class_ids
  std::vector<uint32_t> class_of_path;
  bool finished = false;
  // between hlac_geno_finish_begin and _end
  std::vector<uint64_t> fin_src_off; std::vector<uint32_t> fin_cls_len, fin_cls_reads, fin_cls_colour, fin_cls_owner;   // per class, first-seen order
  DevBuf<uint32_t> fin_hist, fin_dvec; bool fin_begun = false;
  float ms[5] = {0, 0, 0, 0, 0};   // map, intern, finish (all of it), H2D, path resolution alone
  uint64_t launches = 0;
  uint64_t res_bytes = 0, res_merges = 0, res_steps = 0, res_paths = 0;   // work of the last path resolution (hlac_geno_resolve_stats)
  struct Ev { cudaEvent_t a, b; int kind; };
  std::vector<Ev> events;
  // multi-GPU (hlac_geno_attach_comm): the finish merges the ranks' paths, resolves 1/n of the union and sums the histograms
  hlac_comm* comm = nullptr;
  bool merged = false;                 // this session's entries are already the union of all ranks
  double merge_ms = 0, exchange_ms = 0;
  DevBuf<uint64_t> scan_tmp;

  ~hlac_geno() {
    for (auto& e : events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (stream) cudaStreamDestroy(stream);
  }
  size_t begin(int kind) {
    if (events.size() >= 512) drain_events();   // long streaming runs: keep the number of live events bounded
    Ev e; e.kind = kind;
    HLAC_CUDA(cudaEventCreate(&e.a)); HLAC_CUDA(cudaEventCreate(&e.b));
    HLAC_CUDA(cudaEventRecord(e.a, stream));
    events.push_back(e);
    return events.size() - 1;
  }
  void end(size_t k) { HLAC_CUDA(cudaEventRecord(events[k].b, stream)); }
  void drain_events() {
    for (auto& e : events) {
      float t = 0; HLAC_CUDA(cudaEventSynchronize(e.b)); HLAC_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
      ms[e.kind] += t; cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    events.clear();
  }
  PathTable table() { return PathTable{t_keys.p, t_ids.p, t_cap - 1}; }
  PathEntries entries() { return PathEntries{e_hash_a.p, e_hash_b.p, e_first_ord.p, e_count.p, e_path_off.p, e_path_len.p, e_rep.p}; }

  void ensure_table(uint64_t want_entries) {
    uint64_t cap = t_cap ? t_cap : (1u << 16);
    while (cap < want_entries * 2) cap <<= 1;
    if (cap != t_cap) {
      DevBuf<unsigned long long> nk; DevBuf<uint32_t> ni;
      nk.reserve(cap, stream); ni.reserve(cap, stream);
      HLAC_CUDA(cudaMemsetAsync(nk.p, 0, cap * sizeof(unsigned long long), stream));
      if (t_cap) {
        PathTable nt{nk.p, ni.p, cap - 1};
        k_rehash<<<(unsigned)((t_cap + 255) / 256), 256, 0, stream>>>(t_keys.p, t_ids.p, t_cap, nt);
        HLAC_CUDA(cudaGetLastError());
        launches++;
      }
      HLAC_CUDA(cudaStreamSynchronize(stream));
      std::swap(t_keys.p, nk.p); std::swap(t_keys.cap, nk.cap);
      std::swap(t_ids.p, ni.p); std::swap(t_ids.cap, ni.cap);
      t_cap = cap;
    }
    const size_t ne = want_entries;
    e_hash_a.reserve(ne, stream, true, n_entries); e_hash_b.reserve(ne, stream, true, n_entries); e_path_off.reserve(ne, stream, true, n_entries);
    e_first_ord.reserve(ne, stream, true, n_entries); e_count.reserve(ne, stream, true, n_entries);
    e_path_len.reserve(ne, stream, true, n_entries); e_rep.reserve(ne, stream, true, n_entries);
  }

  // Phase 1 + 2 of the multi-rank finish: all-gather every rank's distinct paths (variable-size, NCCL broadcasts in one
  // group), build the union in a fresh hash table on the device, fold first-seen ordinals (min) and read counts (sum), adopt
  // the union as this session's entries and renumber the local records. Collective.
  void merge_over_comm() {
    const int N = comm_nranks(comm), R = comm_rank(comm);
    const double t_begin = now_ms();
    const uint64_t E0 = n_entries;
    // local colours compacted in path order
    DevBuf<uint64_t> l_off; l_off.reserve(E0 + 1, stream);
    exclusive_scan_u32(e_path_len.p, l_off.p, E0, scan_tmp, stream);
    // sizes of every rank: (paths, colour ids)
    DevBuf<uint64_t> d_sizes; d_sizes.reserve(2 * (size_t)N + 2, stream);
    const uint64_t e0 = E0;
    HLAC_CUDA(cudaMemcpyAsync(d_sizes.p + 2 * N, &e0, 8, cudaMemcpyHostToDevice, stream));
    HLAC_CUDA(cudaMemcpyAsync(d_sizes.p + 2 * N + 1, l_off.p + E0, 8, cudaMemcpyDeviceToDevice, stream));
    comm_allgather(comm, d_sizes.p + 2 * N, d_sizes.p, 2, HLAC_COMM_U64, stream);
    std::vector<uint64_t> sizes(2 * (size_t)N);
    HLAC_CUDA(cudaMemcpyAsync(sizes.data(), d_sizes.p, sizes.size() * 8, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    std::vector<size_t> ecnt(N), edis(N), ccnt(N), cdis(N);
    uint64_t T = 0, C = 0;
    for (int r = 0; r < N; r++) { ecnt[r] = sizes[2 * r]; edis[r] = T; T += ecnt[r]; ccnt[r] = sizes[2 * r + 1]; cdis[r] = C; C += ccnt[r]; }
    if (T >= 0xFFFFFFFFull) throw Error(HLAC_ERR_LIMIT, "too many paths over all ranks");
    DevBuf<uint32_t> l_col; l_col.reserve(std::max<uint64_t>(sizes[2 * R + 1], 1), stream);
    if (E0) k_pack_colours<<<(unsigned)(((uint64_t)E0 * 32 + 255) / 256), 256, 0, stream>>>(arena.p, e_path_off.p, e_path_len.p, l_off.p, (uint32_t)E0, l_col.p);
    // gathered arrays
    DevBuf<uint64_t> g_a, g_b, g_coff; DevBuf<unsigned long long> g_ord; DevBuf<uint32_t> g_cnt, g_len, g_col;
    g_a.reserve(std::max<uint64_t>(T, 1), stream); g_b.reserve(std::max<uint64_t>(T, 1), stream); g_ord.reserve(std::max<uint64_t>(T, 1), stream);
    g_cnt.reserve(std::max<uint64_t>(T, 1), stream); g_len.reserve(std::max<uint64_t>(T, 1), stream); g_col.reserve(std::max<uint64_t>(C, 1), stream);
    g_coff.reserve(T + 1, stream);
    comm_allgatherv(comm, e_hash_a.p, g_a.p, ecnt.data(), edis.data(), HLAC_COMM_U64, stream);
    comm_allgatherv(comm, e_hash_b.p, g_b.p, ecnt.data(), edis.data(), HLAC_COMM_U64, stream);
    comm_allgatherv(comm, e_first_ord.p, g_ord.p, ecnt.data(), edis.data(), HLAC_COMM_U64, stream);
    comm_allgatherv(comm, e_count.p, g_cnt.p, ecnt.data(), edis.data(), HLAC_COMM_U32, stream);
    comm_allgatherv(comm, e_path_len.p, g_len.p, ecnt.data(), edis.data(), HLAC_COMM_U32, stream);
    comm_allgatherv(comm, l_col.p, g_col.p, ccnt.data(), cdis.data(), HLAC_COMM_U32, stream);
    exclusive_scan_u32(g_len.p, g_coff.p, T, scan_tmp, stream);   // ranks are concatenated in the same order in both arrays
    // the union table
    uint64_t cap = 1u << 16;
    while (cap < T * 2) cap <<= 1;
    DevBuf<unsigned long long> nk; DevBuf<uint32_t> ni, rep_of_uid, l2u; DevBuf<unsigned long long> cnts;
    nk.reserve(cap, stream); ni.reserve(cap, stream); rep_of_uid.reserve(std::max<uint64_t>(T, 1), stream); cnts.reserve(4, stream);
    HLAC_CUDA(cudaMemsetAsync(nk.p, 0, cap * sizeof(unsigned long long), stream));
    HLAC_CUDA(cudaMemsetAsync(cnts.p, 0, 4 * sizeof(unsigned long long), stream));
    PathTable nt{nk.p, ni.p, cap - 1};
    if (T) k_union_insert<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(nt, g_a.p, T);
    k_union_number<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(nt, cap, cnts.p, rep_of_uid.p);
    unsigned long long c4[4];
    HLAC_CUDA(cudaMemcpyAsync(c4, cnts.p, sizeof c4, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    const uint64_t U = c4[0];
    DevBuf<uint64_t> u_a, u_b, u_off; DevBuf<unsigned long long> u_ord; DevBuf<uint32_t> u_cnt, u_len, u_rep;
    const size_t un = std::max<uint64_t>(U, 1);
    u_a.reserve(un, stream); u_b.reserve(un, stream); u_off.reserve(un, stream); u_ord.reserve(un, stream); u_cnt.reserve(un, stream); u_len.reserve(un, stream); u_rep.reserve(un, stream);
    HLAC_CUDA(cudaMemsetAsync(u_ord.p, 0xFF, un * 8, stream));
    HLAC_CUDA(cudaMemsetAsync(u_cnt.p, 0, un * 4, stream));
    HLAC_CUDA(cudaMemsetAsync(u_rep.p, 0, un * 4, stream));
    if (T) k_union_fill<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(nt, g_a.p, g_b.p, g_ord.p, g_cnt.p, g_len.p, g_coff.p, g_col.p, T, rep_of_uid.p,
                                                                        UnionOut{u_a.p, u_b.p, u_ord.p, u_cnt.p, u_off.p, u_len.p}, cnts.p + 1, cnts.p + 2);
    // local records: local path id -> union id
    l2u.reserve(std::max<uint64_t>(E0, 1), stream);
    if (E0) k_local_to_union<<<(unsigned)((E0 + 255) / 256), 256, 0, stream>>>(nt, e_hash_a.p, (uint32_t)E0, l2u.p);
    if (n_records_ub) k_remap<<<(unsigned)((n_records_ub + 255) / 256), 256, 0, stream>>>(rec_path.p, n_records_ub, l2u.p);
    if (record_reads && reads_pushed) k_remap<<<(unsigned)((reads_pushed + 255) / 256), 256, 0, stream>>>(read_path.p, reads_pushed, l2u.p);
    HLAC_CUDA(cudaGetLastError());
    HLAC_CUDA(cudaMemcpyAsync(c4, cnts.p, sizeof c4, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    if (c4[1]) throw Error(HLAC_ERR_STATE, "path hash collision detected while merging ranks (equal 64-bit hash A, different colour paths)");
    if (c4[2]) throw Error(HLAC_ERR_LIMIT, "more than 2^32 reads on one path");
    // adopt the union
    auto take = [](auto& dst, auto& src) { std::swap(dst.p, src.p); std::swap(dst.cap, src.cap); };
    take(e_hash_a, u_a); take(e_hash_b, u_b); take(e_path_off, u_off); take(e_first_ord, u_ord); take(e_count, u_cnt); take(e_path_len, u_len); take(e_rep, u_rep);
    take(arena, g_col); take(t_keys, nk); take(t_ids, ni);
    t_cap = cap; n_entries = U; arena_used = C;
    const unsigned long long c2[2] = {U, C};
    HLAC_CUDA(cudaMemcpyAsync(ctr.p, c2, sizeof c2, cudaMemcpyHostToDevice, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    launches += 7;
    merged = true;
    merge_ms = now_ms() - t_begin;
  }

  void push_device(const hlac_batch& b, int forward_only) {
    if (b.n == 0) return;
    if (!b.seq || !b.len || !b.cell || !b.umi) throw Error(HLAC_ERR_ARG, "batch has a null array");
    if (b.words_per_read == 0 || b.words_per_read > 16) throw Error(HLAC_ERR_ARG, "words_per_read must be 1..16");
    if (b.n >= (1ull << 32)) throw Error(HLAC_ERR_LIMIT, "batch too large");
    finished = false;
    if (merged) throw Error(HLAC_ERR_STATE, "this session has already merged its paths with the other ranks: no more pushes");
    const uint64_t n = b.n;
    cov.reserve(n, stream); hash_a.reserve(n, stream); hash_b.reserve(n, stream); strand.reserve(n, stream);
    ensure_table(n_entries + n);
    rec_key.reserve(n_records_ub + n, stream, true, n_records_ub);
    rec_path.reserve(n_records_ub + n, stream, true, n_records_ub);
    if (record_reads) read_path.reserve(reads_pushed + n, stream, true, reads_pushed);
    GenoMapArgs a{};
    a.ix = idx->view; a.seq = b.seq; a.len = b.len; a.n = n; a.wpr = b.words_per_read; a.nw = 2 * a.wpr + 2; a.row_stride = (2 * a.nw) | 1;
    a.forward_only = forward_only; a.cov = cov.p; a.hash_a = hash_a.p; a.hash_b = hash_b.p; a.strand = strand.p; a.stats = ctr.p + 4;
    const size_t smem_paths = (size_t)GENO_THREADS * a.row_stride * 4;   // k_geno_paths: one row pair per thread
    const size_t per_warp_words = (size_t)32 * GENO_R * a.row_stride + 32 * GENO_R * 2 + 32 * GENO_R / 2 + 3 * 32 * GENO_R / 4;
    const size_t smem = (GENO_THREADS / 32) * per_warp_words * 4;
    if (smem > 227 * 1024) throw Error(HLAC_ERR_LIMIT, "reads too long for the shared-memory tile of the genotyping kernel");
    auto* const kmap = stride == 3 ? k_geno_map<3> : k_geno_map<1>;
    auto* const kpaths = stride == 3 ? k_geno_paths<3> : k_geno_paths<1>;
    HLAC_CUDA(cudaFuncSetAttribute(kmap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HLAC_CUDA(cudaFuncSetAttribute(kpaths, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_paths));
    int occ = 1;
    HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kmap, GENO_THREADS, smem));
    const uint64_t per_block = (uint64_t)GENO_THREADS * GENO_R;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + per_block - 1) / per_block, (uint64_t)n_sm * std::max(occ, 1));
    size_t e0 = begin(0);
    kmap<<<grid, GENO_THREADS, smem, stream>>>(a);
    HLAC_CUDA(cudaGetLastError());
    end(e0);
    size_t e1 = begin(1);
    const unsigned g1 = (unsigned)((n + 255) / 256);
    k_intern_insert<<<g1, 256, 0, stream>>>(table(), entries(), cov.p, hash_a.p, hash_b.p, n, ctr.p);
    HLAC_CUDA(cudaGetLastError());
    unsigned long long ne = 0;
    HLAC_CUDA(cudaMemcpyAsync(&ne, ctr.p, sizeof ne, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    const uint32_t n_new = (uint32_t)(ne - n_entries);
    launches += 2;
    if (n_new) {
      arena.reserve(arena_used + (uint64_t)n_new * 32 * b.words_per_read, stream, true, arena_used);   // a path has at most L nodes
      GenoPathArgs p{};
      p.ix = idx->view; p.seq = b.seq; p.len = b.len; p.strand = strand.p; p.wpr = a.wpr; p.nw = a.nw; p.row_stride = a.row_stride;
      p.e = entries(); p.first_new = (uint32_t)n_entries; p.n_new = n_new; p.arena = arena.p; p.arena_cursor = ctr.p + 1;
      kpaths<<<(n_new + GENO_THREADS - 1) / GENO_THREADS, GENO_THREADS, smem_paths, stream>>>(p);
      HLAC_CUDA(cudaGetLastError());
      launches++;
    }
    if (verify) batch_pid.reserve(n, stream);
    k_intern_lookup<<<g1, 256, 0, stream>>>(table(), entries(), cov.p, hash_a.p, hash_b.p, b.cell, b.umi, n, reads_pushed, b.ordinal,
                                            rec_key.p, rec_path.p, ctr.p + 2, record_reads ? read_path.p + reads_pushed : nullptr, ctr.p + 3,
                                            verify ? batch_pid.p : nullptr);
    HLAC_CUDA(cudaGetLastError());
    launches++;
    if (verify) {
      auto* const kver = stride == 3 ? k_geno_verify<3> : k_geno_verify<1>;
      HLAC_CUDA(cudaFuncSetAttribute(kver, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_paths));
      GenoPathArgs p{};
      p.ix = idx->view; p.seq = b.seq; p.len = b.len; p.strand = strand.p; p.wpr = a.wpr; p.nw = a.nw; p.row_stride = a.row_stride;
      p.e = entries(); p.arena = arena.p;
      kver<<<g1, GENO_THREADS, smem_paths, stream>>>(p, cov.p, batch_pid.p, n, ctr.p + 3);
      HLAC_CUDA(cudaGetLastError());
      launches++;
    }
    unsigned long long c2[3];
    HLAC_CUDA(cudaMemcpyAsync(c2, ctr.p + 1, sizeof c2, cudaMemcpyDeviceToHost, stream));
    end(e1);
    HLAC_CUDA(cudaStreamSynchronize(stream));
    n_entries = ne; arena_used = c2[0]; n_records_ub = c2[1];
    if (c2[2]) throw Error(HLAC_ERR_STATE, "path hash collision detected (a read's colour path differs from the interned path its hashes matched)");
    reads_pushed += n; last_n = n;
  }
};

extern "C" {

int hlac_geno_new(hlac_index* idx, hlac_geno** out) try {
  if (!idx || !out) throw Error(HLAC_ERR_ARG, "null argument");
  if (idx->device < 0) throw Error(HLAC_ERR_CUDA, "index is host-only (device -1): genotyping needs a CUDA device, there is no CPU fallback");
  auto s = std::make_unique<hlac_geno>();
  s->idx = idx; s->device = idx->device; s->stride = seed_stride();
  if (const char* e = getenv("HLAC_GENO_VERIFY")) s->verify = atoi(e) != 0;
  DeviceGuard g(s->device);
  HLAC_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  s->ctr.reserve(16, s->stream);
  HLAC_CUDA(cudaMemsetAsync(s->ctr.p, 0, 16 * sizeof(unsigned long long), s->stream));
  cudaDeviceProp prop; HLAC_CUDA(cudaGetDeviceProperties(&prop, s->device));
  s->n_sm = prop.multiProcessorCount;
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  *out = s.release();
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_geno_free(hlac_geno* s) {
  if (s) { int prev = 0; cudaGetDevice(&prev); cudaSetDevice(s->device); delete s; cudaSetDevice(prev); }   // the caller's current device is left as it was
  return HLAC_OK;
}

int hlac_geno_attach_comm(hlac_geno* s, hlac_comm* comm) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  if (comm && comm_device(comm) != s->device) throw Error(HLAC_ERR_ARG, "communicator and session live on different devices");
  if (s->merged) throw Error(HLAC_ERR_STATE, "this session has already merged its paths");
  s->comm = comm;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_geno_record_reads(hlac_geno* s, int on) {
  if (!s) { set_last_error("null session"); return HLAC_ERR_ARG; }
  if (s->reads_pushed) { set_last_error("hlac_geno_record_reads must be called before the first push"); return HLAC_ERR_STATE; }
  s->record_reads = on != 0;
  return HLAC_OK;
}

int hlac_geno_push_device(hlac_geno* s, const hlac_batch* b, int forward_only) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  DeviceGuard g(s->device);
  s->push_device(*b, forward_only);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_geno_push(hlac_geno* s, const hlac_batch* b, int forward_only) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  if (b->n == 0) return HLAC_OK;
  DeviceGuard g(s->device);
  size_t e = s->begin(3);
  s->st_seq.upload(b->seq, b->n * b->words_per_read, s->stream);
  s->st_len.upload(b->len, b->n, s->stream);
  s->st_cell.upload(b->cell, b->n, s->stream);
  s->st_umi.upload(b->umi, b->n, s->stream);
  if (b->ordinal) s->st_ord.upload(b->ordinal, b->n, s->stream);
  s->end(e);
  hlac_batch d = *b;
  d.seq = s->st_seq.p; d.len = s->st_len.p; d.cell = s->st_cell.p; d.umi = s->st_umi.p; d.flags = nullptr;
  d.ordinal = b->ordinal ? s->st_ord.p : nullptr;
  s->push_device(d, forward_only);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_geno_last_cov(hlac_geno* s, uint32_t* cov, uint64_t n) try {
  if (!s || !cov) throw Error(HLAC_ERR_ARG, "null argument");
  if (n != s->last_n) throw Error(HLAC_ERR_ARG, "n does not match the last pushed batch");
  DeviceGuard g(s->device);
  HLAC_CUDA(cudaMemcpyAsync(cov, s->cov.p, n * 4, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

static double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
#define HLAC_TICK(label) do { if (dbg) { double t_ = now_s(); fprintf(stderr, "[hlac timing] %-28s %.3f s\n", label, t_ - t_last); t_last = t_; } } while (0)

int hlac_geno_finish_begin(hlac_geno* s, uint32_t** umi_hist_device, uint64_t* n_classes) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null argument");
  DeviceGuard g(s->device);
  const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr; double t_last = now_s();
  const HostIndex& hx = s->idx->host;
  const uint32_t nitems = (uint32_t)hx.tx_names.size();
  // multi-GPU: every rank's paths become one union (identical on all ranks up to numbering); from here on E counts union paths
  hlac_comm* const comm = s->comm;
  const int n_ranks = comm ? comm_nranks(comm) : 1, my_rank = comm ? comm_rank(comm) : 0;
  if (comm && !s->merged) s->merge_over_comm();
  HLAC_TICK("merge paths over the ranks");
  const uint32_t E = (uint32_t)s->n_entries;
  const uint64_t R = s->n_records_ub;
  (void)nitems;
  auto& cls_reads = s->fin_cls_reads;  // counts_reads
  cls_reads.clear(); s->fin_cls_colour.clear(); s->fin_src_off.clear(); s->fin_cls_len.clear();
  s->class_of_path.assign(E, NONE32);
  size_t n_classes_out = 0;
  if (E) {
    // ---- K3a: resolve every distinct path to its class vector (kept on the device) ----
    DevBuf<uint32_t> clen, clast; clen.reserve(E, s->stream); clast.reserve(E, s->stream);
    size_t ev = s->begin(2);
    k_class_len<<<(E + 255) / 256, 256, 0, s->stream>>>(s->arena.p, s->e_path_off.p, s->e_path_len.p, s->idx->col_off.p, E, clen.p, clast.p);
    HLAC_CUDA(cudaGetLastError());
    std::vector<uint32_t> hlen(E), hlast(E);
    HLAC_CUDA(cudaMemcpyAsync(hlen.data(), clen.p, (size_t)E * 4, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(hlast.data(), clast.p, (size_t)E * 4, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    // multi-GPU: this rank resolves the paths with hash A mod n_ranks == rank; olen = vector length of an owned path, else 0
    std::vector<uint32_t> olen_store; std::vector<uint32_t> owner_count(n_ranks, 0);
    if (comm) {
      DevBuf<uint32_t> d_olen, d_oc; d_olen.reserve(E, s->stream); d_oc.reserve(n_ranks, s->stream);
      HLAC_CUDA(cudaMemsetAsync(d_oc.p, 0, (size_t)n_ranks * 4, s->stream));
      k_owned_mask<<<(E + 255) / 256, 256, 0, s->stream>>>(s->e_hash_a.p, E, (uint32_t)n_ranks, (uint32_t)my_rank, clen.p, d_olen.p, d_oc.p);
      HLAC_CUDA(cudaGetLastError());
      olen_store.resize(E);
      HLAC_CUDA(cudaMemcpyAsync(olen_store.data(), d_olen.p, (size_t)E * 4, cudaMemcpyDeviceToHost, s->stream));
      HLAC_CUDA(cudaMemcpyAsync(owner_count.data(), d_oc.p, (size_t)n_ranks * 4, cudaMemcpyDeviceToHost, s->stream));
      HLAC_CUDA(cudaStreamSynchronize(s->stream));
      s->launches++;
    }
    const std::vector<uint32_t>& olen = comm ? olen_store : hlen;
    std::vector<uint64_t> voff(E + 1, 0);
    for (uint32_t i = 0; i < E; i++) voff[i + 1] = voff[i] + olen[i];
    std::vector<uint32_t> hpl(E);   // path lengths (nodes)
    {   // work of this resolution, for the roofline record: merges, vector elements walked, compulsory bytes
      HLAC_CUDA(cudaMemcpyAsync(hpl.data(), s->e_path_len.p, (size_t)E * 4, cudaMemcpyDeviceToHost, s->stream));
      HLAC_CUDA(cudaStreamSynchronize(s->stream));
      const uint64_t bw = ((uint64_t)hx.tx_names.size() + 31) / 32 * 4;
      s->res_bytes = s->res_merges = s->res_steps = 0; s->res_paths = E;
      s->res_paths = 0;
      for (uint32_t i = 0; i < E; i++) {
        if (olen[i] == 0) continue;   // resolved by another rank
        const uint64_t m = hpl[i] ? hpl[i] - 1 : 0;
        s->res_paths++; s->res_merges += m; s->res_steps += m * hlen[i]; s->res_bytes += m * bw + (uint64_t)hpl[i] * 4 + (uint64_t)hlen[i] * 4;
      }
    }
    DevBuf<uint64_t> d_voff, d_vhash, d_vhash2;
    d_voff.upload(voff.data(), E + 1, s->stream); d_vhash.reserve(E, s->stream); d_vhash2.reserve(E, s->stream);
    s->fin_dvec.reserve(std::max<uint64_t>(voff[E], 1), s->stream);
    const size_t ev_res = s->begin(4);
    {
      uint32_t n_max = 1;
      for (uint32_t i = 0; i < E; i++) n_max = std::max(n_max, olen[i]);
      const uint32_t bitset_words = ((uint32_t)hx.tx_names.size() + 31) / 32;
      const bool narrow = hx.tx_names.size() <= 65535;   // tx ids fit in 16 bits
      // colour membership bitsets: once per index, if they fit comfortably (n_colours x ceil(n_tx/32) words); the kernel then
      // reads membership from them and needs no bitset of its own in shared memory
      const bool bits_global = ((s->idx->col_bits.p && s->idx->col_bits_words == bitset_words) ||
                                (size_t)hx.n_colours() * bitset_words * 4 <= (size_t)4 << 30) && !getenv("HLAC_NO_COLOUR_BITSETS");
      auto smem_for = [&](uint32_t cap) { return (bits_global ? 0 : (size_t)((bitset_words + 3u) & ~3u) * 4) + (size_t)((cap + 7u) & ~7u) * (narrow ? 2 + 2 + 2 : 4 + 4 + 2) + 16; };
      const size_t smem_par = smem_for(n_max);
      const char* force = getenv("HLAC_RESOLVE");   // "seq" / "par" for testing both kernels
      const bool par_ok = n_max < 65535 && smem_par <= 220 * 1024;
      if (par_ok && !(force && std::string(force) == "seq")) {
        // data-parallel merges, one block per path, dynamic scheduling. Paths are grouped by vector length so that the
        // bulk of (short) vectors runs with a small shared-memory footprint = several blocks per SM.
        constexpr int RES_THREADS = 256;   // measured: 512 threads for the long-vector groups changes nothing (0.404 vs 0.409 s), 1024 is slower
        auto kern = narrow ? k_resolve_paths_par<uint16_t, RES_THREADS> : k_resolve_paths_par<uint32_t, RES_THREADS>;
        // colour membership bitsets: once per index, if they fit comfortably (n_colours x ceil(n_tx/32) words)
        hlac_index* ixp = s->idx;
        const uint32_t ncol = hx.n_colours();
        if (bits_global && !(ixp->col_bits.p && ixp->col_bits_words == bitset_words)) {
          ixp->col_bits.reserve((size_t)ncol * bitset_words, s->stream);
          HLAC_CUDA(cudaMemsetAsync(ixp->col_bits.p, 0, (size_t)ncol * bitset_words * 4, s->stream));
          k_colour_bitsets<<<(unsigned)(((uint64_t)ncol * 32 + 255) / 256), 256, 0, s->stream>>>(ixp->col_off.p, ixp->col_tx.p, ncol, bitset_words, ixp->col_bits.p);
          HLAC_CUDA(cudaGetLastError());
          ixp->col_bits_words = bitset_words;
          s->launches++;
        }
        const uint32_t* col_bits = bits_global ? ixp->col_bits.p : nullptr;
        // groups: vectors of up to RES_WARP_MAX elements go to the warp-per-path kernel, longer ones to the block-per-path kernel
        struct Group { bool warp; uint32_t cap; std::vector<uint32_t> ids; };
        std::vector<Group> groups;
        constexpr int RES_WARPS = 4;
        const bool use_warp = bits_global && !getenv("HLAC_RESOLVE_NO_WARP");
        uint32_t warp_max = RES_WARP_MAX;
        if (const char* e = getenv("HLAC_RESOLVE_WARP_MAX")) warp_max = std::min<uint32_t>(8192, std::max<uint32_t>(256, (uint32_t)atoi(e)));   // A/B switch
        if (!narrow) warp_max = std::min<uint32_t>(warp_max, 4096);   // 4 warps x 10 bytes x cap of shared memory per block
        // capacity classes: shared memory per path is proportional to the class capacity and decides how many paths an SM
        // works on at once, so the classes are fine (warp kernel: steps of 256 elements; block kernel: x 1.25), not octaves -
        // e.g. the 74 k paths of 1 025..1 641 elements of the config-3 database run 21 instead of 16 to an SM
        const bool fine = !getenv("HLAC_RESOLVE_COARSE");
        if (use_warp) for (uint32_t cap = 256; cap <= warp_max; cap = fine ? cap + 256 : cap * 2) groups.push_back({true, cap, {}});
        const size_t first_block = groups.size();
        if (fine) {
          uint32_t cap = use_warp ? warp_max : 1024;
          while (cap < n_max && groups.size() - first_block < 15) { cap = std::min<uint32_t>(n_max, (cap + cap / 4 + 63) & ~63u); groups.push_back({false, cap, {}}); }
          if (groups.size() == first_block || groups.back().cap < n_max) groups.push_back({false, n_max, {}});
        } else {
          std::vector<uint32_t> caps;
          for (uint32_t cap = n_max; ; cap = (cap + 1) / 2) { caps.push_back(cap); if (cap <= 2048 || caps.size() == 4) break; }
          for (size_t k = caps.size(); k-- > 0;) groups.push_back({false, caps[k], {}});   // ascending capacity
        }
        for (uint32_t i = 0; i < E; i++) {
          if (olen[i] == 0) continue;   // another rank's path
          size_t gsel = (use_warp && olen[i] <= groups[first_block - 1].cap) ? 0 : first_block;
          while (gsel + 1 < groups.size() && olen[i] > groups[gsel].cap) gsel++;
          groups[gsel].ids.push_back(i);
        }
        DevBuf<unsigned int> d_next; d_next.reserve(groups.size(), s->stream);
        HLAC_CUDA(cudaMemsetAsync(d_next.p, 0, groups.size() * 4, s->stream));
        std::vector<std::unique_ptr<DevBuf<uint32_t>>> d_ids;
        auto wkern = narrow ? k_resolve_paths_warp<uint16_t, RES_WARPS> : k_resolve_paths_warp<uint32_t, RES_WARPS>;
        const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr;
        std::vector<cudaEvent_t> gev;
        for (size_t gsel = 0; gsel < groups.size(); gsel++) {   // short vectors first
          Group& g = groups[gsel];
          if (dbg) { cudaEvent_t e; HLAC_CUDA(cudaEventCreate(&e)); HLAC_CUDA(cudaEventRecord(e, s->stream)); gev.push_back(e); }
          if (g.ids.empty()) continue;
          d_ids.emplace_back(new DevBuf<uint32_t>());
          d_ids.back()->upload(g.ids.data(), g.ids.size(), s->stream);
          int occ = 1;
          if (g.warp) {
            const size_t smem_g = (size_t)RES_WARPS * g.cap * (narrow ? 6 : 10);
            HLAC_CUDA(cudaFuncSetAttribute(wkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
            HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wkern, RES_WARPS * 32, smem_g));
            const unsigned grid = (unsigned)std::min<uint64_t>((g.ids.size() + RES_WARPS - 1) / RES_WARPS, (uint64_t)s->n_sm * std::max(occ, 1));
            wkern<<<grid, RES_WARPS * 32, smem_g, s->stream>>>(s->arena.p, s->e_path_off.p, s->e_path_len.p, s->idx->col_off.p, s->idx->col_tx.p,
                                                              (uint32_t)g.ids.size(), d_voff.p, s->fin_dvec.p, d_vhash.p, d_vhash2.p, g.cap,
                                                              bitset_words, d_next.p + gsel, d_ids.back()->p, col_bits);
          } else {
            const size_t smem_g = smem_for(g.cap);
            HLAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_par));
            HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RES_THREADS, smem_g));
            const unsigned grid = (unsigned)std::min<uint64_t>(g.ids.size(), (uint64_t)s->n_sm * std::max(occ, 1));
            kern<<<grid, RES_THREADS, smem_g, s->stream>>>(s->arena.p, s->e_path_off.p, s->e_path_len.p, s->idx->col_off.p, s->idx->col_tx.p,
                                                          (uint32_t)g.ids.size(), d_voff.p, s->fin_dvec.p, d_vhash.p, d_vhash2.p, g.cap,
                                                          bitset_words, d_next.p + gsel, d_ids.back()->p, col_bits);
          }
          HLAC_CUDA(cudaGetLastError());
          s->launches++;
        }
        if (dbg) { cudaEvent_t e; HLAC_CUDA(cudaEventCreate(&e)); HLAC_CUDA(cudaEventRecord(e, s->stream)); gev.push_back(e); }
        HLAC_CUDA(cudaStreamSynchronize(s->stream));   // d_next / d_ids go out of scope
        if (dbg) {
          for (size_t gsel = 0; gsel < groups.size(); gsel++) {
            float ms = 0; cudaEventElapsedTime(&ms, gev[gsel], gev[gsel + 1]);
            uint64_t merges = 0, steps = 0;
            for (uint32_t i : groups[gsel].ids) { merges += hpl[i] - 1; steps += (uint64_t)(hpl[i] - 1) * olen[i]; }
            fprintf(stderr, "[hlac] resolve group %zu (%s, cap %u): %zu paths, %llu merges, %llu element steps, %.3f ms\n", gsel, groups[gsel].warp ? "warp" : "block",
                    groups[gsel].cap, groups[gsel].ids.size(), (unsigned long long)merges, (unsigned long long)steps, ms);
          }
          for (cudaEvent_t e : gev) cudaEventDestroy(e);
        }
      } else {
        if (force && std::string(force) == "par") throw Error(HLAC_ERR_LIMIT, "HLAC_RESOLVE=par but the class vectors do not fit in shared memory");
        if (comm) throw Error(HLAC_ERR_LIMIT, "the sequential path resolution (class vectors beyond shared memory, or HLAC_RESOLVE=seq) is single-GPU only");
        k_resolve_paths<<<(E + 127) / 128, 128, 0, s->stream>>>(s->arena.p, s->e_path_off.p, s->e_path_len.p, s->idx->col_off.p, s->idx->col_tx.p, E,
                                                               d_voff.p, s->fin_dvec.p, d_vhash.p, d_vhash2.p);
      }
    }
    HLAC_CUDA(cudaGetLastError());
    s->end(ev_res);
    if (comm) {   // every rank contributes the fingerprints (hash A, vector hash pair) of the paths it resolved
      const double t0 = now_ms();
      std::vector<size_t> cnt3(n_ranks), dis3(n_ranks); size_t tot = 0;
      for (int r = 0; r < n_ranks; r++) { cnt3[r] = (size_t)owner_count[r] * 3; dis3[r] = tot; tot += cnt3[r]; }
      DevBuf<uint64_t> d_send, d_all; DevBuf<unsigned long long> d_cur;
      d_send.reserve(std::max<size_t>(cnt3[my_rank], 1), s->stream); d_all.reserve(std::max<size_t>(tot, 1), s->stream); d_cur.reserve(1, s->stream);
      HLAC_CUDA(cudaMemsetAsync(d_cur.p, 0, 8, s->stream));
      k_pack_fingerprints<<<(E + 255) / 256, 256, 0, s->stream>>>(s->e_hash_a.p, d_vhash.p, d_vhash2.p, E, (uint32_t)n_ranks, (uint32_t)my_rank, d_send.p, d_cur.p);
      comm_allgatherv(comm, d_send.p, d_all.p, cnt3.data(), dis3.data(), HLAC_COMM_U64, s->stream);
      k_unpack_fingerprints<<<(unsigned)((tot / 3 + 255) / 256), 256, 0, s->stream>>>(s->table(), d_all.p, tot / 3, d_vhash.p, d_vhash2.p);
      HLAC_CUDA(cudaGetLastError());
      HLAC_CUDA(cudaStreamSynchronize(s->stream));
      s->launches += 2;
      s->exchange_ms = now_ms() - t0;
    }
    s->end(ev);
    s->launches += 2;
    std::vector<uint64_t> vhash(E), vhash2(E), ford(E); std::vector<uint32_t> cnt(E);
    HLAC_CUDA(cudaMemcpyAsync(vhash.data(), d_vhash.p, (size_t)E * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(vhash2.data(), d_vhash2.p, (size_t)E * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(ford.data(), s->e_first_ord.p, (size_t)E * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(cnt.data(), s->e_count.p, (size_t)E * 4, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    HLAC_TICK("resolve paths");
    // ---- host: intern equal vectors by (hash pair, length, last colour); exact equality is verified on the device ----
    struct Cls { uint32_t rep; uint64_t ord; uint32_t reads; };
    std::vector<Cls> classes; std::vector<uint32_t> tmp_cls(E), rep_of_path(E);
    {   // open-addressing table over the first vector hash; candidates compare (second hash, length, last colour)
      uint64_t cap = 64; while (cap < 2 * (uint64_t)E) cap <<= 1;
      std::vector<uint32_t> slot(cap, NONE32);   // class index
      classes.reserve(E / 2 + 16);
      for (uint32_t p = 0; p < E; p++) {
        uint64_t h = (vhash[p] * 0x9E3779B97F4A7C15ULL) >> 20 & (cap - 1);
        uint32_t found = NONE32;
        for (;; h = (h + 1) & (cap - 1)) {
          const uint32_t c = slot[h];
          if (c == NONE32) break;
          const uint32_t r = classes[c].rep;
          if (vhash[r] == vhash[p] && vhash2[r] == vhash2[p] && hlen[r] == hlen[p] && hlast[r] == hlast[p]) { found = c; break; }
        }
        if (found == NONE32) { found = (uint32_t)classes.size(); classes.push_back({p, ford[p], 0}); slot[h] = found; }
        classes[found].ord = std::min(classes[found].ord, ford[p]);
        classes[found].reads += cnt[p];
        tmp_cls[p] = found;
        rep_of_path[p] = classes[found].rep;
      }
    }
    if (!comm) {   // (with several ranks the representative's vector may live on another rank: identity then rests on the
                   // 128-bit vector hash + length + last colour alone)
      DevBuf<uint32_t> d_rep; DevBuf<unsigned long long> d_bad;
      d_rep.upload(rep_of_path.data(), E, s->stream); d_bad.reserve(1, s->stream);
      HLAC_CUDA(cudaMemsetAsync(d_bad.p, 0, 8, s->stream));
      k_verify_classes<<<(unsigned)(((uint64_t)E * 32 + 255) / 256), 256, 0, s->stream>>>(s->fin_dvec.p, d_voff.p, d_rep.p, E, d_bad.p);
      HLAC_CUDA(cudaGetLastError());
      unsigned long long bad = 0;
      HLAC_CUDA(cudaMemcpyAsync(&bad, d_bad.p, 8, cudaMemcpyDeviceToHost, s->stream));
      HLAC_CUDA(cudaStreamSynchronize(s->stream));
      s->launches++;
      if (bad) throw Error(HLAC_ERR_STATE, "class vector hash collision detected");
    }
    std::vector<uint64_t> hhash;   // multi-GPU: a class is represented by its path with the smallest hash A (the same path on every
                                   // rank whatever the local numbering), whose owner holds the class vector
    if (comm) {
      hhash.resize(E);
      HLAC_CUDA(cudaMemcpyAsync(hhash.data(), s->e_hash_a.p, (size_t)E * 8, cudaMemcpyDeviceToHost, s->stream));
      HLAC_CUDA(cudaStreamSynchronize(s->stream));
      for (uint32_t p = 0; p < E; p++) { Cls& c = classes[tmp_cls[p]]; if (hhash[p] < hhash[c.rep]) c.rep = p; }
    }
    HLAC_TICK("intern + verify classes");
    // class ids in first-seen order (first-seen ordinals are unique): LSD radix sort of the class indices by ordinal, 16 bits a pass
    std::vector<uint32_t> order(classes.size());
    {
      const size_t nc = classes.size();
      uint64_t max_ord = 0;
      for (const Cls& c : classes) max_ord = std::max(max_ord, c.ord);
      std::vector<uint32_t> other(nc), hist(65536);
      for (uint32_t i = 0; i < nc; i++) order[i] = i;
      for (int shift = 0; shift < 64 && (shift == 0 || (max_ord >> shift) != 0); shift += 16) {
        std::fill(hist.begin(), hist.end(), 0u);
        for (uint32_t i = 0; i < nc; i++) hist[(classes[order[i]].ord >> shift) & 0xFFFFu]++;
        uint32_t run = 0;
        for (uint32_t d = 0; d < 65536; d++) { const uint32_t h = hist[d]; hist[d] = run; run += h; }
        for (uint32_t i = 0; i < nc; i++) other[hist[(classes[order[i]].ord >> shift) & 0xFFFFu]++] = order[i];
        order.swap(other);
      }
    }
    std::vector<uint32_t> rank(classes.size());
    for (uint32_t r = 0; r < order.size(); r++) rank[order[r]] = r;
    cls_reads.resize(classes.size()); s->fin_cls_colour.resize(classes.size()); s->fin_src_off.resize(classes.size()); s->fin_cls_len.resize(classes.size());
    s->fin_cls_owner.assign(classes.size(), 0);
    for (uint32_t c = 0; c < classes.size(); c++) {
      const uint32_t r = classes[c].rep;
      s->fin_cls_colour[rank[c]] = hlast[r];
      s->fin_src_off[rank[c]] = voff[r];   // (valid on the rank that owns the representative)
      s->fin_cls_len[rank[c]] = hlen[r];
      cls_reads[rank[c]] = classes[c].reads;
      if (comm) s->fin_cls_owner[rank[c]] = (uint32_t)(hhash[r] % (uint64_t)n_ranks);
    }
    for (uint32_t p = 0; p < E; p++) s->class_of_path[p] = rank[tmp_cls[p]];
    n_classes_out = classes.size();
    HLAC_TICK("order classes");
    // ---- K3b: group records by (barcode, umi), lowest class id per UMI ----
    s->fin_hist.reserve(std::max<size_t>(classes.size(), 1), s->stream);
    HLAC_CUDA(cudaMemsetAsync(s->fin_hist.p, 0, std::max<size_t>(classes.size(), 1) * 4, s->stream));
    if (R) {
      DevBuf<uint32_t> d_rank_of_path, d_ranks, d_ranks_b; DevBuf<uint64_t> d_keys_a, d_keys_b; RadixSortScratch rs;
      uint32_t* const hist = s->fin_hist.p;
      d_rank_of_path.upload(s->class_of_path.data(), E, s->stream);
      d_ranks.reserve(R, s->stream); d_ranks_b.reserve(R, s->stream); d_keys_a.reserve(R, s->stream); d_keys_b.reserve(R, s->stream);
      size_t e3 = s->begin(2);
      k_map_rank<<<(unsigned)((R + 255) / 256), 256, 0, s->stream>>>(s->rec_path.p, d_rank_of_path.p, R, d_ranks.p);
      HLAC_CUDA(cudaGetLastError());
      // group the records by (barcode, UMI): stable radix sort of a copy of the keys (the records themselves stay as pushed)
      HLAC_CUDA(cudaMemcpyAsync(d_keys_a.p, s->rec_key.p, R * 8, cudaMemcpyDeviceToDevice, s->stream));
      const bool second = radix_sort_pairs(d_keys_a.p, d_keys_b.p, d_ranks.p, d_ranks_b.p, R, 0, 64, rs, s->stream);
      k_umi_min<<<(unsigned)((R + 255) / 256), 256, 0, s->stream>>>(second ? d_keys_b.p : d_keys_a.p, second ? d_ranks_b.p : d_ranks.p, R, hist);
      HLAC_CUDA(cudaGetLastError());
      s->end(e3);
      s->launches += 3;
      HLAC_CUDA(cudaStreamSynchronize(s->stream));
    }
  }
  HLAC_TICK("sort records + umi min");
  if (!E) { s->fin_hist.reserve(1, s->stream); HLAC_CUDA(cudaMemsetAsync(s->fin_hist.p, 0, 4, s->stream)); HLAC_CUDA(cudaStreamSynchronize(s->stream)); }
  if (comm) {   // every (barcode, UMI) group is complete on one rank (reads are partitioned by barcode): the histograms add up
    comm_allreduce_sum(comm, s->fin_hist.p, s->fin_hist.p, std::max<size_t>(n_classes_out, 1), HLAC_COMM_U32, s->stream);
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    HLAC_TICK("all-reduce of the UMI histogram");
  }
  s->fin_begun = true;
  if (umi_hist_device) *umi_hist_device = s->fin_hist.p;
  if (n_classes) *n_classes = n_classes_out;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// Second half of eq_class_counts: reads the (possibly all-reduced) per-class UMI histogram back and assembles the tables.
int hlac_geno_finish_end(hlac_geno* s, hlac_class_tables* out) try {
  if (!s || !out) throw Error(HLAC_ERR_ARG, "null argument");
  if (!s->fin_begun) throw Error(HLAC_ERR_STATE, "hlac_geno_finish_begin has not run");
  DeviceGuard g(s->device);
  const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr; double t_last = now_s();
  auto& cls_reads = s->fin_cls_reads;
  const size_t nc = cls_reads.size();
  const HostIndex& hx = s->idx->host;
  const uint32_t nitems = (uint32_t)hx.tx_names.size();
  std::vector<uint32_t> umi_hist(std::max<size_t>(nc, 1), 0);
  HLAC_CUDA(cudaMemcpyAsync(umi_hist.data(), s->fin_hist.p, umi_hist.size() * 4, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  memset(out, 0, sizeof *out);
  out->nitems = nitems;
  uint64_t total_reads = 0; for (uint32_t c : cls_reads) total_reads += c;
  out->nreads = (uint32_t)total_reads;   // total_reads (mapping.rs:195): every counted read, over all ranks after a path merge
  out->reads_explained = (uint64_t*)calloc(std::max<uint32_t>(nitems, 1), 8);
  // ---- counts_reads (mapping.rs:199-201): class vectors gathered into first-seen order on the device, one D2H ----
  out->n_read_classes = nc;
  out->reads_off = (uint64_t*)calloc(nc + 1, 8); out->reads_cnt = (uint32_t*)calloc(std::max<size_t>(nc, 1), 4);
  for (size_t c = 0; c < nc; c++) { out->reads_off[c + 1] = out->reads_off[c] + s->fin_cls_len[c]; out->reads_cnt[c] = cls_reads[c]; }
  const uint64_t ntx = out->reads_off[nc];
  if (s->lean) { free(out->reads_off); free(out->reads_cnt); out->reads_off = nullptr; out->reads_cnt = nullptr; }
  // large tables go to pinned host memory: a pageable 4 GB D2H was measured at < 1 GB/s (page faults under the DMA staging)
  if (s->lean) out->reads_tx = nullptr;
  else if (ntx * 4 >= (64ull << 20)) HLAC_CUDA(cudaHostAlloc((void**)&out->reads_tx, ntx * 4, cudaHostAllocDefault));
  else out->reads_tx = (uint32_t*)malloc(std::max<uint64_t>(ntx, 1) * 4);
  if (nc) {
    DevBuf<uint64_t> d_src, d_dst; DevBuf<uint32_t> d_out;
    if (!s->lean && !s->comm) {
      d_src.upload(s->fin_src_off.data(), nc, s->stream); d_dst.upload(out->reads_off, nc + 1, s->stream); d_out.reserve(std::max<uint64_t>(ntx, 1), s->stream);
      k_gather_classes<<<(unsigned)nc, 256, 0, s->stream>>>(s->fin_dvec.p, d_src.p, d_dst.p, d_out.p);
      HLAC_CUDA(cudaGetLastError());
      HLAC_CUDA(cudaMemcpyAsync(out->reads_tx, d_out.p, ntx * 4, cudaMemcpyDeviceToHost, s->stream));
    } else if (!s->lean) {
      // multi-GPU: a class vector lives on the rank that resolved the class's representative path. Every rank packs the
      // vectors it holds (in class order), one variable-size all-gather brings them together, a second gather puts them in
      // class order.
      const int N = comm_nranks(s->comm), Rk = comm_rank(s->comm);
      std::vector<size_t> cnt(N, 0), dis(N, 0);
      std::vector<uint64_t> pack_src, pack_dst, all_src(nc);
      for (size_t c = 0; c < nc; c++) {
        const uint32_t o = s->fin_cls_owner[c];
        all_src[c] = cnt[o];   // offset inside the owner's block, made absolute below
        if ((int)o == Rk) { pack_src.push_back(s->fin_src_off[c]); pack_dst.push_back(cnt[o]); }
        cnt[o] += s->fin_cls_len[c];
      }
      size_t tot = 0; for (int r = 0; r < N; r++) { dis[r] = tot; tot += cnt[r]; }
      for (size_t c = 0; c < nc; c++) all_src[c] += dis[s->fin_cls_owner[c]];
      pack_dst.push_back(cnt[Rk]);
      DevBuf<uint64_t> d_ps, d_pd; DevBuf<uint32_t> d_pack, d_all;
      d_pack.reserve(std::max<size_t>(cnt[Rk], 1), s->stream); d_all.reserve(std::max<size_t>(tot, 1), s->stream);
      if (pack_src.size()) {
        d_ps.upload(pack_src.data(), pack_src.size(), s->stream); d_pd.upload(pack_dst.data(), pack_dst.size(), s->stream);
        k_gather_classes<<<(unsigned)pack_src.size(), 256, 0, s->stream>>>(s->fin_dvec.p, d_ps.p, d_pd.p, d_pack.p);
      }
      comm_allgatherv(s->comm, d_pack.p, d_all.p, cnt.data(), dis.data(), HLAC_COMM_U32, s->stream);
      d_src.upload(all_src.data(), nc, s->stream); d_dst.upload(out->reads_off, nc + 1, s->stream); d_out.reserve(std::max<uint64_t>(ntx, 1), s->stream);
      k_gather_classes<<<(unsigned)nc, 256, 0, s->stream>>>(d_all.p, d_src.p, d_dst.p, d_out.p);
      HLAC_CUDA(cudaGetLastError());
      HLAC_CUDA(cudaMemcpyAsync(out->reads_tx, d_out.p, ntx * 4, cudaMemcpyDeviceToHost, s->stream));
      HLAC_CUDA(cudaStreamSynchronize(s->stream));   // the packing buffers go out of scope
    }
    // reads_explained (mapping.rs:196-198): per colour totals, then one pass over the used colour lists on the device
    std::vector<unsigned long long> per_colour(hx.n_colours(), 0);
    for (size_t c = 0; c < nc; c++) per_colour[s->fin_cls_colour[c]] += cls_reads[c];
    std::vector<uint32_t> cols; std::vector<unsigned long long> tots;
    for (uint32_t col = 0; col < per_colour.size(); col++) if (per_colour[col]) { cols.push_back(col); tots.push_back(per_colour[col]); }
    DevBuf<uint32_t> d_cols; DevBuf<unsigned long long> d_tots, d_expl;
    d_cols.upload(cols.data(), cols.size(), s->stream); d_tots.upload(tots.data(), tots.size(), s->stream); d_expl.reserve(std::max<uint32_t>(nitems, 1), s->stream);
    HLAC_CUDA(cudaMemsetAsync(d_expl.p, 0, (size_t)std::max<uint32_t>(nitems, 1) * 8, s->stream));
    k_reads_explained<<<(unsigned)((cols.size() * 32 + 255) / 256), 256, 0, s->stream>>>(d_cols.p, d_tots.p, (uint32_t)cols.size(), s->idx->col_off.p, s->idx->col_tx.p, d_expl.p);
    HLAC_CUDA(cudaGetLastError());
    HLAC_CUDA(cudaMemcpyAsync(out->reads_explained, d_expl.p, (size_t)nitems * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    s->launches += 2;
  }
  uint64_t q = 0;
  HLAC_TICK("assemble counts_reads");
  // counts_umi key = sort + dedup of the UMI's class (mapping.rs:205-207). A class vector is a permutation of the
  // colour list of its path's last node, and colour lists are sorted, duplicate-free and pairwise distinct, so the key
  // IS that colour list: accumulate per colour id, then emit in lexicographic order of the lists.
  // (The reference's HashMap has no order; the classes are emitted in ascending colour id - the index's own, identical on
  // every rank and run.)
  std::vector<uint64_t> by_colour(hx.n_colours(), 0);
  for (size_t c = 0; c < nc; c++) if (umi_hist[c]) by_colour[s->fin_cls_colour[c]] += umi_hist[c];
  std::vector<uint32_t> used;
  uint64_t utx = 0;
  for (uint32_t col = 0; col < by_colour.size(); col++) if (by_colour[col]) { used.push_back(col); utx += hx.col_off[col + 1] - hx.col_off[col]; }
  out->n_umi_classes = used.size();
  out->umi_off = (uint64_t*)calloc(used.size() + 1, 8); out->umi_tx = (uint32_t*)malloc(std::max<uint64_t>(utx, 1) * 4); out->umi_cnt = (uint32_t*)calloc(std::max<size_t>(used.size(), 1), 4);
  q = 0; size_t c = 0;
  for (uint32_t col : used) { q += hx.col_off[col + 1] - hx.col_off[col]; out->umi_cnt[c] = (uint32_t)by_colour[col]; out->umi_off[++c] = q; }
  {   // the lists themselves (tens of MB into fresh pages at IMGT scale: the page faults are the cost, and they parallelise)
    unsigned n_threads = utx >= (1u << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
    if (const char* e = getenv("HLAC_HOST_THREADS")) n_threads = utx >= (1u << 20) ? (unsigned)std::max(1, atoi(e)) : 1u;
    auto copy_range = [&](size_t a, size_t b) {
      for (size_t k = a; k < b; k++) {
        const uint32_t col = used[k];
        memcpy(out->umi_tx + out->umi_off[k], hx.col_tx.data() + hx.col_off[col], (size_t)(hx.col_off[col + 1] - hx.col_off[col]) * 4);
      }
    };
    if (n_threads <= 1) copy_range(0, used.size());
    else {   // chunks of about equal bytes
      std::vector<std::thread> th;
      size_t a = 0;
      for (unsigned t = 0; t < n_threads; t++) {
        const uint64_t want = utx * (t + 1) / n_threads;
        size_t b = a;
        while (b < used.size() && out->umi_off[b] < want) b++;
        if (t + 1 == n_threads) b = used.size();
        if (b > a) th.emplace_back(copy_range, a, b);
        a = b;
      }
      for (auto& x : th) x.join();
    }
  }
  HLAC_TICK("assemble counts_umi");
  s->finished = true; s->fin_begun = false;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_geno_finish(hlac_geno* s, hlac_class_tables* out) {
  int rc = hlac_geno_finish_begin(s, nullptr, nullptr);
  if (rc != HLAC_OK) return rc;
  return hlac_geno_finish_end(s, out);
}

int hlac_geno_set_lean(hlac_geno* s, int on) {
  if (!s) { set_last_error("null session"); return HLAC_ERR_ARG; }
  s->lean = on != 0;
  return HLAC_OK;
}

// The caller (em.rs:361-434) on a finished session: pair_reads_explained (em.rs:36-44) is computed from the class
// vectors that are still on the device (one warp per class marks which candidate alleles it contains), so it also
// works in lean mode where counts_reads never reaches the host.
int hlac_geno_call(hlac_geno* s, const hlac_class_tables* t, const double* theta, hlac_calls* out) try {
  if (!s || !t || !theta || !out) throw Error(HLAC_ERR_ARG, "null argument");
  if (!s->finished) throw Error(HLAC_ERR_STATE, "hlac_geno_finish has not run");
  DeviceGuard g(s->device);
  const auto& names = s->idx->host.tx_names;
  if (names.size() != t->nitems) throw Error(HLAC_ERR_ARG, "class tables do not match the index");
  const Ranked rk = rank_alleles(names, t->nitems, t->reads_explained, theta);
  if (rk.cands.size() > 128) throw Error(HLAC_ERR_LIMIT, "more than 128 candidate alleles");
  const size_t nc = s->fin_cls_reads.size();
  std::vector<unsigned long long> single(128, 0), both(128 * 128, 0);
  if (nc && !rk.cands.empty()) {
    std::vector<uint8_t> cand_idx(t->nitems, 0xFF);
    for (size_t k = 0; k < rk.cands.size(); k++) cand_idx[rk.cands[k]] = (uint8_t)k;
    DevBuf<uint8_t> d_ci; DevBuf<uint64_t> d_mask; DevBuf<uint32_t> d_col, d_reads; DevBuf<unsigned long long> d_acc;
    d_ci.upload(cand_idx.data(), cand_idx.size(), s->stream); d_col.upload(s->fin_cls_colour.data(), nc, s->stream); d_mask.reserve(2 * nc, s->stream);
    d_reads.upload(s->fin_cls_reads.data(), nc, s->stream); d_acc.reserve(128 + 128 * 128, s->stream);
    HLAC_CUDA(cudaMemsetAsync(d_acc.p, 0, (128 + 128 * 128) * sizeof(unsigned long long), s->stream));
    k_class_cand_mask_colour<<<(unsigned)((nc * 32 + 255) / 256), 256, 0, s->stream>>>(d_col.p, (uint32_t)nc, s->idx->col_off.p, s->idx->col_tx.p, d_ci.p, d_mask.p);
    k_pair_matrix<<<(unsigned)((nc + 255) / 256), 256, 0, s->stream>>>(d_mask.p, d_reads.p, (uint32_t)nc, d_acc.p, d_acc.p + 128);
    HLAC_CUDA(cudaGetLastError());
    HLAC_CUDA(cudaMemcpyAsync(single.data(), d_acc.p, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(both.data(), d_acc.p + 128, 128 * 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    s->launches += 2;
  }
  auto pair_explained = [&](uint32_t a, uint32_t b) {
    int ka = rk.cand_of[a], kb = rk.cand_of[b];
    if (ka > kb) std::swap(ka, kb);
    return (uint32_t)(ka == kb ? single[ka] : single[ka] + single[kb] - both[ka * 128 + kb]);
  };
  call_from_ranked(rk, t->nreads, pair_explained, out);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// ---- multi-GPU merge of the interned paths ----------------------------------------------------
// Reads are partitioned by barcode, so every (barcode, UMI) group is complete on one rank, but class identity and the
// first-seen class order are global. Ranks exchange their distinct colour paths (tiny: a few ids each), every rank
// builds the same union (hlac_paths_merge), adopts it (hlac_geno_import_paths) and then only the per-class UMI
// histogram needs an all-reduce between hlac_geno_finish_begin and hlac_geno_finish_end.
int hlac_geno_export_paths(hlac_geno* s, hlac_paths* out) try {
  if (!s || !out) throw Error(HLAC_ERR_ARG, "null argument");
  DeviceGuard g(s->device);
  const uint64_t E = s->n_entries;
  memset(out, 0, sizeof *out);
  out->n_paths = E;
  out->hash_a = (uint64_t*)calloc(std::max<uint64_t>(E, 1), 8); out->hash_b = (uint64_t*)calloc(std::max<uint64_t>(E, 1), 8);
  out->first_ord = (uint64_t*)calloc(std::max<uint64_t>(E, 1), 8); out->count = (uint32_t*)calloc(std::max<uint64_t>(E, 1), 4);
  out->off = (uint64_t*)calloc(E + 1, 8);
  std::vector<uint64_t> poff(E); std::vector<uint32_t> plen(E), arena(std::max<uint64_t>(s->arena_used, 1));
  if (E) {
    HLAC_CUDA(cudaMemcpyAsync(out->hash_a, s->e_hash_a.p, E * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(out->hash_b, s->e_hash_b.p, E * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(out->first_ord, s->e_first_ord.p, E * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(out->count, s->e_count.p, E * 4, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(poff.data(), s->e_path_off.p, E * 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(plen.data(), s->e_path_len.p, E * 4, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaMemcpyAsync(arena.data(), s->arena.p, s->arena_used * 4, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
  }
  for (uint64_t i = 0; i < E; i++) out->off[i + 1] = out->off[i] + plen[i];
  out->colours = (uint32_t*)calloc(std::max<uint64_t>(out->off[E], 1), 4);
  for (uint64_t i = 0; i < E; i++) memcpy(out->colours + out->off[i], arena.data() + poff[i], (size_t)plen[i] * 4);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_paths_merge(const hlac_paths* parts, uint32_t n_parts, hlac_paths* out) try {
  if (!parts || !out) throw Error(HLAC_ERR_ARG, "null argument");
  struct Ref { uint32_t part; uint64_t idx; uint64_t ord; uint64_t count; };
  std::map<uint64_t, Ref> uni;   // ordered by hash A: every rank builds the identical union
  for (uint32_t p = 0; p < n_parts; p++)
    for (uint64_t i = 0; i < parts[p].n_paths; i++) {
      const hlac_paths& P = parts[p];
      auto it = uni.find(P.hash_a[i]);
      if (it == uni.end()) { uni.emplace(P.hash_a[i], Ref{p, i, P.first_ord[i], P.count[i]}); continue; }
      const hlac_paths& Q = parts[it->second.part]; const uint64_t j = it->second.idx;
      const uint64_t la = P.off[i + 1] - P.off[i], lb = Q.off[j + 1] - Q.off[j];
      if (P.hash_b[i] != Q.hash_b[j] || la != lb || memcmp(P.colours + P.off[i], Q.colours + Q.off[j], la * 4) != 0)
        throw Error(HLAC_ERR_STATE, "path hash collision detected while merging ranks");
      it->second.ord = std::min(it->second.ord, P.first_ord[i]);
      it->second.count += P.count[i];
    }
  const uint64_t E = uni.size();
  memset(out, 0, sizeof *out);
  out->n_paths = E;
  out->hash_a = (uint64_t*)calloc(std::max<uint64_t>(E, 1), 8); out->hash_b = (uint64_t*)calloc(std::max<uint64_t>(E, 1), 8);
  out->first_ord = (uint64_t*)calloc(std::max<uint64_t>(E, 1), 8); out->count = (uint32_t*)calloc(std::max<uint64_t>(E, 1), 4);
  out->off = (uint64_t*)calloc(E + 1, 8);
  uint64_t k = 0, tot = 0;
  for (auto& kv : uni) { const hlac_paths& P = parts[kv.second.part]; tot += P.off[kv.second.idx + 1] - P.off[kv.second.idx]; }
  out->colours = (uint32_t*)calloc(std::max<uint64_t>(tot, 1), 4);
  for (auto& kv : uni) {
    const hlac_paths& P = parts[kv.second.part]; const uint64_t i = kv.second.idx, l = P.off[i + 1] - P.off[i];
    if (kv.second.count > 0xFFFFFFFFull) throw Error(HLAC_ERR_LIMIT, "more than 2^32 reads on one path");
    out->hash_a[k] = kv.first; out->hash_b[k] = P.hash_b[i]; out->first_ord[k] = kv.second.ord; out->count[k] = (uint32_t)kv.second.count;
    memcpy(out->colours + out->off[k], P.colours + P.off[i], l * 4);
    out->off[k + 1] = out->off[k] + l;
    k++;
  }
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_geno_import_paths(hlac_geno* s, const hlac_paths* m) try {
  if (!s || !m) throw Error(HLAC_ERR_ARG, "null argument");
  DeviceGuard g(s->device);
  // local id -> merged id
  const uint64_t E0 = s->n_entries, E = m->n_paths;
  if (E >= 0xFFFFFFFFull) throw Error(HLAC_ERR_LIMIT, "too many distinct paths");
  std::vector<uint64_t> local_a(std::max<uint64_t>(E0, 1));
  if (E0) { HLAC_CUDA(cudaMemcpyAsync(local_a.data(), s->e_hash_a.p, E0 * 8, cudaMemcpyDeviceToHost, s->stream)); HLAC_CUDA(cudaStreamSynchronize(s->stream)); }
  std::vector<uint32_t> map(std::max<uint64_t>(E0, 1));
  for (uint64_t i = 0; i < E0; i++) {
    const uint64_t* it = std::lower_bound(m->hash_a, m->hash_a + E, local_a[i]);   // merged paths are sorted by hash A
    if (it == m->hash_a + E || *it != local_a[i]) throw Error(HLAC_ERR_ARG, "merged path set does not contain this session's paths");
    map[i] = (uint32_t)(it - m->hash_a);
  }
  DevBuf<uint32_t> d_map; d_map.upload(map.data(), map.size(), s->stream);
  if (s->n_records_ub) k_remap<<<(unsigned)((s->n_records_ub + 255) / 256), 256, 0, s->stream>>>(s->rec_path.p, s->n_records_ub, d_map.p);
  if (s->record_reads && s->reads_pushed) k_remap<<<(unsigned)((s->reads_pushed + 255) / 256), 256, 0, s->stream>>>(s->read_path.p, s->reads_pushed, d_map.p);
  HLAC_CUDA(cudaGetLastError());
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  // adopt the merged entries
  s->n_entries = 0;   // nothing to keep while growing
  s->ensure_table(std::max<uint64_t>(E, 1));
  std::vector<uint32_t> plen(std::max<uint64_t>(E, 1)), rep(std::max<uint64_t>(E, 1), 0);
  for (uint64_t i = 0; i < E; i++) plen[i] = (uint32_t)(m->off[i + 1] - m->off[i]);
  s->e_hash_a.upload(m->hash_a, E, s->stream); s->e_hash_b.upload(m->hash_b, E, s->stream);
  s->e_first_ord.upload((const unsigned long long*)m->first_ord, E, s->stream); s->e_count.upload(m->count, E, s->stream);
  s->e_path_off.upload(m->off, E, s->stream); s->e_path_len.upload(plen.data(), E, s->stream); s->e_rep.upload(rep.data(), E, s->stream);
  s->arena.upload(m->colours, std::max<uint64_t>(m->off[E], 1), s->stream);
  HLAC_CUDA(cudaMemsetAsync(s->t_keys.p, 0, s->t_cap * sizeof(unsigned long long), s->stream));
  if (E) k_table_build<<<(unsigned)((E + 255) / 256), 256, 0, s->stream>>>(s->table(), s->e_hash_a.p, (uint32_t)E);
  HLAC_CUDA(cudaGetLastError());
  unsigned long long c2[2] = {E, m->off[E]};
  HLAC_CUDA(cudaMemcpyAsync(s->ctr.p, c2, sizeof c2, cudaMemcpyHostToDevice, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  s->n_entries = E; s->arena_used = m->off[E]; s->finished = false; s->launches += 3;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_paths_free(hlac_paths* p) {
  if (p) { free(p->hash_a); free(p->hash_b); free(p->first_ord); free(p->count); free(p->off); free(p->colours); memset(p, 0, sizeof *p); }
  return HLAC_OK;
}

int hlac_geno_read_class_ids(hlac_geno* s, uint32_t* ids, uint64_t n) try {
  if (!s || !ids) throw Error(HLAC_ERR_ARG, "null argument");
  if (!s->record_reads) throw Error(HLAC_ERR_STATE, "per-read path ids were not recorded (hlac_geno_record_reads)");
  if (!s->finished) throw Error(HLAC_ERR_STATE, "hlac_geno_finish has not run");
  if (n != s->reads_pushed) throw Error(HLAC_ERR_ARG, "n must equal the number of reads pushed");
  DeviceGuard g(s->device);
  HLAC_CUDA(cudaMemcpyAsync(ids, s->read_path.p, n * 4, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  for (uint64_t i = 0; i < n; i++) if (ids[i] != NONE32) ids[i] = s->class_of_path[ids[i]];
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_class_tables_free(hlac_class_tables* t) {
  if (t) {
    cudaPointerAttributes at{};
    if (t->reads_tx && cudaPointerGetAttributes(&at, t->reads_tx) == cudaSuccess && at.type == cudaMemoryTypeHost) { cudaFreeHost(t->reads_tx); t->reads_tx = nullptr; }
    else cudaGetLastError();
    free(t->reads_off); free(t->reads_tx); free(t->reads_cnt); free(t->umi_off); free(t->umi_tx); free(t->umi_cnt); free(t->reads_explained);
    memset(t, 0, sizeof *t);
  }
  return HLAC_OK;
}

int hlac_geno_timing(hlac_geno* s, float ms[4], uint64_t* launches) try {
  if (getenv("HLAC_DEBUG_TIMING")) {
    const AllocStats& a = alloc_cache().st;
    fprintf(stderr, "[hlac] device allocations so far: %llu (%llu from the cache) %.1f ms (longest %.1f ms), %llu frees %.1f ms (longest %.1f ms), %zu MB cached\n",
            a.n_malloc, a.n_reused, 1e3 * a.t_malloc, 1e3 * a.max_malloc, a.n_free, 1e3 * a.t_free, 1e3 * a.max_free, alloc_cache().cached_bytes >> 20);
  }
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  s->drain_events();
  if (ms) for (int i = 0; i < 4; i++) ms[i] = s->ms[i];
  if (launches) *launches = s->launches;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// the last path resolution (hlac_geno_finish*): kernel time and the work it did — distinct paths, no-truncate merges,
// vector elements walked (sum over merges of the vector length) and compulsory bytes (per merge one colour membership
// bitset, per path its colour ids in and its class vector out)
int hlac_geno_resolve_stats(hlac_geno* s, float* ms, uint64_t* paths, uint64_t* merges, uint64_t* element_steps, uint64_t* bytes) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  s->drain_events();
  if (ms) *ms = s->ms[4];
  if (paths) *paths = s->res_paths;
  if (merges) *merges = s->res_merges;
  if (element_steps) *element_steps = s->res_steps;
  if (bytes) *bytes = s->res_bytes;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_geno_stats(hlac_geno* s, uint64_t* some_aln, uint64_t* long_aln, uint64_t* n_paths, uint64_t* bitmap_tests, uint64_t* dict_probes) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  unsigned long long c[4];
  HLAC_CUDA(cudaMemcpyAsync(c, s->ctr.p + 4, sizeof c, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  if (some_aln) *some_aln = c[0];
  if (long_aln) *long_aln = c[1];
  if (n_paths) *n_paths = s->n_entries;
  if (bitmap_tests) *bitmap_tests = c[2];
  if (dict_probes) *dict_probes = c[3];
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

}  // extern "C"
