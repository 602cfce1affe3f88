// Counting session: map_and_count_pseudo's hot loop (src/mapping.rs:740-807) and count() (src/mapping.rs:823-909)
// on the GPU.
//
//   K1  k_count_map_q   queue-staged (default; k_count_map = the earlier tile-staged form): filters (mapping.rs:752-763),
//                       lazy 4-way map_read policy (mapping.rs:765-793), class -> (gene, slot) (mapping.rs:795-806), emits one
//                       packed 64-bit key (umi | cell | code) per scored read through a warp-aggregated append.
//   K3a cub radix sort  keys by (cell, umi)  — the "group by cell then UMI" of mapping.rs:829-841
//   K3b k_consensus     one thread per (cell, UMI) run: gene majority + allele decision (mapping.rs:843-896)
//                       accumulated into a dense u32 [n_cells][nrows] matrix.
//
// Laziness: the reference evaluates all four map_read calls eagerly but only consumes them through the if/else-if
// chain at mapping.rs:772-793; map_read is pure, so evaluating a call only when the chain reaches it yields the
// same `aln` and the same metrics.
//
// Class handling: eq_class_to_gene (mapping.rs:679-696) only holds [a1], [a2] and sorted [a1,a2] per gene, and the
// reference's class is the LAST visited node's colour list permuted by no-truncate merges (mapping.rs:283-308
// semantics inside the pinned map_read; SURVEY.md finding #2). For a 1-element list the merges are no-ops; for a
// 2-element sorted list [p,q] they can only ever swap it to [q,p] (which then misses the table), and that happens
// iff some other visited node's colour contains q but not p. So per colour we precompute `cinfo`: the code it
// yields as a final class and a per-gene "would swap" mask; the walk ORs the masks.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "device_index.hpp"
#include "hla_fasta.hpp"

using namespace hlac;

namespace {

constexpr uint32_t CODE_NONE5 = 31;   // 5-bit "no code"
constexpr int CODE_BITS = 5;
constexpr int UMI_BITS = 32;

struct CountMapArgs {
  DevIndexView cds, gen;
  const uint64_t* seq;
  const uint16_t* len;
  const uint32_t* cell;
  const uint32_t* umi;
  const uint8_t* flags;
  uint64_t n;
  uint32_t wpr;                // u64 words per read
  uint32_t nw;                 // u32 words per strand row in shared memory (2*wpr + 2)
  uint32_t row_stride;         // u32 words per thread (both strands, odd)
  int primary_only;
  uint32_t cell_bits;          // key layout: umi << (5 + cell_bits) | cell << 5 | code
  uint32_t n_cells;            // cell indices >= n_cells are a caller error: dropped and counted in metrics[8]
  uint64_t* keys;
  unsigned long long* cursor;  // number of keys appended so far
  unsigned long long* metrics; // [6] + [2] seed-filter tests / dictionary probes + [1] out-of-range cell indices + [1] max UMI key
  uint8_t* codes;              // optional per-read debug output
  const uint64_t* rec;         // packed records (hlac_packed_batch) instead of the five arrays above, or NULL
};

// The session keeps its own copy of each index's node table in which the colour id (word 3 of the node record) is
// replaced by that colour's cinfo word, so a node visit needs no dependent table lookup.
struct CountVis {
  uint32_t swap_or = 0, last = CODE_NONE5;
  __device__ __forceinline__ void visit(uint32_t ci, uint32_t) {
    swap_or |= ci >> 8;
    last = ci;
  }
  __device__ __forceinline__ uint32_t code() const {
    const uint32_t c = last & 31u;
    if (c == CODE_NONE5) return c;
    if (c % 3 == 2 && ((swap_or >> (c / 3)) & 1u)) return CODE_NONE5;   // [p,q] was permuted to [q,p]
    return c;
  }
};

struct SharedWords {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator[](int i) const { return p[i]; }
};
struct BitmapShared {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i]; }
};
struct BitmapGlobal {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return __ldg(p + i); }
};

// Warp-staged kernel. Each WARP owns a tile of 32*R reads and moves it through stages, compacting the surviving reads
// into dense warp-private task queues in shared memory between stages, so that the lanes of a warp always execute the
// same kind of work and no block-wide barrier is needed (ncu history in profiles/: thread-per-read version 6.65 of 32
// lanes active per issued instruction; block-staged version 12.6 lanes but 30 % of stall samples on __syncthreads):
//   load    filters (mapping.rs:752-763), read rows into shared memory, queue 0
//   scan p  p = cds-fwd, cds-rev, gen-fwd, gen-rev: first-seed skip-scan for the reads still without a seed
//           (map_read is None iff its FIRST seed scan fails, so the laziness of mapping.rs:772-790 is a cascade of scans);
//           reads that find a seed go to the walk queue, the others to the next scan queue (revcomp built at p = 1)
//   walk    every read with a seed: left extension + forward walk + class code + coverage policy, key append
// Queue counters are warp-uniform registers (every lane sees the same ballots), so the queues need no atomics.
template <bool SMEM_BM, int R, int MAP_THREADS, int MINB = 1>
__global__ void __launch_bounds__(MAP_THREADS, MINB) k_count_map(const CountMapArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int TILE = 32 * R;
  uint32_t* wbase = smem;
  if (SMEM_BM) {
    const uint32_t nb = a.cds.bitmap_words + a.gen.bitmap_words;
    for (uint32_t i = threadIdx.x; i < a.cds.bitmap_words; i += MAP_THREADS) smem[i] = __ldg(a.cds.bitmap + i);
    for (uint32_t i = threadIdx.x; i < a.gen.bitmap_words; i += MAP_THREADS) smem[a.cds.bitmap_words + i] = __ldg(a.gen.bitmap + i);
    wbase = smem + nb;
    __syncthreads();   // the only block-wide barrier
  }
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per_warp = TILE * a.row_stride + TILE * 2 + TILE / 2 + 3 * TILE / 4;   // u32 words
  uint32_t* rows = wbase + warp * per_warp;
  uint32_t* seed_node = rows + TILE * a.row_stride;
  uint32_t* seed_info = seed_node + TILE;                                   // off(20) | pos(10) | phase(2)
  uint16_t* lens = reinterpret_cast<uint16_t*>(seed_info + TILE);
  uint8_t* q = reinterpret_cast<uint8_t*>(lens + TILE);                     // 3 queues of TILE slots
  unsigned long long m_reads = 0, m_nonprim = 0, m_notcell = 0, m_cds = 0, m_gen = 0, m_none = 0, m_badcell = 0;
  uint32_t n_tests = 0, n_probes = 0, umi_max = 0;
  const uint64_t n_warps = (uint64_t)gridDim.x * (MAP_THREADS / 32);
  for (uint64_t tile = (uint64_t)blockIdx.x * (MAP_THREADS / 32) + warp; tile * TILE < a.n; tile += n_warps) {
    const uint64_t base = tile * TILE;
    unsigned cnt0 = 0, cnt1 = 0, cntw = 0;   // warp-uniform queue lengths
    __syncwarp();
    // ---- load: all global loads of a read are issued before any is consumed ----
#pragma unroll
    for (int r = 0; r < R; r++) {
      const unsigned slot = r * 32 + lane;
      const uint64_t i = base + slot;
      bool keep = false;
      if (i < a.n) {
        const uint8_t fl = __ldg(a.flags + i);
        const uint32_t cell = __ldg(a.cell + i);
        const uint16_t L = (uint16_t)min((uint32_t)__ldg(a.len + i), 32u * a.wpr);   // never read past the packed words
        uint32_t* fw = rows + slot * a.row_stride;
        for (uint32_t w = 0; w < a.wpr; w++) {
          const uint64_t v = __ldg(a.seq + i * a.wpr + w);
          fw[2 * w] = (uint32_t)(v >> 32); fw[2 * w + 1] = (uint32_t)v;
        }
        fw[2 * a.wpr] = 0; fw[2 * a.wpr + 1] = 0;
        lens[slot] = L;
        m_reads++;
        const bool primary = fl & HLAC_FLAG_PRIMARY;
        if (!primary) m_nonprim++;
        if (primary || !a.primary_only) {
          if (cell == HLAC_NOT_A_CELL) m_notcell++;
          else if (cell >= a.n_cells) m_badcell++;
          else keep = true;
        }
        if (a.codes) a.codes[i] = (uint8_t)HLAC_CODE_NONE;
      }
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
      if (keep) q[cnt0 + __popc(ballot & ((1u << lane) - 1))] = (uint8_t)slot;
      cnt0 += __popc(ballot);
    }
    __syncwarp();
    // ---- scan cascade ----
#pragma unroll 1
    for (int p = 0; p < 4; p++) {
      const unsigned nt = (p & 1) ? cnt1 : cnt0;
      uint8_t* qc = q + ((p & 1) ? TILE : 0);
      uint8_t* qn = q + ((p & 1) ? 0 : TILE);
      unsigned cntn = 0;
      const DevIndexView& ix = p < 2 ? a.cds : a.gen;
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
          SharedWords rd{(p & 1) ? rc : fw};
          int pos = 0; uint32_t node = 0, off = 0;
          if (L >= K) {
            if (SMEM_BM) found = find_seed(ix, rd, BitmapShared{smem + (p < 2 ? 0 : a.cds.bitmap_words)}, pos, L - K, node, off, n_tests, n_probes);
            else found = find_seed(ix, rd, BitmapGlobal{ix.bitmap}, pos, L - K, node, off, n_tests, n_probes);
          }
          if (found) { seed_node[slot] = node; seed_info[slot] = (off << 12) | ((uint32_t)pos << 2) | (uint32_t)p; }
        }
        __syncwarp();
        const unsigned bf = __ballot_sync(0xFFFFFFFFu, task && found), bn = __ballot_sync(0xFFFFFFFFu, task && !found);
        const unsigned below = (1u << lane) - 1;
        if (task && found) q[2 * TILE + cntw + __popc(bf & below)] = (uint8_t)slot;
        if (task && !found) qn[cntn + __popc(bn & below)] = (uint8_t)slot;
        cntw += __popc(bf); cntn += __popc(bn);
      }
      if (p & 1) { cnt0 = cntn; cnt1 = 0; } else { cnt1 = cntn; cnt0 = 0; }
      __syncwarp();
    }
    if (lane == 0) m_none += cnt0;   // all four None (mapping.rs:791-793): survivors of p = 3 sit in queue 0
    // ---- walk ----
    for (unsigned t0 = 0; t0 < cntw; t0 += 32) {
      uint32_t code = CODE_NONE5;
      uint64_t key = 0;
      if (t0 + lane < cntw) {
        const unsigned slot = q[2 * TILE + t0 + lane];
        const uint64_t i = base + slot;
        const uint32_t info = seed_info[slot];
        const int p = (int)(info & 3u), pos = (int)((info >> 2) & 1023u);
        const uint32_t off = info >> 12, node = seed_node[slot];
        uint32_t* fw = rows + slot * a.row_stride;
        const int L = lens[slot];
        const DevIndexView& ix = p < 2 ? a.cds : a.gen;
        SharedWords rd{(p & 1) ? fw + a.nw : fw};
        CountVis vis;
        uint32_t cov;
        if (SMEM_BM) cov = walk_from_seed(ix, rd, BitmapShared{smem + (p < 2 ? 0 : a.cds.bitmap_words)}, L, pos, node, off, vis, n_tests, n_probes);
        else cov = walk_from_seed(ix, rd, BitmapGlobal{ix.bitmap}, L, pos, node, off, vis, n_tests, n_probes);
        const bool good = cov > MIN_SCORE_COUNT_PSEUDO;
        if (good) { if (p < 2) m_cds++; else m_gen++; }
        if (p == 0 || good) code = vis.code();   // a CDS-forward hit is used whatever its coverage (mapping.rs:772-775)
        if (code != CODE_NONE5) {
          const uint32_t um = __ldg(a.umi + i);
          umi_max = max(umi_max, um);
          key = ((uint64_t)um << (CODE_BITS + a.cell_bits)) | ((uint64_t)__ldg(a.cell + i) << CODE_BITS) | code;
          if (a.codes) a.codes[i] = (uint8_t)code;
        }
      }
      __syncwarp();
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, code != CODE_NONE5);
      if (ballot) {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(a.cursor, (unsigned long long)__popc(ballot));
        at = __shfl_sync(0xFFFFFFFFu, at, 0);
        if (code != CODE_NONE5) a.keys[at + __popc(ballot & ((1u << lane) - 1))] = key;
      }
    }
  }
  // metrics: warp reduce then one atomic per counter per warp
  unsigned long long m[9] = {m_reads, m_nonprim, m_notcell, m_cds, m_gen, m_none, n_tests, n_probes, m_badcell};
#pragma unroll
  for (int k = 0; k < 9; k++) {
    unsigned long long v = m[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if (lane == 0 && v) atomicAdd(a.metrics + k, v);
  }
  for (int o = 16; o > 0; o >>= 1) umi_max = max(umi_max, __shfl_down_sync(0xFFFFFFFFu, umi_max, o));
  if (lane == 0 && umi_max) atomicMax(a.metrics + 9, (unsigned long long)umi_max);   // how many UMI bits the sort must cover
}

// Queue-staged variant of the same kernel (shape R = 0). Instead of moving one fixed tile through the stages, each
// warp keeps a pool of NS read slots and two warp-private task queues that persist across refills: SCAN entries are
// (slot, p) with p = the orientation/index the chain of mapping.rs:772-790 has reached, WALK entries are slots with
// a seed. A round of either kind only runs when 32 tasks are waiting (or the warp's input is exhausted), so every
// round starts with all lanes busy whatever fraction of the reads survives each stage; scan rounds mix the four
// orientations (same code, per-lane operands). The warp streams a contiguous range of the batch and refills the
// pool with as many reads as there are free slots.
template <bool SMEM_BM, int MAP_THREADS, int MINB, bool PACKED>
__global__ void __launch_bounds__(MAP_THREADS, MINB) k_count_map_q(const CountMapArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr int NS = 64;
  uint32_t* wbase = smem;
  if (SMEM_BM) {
    const uint32_t nb = a.cds.bitmap_words + a.gen.bitmap_words;
    for (uint32_t i = threadIdx.x; i < a.cds.bitmap_words; i += MAP_THREADS) smem[i] = __ldg(a.cds.bitmap + i);
    for (uint32_t i = threadIdx.x; i < a.gen.bitmap_words; i += MAP_THREADS) smem[a.cds.bitmap_words + i] = __ldg(a.gen.bitmap + i);
    wbase = smem + nb;
    __syncthreads();
  }
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned below = (1u << lane) - 1;
  const uint32_t per_warp = NS * a.row_stride + NS * 3 + NS / 2 + 3 * NS / 4;   // u32 words
  uint32_t* rows = wbase + warp * per_warp;
  uint32_t* seed_node = rows + NS * a.row_stride;
  uint32_t* seed_info = seed_node + NS;                                     // off(20) | pos(10) | p(2)
  uint32_t* ridx = seed_info + NS;                                          // read index within the batch
  uint16_t* lens = reinterpret_cast<uint16_t*>(ridx + NS);
  uint8_t* qs = reinterpret_cast<uint8_t*>(lens + NS);                      // scan queue: slot | p << 6
  uint8_t* qw = qs + NS;                                                    // walk queue
  uint8_t* fr = qw + NS;                                                    // free-slot stack
  fr[lane] = (uint8_t)lane; fr[lane + 32] = (uint8_t)(lane + 32);
  unsigned ns = 0, nwk = 0, nfree = NS;                                     // warp-uniform
  unsigned long long m_reads = 0, m_nonprim = 0, m_notcell = 0, m_cds = 0, m_gen = 0, m_none = 0, m_badcell = 0;
  uint32_t n_tests = 0, n_probes = 0, umi_max = 0;
  const uint64_t n_warps = (uint64_t)gridDim.x * (MAP_THREADS / 32);
  const uint64_t per = (a.n + n_warps - 1) / n_warps;
  uint64_t cur = min(a.n, ((uint64_t)blockIdx.x * (MAP_THREADS / 32) + warp) * per);
  const uint64_t hi = min(a.n, cur + per);
  __syncwarp();
  for (;;) {
    const bool more = cur < hi;
    int what;   // 0 walk, 1 scan, 2 refill
    if (nwk >= 32) what = 0;
    else if (ns >= 32) what = 1;
    else if (more) what = 2;
    else if (ns > 0) what = 1;
    else if (nwk > 0) what = 0;
    else break;
    if (what == 2) {
      // ---- refill: filters (mapping.rs:752-763) + read rows into free slots ----
      const unsigned c = (unsigned)min((uint64_t)min(32u, nfree), hi - cur);
      bool keep = false;
      unsigned slot = 0;
      if (lane < c) {
        const uint64_t i = cur + lane;
        slot = fr[nfree - c + lane];
        uint8_t fl; uint32_t cell; uint16_t L;
        const uint64_t* words;
        if (PACKED) {   // one record per read: sequence words, then umi | primary | len | cell
          words = a.rec + i * (a.wpr + 1);
          const uint32_t meta = (uint32_t)__ldg(words + a.wpr);
          fl = (meta >> 31) ? HLAC_FLAG_PRIMARY : 0;
          cell = meta & HLAC_PACKED_NOT_A_CELL; if (cell == HLAC_PACKED_NOT_A_CELL) cell = HLAC_NOT_A_CELL;
          L = (uint16_t)min((meta >> 23) & 0xFFu, 32u * a.wpr);
        } else {
          words = a.seq + i * a.wpr;
          fl = __ldg(a.flags + i);
          cell = __ldg(a.cell + i);
          L = (uint16_t)min((uint32_t)__ldg(a.len + i), 32u * a.wpr);
        }
        uint32_t* fw = rows + slot * a.row_stride;
        for (uint32_t w = 0; w < a.wpr; w++) {
          const uint64_t v = __ldg(words + w);
          fw[2 * w] = (uint32_t)(v >> 32); fw[2 * w + 1] = (uint32_t)v;
        }
        fw[2 * a.wpr] = 0; fw[2 * a.wpr + 1] = 0;
        lens[slot] = L; ridx[slot] = (uint32_t)i;
        make_revcomp(SharedWords{fw}, fw + a.nw, L, (int)a.nw);   // here every lane is busy; in a mixed scan round only the p = 1 lanes would be
        m_reads++;
        const bool primary = fl & HLAC_FLAG_PRIMARY;
        if (!primary) m_nonprim++;
        if (primary || !a.primary_only) {
          if (cell == HLAC_NOT_A_CELL) m_notcell++;
          else if (cell >= a.n_cells) m_badcell++;
          else keep = true;
        }
        if (a.codes) a.codes[i] = (uint8_t)HLAC_CODE_NONE;
      }
      __syncwarp();
      nfree -= c; cur += c;
      const unsigned bk = __ballot_sync(0xFFFFFFFFu, keep), bd = __ballot_sync(0xFFFFFFFFu, lane < c && !keep);
      if (keep) qs[ns + __popc(bk & below)] = (uint8_t)slot;                 // p = 0
      if (lane < c && !keep) fr[nfree + __popc(bd & below)] = (uint8_t)slot;
      ns += __popc(bk); nfree += __popc(bd);
      __syncwarp();
    } else if (what == 1) {
      // ---- scan round: first-seed skip-scan of (slot, p) ----
      const unsigned cnt = min(32u, ns), qb = ns - cnt;
      const bool task = lane < cnt;
      bool found = false;
      unsigned slot = 0, p = 0;
      if (task) {
        const unsigned e = qs[qb + lane];
        slot = e & 63u; p = e >> 6;
        uint32_t* fw = rows + slot * a.row_stride;
        const int L = lens[slot];
        SharedWords rd{(p & 1) ? fw + a.nw : fw};
        const DevIndexView& ix = p < 2 ? a.cds : a.gen;
        int pos = 0; uint32_t node = 0, off = 0;
        if (L >= K) {
          if (SMEM_BM) found = find_seed(ix, rd, BitmapShared{smem + (p < 2 ? 0 : a.cds.bitmap_words)}, pos, L - K, node, off, n_tests, n_probes);
          else found = find_seed(ix, rd, BitmapGlobal{ix.bitmap}, pos, L - K, node, off, n_tests, n_probes);
        }
        if (found) { seed_node[slot] = node; seed_info[slot] = (off << 12) | ((uint32_t)pos << 2) | p; }
      }
      __syncwarp();
      ns = qb;
      const bool again = task && !found && p < 3, dead = task && !found && p == 3;
      const unsigned bf = __ballot_sync(0xFFFFFFFFu, task && found), bn = __ballot_sync(0xFFFFFFFFu, again),
                     bz = __ballot_sync(0xFFFFFFFFu, dead);
      if (task && found) qw[nwk + __popc(bf & below)] = (uint8_t)slot;
      if (again) qs[ns + __popc(bn & below)] = (uint8_t)(slot | ((p + 1) << 6));
      if (dead) fr[nfree + __popc(bz & below)] = (uint8_t)slot;             // all four None (mapping.rs:791-793)
      nwk += __popc(bf); ns += __popc(bn); nfree += __popc(bz);
      if (lane == 0) m_none += __popc(bz);
      __syncwarp();
    } else {
      // ---- walk round: left extension + forward walk + class code + coverage policy, key append ----
      const unsigned cnt = min(32u, nwk), qb = nwk - cnt;
      const bool task = lane < cnt;
      uint32_t code = CODE_NONE5;
      uint64_t key = 0;
      unsigned slot = 0;
      if (task) {
        slot = qw[qb + lane];
        const uint64_t i = ridx[slot];
        const uint32_t info = seed_info[slot];
        const int p = (int)(info & 3u), pos = (int)((info >> 2) & 1023u);
        const uint32_t off = info >> 12, node = seed_node[slot];
        uint32_t* fw = rows + slot * a.row_stride;
        const int L = lens[slot];
        const DevIndexView& ix = p < 2 ? a.cds : a.gen;
        SharedWords rd{(p & 1) ? fw + a.nw : fw};
        CountVis vis;
        uint32_t cov;
        if (SMEM_BM) cov = walk_from_seed(ix, rd, BitmapShared{smem + (p < 2 ? 0 : a.cds.bitmap_words)}, L, pos, node, off, vis, n_tests, n_probes);
        else cov = walk_from_seed(ix, rd, BitmapGlobal{ix.bitmap}, L, pos, node, off, vis, n_tests, n_probes);
        const bool good = cov > MIN_SCORE_COUNT_PSEUDO;
        if (good) { if (p < 2) m_cds++; else m_gen++; }
        if (p == 0 || good) code = vis.code();   // a CDS-forward hit is used whatever its coverage (mapping.rs:772-775)
        if (code != CODE_NONE5) {
          uint32_t um, cl;
          if (PACKED) { const uint64_t meta = __ldg(a.rec + i * (a.wpr + 1) + a.wpr); um = (uint32_t)(meta >> 32); cl = (uint32_t)meta & HLAC_PACKED_NOT_A_CELL; }
          else { um = __ldg(a.umi + i); cl = __ldg(a.cell + i); }
          umi_max = max(umi_max, um);
          key = ((uint64_t)um << (CODE_BITS + a.cell_bits)) | ((uint64_t)cl << CODE_BITS) | code;
          if (a.codes) a.codes[i] = (uint8_t)code;
        }
      }
      __syncwarp();
      nwk = qb;
      if (task) fr[nfree + lane] = (uint8_t)slot;
      nfree += cnt;
      const unsigned ballot = __ballot_sync(0xFFFFFFFFu, code != CODE_NONE5);
      if (ballot) {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(a.cursor, (unsigned long long)__popc(ballot));
        at = __shfl_sync(0xFFFFFFFFu, at, 0);
        if (code != CODE_NONE5) a.keys[at + __popc(ballot & below)] = key;
      }
      __syncwarp();
    }
  }
  unsigned long long m[9] = {m_reads, m_nonprim, m_notcell, m_cds, m_gen, m_none, n_tests, n_probes, m_badcell};
#pragma unroll
  for (int k = 0; k < 9; k++) {
    unsigned long long v = m[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if (lane == 0 && v) atomicAdd(a.metrics + k, v);
  }
  for (int o = 16; o > 0; o >>= 1) umi_max = max(umi_max, __shfl_down_sync(0xFFFFFFFFu, umi_max, o));
  if (lane == 0 && umi_max) atomicMax(a.metrics + 9, (unsigned long long)umi_max);
}

// count() (mapping.rs:842-896) for one (cell, UMI) run of the sorted keys.
__global__ void k_consensus(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ dense, uint32_t nrows,
                            const uint32_t* __restrict__ gene_row0, uint32_t cell_bits) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k0 = keys[i], grp = k0 >> CODE_BITS;
  if (i > 0 && (keys[i - 1] >> CODE_BITS) == grp) return;   // not a run head
  // pass 1: majority-vote candidate gene (only a gene with > 50 % of the UMI's reads can win, mapping.rs:874-875)
  uint32_t cand = 0, votes = 0, nreads = 0;
  for (uint64_t j = i; j < n; j++) {
    const uint64_t k = keys[j];
    if ((k >> CODE_BITS) != grp) break;
    const uint32_t g = (uint32_t)(k & 31u) / 3;
    if (votes == 0) { cand = g; votes = 1; }
    else if (g == cand) votes++;
    else votes--;
    nreads++;
  }
  // pass 2: the candidate's (allele 1, allele 2, gene) counts
  uint32_t f[3] = {0, 0, 0};
  for (uint64_t j = i; j < i + nreads; j++) {
    const uint32_t c = (uint32_t)(keys[j] & 31u);
    if (c / 3 == cand) f[c % 3]++;
  }
  const double nr = (double)nreads;
  const uint32_t ng = f[0] + f[1] + f[2];
  if (!((double)ng > GENE_CONSENSUS_THRESHOLD * nr)) return;
  uint32_t slot;
  if ((double)f[0] > ALLELE_CONSENSUS_THRESHOLD * nr && ALLELE_CONSENSUS_THRESHOLD * nr > (double)f[1]) slot = 0;
  else if ((double)f[1] > ALLELE_CONSENSUS_THRESHOLD * nr && ALLELE_CONSENSUS_THRESHOLD * nr > (double)f[0]) slot = 1;
  else slot = 2;
  const uint32_t cell = (uint32_t)(k0 >> CODE_BITS) & ((1u << cell_bits) - 1u);
  atomicAdd(dense + (size_t)cell * nrows + gene_row0[cand] + slot, 1u);
}

// ---- dense matrix -> COO on the device (cells ascending, rows ascending = linear order of the dense array) ----
constexpr int COO_THREADS = 256, COO_PER_THREAD = 8, COO_TILE = COO_THREADS * COO_PER_THREAD;
__global__ void k_coo_count(const uint32_t* __restrict__ dense, uint64_t n, uint32_t* block_nnz) {
  __shared__ uint32_t warp_cnt[COO_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * COO_TILE + (uint64_t)threadIdx.x * COO_PER_THREAD;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < COO_PER_THREAD; k++) if (base + k < n && dense[base + k]) c++;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < COO_THREADS / 32; w++) t += warp_cnt[w]; block_nnz[blockIdx.x] = t; }
}
// single block: exclusive scan of the per-block counts; total to block_off[nblocks]
__global__ void k_coo_scan(const uint32_t* __restrict__ block_nnz, uint32_t nblocks, uint64_t* block_off) {
  __shared__ uint64_t carry;
  __shared__ uint64_t part[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024) {
    const uint32_t i = b0 + threadIdx.x;
    const uint64_t v = i < nblocks ? block_nnz[i] : 0;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan
      const uint64_t add = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < nblocks) block_off[i] = carry + part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_off[nblocks] = carry;
}
__global__ void k_coo_write(const uint32_t* __restrict__ dense, uint64_t n, uint32_t nrows, const uint64_t* __restrict__ block_off,
                            uint32_t* row, uint32_t* col, uint32_t* val) {
  __shared__ uint32_t warp_off[COO_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * COO_TILE + (uint64_t)threadIdx.x * COO_PER_THREAD;
  uint32_t v[COO_PER_THREAD]; uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < COO_PER_THREAD; k++) { v[k] = base + k < n ? dense[base + k] : 0; c += v[k] != 0; }
  uint32_t incl = c;   // inclusive scan of the per-thread counts inside the warp
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((threadIdx.x & 31) >= (unsigned)o) incl += t; }
  if ((threadIdx.x & 31) == 31) warp_off[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (unsigned w = 0; w < (threadIdx.x >> 5); w++) woff += warp_off[w];
  uint64_t at = block_off[blockIdx.x] + woff + incl - c;
#pragma unroll
  for (int k = 0; k < COO_PER_THREAD; k++) if (v[k]) {
    const uint64_t e = base + k;
    row[at] = (uint32_t)(e % nrows); col[at] = (uint32_t)(e / nrows); val[at] = v[k]; at++;
  }
}

// per-colour info word for the counting kernel: bits 0-4 code (gene*3+slot or 31), bits 8-15 per-gene swap mask
std::vector<uint32_t> make_cinfo(const HostIndex& ix, const RowLayout& rows) {
  std::vector<uint32_t> out(ix.n_colours());
  for (uint32_t c = 0; c < ix.n_colours(); c++) {
    const uint32_t* tx = &ix.col_tx[ix.col_off[c]];
    const size_t n = ix.col_off[c + 1] - ix.col_off[c];
    auto has = [&](uint32_t t) { return std::binary_search(tx, tx + n, t); };
    uint32_t code = CODE_NONE5, swap = 0;
    for (uint32_t g = 0; g < rows.n_genes; g++) {
      const uint32_t a1 = rows.a1[g], a2 = rows.a2[g];
      if (n == 1 && tx[0] == a1) code = g * 3 + 0;
      if (a2 != NONE32) {
        if (n == 1 && tx[0] == a2) code = g * 3 + 1;
        const uint32_t lo = std::min(a1, a2), hi = std::max(a1, a2);
        if (n == 2 && tx[0] == lo && tx[1] == hi) code = g * 3 + 2;
        if (has(hi) && !has(lo)) swap |= 1u << g;
      }
    }
    out[c] = code | (swap << 8);
  }
  return out;
}

}  // namespace

struct hlac_count {
  hlac_index* cds = nullptr;
  hlac_index* gen = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  RowLayout rows;
  uint32_t n_cells = 0;
  int primary_only = 0;
  DevBuf<uint32_t> nodes_cds, nodes_gen, gene_row0, dense;   // node tables with cinfo in place of the colour id
  DevBuf<uint64_t> keys, keys_sorted;
  uint64_t host_reads = 0, host_non_primary = 0, host_not_cell = 0;   // reads the host driver filtered out itself (hlac_count_add_filtered)
  struct Stage {   // one staging set for host batches; two of them so the H2D of batch i+1 overlaps the kernel of batch i
    DevBuf<uint64_t> seq, rec; DevBuf<uint16_t> len; DevBuf<uint32_t> cell, umi; DevBuf<uint8_t> flags;
    cudaEvent_t copied = nullptr, consumed = nullptr;
  } stage[2];
  cudaStream_t copy_stream = nullptr;
  uint64_t n_push = 0;
  DevBuf<uint8_t> codes;
  DevBuf<uint32_t> coo_block_nnz, coo_dev; DevBuf<uint64_t> coo_block_off;   // device-side COO extraction
  uint32_t* coo_host = nullptr; size_t coo_host_cap = 0;                      // pinned, 3 arrays back to back, session-owned
  DevBuf<unsigned long long> counters;   // [0] cursor, [1..8] metrics
  DevBuf<uint8_t> cub_tmp;
  uint64_t reads_pushed = 0, last_n = 0;
  bool smem_bm = false;
  size_t bitmap_smem_bytes = 0;
  int blocks_per_sm = 1, n_sm = 1;
  bool want_codes = false;
  bool reduced = false;
  uint32_t cell_bits = 1;
  uint32_t cfg_wpr = 0;
  int cfg_occ = 0, cfg_r = 0, cfg_threads = 0;
  bool cfg_bm = false;
  void* cfg_kern = nullptr; void* cfg_kern_packed = nullptr;
  uint64_t n_keys = 0;
  // timing
  struct Ev { cudaEvent_t a, b; int kind; };
  std::vector<Ev> events;
  float ms[4] = {0, 0, 0, 0};
  uint64_t launches = 0;

  ~hlac_count() {
    for (auto& e : events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (coo_host) cudaFreeHost(coo_host);
    for (auto& st : stage) { if (st.copied) cudaEventDestroy(st.copied); if (st.consumed) cudaEventDestroy(st.consumed); }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (stream && own_stream) cudaStreamDestroy(stream);
  }
  Ev& begin(int kind, cudaStream_t on = nullptr) {
    if (events.size() >= 512) drain_events();   // long streaming runs: keep the number of live events bounded
    Ev e; e.kind = kind;
    HLAC_CUDA(cudaEventCreate(&e.a)); HLAC_CUDA(cudaEventCreate(&e.b));
    HLAC_CUDA(cudaEventRecord(e.a, on ? on : stream));
    events.push_back(e);
    return events.back();
  }
  void end(Ev& e, cudaStream_t on = nullptr) { HLAC_CUDA(cudaEventRecord(e.b, on ? on : stream)); }
  void drain_events() {
    for (auto& e : events) {
      float t = 0; HLAC_CUDA(cudaEventSynchronize(e.b)); HLAC_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
      ms[e.kind] += t; cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    events.clear();
  }
  void reset() {
    HLAC_CUDA(cudaMemsetAsync(counters.p, 0, 11 * sizeof(unsigned long long), stream));
    reads_pushed = 0; last_n = 0; reduced = false; n_keys = 0; host_reads = host_non_primary = host_not_cell = 0;
    drain_events(); ms[0] = ms[1] = ms[2] = ms[3] = 0; launches = 0;
  }

  void launch_map(const hlac_batch& b /* device pointers */, const uint64_t* rec = nullptr /* packed records instead of b's arrays */) {
    if (b.n == 0) return;
    if (b.words_per_read == 0 || b.words_per_read > 16) throw Error(HLAC_ERR_ARG, "words_per_read must be 1..16");
    keys.reserve(reads_pushed + b.n, stream, true, reads_pushed);
    if (want_codes) codes.reserve(b.n, stream);
    CountMapArgs a{};
    a.cds = cds->view; a.gen = gen->view;
    a.cds.nodes = reinterpret_cast<const uint4*>(nodes_cds.p); a.gen.nodes = reinterpret_cast<const uint4*>(nodes_gen.p);
    a.seq = b.seq; a.len = b.len; a.cell = b.cell; a.umi = b.umi; a.flags = b.flags;
    a.n = b.n; a.wpr = b.words_per_read; a.nw = 2 * b.words_per_read + 2; a.row_stride = (2 * a.nw) | 1;
    a.primary_only = primary_only; a.n_cells = n_cells; a.cell_bits = cell_bits;
    a.keys = keys.p; a.cursor = counters.p; a.metrics = counters.p + 1;
    a.codes = want_codes ? codes.p : nullptr;
    a.rec = rec;
    // launch shape: (bitmaps in shared memory?, reads per lane per warp tile R, threads per block). Defaults below;
    // HLAC_COUNT_SHAPE="smem,R,threads" overrides them for tuning runs.
    using Kern = void (*)(const CountMapArgs);
    auto smem_for = [&](bool bm, int r, int threads) {
      const size_t per_warp_words = r == 0 ? (size_t)64 * a.row_stride + 64 * 3 + 64 / 2 + 3 * 64 / 4   // queue-staged kernel
                                           : (size_t)32 * r * a.row_stride + 32 * r * 2 + 32 * r / 2 + 3 * 32 * r / 4;
      return (bm ? bitmap_smem_bytes : 0) + (threads / 32) * per_warp_words * 4;
    };
    if (cfg_wpr != b.words_per_read) {
      struct Shape { bool bm; int r, threads; Kern k; int minb = 1; Kern k_packed = nullptr; };
      static const Shape shapes[] = {
#define HLAC_SHAPE(BM, RR, TT) {BM, RR, TT, k_count_map<BM, RR, TT>}
          // defaults first: measured best on B200 (profiles/r01_count_map_tuning.md) — the queue-staged kernel (R = 0) at
          // 1024 resident threads per SM, then the tile-staged shapes (occupancy beats tile size)
#define HLAC_QSHAPE(BM, TT, MB) {BM, 0, TT, k_count_map_q<BM, TT, MB, false>, 1, k_count_map_q<BM, TT, MB, true>}
          HLAC_QSHAPE(true, 1024, 1), HLAC_QSHAPE(true, 512, 2), HLAC_QSHAPE(false, 256, 4), HLAC_QSHAPE(false, 128, 8),
#undef HLAC_QSHAPE
          HLAC_SHAPE(true, 2, 1024), HLAC_SHAPE(true, 1, 512), HLAC_SHAPE(true, 1, 256), HLAC_SHAPE(true, 4, 512), HLAC_SHAPE(true, 2, 512),
          HLAC_SHAPE(false, 2, 128), HLAC_SHAPE(false, 1, 128), HLAC_SHAPE(false, 2, 256), HLAC_SHAPE(false, 1, 256), HLAC_SHAPE(false, 4, 128),
          HLAC_SHAPE(true, 4, 256), HLAC_SHAPE(true, 8, 256), HLAC_SHAPE(false, 4, 256), HLAC_SHAPE(false, 8, 128),
          {false, 1, 128, k_count_map<false, 1, 128, 12>, 12}, {false, 1, 128, k_count_map<false, 1, 128, 16>, 16},
          {true, 1, 512, k_count_map<true, 1, 512, 3>, 3}};
#undef HLAC_SHAPE
      const Shape* pick = nullptr;
      int want_bm = -1, want_r = 0, want_t = 0, want_minb = 1;
      if (const char* e = getenv("HLAC_COUNT_SHAPE")) sscanf(e, "%d,%d,%d,%d", &want_bm, &want_r, &want_t, &want_minb);
      for (const Shape& sh : shapes) {
        if (want_bm >= 0) { if ((int)sh.bm != want_bm || sh.r != want_r || sh.threads != want_t || sh.minb != want_minb) continue; }
        else if (sh.minb != 1) continue;
        else if (sh.bm != smem_bm) continue;
        if (sh.bm && !smem_bm) continue;
        if (smem_for(sh.bm, sh.r, sh.threads) > 227 * 1024) continue;
        pick = &sh; break;
      }
      if (!pick) throw Error(HLAC_ERR_LIMIT, "no launch shape of the count kernel fits (reads too long, or bad HLAC_COUNT_SHAPE)");
      cfg_bm = pick->bm; cfg_r = pick->r; cfg_threads = pick->threads; cfg_kern = (void*)pick->k; cfg_kern_packed = (void*)pick->k_packed;
      if (pick->k_packed) HLAC_CUDA(cudaFuncSetAttribute(pick->k_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(cfg_bm, cfg_r, cfg_threads)));
      HLAC_CUDA(cudaFuncSetAttribute(pick->k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(cfg_bm, cfg_r, cfg_threads)));
      HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg_occ, pick->k, cfg_threads, smem_for(cfg_bm, cfg_r, cfg_threads)));
      if (cfg_occ < 1) throw Error(HLAC_ERR_LIMIT, "count map kernel does not fit on the device");
      cfg_wpr = b.words_per_read;
    }
    if (rec && !cfg_kern_packed) throw Error(HLAC_ERR_ARG, "packed records need the queue-staged count kernel (HLAC_COUNT_SHAPE selects a tile-staged shape)");
    Kern kern = (Kern)(rec ? cfg_kern_packed : cfg_kern);
    const size_t smem = smem_for(cfg_bm, cfg_r, cfg_threads);
    const int MAP_R = cfg_r ? cfg_r : 2, MAP_THREADS = cfg_threads;
    if (cfg_r == 0 && b.n >= (1ull << 32)) throw Error(HLAC_ERR_LIMIT, "the queue-staged count kernel takes batches below 2^32 reads");
    const int occ = cfg_occ;
    const uint64_t need = (b.n + (uint64_t)MAP_THREADS * MAP_R - 1) / ((uint64_t)MAP_THREADS * MAP_R);
    const unsigned grid = (unsigned)std::min<uint64_t>(need, (uint64_t)n_sm * occ);
    Ev& e = begin(0);
    kern<<<grid, MAP_THREADS, smem, stream>>>(a);
    HLAC_CUDA(cudaGetLastError());
    end(e);
    launches++;
    reads_pushed += b.n; last_n = b.n; reduced = false;
  }
};

extern "C" {

int hlac_count_new(hlac_index* cds, hlac_index* genomic, uint32_t n_cells, int primary_only, hlac_count** out) try {
  if (!cds || !genomic || !out) throw Error(HLAC_ERR_ARG, "null argument");
  if (cds->device < 0 || genomic->device < 0) throw Error(HLAC_ERR_CUDA, "index is host-only (device -1): counting needs a CUDA device, there is no CPU fallback");
  if (cds->device != genomic->device) throw Error(HLAC_ERR_ARG, "CDS and genomic index live on different devices");
  if (n_cells == 0 || n_cells >= (1u << 27)) throw Error(HLAC_ERR_LIMIT, "n_cells must be in 1..2^27-1");
  auto s = std::make_unique<hlac_count>();
  s->cds = cds; s->gen = genomic; s->device = cds->device; s->n_cells = n_cells; s->primary_only = primary_only;
  while ((1ull << s->cell_bits) < n_cells) s->cell_bits++;
  DeviceGuard g(s->device);
  s->rows = build_row_layout(cds->host.tx_names);
  if (s->rows.n_genes > 8) throw Error(HLAC_ERR_LIMIT, "more than 8 genes in the CDS FASTA");
  HLAC_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  auto node_table = [&](const HostIndex& h, DevBuf<uint32_t>& dst) {
    const auto ci = make_cinfo(h, s->rows);
    std::vector<uint32_t> rec((size_t)h.n_nodes() * 12);
    for (uint32_t i = 0; i < h.n_nodes(); i++) {
      uint32_t* r = &rec[(size_t)i * 12];
      r[0] = h.start[i]; r[1] = h.len[i]; r[2] = h.exts[i]; r[3] = ci[h.colour[i]];
      for (int b = 0; b < 4; b++) { r[4 + b] = h.radj[(size_t)i * 4 + b]; r[8 + b] = h.ladj[(size_t)i * 4 + b]; }
    }
    dst.upload(rec.data(), rec.size(), s->stream);
  };
  node_table(cds->host, s->nodes_cds); node_table(genomic->host, s->nodes_gen);
  std::vector<uint32_t> r0 = s->rows.gene_row0; r0.resize(8, 0);
  s->gene_row0.upload(r0.data(), r0.size(), s->stream);
  s->counters.reserve(16, s->stream);
  HLAC_CUDA(cudaMemsetAsync(s->counters.p, 0, 16 * sizeof(unsigned long long), s->stream));
  s->dense.reserve((size_t)n_cells * std::max<uint32_t>(s->rows.nrows, 1), s->stream);
  cudaDeviceProp prop; HLAC_CUDA(cudaGetDeviceProperties(&prop, s->device));
  s->n_sm = prop.multiProcessorCount;
  s->bitmap_smem_bytes = ((size_t)cds->view.bitmap_words + genomic->view.bitmap_words) * 4;
  s->smem_bm = s->bitmap_smem_bytes <= 96 * 1024;
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  *out = s.release();
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_free(hlac_count* s) {
  if (s) { cudaSetDevice(s->device); delete s; }
  return HLAC_OK;
}
uint32_t hlac_count_nrows(const hlac_count* s) { return s ? s->rows.nrows : 0; }
const char* hlac_count_row_name(const hlac_count* s, uint32_t i) { return (s && i < s->rows.rownames.size()) ? s->rows.rownames[i].c_str() : nullptr; }

int hlac_count_reserve(hlac_count* s, uint64_t total_reads) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  s->keys.reserve(total_reads, s->stream, true, s->reads_pushed);
  s->keys_sorted.reserve(total_reads, s->stream);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

static void validate_batch(const hlac_batch* b) {
  if (b->n && (!b->seq || !b->len || !b->cell || !b->umi || !b->flags)) throw Error(HLAC_ERR_ARG, "batch has a null array");
  if (b->n >= (1ull << 40)) throw Error(HLAC_ERR_LIMIT, "batch too large");
}

int hlac_count_push_device(hlac_count* s, const hlac_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_batch(b);
  DeviceGuard g(s->device);
  s->launch_map(*b);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_push(hlac_count* s, const hlac_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_batch(b);
  if (b->n == 0) return HLAC_OK;
  DeviceGuard g(s->device);
  auto& st = s->stage[s->n_push++ & 1];
  if (!s->copy_stream) HLAC_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  if (!st.copied) { HLAC_CUDA(cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming)); HLAC_CUDA(cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming)); }
  else HLAC_CUDA(cudaStreamWaitEvent(s->copy_stream, st.consumed, 0));   // the kernel that read this set two pushes ago
  auto& e = s->begin(3, s->copy_stream);
  st.seq.upload(b->seq, b->n * b->words_per_read, s->copy_stream);
  st.len.upload(b->len, b->n, s->copy_stream);
  st.cell.upload(b->cell, b->n, s->copy_stream);
  st.umi.upload(b->umi, b->n, s->copy_stream);
  st.flags.upload(b->flags, b->n, s->copy_stream);
  s->end(e, s->copy_stream);
  HLAC_CUDA(cudaEventRecord(st.copied, s->copy_stream));
  HLAC_CUDA(cudaStreamWaitEvent(s->stream, st.copied, 0));
  hlac_batch d = *b;
  d.seq = st.seq.p; d.len = st.len.p; d.cell = st.cell.p; d.umi = st.umi.p; d.flags = st.flags.p;
  s->launch_map(d);
  HLAC_CUDA(cudaEventRecord(st.consumed, s->stream));
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

static void validate_packed(const hlac_count* s, const hlac_packed_batch* b) {
  if (b->n && !b->rec) throw Error(HLAC_ERR_ARG, "null record array");
  if (b->n && (b->words_per_read == 0 || b->words_per_read > 7)) throw Error(HLAC_ERR_ARG, "packed records hold reads of <= 224 bases (words_per_read 1..7; len is an 8-bit field)");
  if (b->n >= (1ull << 32)) throw Error(HLAC_ERR_LIMIT, "batch too large");
  if (s->n_cells >= HLAC_PACKED_NOT_A_CELL) throw Error(HLAC_ERR_LIMIT, "packed records carry 23-bit cell indices: n_cells must be below 8388607");
}

int hlac_count_push_packed_device(hlac_count* s, const hlac_packed_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_packed(s, b);
  DeviceGuard g(s->device);
  hlac_batch d{}; d.n = b->n; d.words_per_read = b->words_per_read;
  s->launch_map(d, b->rec);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_push_packed(hlac_count* s, const hlac_packed_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_packed(s, b);
  if (b->n == 0) return HLAC_OK;
  DeviceGuard g(s->device);
  auto& st = s->stage[s->n_push++ & 1];
  if (!s->copy_stream) HLAC_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  if (!st.copied) { HLAC_CUDA(cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming)); HLAC_CUDA(cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming)); }
  else HLAC_CUDA(cudaStreamWaitEvent(s->copy_stream, st.consumed, 0));   // the kernel that read this set two pushes ago
  auto& e = s->begin(3, s->copy_stream);
  st.rec.upload(b->rec, b->n * (b->words_per_read + 1), s->copy_stream);   // ONE copy per push
  s->end(e, s->copy_stream);
  HLAC_CUDA(cudaEventRecord(st.copied, s->copy_stream));
  HLAC_CUDA(cudaStreamWaitEvent(s->stream, st.copied, 0));
  hlac_batch d{}; d.n = b->n; d.words_per_read = b->words_per_read;
  s->launch_map(d, st.rec.p);
  HLAC_CUDA(cudaEventRecord(st.consumed, s->stream));
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_pack_records(const hlac_batch* b, uint64_t* rec) try {
  if (!b || (b->n && (!rec || !b->seq || !b->len || !b->cell || !b->umi || !b->flags))) throw Error(HLAC_ERR_ARG, "null argument");
  const uint32_t wpr = b->words_per_read;
  if (b->n && (wpr == 0 || wpr > 7)) throw Error(HLAC_ERR_ARG, "packed records hold reads of <= 224 bases (words_per_read 1..7)");
  for (uint64_t i = 0; i < b->n; i++) {
    if (b->len[i] > 255) throw Error(HLAC_ERR_LIMIT, "read longer than 255 bases does not fit a packed record");
    uint32_t cell = b->cell[i];
    if (cell == HLAC_NOT_A_CELL) cell = HLAC_PACKED_NOT_A_CELL;
    else if (cell >= HLAC_PACKED_NOT_A_CELL) throw Error(HLAC_ERR_LIMIT, "cell index does not fit the 23-bit field of a packed record");
    uint64_t* r = rec + i * (wpr + 1);
    for (uint32_t w = 0; w < wpr; w++) r[w] = b->seq[i * wpr + w];
    r[wpr] = HLAC_PACKED_META(b->umi[i], b->flags[i] & HLAC_FLAG_PRIMARY, b->len[i], cell);
  }
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_sync(hlac_count* s) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  if (s->copy_stream) HLAC_CUDA(cudaStreamSynchronize(s->copy_stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_set_stream(hlac_count* s, void* cuda_stream) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  s->drain_events();
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  s->stream = (cudaStream_t)cuda_stream;
  s->own_stream = false;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_record_codes(hlac_count* s, int on) {
  if (!s) { set_last_error("null session"); return HLAC_ERR_ARG; }
  s->want_codes = on != 0;
  return HLAC_OK;
}

int hlac_count_last_codes(hlac_count* s, uint8_t* codes, uint64_t n) try {
  if (!s || !codes) throw Error(HLAC_ERR_ARG, "null argument");
  DeviceGuard g(s->device);
  if (!s->want_codes) throw Error(HLAC_ERR_STATE, "per-read codes are not being recorded (hlac_count_record_codes)");
  if (n != s->last_n) throw Error(HLAC_ERR_ARG, "n does not match the last pushed batch");
  HLAC_CUDA(cudaMemcpyAsync(codes, s->codes.p, n, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_reduce(hlac_count* s, uint32_t** dense_device, uint64_t* n_elems) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  const size_t nd = (size_t)s->n_cells * std::max<uint32_t>(s->rows.nrows, 1);
  unsigned long long nk = 0, umi_max = 0;
  HLAC_CUDA(cudaMemcpyAsync(&nk, s->counters.p, sizeof nk, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaMemcpyAsync(&umi_max, s->counters.p + 10, sizeof umi_max, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaMemsetAsync(s->dense.p, 0, nd * sizeof(uint32_t), s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  s->n_keys = nk;
  if (nk > 0) {
    s->keys_sorted.reserve(nk, s->stream);
    // keys are umi << (5 + cell_bits) | cell << 5 | code: only the UMI bits actually seen need sorting (10-bp UMIs: 21)
    int umi_bits = 1;
    while (umi_bits < UMI_BITS && (umi_max >> umi_bits)) umi_bits++;
    const int end_bit = CODE_BITS + (int)s->cell_bits + umi_bits;
    size_t tmp = 0;
    HLAC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, s->keys.p, s->keys_sorted.p, (uint64_t)nk, CODE_BITS, end_bit, s->stream));
    s->cub_tmp.reserve(tmp, s->stream);
    auto& e1 = s->begin(1);
    HLAC_CUDA(cub::DeviceRadixSort::SortKeys(s->cub_tmp.p, tmp, s->keys.p, s->keys_sorted.p, (uint64_t)nk, CODE_BITS, end_bit, s->stream));
    s->end(e1);
    auto& e2 = s->begin(2);
    k_consensus<<<(unsigned)((nk + 255) / 256), 256, 0, s->stream>>>(s->keys_sorted.p, nk, s->dense.p, s->rows.nrows, s->gene_row0.p, s->cell_bits);
    HLAC_CUDA(cudaGetLastError());
    s->end(e2);
    s->launches += 1;   // k_consensus; the cub sort passes are library kernels and not counted
  }
  s->reduced = true;
  if (dense_device) *dense_device = s->dense.p;
  if (n_elems) *n_elems = nd;
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_entries(hlac_count* s, hlac_coo* out, hlac_metrics* metrics) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  if (!s->reduced) throw Error(HLAC_ERR_STATE, "hlac_count_reduce has not run since the last push");
  DeviceGuard g(s->device);
  const uint32_t nrows = s->rows.nrows;
  const uint64_t nd = (uint64_t)s->n_cells * std::max<uint32_t>(nrows, 1);
  unsigned long long ctr[10];
  HLAC_CUDA(cudaMemcpyAsync(ctr, s->counters.p, sizeof ctr, cudaMemcpyDeviceToHost, s->stream));
  uint64_t nnz = 0;
  if (out && nrows) {
    // dense -> COO on the device: per-block non-zero counts, one-block scan, ordered write; only the non-zeros cross PCIe
    const uint32_t nb = (uint32_t)((nd + COO_TILE - 1) / COO_TILE);
    s->coo_block_nnz.reserve(nb, s->stream); s->coo_block_off.reserve(nb + 1, s->stream);
    k_coo_count<<<nb, COO_THREADS, 0, s->stream>>>(s->dense.p, nd, s->coo_block_nnz.p);
    k_coo_scan<<<1, 1024, 0, s->stream>>>(s->coo_block_nnz.p, nb, s->coo_block_off.p);
    HLAC_CUDA(cudaGetLastError());
    HLAC_CUDA(cudaMemcpyAsync(&nnz, s->coo_block_off.p + nb, 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    if (nnz) {
      s->coo_dev.reserve(3 * nnz, s->stream);
      if (s->coo_host_cap < 3 * nnz) {
        if (s->coo_host) cudaFreeHost(s->coo_host);
        s->coo_host = nullptr; s->coo_host_cap = 0;
        const size_t cap = 3 * nnz + 3 * nnz / 4 + 1024;
        HLAC_CUDA(cudaHostAlloc((void**)&s->coo_host, cap * 4, cudaHostAllocDefault));
        s->coo_host_cap = cap;
      }
      k_coo_write<<<nb, COO_THREADS, 0, s->stream>>>(s->dense.p, nd, nrows, s->coo_block_off.p, s->coo_dev.p, s->coo_dev.p + nnz, s->coo_dev.p + 2 * nnz);
      HLAC_CUDA(cudaGetLastError());
      HLAC_CUDA(cudaMemcpyAsync(s->coo_host, s->coo_dev.p, 3 * nnz * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    s->launches += 3;
  }
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  if (ctr[9]) throw Error(HLAC_ERR_ARG, std::to_string(ctr[9]) + " reads carried a cell index >= n_cells (the reference panics in TriMat::add_triplet, main.rs:279)");
  if (metrics) {
    metrics->num_reads = ctr[1] + s->host_reads; metrics->num_non_primary = ctr[2] + s->host_non_primary; metrics->num_not_cell_bc = ctr[3] + s->host_not_cell;
    metrics->num_cds_align = ctr[4]; metrics->num_gen_align = ctr[5]; metrics->num_not_aligned = ctr[6];
  }
  if (out) {
    // the arrays live in the session's pinned buffer: valid until the next hlac_count_entries / _reset / _free on this
    // session; hlac_coo_free does not free borrowed arrays
    out->n = nnz; out->nrows = nrows; out->borrowed = 1;
    out->row = s->coo_host; out->col = s->coo_host ? s->coo_host + nnz : nullptr; out->val = s->coo_host ? s->coo_host + 2 * nnz : nullptr;
  }
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// Reads whose barcode is not in the whitelist are never scored (mapping.rs:760-763): a host driver may drop them before the
// device and report only how they count in the metrics.
int hlac_count_add_filtered(hlac_count* s, uint64_t n_reads, uint64_t n_non_primary, uint64_t n_not_cell) {
  if (!s) { set_last_error("null session"); return HLAC_ERR_ARG; }
  s->host_reads += n_reads; s->host_non_primary += n_non_primary; s->host_not_cell += n_not_cell;
  return HLAC_OK;
}

int hlac_count_finish(hlac_count* s, hlac_coo* out, hlac_metrics* metrics) {
  int rc = hlac_count_reduce(s, nullptr, nullptr);
  if (rc != HLAC_OK) return rc;
  return hlac_count_entries(s, out, metrics);
}

int hlac_count_reset(hlac_count* s) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  s->reset();
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_coo_free(hlac_coo* c) {
  if (c) { if (!c->borrowed) { free(c->row); free(c->col); free(c->val); } c->row = c->col = c->val = nullptr; c->n = 0; }
  return HLAC_OK;
}

int hlac_count_timing(hlac_count* s, float ms[4], uint64_t* launches) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  s->drain_events();
  if (ms) for (int i = 0; i < 4; i++) ms[i] = s->ms[i];
  if (launches) *launches = s->launches;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// seed-filter counters since the last reset: l-mer bitmap tests and dictionary probes (SURVEY.md §8d "probes/read")
int hlac_count_probe_stats(hlac_count* s, uint64_t* bitmap_tests, uint64_t* dict_probes) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  unsigned long long c[2];
  HLAC_CUDA(cudaMemcpyAsync(c, s->counters.p + 7, sizeof c, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  if (bitmap_tests) *bitmap_tests = c[0];
  if (dict_probes) *dict_probes = c[1];
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

}  // extern "C"
