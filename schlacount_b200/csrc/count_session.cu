// Counting session: map_and_count_pseudo's hot loop (src/mapping.rs:740-807) and count() (src/mapping.rs:823-909)
// on the GPU.
//
//   K1  k_count_map_q       queue-staged: filters (mapping.rs:752-763), lazy 4-way map_read policy (mapping.rs:765-793),
//                           class -> (gene, slot) (mapping.rs:795-806); every scored read leaves as one packed 64-bit key
//                           (umi | cell | code) appended to the hash BUCKET of its (cell, UMI) molecule.
//   K3  k_group_consensus   one block per bucket: the "group by cell then UMI" of mapping.rs:829-841 as a shared-memory
//                           hash grouping (no global sort: count() only needs the reads of a molecule together, and the
//                           result is accumulated into a dense matrix, so no order is needed anywhere), then the gene
//                           majority + allele decision of mapping.rs:843-896 per molecule -> dense u32 [n_cells][nrows].
//   K4  k_coo_count/_scan/_write  dense matrix -> ordered COO triplets written straight into host-mapped pinned memory,
//                           together with the counters: the host synchronises ONCE per finish.
// Fallback (a bucket overflowed: a molecule deeper than a bucket, or a hostile hash pattern): all keys are gathered, sorted
// (radix_sort.cuh) and reduced by k_consensus, one thread per (cell, UMI) run — kept for exactness, never on the normal path.
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
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "comm.hpp"
#include "device_index.hpp"
#include "hla_fasta.hpp"
#include "radix_sort.cuh"

using namespace hlac;

namespace {

constexpr uint32_t CODE_NONE5 = 31;   // 5-bit "no code"
constexpr int CODE_BITS = 5;
constexpr int UMI_BITS = 32;
constexpr uint32_t BUCKET_CAP = 2048;      // keys per bucket (16 KB of shared memory in k_group_consensus)
constexpr uint32_t BUCKET_TARGET = 1024;   // buckets are provisioned for at most this many keys on average

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
  return x;
}

// where the keys of the scored reads go: bucket = hash of the molecule (umi, cell), so a molecule is complete in one bucket
struct BucketStore {
  uint64_t* keys;                 // [n_buckets][BUCKET_CAP]
  uint32_t* fill;                 // [n_buckets] keys offered to the bucket (> BUCKET_CAP: the excess went to `ovf`)
  uint32_t mask;                  // n_buckets - 1 (power of two)
  uint64_t* ovf;                  // overflow list
  unsigned long long* ovf_cursor;
  uint64_t ovf_cap;
};
__device__ __forceinline__ void bucket_append(const BucketStore& bs, uint64_t key) {
  const uint32_t b = (uint32_t)mix64(key >> CODE_BITS) & bs.mask;
  const uint32_t at = atomicAdd(bs.fill + b, 1u);
  if (at < BUCKET_CAP) bs.keys[(size_t)b * BUCKET_CAP + at] = key;
  else { const unsigned long long o = atomicAdd(bs.ovf_cursor, 1ULL); if (o < bs.ovf_cap) bs.ovf[o] = key; }
}

struct CountMapArgs {
  DevIndexView cds, gen;
  const uint64_t* seq;
  const uint16_t* len;
  const uint32_t* cell;
  const uint32_t* umi;
  const uint8_t* flags;
  uint64_t n;
  uint32_t wpr;                // u64 words per read
  uint32_t nw;                 // u32 words per strand row in shared memory (2*wpr + 1)
  uint32_t row_stride;         // u32 words per slot (both strands + the read index, odd)
  int primary_only;
  int scan_budget;   // l-mer tests of an orientation's first look (k_count_map_q scan rounds)
  uint32_t cell_bits;          // key layout: umi << (5 + cell_bits) | cell << 5 | code
  uint32_t n_cells;            // cell indices >= n_cells are a caller error: dropped and counted in metrics[8]
  BucketStore bk;
  unsigned long long* metrics; // [6] + [2] seed-filter tests / dictionary probes + [1] out-of-range cell indices + [1] max UMI key
  uint8_t* codes;              // optional per-read debug output
  const uint64_t* rec;         // packed records (hlac_packed_batch) instead of the five arrays above, or NULL
  const uint32_t* wire;        // wire records (hlac_wire_batch): bit-packed u32 words, or NULL
  uint32_t wire_words, wire_len, wire_cell_bits, wire_umi_bits;
};

// nbits (1..32) of a big-endian bit stream held in u32 words, starting at bit `pos` (bit 0 = MSB of word 0)
__device__ __forceinline__ uint32_t wire_bits(const uint32_t* __restrict__ w, uint32_t pos, uint32_t nbits, uint32_t n_words) {
  const uint32_t i = pos >> 5, sh = pos & 31;
  const uint32_t hi = __ldg(w + i), lo = (i + 1 < n_words) ? __ldg(w + i + 1) : 0u;
  return __funnelshift_l(lo, hi, sh) >> (32 - nbits);
}

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

struct BitmapGlobal {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return __ldg(p + i); }
};

// ---- bulk-copy (TMA) + mbarrier primitives for the record prefetch --------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Queue-staged map kernel (ncu history of its predecessors in profiles/: thread-per-read 6.65 of 32 lanes active per issued
// instruction, block-staged 12.6 lanes but barrier-bound, warp-staged tiles 12.2, queue-staged 14.65). Each
// warp keeps a pool of NS read slots and three warp-private task queues that persist across refills: SCAN entries are
// (slot, p) with p = the orientation/index the chain of mapping.rs:772-790 has reached (their first look, bounded by
// scan_budget l-mer tests), LONG entries are the same tasks still scanning after their first look, WALK entries are slots
// with a seed. A round of any kind only runs when 32 tasks are waiting (or nothing else can be done), so every
// round starts with all lanes busy whatever fraction of the reads survives each stage; scan rounds mix the four
// orientations (same code, per-lane operands) but not short and long scans. The warp streams a contiguous range of the
// batch and refills the pool when at least 16 slots are free.
//
// Shared memory is addressed through explicit 32-bit shared-window addresses (device_index.cuh). One slot is
// row_stride = 2 * nw + 1 words (odd: lanes on different slots never share a bank): the forward row, the reverse-complement
// row (nw = 2 * wpr + 1 words each: the packed bases plus one word the 16-base extraction may touch but whose content
// never reaches a result) and the read's index in the batch; the spare word of the forward row holds the seed's node, the
// spare word of the reverse row the seed's (offset, position, orientation).
//
// PACKED && TMA: the warp's contiguous records are prefetched 32 at a time by ONE bulk asynchronous copy
// (cp.async.bulk, completion on a per-warp mbarrier) into a per-warp staging tile; the copy of the next tile is issued as
// soon as the current one has been handed to slots, and lands while the warp runs scan / walk rounds, so a refill reads
// shared memory instead of waiting on 32 dependent global loads.
//
// Shared memory doubles as the L1 carve-out, and the walk's node records / node sequences live in L1: the r02 kernel got
// 10 % faster by shrinking the slots alone (215 -> 170 KB requested, L1 28 -> 60 KB), and the bulk prefetch, which costs
// 32 KB of staging, was measured SLOWER on config 2 for that reason (1.25 vs 1.17 ms per 10 M reads). Hence two more
// knobs: NS (slots per warp: fewer slots = more L1) and IDX_SMEM (for the small indexes of known-genotype counting the
// node tables and node sequences are staged in shared memory themselves, so only dictionary probes go through L1).
__host__ __device__ constexpr uint32_t map_nw(uint32_t wpr) { return 2 * wpr + 1; }
__host__ __device__ constexpr uint32_t map_row_stride(uint32_t wpr) { return 2 * map_nw(wpr) + 1; }
__host__ __device__ constexpr uint32_t map_warp_bytes(uint32_t wpr, bool tma, uint32_t ns) {
  // slots + lens u16 + four byte queues, rounded to 16 B; staging tile (32 records of (wpr + 1) u64) + mbarrier
  return ((ns * map_row_stride(wpr) * 4 + ns * 2 + ns * 4 + 15) & ~15u) + (tma ? 32 * (wpr + 1) * 8 + 16 : 0);
}
__host__ __device__ constexpr uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

enum { IN_SOA = 0, IN_PACKED = 1, IN_WIRE = 2 };
template <bool SMEM_BM, int MAP_THREADS, int MINB, int INPUT, int STRIDE, bool TMA, int NS = 64, bool IDX_SMEM = false>
__global__ void __launch_bounds__(MAP_THREADS, MINB) k_count_map_q(const CountMapArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr bool PACKED = INPUT == IN_PACKED;
  static_assert(!TMA || PACKED, "the bulk prefetch reads packed records");
  static_assert(NS <= 64 && NS % 8 == 0 && NS >= 32, "slot ids are 6-bit queue entries");
  static_assert(!IDX_SMEM || SMEM_BM, "the staged index sits behind the staged bitmaps");
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  uint32_t bm_bytes = 0;
  uint32_t g_cds_nodes = 0, g_cds_seq = 0, g_gen_nodes = 0, g_gen_seq = 0;   // IDX_SMEM: shared addresses of the staged graphs
  if (SMEM_BM) {
    const uint32_t nb = a.cds.bitmap_words + a.gen.bitmap_words;
    for (uint32_t i = threadIdx.x; i < a.cds.bitmap_words; i += MAP_THREADS) smem[i] = __ldg(a.cds.bitmap + i);
    for (uint32_t i = threadIdx.x; i < a.gen.bitmap_words; i += MAP_THREADS) smem[a.cds.bitmap_words + i] = __ldg(a.gen.bitmap + i);
    bm_bytes = align16(nb * 4);
    if (IDX_SMEM) {
      const uint32_t w0 = bm_bytes / 4, w1 = w0 + a.cds.n_nodes * 12, w2 = w1 + a.gen.n_nodes * 12, w3 = w2 + align16(a.cds.seq_words * 4) / 4,
                     w4 = w3 + align16(a.gen.seq_words * 4) / 4;
      const uint32_t* src;
      src = reinterpret_cast<const uint32_t*>(a.cds.nodes); for (uint32_t i = threadIdx.x; i < a.cds.n_nodes * 12; i += MAP_THREADS) smem[w0 + i] = __ldg(src + i);
      src = reinterpret_cast<const uint32_t*>(a.gen.nodes); for (uint32_t i = threadIdx.x; i < a.gen.n_nodes * 12; i += MAP_THREADS) smem[w1 + i] = __ldg(src + i);
      for (uint32_t i = threadIdx.x; i < a.cds.seq_words; i += MAP_THREADS) smem[w2 + i] = __ldg(a.cds.seq32 + i);
      for (uint32_t i = threadIdx.x; i < a.gen.seq_words; i += MAP_THREADS) smem[w3 + i] = __ldg(a.gen.seq32 + i);
      g_cds_nodes = sbase + w0 * 4; g_gen_nodes = sbase + w1 * 4; g_cds_seq = sbase + w2 * 4; g_gen_seq = sbase + w3 * 4;
      bm_bytes = w4 * 4;
    }
    __syncthreads();
  }
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned below = (1u << lane) - 1;
  const uint32_t nw = a.nw, rs4 = a.row_stride * 4;
  const uint32_t wb = sbase + bm_bytes + warp * map_warp_bytes(a.wpr, TMA, NS);   // this warp's region
  const uint32_t a_lens = wb + NS * rs4;            // u16[NS]
  const uint32_t a_qs = a_lens + NS * 2;            // scan queue (first look): slot | p << 6
  const uint32_t a_qw = a_qs + NS;                  // walk queue
  const uint32_t a_fr = a_qw + NS;                  // free-slot stack
  const uint32_t a_ql = a_fr + NS;                  // scan queue (long scans): slot | p << 6, scan state in the reverse row's spare word
  const uint32_t a_stage = wb + ((NS * rs4 + NS * 2 + NS * 4 + 15) & ~15u);   // TMA: 32 records
  const uint32_t a_bar = a_stage + 32 * (a.wpr + 1) * 8;
  const uint32_t bm_cds = sbase, bm_gen = sbase + a.cds.bitmap_words * 4;
  for (unsigned k = lane; k < (unsigned)NS; k += 32) sts8(a_fr + k, k);
  unsigned ns = 0, nl = 0, nwk = 0, nfree = NS;                             // warp-uniform
  uint32_t m_reads = 0, m_nonprim = 0, m_notcell = 0, m_cds = 0, m_gen = 0, m_none = 0, m_badcell = 0;   // per thread: a batch is < 2^32 reads
  uint32_t n_tests = 0, n_probes = 0, umi_max = 0;
  const uint64_t n_warps = (uint64_t)gridDim.x * (MAP_THREADS / 32);
  const uint64_t per = ((a.n + n_warps - 1) / n_warps + 31) & ~31ull;      // whole 32-record tiles per warp
  uint64_t cur = min(a.n, ((uint64_t)blockIdx.x * (MAP_THREADS / 32) + warp) * per);
  const uint64_t hi = min(a.n, cur + per);
  const uint32_t rec_bytes = (a.wpr + 1) * 8;
  // bulk prefetch state (warp-uniform): records of the staged tile not yet handed out, the tile in flight
  uint32_t st_n = 0, st_next = 0, fl_n = 0, phase = 0;
  uint64_t fl_first = cur;   // first record of the tile to request next
  if (TMA) {
    if (lane == 0) mbar_init(a_bar, 1);
    __syncwarp();
    if (fl_first < hi) {
      fl_n = (uint32_t)min((uint64_t)32, hi - fl_first);
      if (lane == 0) bulk_g2s(a_stage, a.rec + fl_first * (a.wpr + 1), fl_n * rec_bytes, a_bar);
      fl_first += fl_n;
    }
  }
  __syncwarp();
  for (;;) {
    const bool more = cur < hi;
    int what;   // 0 walk, 1 scan (first look), 2 refill, 3 scan (long)
    if (nwk >= 32) what = 0;
    else if (nl >= 32) what = 3;
    else if (ns >= 32) what = 1;
    else if (more && nfree >= 16) what = 2;   // no full round waiting: with 64 slots fewer than 16 free ones leave some queue >= 17 tasks
    else if (nwk + nl + ns > 0) what = (nwk >= nl && nwk >= ns) ? 0 : (nl >= ns ? 3 : 1);
    else if (more) what = 2;                  // every slot is free
    else break;
    if (what == 2) {
      // ---- refill: filters (mapping.rs:752-763) + read rows into free slots ----
      unsigned c;
      if (TMA) {
        if (st_next == st_n) {   // staging tile used up: the one in flight has (long) landed
          mbar_wait(a_bar, phase); phase ^= 1;
          st_n = fl_n; st_next = 0; fl_n = 0;
        }
        c = min(min(32u, nfree), st_n - st_next);
      } else c = (unsigned)min((uint64_t)min(32u, nfree), hi - cur);
      bool keep = false;
      unsigned slot = 0;
      if (lane < c) {
        const uint64_t i = cur + lane;
        slot = lds8(a_fr + nfree - c + lane);
        const uint32_t fw = wb + slot * rs4, rc = fw + nw * 4;
        uint8_t fl; uint32_t cell; uint32_t L;
        if (PACKED) {   // one record per read: sequence words, then umi | primary | len | cell
          uint32_t meta;
          if (TMA) {
            const uint32_t src = a_stage + (st_next + lane) * rec_bytes;
            if ((a.wpr & 1) && a.wpr == 3) {   // 32-byte records: two 128-bit loads
              const uint4 q0 = lds128(src), q1 = lds128(src + 16);
              sts32(fw, q0.y); sts32(fw + 4, q0.x); sts32(fw + 8, q0.w); sts32(fw + 12, q0.z); sts32(fw + 16, q1.y); sts32(fw + 20, q1.x);
              meta = q1.z;
            } else {
              for (uint32_t w = 0; w < a.wpr; w++) { const uint2 v = lds64(src + 8 * w); sts32(fw + 8 * w, v.y); sts32(fw + 8 * w + 4, v.x); }
              meta = lds32(src + 8 * a.wpr);
            }
          } else {
            const uint64_t* words = a.rec + i * (a.wpr + 1);
            meta = (uint32_t)__ldg(words + a.wpr);
            for (uint32_t w = 0; w < a.wpr; w++) { const uint64_t v = __ldg(words + w); sts32(fw + 8 * w, (uint32_t)(v >> 32)); sts32(fw + 8 * w + 4, (uint32_t)v); }
          }
          fl = (meta >> 31) ? HLAC_FLAG_PRIMARY : 0;
          cell = meta & HLAC_PACKED_NOT_A_CELL; if (cell == HLAC_PACKED_NOT_A_CELL) cell = HLAC_NOT_A_CELL;
          L = min((meta >> 23) & 0xFFu, 32u * a.wpr);
        } else if (INPUT == IN_WIRE) {   // bit-packed: bases, then primary (1), cell (cell_bits, all ones = not a cell), umi
          const uint32_t* words = a.wire + i * a.wire_words;
          const uint32_t nsw = (2 * a.wire_len + 31) >> 5;   // words carrying bases (the last may end in meta bits: never read as bases)
          for (uint32_t w = 0; w < 2 * a.wpr; w++) sts32(fw + 4 * w, w < nsw ? __ldg(words + w) : 0u);
          const uint32_t pos = 2 * a.wire_len;
          fl = wire_bits(words, pos, 1, a.wire_words) ? HLAC_FLAG_PRIMARY : 0;
          cell = wire_bits(words, pos + 1, a.wire_cell_bits, a.wire_words);
          if (cell == (a.wire_cell_bits == 32 ? 0xFFFFFFFFu : (1u << a.wire_cell_bits) - 1u)) cell = HLAC_NOT_A_CELL;
          L = a.wire_len;
        } else {
          const uint64_t* words = a.seq + i * a.wpr;
          fl = __ldg(a.flags + i);
          cell = __ldg(a.cell + i);
          L = min((uint32_t)__ldg(a.len + i), 32u * a.wpr);   // never read past the packed words
          for (uint32_t w = 0; w < a.wpr; w++) { const uint64_t v = __ldg(words + w); sts32(fw + 8 * w, (uint32_t)(v >> 32)); sts32(fw + 8 * w + 4, (uint32_t)v); }
        }
        sts32(fw + 8 * a.wpr, 0);
        sts16(a_lens + 2 * slot, L); sts32(fw + rs4 - 4, (uint32_t)i);
        make_revcomp_s(SmemWords{fw}, rc, (int)L, (int)nw);   // here every lane is busy; in a mixed scan round only the p = 1 lanes would be
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
      if (TMA) {
        st_next += c;
        if (st_next == st_n && fl_first < hi) {   // tile handed out: request the next one, it lands during the rounds that follow
          fl_n = (uint32_t)min((uint64_t)32, hi - fl_first);
          if (lane == 0) bulk_g2s(a_stage, a.rec + fl_first * (a.wpr + 1), fl_n * rec_bytes, a_bar);
          fl_first += fl_n;
        }
      }
      const unsigned bk = __ballot_sync(0xFFFFFFFFu, keep), bd = __ballot_sync(0xFFFFFFFFu, lane < c && !keep);
      if (keep) sts8(a_qs + ns + __popc(bk & below), slot);                 // p = 0
      if (lane < c && !keep) sts8(a_fr + nfree + __popc(bd & below), slot);
      ns += __popc(bk); nfree += __popc(bd);
      __syncwarp();
    } else if (what != 0) {
      // ---- scan round: first-seed skip-scan of (slot, p). An orientation either seeds within its first window (three l-mer
      // tests and a dictionary probe) or, mostly, scans the whole read in vain (~20 tests), so a task's first look is bounded
      // by a.scan_budget tests; what is still scanning then moves to the long queue, whose rounds run to the end. Rounds of
      // either kind are homogeneous: lanes that seeded do not sit out the other lanes' long scans.
      const bool lng = what == 3;
      const unsigned have = lng ? nl : ns;
      const unsigned cnt = min(32u, have), qb = have - cnt;
      const bool task = lane < cnt;
      int res = SEED_NONE;
      unsigned slot = 0, p = 0;
      if (task) {
        const unsigned e = lds8((lng ? a_ql : a_qs) + qb + lane);
        slot = e & 63u; p = e >> 6;
        const uint32_t fw = wb + slot * rs4, rc = fw + nw * 4;
        const int L = (int)lds16(a_lens + 2 * slot);
        SmemWords rd{(p & 1) ? rc : fw};
        const DevIndexView& ix = p < 2 ? a.cds : a.gen;
        int pos = 0, q = K - (int)ix.lmer; uint32_t node = 0, off = 0;
        if (lng) { const uint32_t st = lds32(rc + 8 * a.wpr); pos = (int)(st & 0xFFFFu); q = (int)(st >> 16); }
        if (L >= K) {
          const int budget = lng ? 0x7FFFFFFF : a.scan_budget;
          if (SMEM_BM) res = find_seed_step<STRIDE, true>(ix, rd, SmemBitmap{p < 2 ? bm_cds : bm_gen}, pos, q, L - K, budget, node, off, n_tests, n_probes);
          else res = find_seed_step<STRIDE, true>(ix, rd, BitmapGlobal{ix.bitmap}, pos, q, L - K, budget, node, off, n_tests, n_probes);
        }
        if (res == SEED_FOUND) { sts32(fw + 8 * a.wpr, node); sts32(rc + 8 * a.wpr, (off << 12) | ((uint32_t)pos << 2) | p); }
        else if (res == SEED_MORE) sts32(rc + 8 * a.wpr, (uint32_t)pos | ((uint32_t)q << 16));
      }
      __syncwarp();
      if (lng) nl = qb; else ns = qb;
      const bool found = task && res == SEED_FOUND, cont = task && res == SEED_MORE;
      const bool again = task && res == SEED_NONE && p < 3, dead = task && res == SEED_NONE && p == 3;
      const unsigned bf = __ballot_sync(0xFFFFFFFFu, found), bn = __ballot_sync(0xFFFFFFFFu, again),
                     bz = __ballot_sync(0xFFFFFFFFu, dead), bc = __ballot_sync(0xFFFFFFFFu, cont);
      if (found) sts8(a_qw + nwk + __popc(bf & below), slot);
      if (again) sts8(a_qs + ns + __popc(bn & below), slot | ((p + 1) << 6));
      if (cont) sts8(a_ql + nl + __popc(bc & below), slot | (p << 6));
      if (dead) sts8(a_fr + nfree + __popc(bz & below), slot);             // all four None (mapping.rs:791-793)
      nwk += __popc(bf); ns += __popc(bn); nl += __popc(bc); nfree += __popc(bz);
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
        slot = lds8(a_qw + qb + lane);
        const uint32_t fw = wb + slot * rs4, rc = fw + nw * 4;
        const uint64_t i = lds32(fw + rs4 - 4);
        const uint32_t info = lds32(rc + 8 * a.wpr), node = lds32(fw + 8 * a.wpr);
        const int p = (int)(info & 3u), pos = (int)((info >> 2) & 1023u);
        const uint32_t off = info >> 12;
        const int L = (int)lds16(a_lens + 2 * slot);
        const DevIndexView& ix = p < 2 ? a.cds : a.gen;
        SmemWords rd{(p & 1) ? rc : fw};
        CountVis vis;
        uint32_t cov;
        if (IDX_SMEM) cov = walk_from_seed<STRIDE>(ix, SmemGraph{p < 2 ? g_cds_nodes : g_gen_nodes, p < 2 ? g_cds_seq : g_gen_seq}, rd, SmemBitmap{p < 2 ? bm_cds : bm_gen}, L, pos, node, off, vis, n_tests, n_probes);
        else if (SMEM_BM) cov = walk_from_seed<STRIDE>(ix, GlobalGraph{ix.nodes, ix.seq32}, rd, SmemBitmap{p < 2 ? bm_cds : bm_gen}, L, pos, node, off, vis, n_tests, n_probes);
        else cov = walk_from_seed<STRIDE>(ix, GlobalGraph{ix.nodes, ix.seq32}, rd, BitmapGlobal{ix.bitmap}, L, pos, node, off, vis, n_tests, n_probes);
        const bool good = cov > MIN_SCORE_COUNT_PSEUDO;
        if (good) { if (p < 2) m_cds++; else m_gen++; }
        if (p == 0 || good) code = vis.code();   // a CDS-forward hit is used whatever its coverage (mapping.rs:772-775)
        if (code != CODE_NONE5) {
          uint32_t um, cl;
          if (PACKED) { const uint64_t meta = __ldg(a.rec + i * (a.wpr + 1) + a.wpr); um = (uint32_t)(meta >> 32); cl = (uint32_t)meta & HLAC_PACKED_NOT_A_CELL; }
          else if (INPUT == IN_WIRE) {
            const uint32_t* words = a.wire + i * a.wire_words;
            cl = wire_bits(words, 2 * a.wire_len + 1, a.wire_cell_bits, a.wire_words);
            um = wire_bits(words, 2 * a.wire_len + 1 + a.wire_cell_bits, a.wire_umi_bits, a.wire_words);
          }
          else { um = __ldg(a.umi + i); cl = __ldg(a.cell + i); }
          umi_max = max(umi_max, um);
          key = ((uint64_t)um << (CODE_BITS + a.cell_bits)) | ((uint64_t)cl << CODE_BITS) | code;
          if (a.codes) a.codes[i] = (uint8_t)code;
        }
      }
      __syncwarp();
      nwk = qb;
      if (task) sts8(a_fr + nfree + lane, slot);
      nfree += cnt;
      if (code != CODE_NONE5) bucket_append(a.bk, key);
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

// ---- count() (mapping.rs:842-896) -------------------------------------------------------------------------------
// the decision for one molecule from its per-candidate counts: nreads scored reads, f = (allele 1, allele 2, gene) reads of
// the majority candidate gene. Returns the row slot or -1 (no gene above 50 %, mapping.rs:874-875). f64 compares as in
// the reference (mapping.rs:879-889).
__device__ __forceinline__ int consensus_slot(uint32_t nreads, const uint32_t f[3]) {
  const double nr = (double)nreads;
  const uint32_t ng = f[0] + f[1] + f[2];
  if (!((double)ng > GENE_CONSENSUS_THRESHOLD * nr)) return -1;
  if ((double)f[0] > ALLELE_CONSENSUS_THRESHOLD * nr && ALLELE_CONSENSUS_THRESHOLD * nr > (double)f[1]) return 0;
  if ((double)f[1] > ALLELE_CONSENSUS_THRESHOLD * nr && ALLELE_CONSENSUS_THRESHOLD * nr > (double)f[0]) return 1;
  return 2;
}

// One block per bucket (persistent over the bucket list). The keys of the bucket are grouped by molecule with an
// open-addressing table in shared memory: the first key of a molecule to claim a slot becomes its representative, later
// keys push themselves onto the representative's list (atomicExch); then one thread per molecule walks its list twice —
// Boyer-Moore majority gene (only a gene with > 50 % of the molecule's reads can win, and the vote finds it whatever the
// order), then the candidate's three counts — and adds 1 to the dense matrix. Nothing here depends on the order of the
// keys, and the dense matrix is order-free, so no sort is needed.
// (A variant that brought each bucket in by one bulk asynchronous copy - cp.async.bulk + mbarrier, double-buffered so that
// the next bucket lands behind the current one's grouping - was measured SLOWER on config 2: 0.180 vs 0.147 ms per 4.3 M keys.
// A bucket is ~2 KB, its coalesced load costs one L2 round trip of the 20-odd a block pays, and the second key buffer takes
// the kernel from 5 to 4 blocks per SM; the kernel is bound by its shared-memory atomics, not by the loads.)
constexpr int GROUP_THREADS = 256;
__global__ void __launch_bounds__(GROUP_THREADS) k_group_consensus(const BucketStore bs, uint32_t* __restrict__ dense, uint32_t nrows,
                                                                   const uint32_t* __restrict__ gene_row0, uint32_t cell_bits,
                                                                   unsigned long long* n_keys_out) {
  __shared__ uint64_t k[BUCKET_CAP];
  __shared__ uint32_t lnk[BUCKET_CAP];          // representative: head of its member list; member: next member (index + 1, 0 = end)
  __shared__ uint32_t ht[2 * BUCKET_CAP];       // index + 1 of the molecule's representative key
  if (*bs.ovf_cursor) return;                   // a bucket overflowed: the host takes the sort path for this step
  const unsigned t = threadIdx.x;
  unsigned long long my_keys = 0;
  for (uint32_t b = blockIdx.x; b <= bs.mask; b += gridDim.x) {
    const uint32_t n = min(bs.fill[b], BUCKET_CAP);
    if (n == 0) continue;
    uint32_t hsz = 64;
    while (hsz < 2 * n) hsz <<= 1;
    const uint32_t hmask = hsz - 1;
    __syncthreads();                            // the previous bucket's lists have been read
    const uint64_t* src = bs.keys + (size_t)b * BUCKET_CAP;
    for (uint32_t i = t; i < n; i += GROUP_THREADS) { k[i] = src[i]; lnk[i] = 0; }
    for (uint32_t i = t; i < hsz; i += GROUP_THREADS) ht[i] = 0;
    if (t == 0) my_keys += n;
    __syncthreads();
    for (uint32_t i = t; i < n; i += GROUP_THREADS) {
      const uint64_t g = k[i] >> CODE_BITS;
      uint32_t h = (uint32_t)(mix64(g) >> 32) & hmask;
      for (;;) {
        uint32_t cur = *(volatile uint32_t*)(ht + h);
        if (cur == 0) { cur = atomicCAS(ht + h, 0u, i + 1); if (cur == 0) break; }   // claimed: this key represents the molecule
        if ((k[cur - 1] >> CODE_BITS) == g) { lnk[i] = atomicExch(lnk + cur - 1, i + 1); break; }
        h = (h + 1) & hmask;
      }
    }
    __syncthreads();
    for (uint32_t h = t; h < hsz; h += GROUP_THREADS) {
      const uint32_t rep = ht[h];
      if (rep == 0) continue;
      // pass 1: majority-vote candidate gene
      uint32_t cand = (uint32_t)(k[rep - 1] & 31u) / 3, votes = 1, nreads = 1;
      for (uint32_t m = lnk[rep - 1]; m; m = lnk[m - 1]) {
        const uint32_t g = (uint32_t)(k[m - 1] & 31u) / 3;
        if (votes == 0) { cand = g; votes = 1; }
        else if (g == cand) votes++;
        else votes--;
        nreads++;
      }
      // pass 2: the candidate's (allele 1, allele 2, gene) counts
      uint32_t f[3] = {0, 0, 0};
      { const uint32_t c = (uint32_t)(k[rep - 1] & 31u); if (c / 3 == cand) f[c % 3]++; }
      for (uint32_t m = lnk[rep - 1]; m; m = lnk[m - 1]) { const uint32_t c = (uint32_t)(k[m - 1] & 31u); if (c / 3 == cand) f[c % 3]++; }
      const int slot = consensus_slot(nreads, f);
      if (slot < 0) continue;
      const uint32_t cell = (uint32_t)(k[rep - 1] >> CODE_BITS) & ((1u << cell_bits) - 1u);
      atomicAdd(dense + (size_t)cell * nrows + gene_row0[cand] + slot, 1u);
    }
  }
  if (t == 0 && my_keys) atomicAdd(n_keys_out, my_keys);
}

// the buckets double: every key of old bucket b goes to new bucket b or b + old_n (one more hash bit)
__global__ void k_rebucket(const BucketStore from, const BucketStore to) {
  for (uint32_t b = blockIdx.x; b <= from.mask; b += gridDim.x) {
    const uint32_t n = min(from.fill[b], BUCKET_CAP);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) bucket_append(to, from.keys[(size_t)b * BUCKET_CAP + i]);
  }
}

// ---- fallback path: gather every key into one array, radix sort, one thread per (cell, UMI) run ------------------
__global__ void k_gather_keys(const BucketStore bs, uint64_t* out, unsigned long long* cursor) {
  __shared__ unsigned long long base;
  for (uint32_t b = blockIdx.x; b <= bs.mask; b += gridDim.x) {
    const uint32_t n = min(bs.fill[b], BUCKET_CAP);
    __syncthreads();
    if (threadIdx.x == 0) base = n ? atomicAdd(cursor, (unsigned long long)n) : 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) out[base + i] = bs.keys[(size_t)b * BUCKET_CAP + i];
  }
}
__global__ void k_consensus(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ dense, uint32_t nrows,
                            const uint32_t* __restrict__ gene_row0, uint32_t cell_bits) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k0 = keys[i], grp = k0 >> CODE_BITS;
  if (i > 0 && (keys[i - 1] >> CODE_BITS) == grp) return;   // not a run head
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
  uint32_t f[3] = {0, 0, 0};
  for (uint64_t j = i; j < i + nreads; j++) {
    const uint32_t c = (uint32_t)(keys[j] & 31u);
    if (c / 3 == cand) f[c % 3]++;
  }
  const int slot = consensus_slot(nreads, f);
  if (slot < 0) return;
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
// single block: exclusive scan of the per-block counts; total to block_off[nblocks]. It also publishes the step's scalars
// to the host-mapped result header: [0] nnz, [1..] a copy of the session counters.
constexpr int N_COUNTERS = 16;
__global__ void k_coo_scan(const uint32_t* __restrict__ block_nnz, uint32_t nblocks, uint64_t* block_off,
                           const unsigned long long* __restrict__ counters, unsigned long long* host_hdr) {
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
  if (threadIdx.x == 0) { block_off[nblocks] = carry; if (host_hdr) host_hdr[0] = carry; }
  if (host_hdr && threadIdx.x < N_COUNTERS) host_hdr[1 + threadIdx.x] = counters[threadIdx.x];
}
// row / col / val may point into host-mapped pinned memory (zero-copy stores: only the non-zeros cross PCIe, and the host
// needs no second synchronisation to learn how many there are). cap = capacity of the three arrays (entries beyond it are
// dropped; the host sees nnz > cap and reports it).
__global__ void __launch_bounds__(COO_THREADS) k_coo_write(const uint32_t* __restrict__ dense, uint64_t n, uint32_t nrows, const uint64_t* __restrict__ block_off,
                            uint32_t* row, uint32_t* col, uint32_t* val, uint64_t cap) {
  __shared__ uint32_t warp_off[COO_THREADS / 32];
  __shared__ uint32_t st_row[COO_TILE], st_col[COO_TILE], st_val[COO_TILE];   // the block's triplets, staged so that the copy-out is coalesced
  const uint64_t base = (uint64_t)blockIdx.x * COO_TILE + (uint64_t)threadIdx.x * COO_PER_THREAD;
  uint32_t v[COO_PER_THREAD]; uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < COO_PER_THREAD; k++) { v[k] = base + k < n ? dense[base + k] : 0; c += v[k] != 0; }
  uint32_t incl = c;   // inclusive scan of the per-thread counts inside the warp
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((threadIdx.x & 31) >= (unsigned)o) incl += t; }
  if ((threadIdx.x & 31) == 31) warp_off[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0, total = 0;
  for (unsigned w = 0; w < COO_THREADS / 32; w++) { if (w < (threadIdx.x >> 5)) woff += warp_off[w]; total += warp_off[w]; }
  uint32_t at = woff + incl - c;
#pragma unroll
  for (int k = 0; k < COO_PER_THREAD; k++) if (v[k]) {
    const uint64_t e = base + k;
    st_row[at] = (uint32_t)(e % nrows); st_col[at] = (uint32_t)(e / nrows); st_val[at] = v[k];
    at++;
  }
  __syncthreads();
  const uint64_t out0 = block_off[blockIdx.x];
  for (uint32_t i = threadIdx.x; i < total; i += COO_THREADS) {
    const uint64_t o = out0 + i;
    if (o < cap) { row[o] = st_row[i]; col[o] = st_col[i]; val[o] = st_val[i]; }
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
  int stride = 1;                        // seed-scan stride (hlac_set_seed_stride) the session was created with
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  RowLayout rows;
  uint32_t n_cells = 0;
  int primary_only = 0;
  int scan_budget = 7;   // HLAC_SCAN_BUDGET (measured on config 2: 3 -> 1.15 ms, 5 -> 1.07, 6-8 -> 1.025, 12 -> 1.05, unbounded -> 1.10)
  DevBuf<uint32_t> nodes_cds, nodes_gen, gene_row0, dense;   // node tables with cinfo in place of the colour id
  // bucket store of the scored reads' keys; `keys` = overflow list (and the gather target of the fallback path)
  DevBuf<uint64_t> bk_keys, keys, keys_sorted, keys_tmp;
  DevBuf<uint32_t> bk_fill;
  uint32_t n_buckets = 0;
  uint64_t host_reads = 0, host_non_primary = 0, host_not_cell = 0;   // reads the host driver filtered out itself (hlac_count_add_filtered)
  struct Stage {   // one staging set for host batches; two of them so the H2D of batch i+1 overlaps the kernel of batch i
    DevBuf<uint64_t> seq, rec; DevBuf<uint16_t> len; DevBuf<uint32_t> cell, umi, wire; DevBuf<uint8_t> flags;
    cudaEvent_t copied = nullptr, consumed = nullptr;
  } stage[2];
  cudaStream_t copy_stream = nullptr;
  uint64_t n_push = 0;
  DevBuf<uint8_t> codes;
  DevBuf<uint32_t> coo_block_nnz; DevBuf<uint64_t> coo_block_off;             // device-side COO extraction
  uint32_t* coo_host = nullptr; size_t coo_cap = 0;                           // pinned + mapped: row[cap] col[cap] val[cap], written by k_coo_write
  unsigned long long* host_hdr = nullptr;                                     // pinned + mapped: [0] nnz, [1..16] counters
  // [0] keys grouped, [1..6] metrics, [7] bitmap tests, [8] dictionary probes, [9] bad cells, [10] max UMI key, [11] overflow cursor,
  // [12] gather cursor (fallback)
  // [13..15] reads the host driver filtered out itself (uploaded at finish so that a merge sums them like the others)
  DevBuf<unsigned long long> counters, counters_merged;
  RadixSortScratch rs_scratch;
  uint64_t reads_pushed = 0, last_n = 0;
  bool smem_bm = false;
  size_t bitmap_smem_bytes = 0;
  int n_sm = 1;
  bool want_codes = false;
  bool reduced = false;
  uint32_t cell_bits = 1;
  uint32_t cfg_wpr = 0;
  int cfg_occ = 0;
  const void* cfg_shape = nullptr;
  bool cfg_tma_fits = false, last_tma = false; size_t last_smem = 0;
  int group_occ = 0;
  uint64_t fallbacks = 0;                // finishes that had to take the sort path
  hlac_comm* comm = nullptr; int comm_root = -1;
  // timing
  struct Ev { cudaEvent_t a, b; int kind; };
  std::vector<Ev> events;
  float ms[4] = {0, 0, 0, 0};
  uint64_t launches = 0;

  ~hlac_count() {
    for (auto& e : events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (coo_host) cudaFreeHost(coo_host);
    if (host_hdr) cudaFreeHost(host_hdr);
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
  size_t nd() const { return (size_t)n_cells * std::max<uint32_t>(rows.nrows, 1); }
  BucketStore store() { return BucketStore{bk_keys.p, bk_fill.p, n_buckets - 1, keys.p, counters.p + 11, keys.cap}; }
  void reset() {
    HLAC_CUDA(cudaMemsetAsync(counters.p, 0, N_COUNTERS * sizeof(unsigned long long), stream));
    if (n_buckets) HLAC_CUDA(cudaMemsetAsync(bk_fill.p, 0, (size_t)n_buckets * 4, stream));
    reads_pushed = 0; last_n = 0; reduced = false; host_reads = host_non_primary = host_not_cell = 0;
    drain_events(); ms[0] = ms[1] = ms[2] = ms[3] = 0; launches = 0;
  }

  // make room for `total` reads in all: buckets provisioned for <= BUCKET_TARGET keys on average (every read could be
  // scored), the overflow list for every read. Growing re-buckets what is already there (one more hash bit per doubling).
  void ensure_capacity(uint64_t total) {
    keys.reserve(std::max<uint64_t>(total, 1), stream, true, keys.cap);   // keeps the overflow entries written so far
    uint64_t want = 64;
    while (want * BUCKET_TARGET < total) want <<= 1;
    if (want > (1ull << 30)) throw Error(HLAC_ERR_LIMIT, "too many reads for one counting session");
    if (want <= n_buckets) return;
    DevBuf<uint64_t> nk; DevBuf<uint32_t> nf;
    nk.reserve((size_t)want * BUCKET_CAP, stream); nf.reserve(want, stream);
    HLAC_CUDA(cudaMemsetAsync(nf.p, 0, (size_t)want * 4, stream));
    if (n_buckets && reads_pushed) {
      const BucketStore from = store();
      BucketStore to{nk.p, nf.p, (uint32_t)want - 1, keys.p, counters.p + 11, keys.cap};
      k_rebucket<<<std::min<uint32_t>(n_buckets, (uint32_t)n_sm * 8), 256, 0, stream>>>(from, to);
      HLAC_CUDA(cudaGetLastError());
      launches++;
    }
    HLAC_CUDA(cudaStreamSynchronize(stream));
    std::swap(bk_keys.p, nk.p); std::swap(bk_keys.cap, nk.cap);
    std::swap(bk_fill.p, nf.p); std::swap(bk_fill.cap, nf.cap);
    n_buckets = (uint32_t)want;
  }

  using Kern = void (*)(const CountMapArgs);
  // launch shapes: (bitmaps in shared memory?, threads per block, slots per warp, index graphs in shared memory?) with the
  // three input variants (five arrays, packed records, packed records through the bulk prefetch)
  struct Shape { bool bm; int threads, ns; bool idx; Kern k, k_packed, k_tma, k_wire; };
  template <int S>
  static const Shape* shapes(int& n) {
    // in order of preference (measured on config 2, profiles/r02_count_map_tuning.md); the first one that fits is used
    static const Shape sh[] = {
#define HLAC_QSHAPE(BM, TT, MB, NSL, IDX) {BM, TT, NSL, IDX, k_count_map_q<BM, TT, MB, IN_SOA, S, false, NSL, IDX>, k_count_map_q<BM, TT, MB, IN_PACKED, S, false, NSL, IDX>, \
                                          k_count_map_q<BM, TT, MB, IN_PACKED, S, true, NSL, IDX>, k_count_map_q<BM, TT, MB, IN_WIRE, S, false, NSL, IDX>}
        HLAC_QSHAPE(true, 1024, 1, 64, true), HLAC_QSHAPE(true, 1024, 1, 48, true), HLAC_QSHAPE(true, 1024, 1, 56, false), HLAC_QSHAPE(true, 1024, 1, 64, false),
        HLAC_QSHAPE(true, 1024, 1, 40, false), HLAC_QSHAPE(true, 512, 2, 64, false), HLAC_QSHAPE(false, 256, 4, 64, false), HLAC_QSHAPE(false, 128, 8, 64, false),
#undef HLAC_QSHAPE
    };
    n = (int)(sizeof sh / sizeof sh[0]);
    return sh;
  }

  void launch_map(const hlac_batch& b /* device pointers */, const uint64_t* rec = nullptr /* packed records instead of b's arrays */,
                  const hlac_wire_batch* wire = nullptr /* wire records (device pointer inside) instead of both */) {
    if (b.n == 0) return;
    if (b.words_per_read == 0 || b.words_per_read > 16) throw Error(HLAC_ERR_ARG, "words_per_read must be 1..16");
    if (b.n >= (1ull << 32)) throw Error(HLAC_ERR_LIMIT, "the count kernel takes batches below 2^32 reads");
    ensure_capacity(reads_pushed + b.n);
    if (want_codes) codes.reserve(b.n, stream);
    CountMapArgs a{};
    a.cds = cds->view; a.gen = gen->view;
    a.cds.nodes = reinterpret_cast<const uint4*>(nodes_cds.p); a.gen.nodes = reinterpret_cast<const uint4*>(nodes_gen.p);
    a.seq = b.seq; a.len = b.len; a.cell = b.cell; a.umi = b.umi; a.flags = b.flags;
    a.n = b.n; a.wpr = b.words_per_read; a.nw = map_nw(a.wpr); a.row_stride = map_row_stride(a.wpr);
    a.primary_only = primary_only; a.scan_budget = scan_budget; a.n_cells = n_cells; a.cell_bits = cell_bits;
    a.bk = store(); a.metrics = counters.p + 1;
    a.codes = want_codes ? codes.p : nullptr;
    a.rec = rec;
    if (wire) { a.wire = wire->rec; a.wire_words = wire->words_per_record; a.wire_len = wire->read_len; a.wire_cell_bits = wire->cell_bits; a.wire_umi_bits = wire->umi_bits; }
    // launch shape: (bitmaps in shared memory?, threads per block); HLAC_COUNT_SHAPE="smem,threads" overrides for tuning runs
    // the bulk (TMA) prefetch needs 16-byte record granularity: (wpr + 1) even, a 16-byte aligned record array. It is opt-in
    // (HLAC_COUNT_TMA=1): its staging tiles come out of the L1 carve-out and it was measured slower on config 2.
    const char* tma_env = getenv("HLAC_COUNT_TMA");
    const bool tma_ok = rec && (a.wpr & 1) && ((uintptr_t)rec & 15) == 0 && tma_env && atoi(tma_env) != 0;
    const size_t idx_bytes = (size_t)12 * 4 * (cds->view.n_nodes + gen->view.n_nodes) + align16(cds->view.seq_words * 4) + align16(gen->view.seq_words * 4);
    auto smem_for = [&](const Shape& sh, bool tma) {
      return (sh.bm ? (bitmap_smem_bytes + 15) & ~(size_t)15 : 0) + (sh.idx ? idx_bytes : 0) + (size_t)(sh.threads / 32) * map_warp_bytes(a.wpr, tma, sh.ns);
    };
    if (cfg_wpr != b.words_per_read) {
      int ns = 0;
      const Shape* sh = stride == 3 ? shapes<3>(ns) : shapes<1>(ns);
      const Shape* pick = nullptr;
      int want_bm = -1, want_t = 0, want_ns = 64, want_idx = 0;   // HLAC_COUNT_SHAPE="smem,threads,slots,idx" forces a shape (tuning runs)
      if (const char* e = getenv("HLAC_COUNT_SHAPE")) sscanf(e, "%d,%d,%d,%d", &want_bm, &want_t, &want_ns, &want_idx);
      for (int i = 0; i < ns; i++) {
        if (want_bm >= 0) { if ((int)sh[i].bm != want_bm || sh[i].threads != want_t || sh[i].ns != want_ns || (int)sh[i].idx != want_idx) continue; }
        else if (sh[i].bm != smem_bm) continue;
        if (sh[i].bm && !smem_bm) continue;
        if (smem_for(sh[i], false) > 227 * 1024) continue;
        pick = &sh[i]; break;
      }
      if (!pick) throw Error(HLAC_ERR_LIMIT, "no launch shape of the count kernel fits (reads too long, or bad HLAC_COUNT_SHAPE)");
      cfg_shape = pick;
      cfg_tma_fits = smem_for(*pick, true) <= 227 * 1024;
      HLAC_CUDA(cudaFuncSetAttribute(pick->k_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(*pick, false)));
      HLAC_CUDA(cudaFuncSetAttribute(pick->k_wire, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(*pick, false)));
      HLAC_CUDA(cudaFuncSetAttribute(pick->k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(*pick, false)));
      if (cfg_tma_fits) HLAC_CUDA(cudaFuncSetAttribute(pick->k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(*pick, true)));
      HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg_occ, pick->k, pick->threads, smem_for(*pick, false)));
      if (cfg_occ < 1) throw Error(HLAC_ERR_LIMIT, "count map kernel does not fit on the device");
      if (cfg_tma_fits) {
        int occ_t = 0;
        HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, pick->k_tma, pick->threads, smem_for(*pick, true)));
        if (occ_t < cfg_occ) cfg_tma_fits = false;   // never trade resident warps for the prefetch
      }
      cfg_wpr = b.words_per_read;
    }
    const Shape& shp = *(const Shape*)cfg_shape;
    const int cfg_threads = shp.threads;
    const bool tma = tma_ok && cfg_tma_fits;
    Kern kern = wire ? shp.k_wire : rec ? (tma ? shp.k_tma : shp.k_packed) : shp.k;
    const size_t smem = smem_for(shp, tma);
    const uint64_t need = (b.n + (uint64_t)cfg_threads * 2 - 1) / ((uint64_t)cfg_threads * 2);
    const unsigned grid = (unsigned)std::min<uint64_t>(need, (uint64_t)n_sm * cfg_occ);
    last_tma = tma; last_smem = smem;
    Ev& e = begin(0);
    kern<<<grid, cfg_threads, smem, stream>>>(a);
    HLAC_CUDA(cudaGetLastError());
    end(e);
    launches++;
    reads_pushed += b.n; last_n = b.n; reduced = false;
  }

  // count() into the dense matrix, enqueued only (no host synchronisation)
  void enqueue_group() {
    HLAC_CUDA(cudaMemsetAsync(dense.p, 0, nd() * sizeof(uint32_t), stream));
    if (!n_buckets) return;
    if (!group_occ) HLAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&group_occ, k_group_consensus, GROUP_THREADS, 0));
    Ev& e = begin(1);
    k_group_consensus<<<std::min<uint32_t>(n_buckets, (uint32_t)(n_sm * std::max(group_occ, 1))), GROUP_THREADS, 0, stream>>>(
        store(), dense.p, rows.nrows, gene_row0.p, cell_bits, counters.p);
    HLAC_CUDA(cudaGetLastError());
    end(e);
    launches++;
  }
  // the round-1 path: gather all keys, radix sort, one thread per run. Synchronous; used when a bucket overflowed.
  void fallback_group() {
    unsigned long long c[N_COUNTERS];
    HLAC_CUDA(cudaMemcpyAsync(c, counters.p, sizeof c, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    const uint64_t n_ovf = c[11];
    if (n_ovf > keys.cap) throw Error(HLAC_ERR_STATE, "overflow list overran (internal error)");
    HLAC_CUDA(cudaMemsetAsync(dense.p, 0, nd() * sizeof(uint32_t), stream));
    // work arrays of the sort (the overflow list itself stays intact: a later finish of the same session needs it again)
    const uint64_t upper = n_ovf + std::min<uint64_t>(reads_pushed, (uint64_t)n_buckets * BUCKET_CAP);
    keys_sorted.reserve(std::max<uint64_t>(upper, 1), stream); keys_tmp.reserve(std::max<uint64_t>(upper, 1), stream);
    if (n_ovf) HLAC_CUDA(cudaMemcpyAsync(keys_sorted.p, keys.p, n_ovf * 8, cudaMemcpyDeviceToDevice, stream));
    unsigned long long cur = n_ovf;   // bucket contents are appended behind the overflow entries
    HLAC_CUDA(cudaMemcpyAsync(counters.p + 12, &cur, 8, cudaMemcpyHostToDevice, stream));
    k_gather_keys<<<std::min<uint32_t>(n_buckets, (uint32_t)n_sm * 8), 256, 0, stream>>>(store(), keys_sorted.p, counters.p + 12);
    HLAC_CUDA(cudaGetLastError());
    HLAC_CUDA(cudaMemcpyAsync(&cur, counters.p + 12, 8, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    const uint64_t nk = cur;
    if (nk) {
      int umi_bits = 1;
      while (umi_bits < UMI_BITS && (c[10] >> umi_bits)) umi_bits++;
      const int end_bit = CODE_BITS + (int)cell_bits + umi_bits;
      Ev& e1 = begin(1);
      const bool second = radix_sort_pairs<uint64_t, uint32_t>(keys_sorted.p, keys_tmp.p, nullptr, nullptr, (uint64_t)nk, CODE_BITS, end_bit, rs_scratch, stream);
      k_consensus<<<(unsigned)((nk + 255) / 256), 256, 0, stream>>>(second ? keys_tmp.p : keys_sorted.p, nk, dense.p, rows.nrows, gene_row0.p, cell_bits);
      HLAC_CUDA(cudaGetLastError());
      end(e1);
      launches += 2;
    }
    unsigned long long nk_ull = nk;
    HLAC_CUDA(cudaMemcpyAsync(counters.p, &nk_ull, 8, cudaMemcpyHostToDevice, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
    fallbacks++;
  }
  // dense -> COO into the host-mapped buffers, enqueued only
  void enqueue_coo(bool merged_by_caller = false) {
    const uint64_t n = nd();
    const uint32_t nrows = rows.nrows;
    // nnz <= min(matrix size, molecules <= reads pushed); a merged matrix (communicator, or a caller that all-reduced the
    // dense matrix between hlac_count_reduce and hlac_count_entries) holds every rank's molecules: bounded by its size only
    const uint64_t cap = (comm || merged_by_caller) ? n : std::min<uint64_t>(n, std::max<uint64_t>(reads_pushed, 1));
    if (!host_hdr) {
      HLAC_CUDA(cudaHostAlloc((void**)&host_hdr, (1 + N_COUNTERS) * 8, cudaHostAllocMapped | cudaHostAllocPortable));
      memset(host_hdr, 0, (1 + N_COUNTERS) * 8);
    }
    if (nrows && coo_cap < cap) {
      HLAC_CUDA(cudaStreamSynchronize(stream));
      if (coo_host) cudaFreeHost(coo_host);
      coo_host = nullptr; coo_cap = 0;
      const size_t want = std::min<uint64_t>(n, cap + cap / 2 + 1024);
      HLAC_CUDA(cudaHostAlloc((void**)&coo_host, want * 12, cudaHostAllocMapped | cudaHostAllocPortable));
      coo_cap = want;
    }
    const uint32_t nb = nrows ? (uint32_t)((n + COO_TILE - 1) / COO_TILE) : 0;
    coo_block_nnz.reserve(std::max<uint32_t>(nb, 1), stream); coo_block_off.reserve(nb + 1, stream);
    Ev& e = begin(2);
    if (nb) k_coo_count<<<nb, COO_THREADS, 0, stream>>>(dense.p, n, coo_block_nnz.p);
    k_coo_scan<<<1, 1024, 0, stream>>>(coo_block_nnz.p, nb, coo_block_off.p, comm ? counters_merged.p : counters.p, host_hdr);
    if (nb) k_coo_write<<<nb, COO_THREADS, 0, stream>>>(dense.p, n, nrows, coo_block_off.p, coo_host, coo_host + coo_cap, coo_host + 2 * coo_cap, coo_cap);
    HLAC_CUDA(cudaGetLastError());
    end(e);
    launches += nb ? 3 : 1;
  }
  // sum the dense matrices (disjoint columns per rank) and the counters over the communicator
  // (the counters go to a separate buffer: the local ones stay valid for a redo through the sort path)
  void enqueue_merge() {
    const unsigned long long hc[3] = {host_reads, host_non_primary, host_not_cell};
    HLAC_CUDA(cudaMemcpyAsync(counters.p + 13, hc, sizeof hc, cudaMemcpyHostToDevice, stream));
    if (!comm) return;
    counters_merged.reserve(N_COUNTERS, stream);
    comm_group_begin(comm);
    if (comm_root < 0) comm_allreduce_sum(comm, dense.p, dense.p, nd(), HLAC_COMM_U32, stream);
    else comm_reduce_sum(comm, dense.p, dense.p, nd(), HLAC_COMM_U32, comm_root, stream);
    comm_allreduce_sum(comm, counters.p, counters_merged.p, N_COUNTERS, HLAC_COMM_U64, stream);
    comm_group_end(comm);
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
  if (const char* e = getenv("HLAC_SCAN_BUDGET")) s->scan_budget = std::max(1, atoi(e));
  s->stride = seed_stride();
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
  s->counters.reserve(N_COUNTERS, s->stream);
  HLAC_CUDA(cudaMemsetAsync(s->counters.p, 0, N_COUNTERS * sizeof(unsigned long long), s->stream));
  s->dense.reserve(s->nd(), s->stream);
  cudaDeviceProp prop; HLAC_CUDA(cudaGetDeviceProperties(&prop, s->device));
  s->n_sm = prop.multiProcessorCount;
  s->bitmap_smem_bytes = ((size_t)cds->view.bitmap_words + genomic->view.bitmap_words) * 4;
  s->smem_bm = s->bitmap_smem_bytes <= 96 * 1024;
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  *out = s.release();
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_free(hlac_count* s) {
  if (s) { int prev = 0; cudaGetDevice(&prev); cudaSetDevice(s->device); delete s; cudaSetDevice(prev); }   // the caller's current device is left as it was
  return HLAC_OK;
}
uint32_t hlac_count_nrows(const hlac_count* s) { return s ? s->rows.nrows : 0; }
const char* hlac_count_row_name(const hlac_count* s, uint32_t i) { return (s && i < s->rows.rownames.size()) ? s->rows.rownames[i].c_str() : nullptr; }

int hlac_count_reserve(hlac_count* s, uint64_t total_reads) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  s->ensure_capacity(total_reads);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_attach_comm(hlac_count* s, hlac_comm* comm, int root) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  if (comm) {
    if (comm_device(comm) != s->device) throw Error(HLAC_ERR_ARG, "communicator and session live on different devices");
    if (root >= comm_nranks(comm)) throw Error(HLAC_ERR_ARG, "root is not a rank of the communicator");
  }
  s->comm = comm; s->comm_root = comm ? root : -1;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

static void validate_batch(const hlac_batch* b) {
  if (b->n && (!b->seq || !b->len || !b->cell || !b->umi || !b->flags)) throw Error(HLAC_ERR_ARG, "batch has a null array");
  if (b->n >= (1ull << 32)) throw Error(HLAC_ERR_LIMIT, "batch too large");
}

int hlac_count_push_device(hlac_count* s, const hlac_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_batch(b);
  DeviceGuard g(s->device);
  s->launch_map(*b);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// staging set for the next host batch: waits (on the copy stream) for the kernel that read this set two pushes ago
static hlac_count::Stage& next_stage(hlac_count* s) {
  auto& st = s->stage[s->n_push++ & 1];
  if (!s->copy_stream) HLAC_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  if (!st.copied) { HLAC_CUDA(cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming)); HLAC_CUDA(cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming)); }
  else HLAC_CUDA(cudaStreamWaitEvent(s->copy_stream, st.consumed, 0));
  return st;
}

int hlac_count_push(hlac_count* s, const hlac_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_batch(b);
  if (b->n == 0) return HLAC_OK;
  DeviceGuard g(s->device);
  auto& st = next_stage(s);
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
  auto& st = next_stage(s);
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

// ---- wire records: the bit-packed host->device format (hlac_wire_batch) ------------------------------------------------
static void validate_wire(const hlac_count* s, const hlac_wire_batch* b) {
  if (b->n && !b->rec) throw Error(HLAC_ERR_ARG, "null record array");
  if (b->n >= (1ull << 32)) throw Error(HLAC_ERR_LIMIT, "batch too large");
  if (b->read_len == 0 || b->read_len > 512) throw Error(HLAC_ERR_ARG, "read_len must be 1..512");
  if (b->cell_bits == 0 || b->cell_bits > 27 || b->umi_bits == 0 || b->umi_bits > 32) throw Error(HLAC_ERR_ARG, "cell_bits must be 1..27 and umi_bits 1..32");
  if (b->words_per_record != hlac_wire_words(b->read_len, b->cell_bits, b->umi_bits)) throw Error(HLAC_ERR_ARG, "words_per_record does not match the field widths");
  if ((uint64_t)s->n_cells > (1ull << b->cell_bits) - 1) throw Error(HLAC_ERR_LIMIT, "cell_bits too narrow for this session's columns (the all-ones value is 'not a cell')");
}

uint32_t hlac_wire_words(uint32_t read_len, uint32_t cell_bits, uint32_t umi_bits) { return (2 * read_len + 1 + cell_bits + umi_bits + 31) / 32; }

int hlac_count_push_wire_device(hlac_count* s, const hlac_wire_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_wire(s, b);
  DeviceGuard g(s->device);
  hlac_batch d{}; d.n = b->n; d.words_per_read = (b->read_len + 31) / 32;
  s->launch_map(d, nullptr, b);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_count_push_wire(hlac_count* s, const hlac_wire_batch* b) try {
  if (!s || !b) throw Error(HLAC_ERR_ARG, "null argument");
  validate_wire(s, b);
  if (b->n == 0) return HLAC_OK;
  DeviceGuard g(s->device);
  auto& st = next_stage(s);
  auto& e = s->begin(3, s->copy_stream);
  st.wire.upload(b->rec, b->n * b->words_per_record, s->copy_stream);   // ONE copy per push, 28 bytes per 91-bp read
  s->end(e, s->copy_stream);
  HLAC_CUDA(cudaEventRecord(st.copied, s->copy_stream));
  HLAC_CUDA(cudaStreamWaitEvent(s->stream, st.copied, 0));
  hlac_batch d{}; d.n = b->n; d.words_per_read = (b->read_len + 31) / 32;
  hlac_wire_batch w = *b; w.rec = st.wire.p;
  s->launch_map(d, nullptr, &w);
  HLAC_CUDA(cudaEventRecord(st.consumed, s->stream));
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// host helper: hlac_batch (host pointers, every read read_len bases) -> wire records
int hlac_pack_wire(const hlac_batch* b, uint32_t read_len, uint32_t cell_bits, uint32_t umi_bits, uint32_t* rec) try {
  if (!b || (b->n && (!rec || !b->seq || !b->len || !b->cell || !b->umi || !b->flags))) throw Error(HLAC_ERR_ARG, "null argument");
  if (read_len == 0 || read_len > 32 * b->words_per_read || cell_bits == 0 || cell_bits > 27 || umi_bits == 0 || umi_bits > 32) throw Error(HLAC_ERR_ARG, "bad field widths");
  const uint32_t W = hlac_wire_words(read_len, cell_bits, umi_bits), wpr = b->words_per_read;
  const uint32_t not_cell = (1u << cell_bits) - 1u;
  for (uint64_t i = 0; i < b->n; i++) {
    if (b->len[i] != read_len) throw Error(HLAC_ERR_LIMIT, "wire records hold reads of one length per batch (read " + std::to_string(i) + " has " + std::to_string(b->len[i]) + " bases)");
    uint32_t cell = b->cell[i];
    if (cell == HLAC_NOT_A_CELL) cell = not_cell;
    else if (cell >= not_cell) throw Error(HLAC_ERR_LIMIT, "cell index does not fit cell_bits");
    if (umi_bits < 32 && (b->umi[i] >> umi_bits)) throw Error(HLAC_ERR_LIMIT, "UMI key does not fit umi_bits");
    uint32_t* r = rec + i * W;
    for (uint32_t w = 0; w < W; w++) r[w] = 0;
    const uint32_t nsw = (2 * read_len + 31) / 32;
    for (uint32_t w = 0; w < nsw; w++) { const uint64_t v = b->seq[i * wpr + (w >> 1)]; r[w] = (w & 1) ? (uint32_t)v : (uint32_t)(v >> 32); }
    if ((2 * read_len) & 31) r[nsw - 1] &= ~(0xFFFFFFFFu >> ((2 * read_len) & 31));   // bits beyond the read stay free for the fields
    auto put = [&](uint32_t pos, uint32_t nbits, uint32_t v) {   // big-endian bit stream
      for (uint32_t k = 0; k < nbits; k++) if ((v >> (nbits - 1 - k)) & 1u) r[(pos + k) >> 5] |= 0x80000000u >> ((pos + k) & 31);
    };
    put(2 * read_len, 1, (b->flags[i] & HLAC_FLAG_PRIMARY) ? 1u : 0u);
    put(2 * read_len + 1, cell_bits, cell);
    put(2 * read_len + 1 + cell_bits, umi_bits, b->umi[i]);
  }
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

// count() of this session's reads into its dense matrix. Returns with the matrix complete on the session stream (one
// synchronisation, to learn whether a bucket overflowed), so that a caller that shards reads by barcode itself can run its
// own all-reduce on it before hlac_count_entries. hlac_count_finish is the fused form (and merges by itself when a
// communicator is attached).
int hlac_count_reduce(hlac_count* s, uint32_t** dense_device, uint64_t* n_elems) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  HLAC_CUDA(cudaMemsetAsync(s->counters.p, 0, 8, s->stream));
  s->enqueue_group();
  unsigned long long ovf = 0;
  HLAC_CUDA(cudaMemcpyAsync(&ovf, s->counters.p + 11, 8, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  if (ovf) s->fallback_group();
  s->reduced = true;
  if (dense_device) *dense_device = s->dense.p;
  if (n_elems) *n_elems = s->nd();
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// results out of the host-mapped header + COO buffers (after the stream has been synchronised)
static void publish(hlac_count* s, hlac_coo* out, hlac_metrics* metrics, bool with_entries) {
  const unsigned long long* ctr = s->host_hdr + 1;
  const uint64_t nnz = with_entries ? s->host_hdr[0] : 0;
  if (ctr[9]) throw Error(HLAC_ERR_ARG, std::to_string(ctr[9]) + " reads carried a cell index >= n_cells (the reference panics in TriMat::add_triplet, main.rs:279)");
  if (nnz > s->coo_cap) throw Error(HLAC_ERR_STATE, "COO buffer overran (internal error)");
  if (metrics) {
    metrics->num_reads = ctr[1] + ctr[13]; metrics->num_non_primary = ctr[2] + ctr[14]; metrics->num_not_cell_bc = ctr[3] + ctr[15];
    metrics->num_cds_align = ctr[4]; metrics->num_gen_align = ctr[5]; metrics->num_not_aligned = ctr[6];
  }
  if (out) {
    // the arrays live in the session's pinned buffer: valid until the next hlac_count_entries / _finish / _free on this
    // session; hlac_coo_free does not free borrowed arrays
    out->n = nnz; out->nrows = s->rows.nrows; out->borrowed = 1;
    out->row = nnz ? s->coo_host : nullptr; out->col = nnz ? s->coo_host + s->coo_cap : nullptr; out->val = nnz ? s->coo_host + 2 * s->coo_cap : nullptr;
  }
}

int hlac_count_entries(hlac_count* s, hlac_coo* out, hlac_metrics* metrics) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  if (!s->reduced) throw Error(HLAC_ERR_STATE, "hlac_count_reduce has not run since the last push");
  DeviceGuard g(s->device);
  { hlac_comm* keep = s->comm; s->comm = nullptr; s->enqueue_merge(); s->enqueue_coo(true); s->comm = keep; }   // the caller merged (or not) itself
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  publish(s, out, metrics, true);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// Reads whose barcode is not in the whitelist are never scored (mapping.rs:760-763): a host driver may drop them before the
// device and report only how they count in the metrics.
int hlac_count_add_filtered(hlac_count* s, uint64_t n_reads, uint64_t n_non_primary, uint64_t n_not_cell) {
  if (!s) { set_last_error("null session"); return HLAC_ERR_ARG; }
  s->host_reads += n_reads; s->host_non_primary += n_non_primary; s->host_not_cell += n_not_cell;
  return HLAC_OK;
}

// group + consensus -> (merge over the communicator) -> COO + counters into host-mapped memory: everything is enqueued
// back to back and the host synchronises once. Only if some bucket overflowed (reported in the same header) is the step
// redone through the sort path.
int hlac_count_finish(hlac_count* s, hlac_coo* out, hlac_metrics* metrics) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  const bool extract = !s->comm || s->comm_root < 0 || comm_rank(s->comm) == s->comm_root;
  HLAC_CUDA(cudaMemsetAsync(s->counters.p, 0, 8, s->stream));
  s->enqueue_group();
  for (int attempt = 0;; attempt++) {
    s->enqueue_merge();
    s->enqueue_coo();
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    if (s->host_hdr[1 + 11] == 0 || attempt == 1) break;   // [11]: overflowed keys, summed over the ranks when merged
    // some rank overflowed a bucket: every rank recomputes its own matrix (the in-place all-reduce consumed it), the
    // overflowing ranks through the sort path
    unsigned long long ovf = 0;
    HLAC_CUDA(cudaMemcpyAsync(&ovf, s->counters.p + 11, 8, cudaMemcpyDeviceToHost, s->stream));
    HLAC_CUDA(cudaStreamSynchronize(s->stream));
    if (ovf) s->fallback_group();
    else { HLAC_CUDA(cudaMemsetAsync(s->counters.p, 0, 8, s->stream)); s->enqueue_group(); }
  }
  s->reduced = true;
  publish(s, extract ? out : nullptr, metrics, extract);
  if (!extract && out) { out->n = 0; out->nrows = s->rows.nrows; out->borrowed = 1; out->row = out->col = out->val = nullptr; }
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

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

// the launch shape the map kernel last ran with, as text: "smem_bitmaps=1 threads=1024 slots=64 idx_smem=1 tma=0 smem_bytes=..."
int hlac_count_shape(hlac_count* s, char* out, size_t cap) try {
  if (!s || !out || cap == 0) throw Error(HLAC_ERR_ARG, "null argument");
  const hlac_count::Shape* sh = (const hlac_count::Shape*)s->cfg_shape;
  char buf[256];
  if (!sh) snprintf(buf, sizeof buf, "none (nothing pushed yet)");
  else snprintf(buf, sizeof buf, "smem_bitmaps=%d threads=%d slots=%d idx_smem=%d tma=%d smem_bytes=%zu blocks_per_sm=%d stride=%d", (int)sh->bm, sh->threads, sh->ns,
                (int)sh->idx, (int)s->last_tma, s->last_smem, s->cfg_occ, s->stride);
  snprintf(out, cap, "%s", buf);
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

// grouped keys (= scored reads) of the last finish / reduce, buckets in use, and how many finishes took the sort path
int hlac_count_group_stats(hlac_count* s, uint64_t* n_keys, uint32_t* n_buckets, uint64_t* fallbacks) try {
  if (!s) throw Error(HLAC_ERR_ARG, "null session");
  DeviceGuard g(s->device);
  unsigned long long c = 0;
  HLAC_CUDA(cudaMemcpyAsync(&c, s->counters.p, sizeof c, cudaMemcpyDeviceToHost, s->stream));
  HLAC_CUDA(cudaStreamSynchronize(s->stream));
  if (n_keys) *n_keys = c;
  if (n_buckets) *n_buckets = s->n_buckets;
  if (fallbacks) *fallbacks = s->fallbacks;
  return HLAC_OK;
} catch (const Error& e) { set_last_error(e.what()); return e.code; } catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

}  // extern "C"
