// Device-side view of a HostIndex and the per-read pseudoalignment walk (the `map_read` kernel body).
//
// Replaces [EXT] debruijn_mapping::pseudoaligner::Pseudoaligner::map_read (call sites src/mapping.rs:341,351,383,
// 765-770; algorithm: SURVEY.md Appendix A). Design (B200-first, not a translation):
//   * the MPHF + verify lookup becomes an exact open-addressing dictionary of 16-byte slots (one LDG.128 per probe);
//   * the linear seed scan becomes a skip-scan: an l-mer presence bitmap (shared memory for small indexes, L2 for
//     large ones) proves "no dictionary k-mer covers this l-mer", which lets the scan jump K-l+1 positions at once.
//     It is only a necessary-condition filter; every candidate is still verified in the exact dictionary, so the
//     first hit position is identical to the reference's position-by-position scan (the scan is stateless);
//   * graph edges are a precomputed adjacency table instead of hashing the extended terminal k-mer;
//   * base-by-base extension becomes 16-base XOR/popcount chunks with exact "stop at the 3rd mismatch" semantics,
//     in ONE flat loop (compare a chunk, then hop / re-seed if the node is exhausted) so that the lanes of a warp,
//     each on its own read, do not serialise on nested loops with different trip counts.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "common.hpp"

namespace hlac {

// Seed-scan stride (template parameter S of find_seed / walk_from_seed / map_read_walk): the pinned debruijn_mapping
// rev is not vendored and the reference's goldens reproduce for strides 1..3 (SURVEY.md finding #4), so the stride is a
// property of the session (hlac_set_seed_stride, default DEFAULT_SEED_STRIDE) and the kernels are instantiated per
// stride. Candidate positions of one scan are start, start+S, start+2S, ...

struct DevIndexView {
  const uint32_t* seq32;     // node sequences, 16 bases per word
  const uint4* nodes;        // 3 per node: {start, len, exts, colour} {radj[4]} {ladj[4]}
  const uint4* dict;         // {kmer lo, kmer hi, node, offset}; hi == NONE32: empty
  const uint32_t* bitmap;    // l-mer presence bits (global copy)
  uint32_t dict_mask;
  uint32_t dict_bits;
  uint32_t lmer;
  uint32_t bitmap_words;
  uint32_t n_nodes;
  uint32_t seq_words;        // u32 words of seq32 (incl. the padding the 16-base extraction may read)
};

// (w[i] : w[i+1]) << 2*(pos%16), high word: the 16 bases starting at `pos`
template <typename P>
__device__ __forceinline__ uint32_t ext16(P w, int pos) {
  const int i = pos >> 4;
  return __funnelshift_l(w[i + 1], w[i], (pos & 15) * 2);
}
template <typename P>
__device__ __forceinline__ uint64_t kmer48(P w, int pos) {
  const int i = pos >> 4, s = (pos & 15) * 2;
  const uint32_t a = w[i], b = w[i + 1], c = w[i + 2];
  const uint32_t hi = __funnelshift_l(b, a, s), lo = __funnelshift_l(c, b, s);
  return (((uint64_t)hi << 32) | lo) >> (64 - 2 * K);
}
template <typename P>
__device__ __forceinline__ uint32_t base_at(P w, int pos) {
  return (w[pos >> 4] >> (30 - 2 * (pos & 15))) & 3u;
}

// Shared memory through explicit 32-bit shared-window addresses (ld.shared / st.shared). With ordinary pointers into an
// `extern __shared__` array and the register file full (64 registers x 1024 threads), ptxas re-materialises the window base
// (S2R SR_CgaCtaId + MOV + LEA) at every use - inside the seed-scan loop that was 3 of 28 instructions per l-mer test and
// an S2R latency on the critical path of each iteration (SASS of the r01 kernel). An address computed once with
// __cvta_generic_to_shared stays a plain register.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
struct SmemWords {   // a row of u32 words at shared address a
  uint32_t a;
  __device__ __forceinline__ uint32_t operator[](int i) const { return lds32(a + 4u * (uint32_t)i); }
};
struct SmemBitmap {
  uint32_t a;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return lds32(a + 4u * i); }
};

// read-only global words through the non-coherent path
struct GWords {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator[](int i) const { return __ldg(p + i); }
};

// Where a walk reads node records and node sequences from: the read-only global path (L1 / L2), or - for the small
// indexes of known-genotype counting - a copy staged in shared memory next to the l-mer bitmaps.
struct GlobalGraph {
  const uint4* nodes; const uint32_t* seq32;
  __device__ __forceinline__ uint4 n0(uint32_t node) const { return __ldg(nodes + 3 * (size_t)node); }
  __device__ __forceinline__ uint4 ra(uint32_t node) const { return __ldg(nodes + 3 * (size_t)node + 1); }
  __device__ __forceinline__ uint4 la(uint32_t node) const { return __ldg(nodes + 3 * (size_t)node + 2); }
  __device__ __forceinline__ GWords seq() const { return GWords{seq32}; }
};
struct SmemGraph {
  uint32_t a_nodes, a_seq;
  __device__ __forceinline__ uint4 n0(uint32_t node) const { return lds128(a_nodes + 48u * node); }
  __device__ __forceinline__ uint4 ra(uint32_t node) const { return lds128(a_nodes + 48u * node + 16u); }
  __device__ __forceinline__ uint4 la(uint32_t node) const { return lds128(a_nodes + 48u * node + 32u); }
  __device__ __forceinline__ SmemWords seq() const { return SmemWords{a_seq}; }
};

__device__ __forceinline__ bool dict_get(const DevIndexView& ix, uint64_t kmer, uint32_t& node, uint32_t& off) {
  uint32_t h = (uint32_t)((kmer * 0x9E3779B97F4A7C15ULL) >> (64 - ix.dict_bits));
  const uint32_t lo = (uint32_t)kmer, hi = (uint32_t)(kmer >> 32);
  for (;;) {
    const uint4 s = __ldg(ix.dict + h);
    if (s.y == NONE32) return false;
    if (s.x == lo && s.y == hi) { node = s.z; off = s.w; return true; }
    h = (h + 1) & ix.dict_mask;
  }
}

// Seed scan: first pos' >= pos (pos' <= last) whose k-mer is in the dictionary. BM(idx) returns one bitmap word.
// A window [p, p+K) can only match if every l-mer inside it is present in the index; l-mers are tested right to left
// and an absent one at offset o rules out every window covering it: the scan jumps o+1 positions.
// (A variant that tested four jump targets speculatively per round was measured slower on B200: 15.6 instead of 9.6
// bitmap tests per read and more registers; so was a warp-synchronous loop whose condition is a vote — every iteration
// then waits for the few lanes that probe the dictionary; see profiles/r01_count_map_tuning.md.)
// With a stride S > 1 only every S-th position (counted from the scan start) is a candidate, so a jump lands on the next
// candidate at or beyond the first position the absent l-mer does not rule out.
// find_seed_step is the resumable form: (pos, q) is the scan state (q = read position of the l-mer under test, K - l beyond
// pos when a window is entered), `budget` bounds the l-mer tests of this call. Returns SEED_FOUND, SEED_NONE (scan exhausted) or
// SEED_MORE (budget used up; call again with the same pos, q).
enum { SEED_NONE = 0, SEED_FOUND = 1, SEED_MORE = 2 };
template <int S, bool BOUNDED, typename RD, typename BM>
__device__ __forceinline__ int find_seed_step(const DevIndexView& ix, RD rd, BM bm, int& pos, int& q, int last, int budget, uint32_t& node,
                                              uint32_t& off, uint32_t& n_tests, uint32_t& n_probes) {
  const int l = (int)ix.lmer;
  const int sh = 32 - 2 * l;
  const int o0 = K - l;
  // one flat loop, one l-mer test per iteration (no nested per-window loop: lanes of a warp advance independently)
  while (pos <= last) {
    if (BOUNDED && budget-- <= 0) return SEED_MORE;
    const uint32_t v = ext16(rd, q) >> sh;
    n_tests++;
    if (!((bm(v >> 5) >> (v & 31)) & 1u)) {   // no k-mer can cover this l-mer: every window up to q is ruled out
      pos = S == 1 ? q + 1 : pos + ((q - pos + S) / S) * S;
      q = pos + o0;
      continue;
    }
    if (q != pos) { q = max(q - l, pos); continue; }
    n_probes++;
    if (dict_get(ix, kmer48(rd, pos), node, off)) return SEED_FOUND;
    pos += S;
    q = pos + o0;
  }
  return SEED_NONE;
}
template <int S, typename RD, typename BM>
__device__ __forceinline__ bool find_seed(const DevIndexView& ix, RD rd, BM bm, int& pos, int last, uint32_t& node,
                                          uint32_t& off, uint32_t& n_tests, uint32_t& n_probes) {
  int q = pos + K - (int)ix.lmer;   // right to left inside the window [pos, pos + K)
  return find_seed_step<S, false>(ix, rd, bm, pos, q, last, 0, node, off, n_tests, n_probes) == SEED_FOUND;
}

// (r02, measured and removed: the same scan as a warp-collective call - one straight-line, predicated l-mer test per
// iteration, the loop condition a REDUX vote, the lanes whose window passed leaving together to probe the dictionary together.
// It fixes what the ncu capture shows here - the flat loop's lanes only reconverge at the loop exit, so the loop runs at ~15
// active lanes and the probe at 3 - but every iteration then costs ~45 instructions for the whole warp instead of ~25 for the
// lanes that take it, and the round still lasts as long as its slowest lane: 1.15 ms per 10 M reads against 1.025.)

// Left extension: only when the first seed sits at or beyond floor(0.2*L). Walks the read
// leftwards from the seed through predecessor nodes; returns the coverage it adds and visits the nodes it enters.
template <typename G, typename RD, typename V>
__device__ __forceinline__ int left_extend(G gr, RD rd, int L, int pos, uint32_t node, uint32_t off, const uint4& n0, V& vis) {
  int cov = 0;
  if (pos >= (L * LEFT_EXTEND_NUM) / LEFT_EXTEND_DEN) {
    int lp = pos - 1;
    uint32_t pn = node;
    uint4 p0 = n0;
    int po = off > 0 ? (int)off - 1 : 0;
    const auto ns = gr.seq();
    for (;;) {
      const int m = min(lp + 1, po + 1);
      int matched = 0, snp = 0;
      bool brk = false;
      // read bases lp, lp-1, ... against node bases po, po-1, ...: <= 16 at a time, XOR + popcount, exact "stop at the third
      // mismatch" (tolerated mismatches count as covered, the breaking one does not) - the same arithmetic as the forward
      // walk, mirrored: the base nearest the seed sits in the low bits
      while (matched < m) {
        const int c = min(16, m - matched);
        const uint32_t wr = ext16(rd, lp - matched - c + 1) >> (32 - 2 * c);
        const uint32_t wn = ext16(ns, (int)p0.x + po - matched - c + 1) >> (32 - 2 * c);
        uint32_t x = wr ^ wn;
        x = (x | (x >> 1)) & 0x55555555u;
        const int pc = __popc(x);
        if (snp + pc > ALLOWED_MISMATCH) {
          for (int t = snp; t < ALLOWED_MISMATCH; t++) x &= x - 1;   // drop the mismatches still tolerated
          matched += (__ffs((int)x) - 1) >> 1;                        // bases before the breaking one
          brk = true;
          break;
        }
        snp += pc; matched += c;
      }
      cov += matched;
      if (lp + 1 - matched == 0 || brk) break;
      lp -= matched;
      const uint32_t b = base_at(rd, lp);
      const uint4 la = gr.la(pn);
      const uint32_t nxt = b == 0 ? la.x : b == 1 ? la.y : b == 2 ? la.z : la.w;
      if (nxt == NONE32) break;
      pn = nxt;
      p0 = gr.n0(pn);
      po = (int)p0.y - K;
      vis.visit(p0.w, pn);
    }
  }
  return cov;
}

// map_read_to_nodes after the first seed: optional left extension, then the forward walk.
// (pos, node, off) is the first verified seed. V::visit(colour, node) is called for every visited node in the
// reference's push order (left-extension nodes first, then the forward walk); the last call is the class-defining node.
template <int S, typename G, typename RD, typename BM, typename V>
__device__ __forceinline__ uint32_t walk_from_seed(const DevIndexView& ix, G gr, RD rd, BM bm, int L, int pos, uint32_t node, uint32_t off,
                                                   V& vis, uint32_t& n_tests, uint32_t& n_probes) {
  const int last = L - K;
  int cov = 0;
  uint4 n0 = gr.n0(node);
  uint4 ra = gr.ra(node);   // right-edge table, fetched with the node so its latency hides behind the compare
  cov += left_extend(gr, rd, L, pos, node, off, n0, vis);
  // ---- forward walk, flattened: one loop whose body compares <= 16 bases inside the current node and then, if the node
  // is exhausted (or the per-node mismatch budget is), takes the edge or re-seeds. Equivalent to the reference's
  // node loop with an inner base loop: a hop is `pos -= K-1; cov -= K-1` followed by `pos += K; cov += K` = +1/+1,
  // the per-node budget `snp` resets at every node entry, tolerated mismatches count as covered, the breaking one
  // does not. ----
  const auto ns = gr.seq();
  pos += K;
  cov += K;
  vis.visit(n0.w, node);
  uint32_t g = n0.x + off + K;          // next node base to compare
  int rem = (int)n0.y - (int)off - K;   // bases left in the node
  int snp = 0;
  for (;;) {
    const int c = min(min(16, L - pos), rem);
    bool brk = false;
    if (c > 0) {
      uint32_t x = ext16(rd, pos) ^ ext16(ns, (int)g);
      x = (x | (x >> 1)) & 0x55555555u;
      if (c < 16) x &= ~(0xFFFFFFFFu >> (2 * c));
      const int pc = __popc(x);
      if (snp + pc > ALLOWED_MISMATCH) {
        for (int t = snp; t < ALLOWED_MISMATCH; t++) x &= ~(0x80000000u >> __clz(x));
        const int adv = __clz(x) >> 1;
        pos += adv; cov += adv;
        brk = true;
      } else {
        snp += pc; pos += c; cov += c; g += c; rem -= c;
      }
    }
    if (pos >= L) break;
    if (!brk && rem > 0) continue;
    uint32_t nxt = NONE32;
    if (!brk) {
      const uint32_t b = base_at(rd, pos);
      nxt = b == 0 ? ra.x : b == 1 ? ra.y : b == 2 ? ra.z : ra.w;
    }
    if (nxt != NONE32) {                 // exact base at the node boundary selects the edge
      node = nxt;
      n0 = gr.n0(node);
      ra = gr.ra(node);
      pos += 1; cov += 1;
      g = n0.x + K; rem = (int)n0.y - K;
    } else {                             // re-seed scanning forward from the current position
      if (pos > last) break;
      if (!find_seed<S>(ix, rd, bm, pos, last, node, off, n_tests, n_probes)) break;
      n0 = gr.n0(node);
      ra = gr.ra(node);
      pos += K; cov += K;
      g = n0.x + off + K; rem = (int)n0.y - (int)off - K;
    }
    snp = 0;
    vis.visit(n0.w, node);
  }
  return (uint32_t)cov;
}

// map_read_to_nodes, whole: None iff the first seed scan finds nothing (or L < K).
template <int S, typename RD, typename BM, typename V>
__device__ __forceinline__ bool map_read_walk(const DevIndexView& ix, RD rd, BM bm, int L, V& vis, uint32_t& cov_out,
                                              uint32_t& n_tests, uint32_t& n_probes) {
  int pos = 0;
  uint32_t node = 0, off = 0;
  if (L < K || !find_seed<S>(ix, rd, bm, pos, L - K, node, off, n_tests, n_probes)) return false;
  cov_out = walk_from_seed<S>(ix, GlobalGraph{ix.nodes, ix.seq32}, rd, bm, L, pos, node, off, vis, n_tests, n_probes);
  return true;
}

// 2-bit reverse complement of 16 packed bases
__device__ __forceinline__ uint32_t revcomp16(uint32_t v) {
  uint32_t r = __brev(~v);
  return ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
}

// Fill the reverse-complement row from the forward row (both 16 bases per word, nw words, zero padded).
// bio::alphabets::dna::revcomp over the packed read (mapping.rs:351-353, 767-768): N was already packed as A.
template <typename P>
__device__ __forceinline__ void make_revcomp(P fwd, uint32_t* rc, int L, int nw) {
  for (int j = 0; j < nw; j++) {
    const int st = L - 16 * (j + 1);
    uint32_t out = 0;
    if (st >= 0) out = revcomp16(ext16(fwd, st));
    else if (st > -16) {
      const int nb = 16 + st;                                  // valid bases in this last partial word
      out = revcomp16(ext16(fwd, 0) >> (2 * (16 - nb))) & ~(0xFFFFFFFFu >> (2 * nb));
    }
    rc[j] = out;
  }
}

// the same, writing the row through its shared-window address
template <typename P>
__device__ __forceinline__ void make_revcomp_s(P fwd, uint32_t rc_addr, int L, int nw) {
  for (int j = 0; j < nw; j++) {
    const int st = L - 16 * (j + 1);
    uint32_t out = 0;
    if (st >= 0) out = revcomp16(ext16(fwd, st));
    else if (st > -16) {
      const int nb = 16 + st;
      out = revcomp16(ext16(fwd, 0) >> (2 * (16 - nb))) & ~(0xFFFFFFFFu >> (2 * nb));
    }
    sts32(rc_addr + 4u * (uint32_t)j, out);
  }
}

}  // namespace hlac
