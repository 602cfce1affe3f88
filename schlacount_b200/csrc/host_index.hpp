// Host-side form of a Pseudoaligner<Kmer24> (mapping.rs:318; [EXT] debruijn_mapping::pseudoaligner) plus the
// derived lookup structures the GPU kernels use. Product code.
#pragma once
#include <string>
#include <vector>

#include "common.hpp"

namespace hlac {

using Bases = std::vector<uint8_t>;  // 2-bit codes, one per byte (host staging only)

struct HostIndex {
  // ---- the graph, as serialised in the .idx ----
  std::vector<uint32_t> seq32;   // node sequences concatenated, 16 bases per u32, base j at bits 31-2(j%16)..30-2(j%16)
  uint64_t n_bases = 0;
  std::vector<uint32_t> start;   // per node: first base in seq32
  std::vector<uint32_t> len;     // per node: bases (>= K)
  std::vector<uint8_t> exts;     // per node: high nibble = right ext mask, low nibble = left ext mask
  std::vector<uint32_t> colour;  // per node: colour (eq-class) id
  std::vector<uint64_t> col_off; // colours as CSR of sorted tx ids
  std::vector<uint32_t> col_tx;
  std::vector<std::string> tx_names;

  // ---- derived (finalize) ----
  std::vector<uint32_t> radj, ladj;  // 4 per node: successor / predecessor node by base, NONE32 if no edge
  std::vector<uint32_t> dict;        // open addressing, 4 u32 per slot {kmer lo, kmer hi, node, offset}; hi==NONE32 empty
  uint32_t dict_bits = 0;
  uint64_t n_kmers = 0;
  std::vector<uint32_t> bitmap;      // presence bitmap over l-mers occurring anywhere in a node
  uint32_t lmer = 0;

  uint32_t n_nodes() const { return (uint32_t)len.size(); }
  uint32_t n_colours() const { return col_off.empty() ? 0 : (uint32_t)col_off.size() - 1; }
  inline uint8_t base(uint64_t g) const { return (seq32[g >> 4] >> (30 - 2 * (g & 15))) & 3; }
  uint64_t kmer_at(uint64_t g) const { uint64_t v = 0; for (int i = 0; i < K; i++) v = (v << 2) | base(g + i); return v; }

  void append_bases(const uint8_t* b, size_t n);  // append to seq32 / n_bases
  void finalize();                                // adjacency, dictionary, l-mer bitmap

  static void load_idx(const std::string& path, HostIndex& out);   // bincode (Appendix B)
  // the same file, written: utils::write_obj(&index, hla_index) (hla.rs:260). Nodes, colours and the three boomphf
  // tables are emitted in the reference's deterministic order, so the fixture's own inputs reproduce the fixture byte
  // for byte (tests/test_host_cpu.py).
  void write_idx(const std::string& path) const;
  // stable sort of (key, value) pairs by key; build() uses it for the k-mer occurrences when given (a device sort)
  using PairSorter = void (*)(uint64_t* keys, uint64_t* vals, size_t n);
  static void build(const std::vector<Bases>& seqs, const std::vector<std::string>& names, HostIndex& out, PairSorter sorter = nullptr);  // build_index
};

// build_index as kernels on `device` (device_build.cu): fills `out` exactly as HostIndex::build would (same node and colour
// numbering). Returns false when it declines (tiny input, closed cycles of interior links): the caller then builds on the host.
// `keep` (optional) receives the device buffers of the parts that hlac_index::upload would otherwise copy back up (the
// dictionary is then not downloaded at all: nothing on the host reads it).
struct DeviceIndexParts;
bool device_build_index(const std::vector<Bases>& seqs, const std::vector<std::string>& names, HostIndex& out, int device, DeviceIndexParts* keep = nullptr);

// multiplicative hash shared by host build and device probe
inline uint32_t dict_hash(uint64_t kmer, uint32_t bits) { return (uint32_t)((kmer * 0x9E3779B97F4A7C15ULL) >> (64 - bits)); }

}  // namespace hlac
