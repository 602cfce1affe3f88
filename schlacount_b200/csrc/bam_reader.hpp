// Minimal BGZF/BAM/BAI reader for the host driver (replaces BamSeqReader over rust-htslib, src/mapping.rs:39-62,231-279).
// Stays on the host by design (north-star): decode, region selection, CB/UB extraction. Product code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <deque>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "common.hpp"

namespace hlac {

struct BamRecord {          // what BamCrRead needs (mapping.rs:64-75)
  std::string barcode, umi; // CB / UB aux strings
  std::vector<uint8_t> seq; // 2-bit codes (from_acgt_bytes: anything but ACGT -> A)
  bool primary = true;      // !(secondary || supplementary) (mapping.rs:247-249)
};

class InflatePool;   // worker threads that inflate BGZF blocks ahead of the record parser

class BamReader {
 public:
  explicit BamReader(const std::string& path);   // opens path and path + ".bai"
  ~BamReader();
  // IndexedReader::fetch((chrom, start, end)) (mapping.rs:52-57): records overlapping [start, end) on `chrom`
  void fetch(const std::string& chrom, uint64_t start, uint64_t end);
  void fetch_unmapped();                          // fetch_str(b"*") (mapping.rs:59-61, 378-379)
  // next record carrying both CB and UB (mapping.rs:254-270 skips the others). false at end of the selection.
  bool next(BamRecord& out);
  const std::vector<std::string>& ref_names() const { return ref_names_; }

 private:
  bool read_block();                              // next inflated BGZF block into block_ (blocks are inflated ahead, in parallel)
  struct Pending { uint64_t coff, bsize; std::future<std::vector<uint8_t>> data; };
  bool enqueue_block();                           // read one raw block at read_coff_ and hand it to the pool
  std::deque<Pending> pending_;
  std::unique_ptr<InflatePool> pool_;
  uint64_t read_coff_ = 0;                        // file offset of the next raw block to read
  bool raw_eof_ = false;
  bool read_bytes(void* dst, size_t n);
  void seek_virtual(uint64_t voff);
  FILE* f_ = nullptr;
  std::vector<uint8_t> block_, cbuf_, rec_;
  size_t block_pos_ = 0;
  uint64_t first_record_voff_ = 0, block_coff_ = 0, next_coff_ = 0;
  std::vector<std::string> ref_names_;
  std::vector<std::vector<uint64_t>> linear_;     // per reference: 16 kb linear index
  uint64_t unmapped_voff_ = 0;                    // a virtual offset at or before the first tid == -1 record
  int sel_tid_ = -2;                              // -2: nothing selected, -1: unmapped
  int64_t sel_beg_ = 0, sel_end_ = 0;
  bool done_ = true;
};

}  // namespace hlac
