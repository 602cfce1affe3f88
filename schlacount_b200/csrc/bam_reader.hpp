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

  // ---- the same selection in two halves, so that the expensive half can run on many threads -------------------------
  // frame(): find up to max_records raw records of the selection (fixed-field filters only: reference id, position
  // beyond the region end). Whole inflated blocks are appended to the chunk's buffer and the record boundaries are walked
  // in place: recs[i] = {pointer to record i's body (after its block_size word), size}, pointing into chunk.buf. The bytes
  // of a record cut by the end of the chunk are carried over to the next call. Sequential and cheap (one memcpy per 64 KB
  // block); uses the reader's state, so one frame() at a time - but it may run on its own thread while other threads
  // parse() the previous chunk. Returns the number of records found; 0 at the end of the selection.
  struct RawRec { const uint8_t* p; size_t n; };
  struct Chunk { std::vector<uint8_t> buf; std::vector<RawRec> recs; };
  size_t frame(Chunk& chunk, size_t max_records);
  // parse(): overlap test (CIGAR) and the CB / UB aux scan of one framed record. const and thread-safe. false = the
  // reference drops the record (no overlap with the region, or CB / UB missing: mapping.rs:254-270).
  struct RecView {
    const uint8_t* seq4 = nullptr;   // BAM 4-bit bases, two per byte
    int32_t l_seq = 0;
    const char* cb = nullptr; uint32_t cb_len = 0;
    const char* ub = nullptr; uint32_t ub_len = 0;
    bool primary = true;
  };
  bool parse(const uint8_t* rec, size_t size, RecView& out) const;
  // 4-bit BAM bases -> 2-bit words (32 bases per u64, first base in the top bits; anything but ACGT -> A), nwords zero padded
  static void pack_2bit(const uint8_t* seq4, int32_t l_seq, uint64_t* words, uint32_t nwords);

 private:
  bool read_block();                              // next inflated BGZF block into block_ (blocks are inflated ahead, in parallel)
  struct Pending { uint64_t coff, bsize; std::future<std::vector<uint8_t>> data; };
  bool enqueue_block();                           // read one raw block at read_coff_ and hand it to the pool
  std::deque<Pending> pending_;
  std::unique_ptr<InflatePool> pool_;
  uint64_t read_coff_ = 0;                        // file offset of the next raw block to read
  bool raw_eof_ = false;
  uint64_t file_pos_ = ~0ull;                     // where the FILE* stands (~0: unknown)
  bool read_bytes(void* dst, size_t n);
  void seek_virtual(uint64_t voff);
  FILE* f_ = nullptr;
  std::vector<uint8_t> block_, cbuf_, rec_, carry_;   // carry_: inflated bytes framed past the last chunk's final record
  Chunk one_;                                     // next(): records framed ahead
  size_t one_pos_ = 0;
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
