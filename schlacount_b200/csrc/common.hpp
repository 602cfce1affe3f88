// Shared host-side declarations for the hlacount library (product code; never includes anything from oracle/).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hlacount.h"

namespace hlac {

constexpr int K = HLAC_K;                     // Kmer24 (pinned by test/fake_db/hla_nuc.fasta.idx)
constexpr uint64_t KMER_MASK = (1ULL << (2 * K)) - 1;
// map_read constants [EXT debruijn_mapping @ae4aadbc; SURVEY.md Appendix A]. Unpinned ones are named here.
constexpr int ALLOWED_MISMATCH = 2;           // README.md:108
constexpr int DEFAULT_SEED_STRIDE = 1;        // unpinned by the goldens (1..3 reproduce both); hlac_set_seed_stride selects 1 or 3
constexpr int LEFT_EXTEND_NUM = 1, LEFT_EXTEND_DEN = 5;  // floor(0.2 * L)
// src/config.rs:4-21
constexpr uint32_t MIN_SCORE_CALL = 40;
constexpr uint32_t MIN_SCORE_COUNT_PSEUDO = 60;
constexpr double GENE_CONSENSUS_THRESHOLD = 0.5;
constexpr double ALLELE_CONSENSUS_THRESHOLD = 0.1;
constexpr uint32_t EM_ITERS = 2000;
constexpr double EM_REL_TH = 5e-4;
constexpr double EM_ABS_TH = 5e-3;
constexpr double EM_CARE_TH = 1e-5;
constexpr uint32_t MIN_READS_CALL = 100;
constexpr double HOMOZYGOUS_TH = 0.15;
constexpr uint32_t PAIRS_TO_OUTPUT = 5;
constexpr uint32_t WEIGHTS_TO_OUTPUT = 10;

constexpr uint32_t NONE32 = 0xFFFFFFFFu;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);
int seed_stride();   // the process-wide seed-scan stride new sessions adopt (hlac_set_seed_stride / HLAC_SEED_STRIDE)

// DnaString::from_acgt_bytes [EXT]: ACGT either case -> 0..3, everything else -> A
inline uint8_t base_code(char c) {
  switch (c) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
  }
}

}  // namespace hlac
