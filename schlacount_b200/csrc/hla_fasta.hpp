// HLA FASTA / allele model on the host (src/hla.rs). Defines the tx-id numbering every class depends on. Product code.
#pragma once
#include <set>
#include <string>
#include <vector>

#include "common.hpp"
#include "host_index.hpp"

namespace hlac {

struct Allele {           // src/hla.rs:23-31; optional fields: -1 = None (None < Some in the derived Ord)
  std::string gene;
  uint16_t f1 = 0;
  int32_t f2 = -1, f3 = -1, f4 = -1;
  std::string name;
  bool operator<(const Allele& o) const;
};
bool parse_allele(const std::string& s, Allele& out);   // src/hla.rs:59-92

struct FastaRecord { std::string id, desc, seq; };
std::vector<FastaRecord> read_fasta(const std::string& path);
void write_fasta_record(std::ostream& o, const FastaRecord& r);   // bio::io::fasta::Writer::write_record layout

// src/hla.rs:102-209. use_filter=false: longest 3-field representative that reaches its 2-field group's max length.
void read_hla_cds(const std::string& fasta, const std::set<std::string>& allele_set, bool use_filter,
                  std::vector<Bases>& seqs, std::vector<std::string>& tx_ids);
// src/hla.rs:236-250
std::set<std::string> load_allele_status(const std::string& path);

// src/mapping.rs:668-728: matrix rows and the class -> (gene, slot) table, in the compact form the kernels use.
struct RowLayout {
  std::vector<std::string> rownames;      // labels.tsv lines
  uint32_t nrows = 0;
  uint32_t n_genes = 0;
  std::vector<uint32_t> gene_row0;        // genes_to_rownums
  std::vector<uint32_t> a1, a2;           // per gene: tx id of the first / second allele (a2 = NONE32 if single)
};
RowLayout build_row_layout(const std::vector<std::string>& tx_names);

}  // namespace hlac
