#include "hla_fasta.hpp"

#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>
#include <ostream>
#include <sstream>
#include <tuple>

namespace hlac {

bool Allele::operator<(const Allele& o) const {
  return std::tie(gene, f1, f2, f3, f4, name) < std::tie(o.gene, o.f1, o.f2, o.f3, o.f4, o.name);
}

// "^[A-Z0-9]+[*][0-9]+(:[0-9]+)*[A-Z]?$" (src/hla.rs:49), fields from "[0-9]+(:[0-9]+)*" (src/hla.rs:50)
bool parse_allele(const std::string& s, Allele& out) {
  auto up = [](char c) { return c >= 'A' && c <= 'Z'; };
  auto dg = [](char c) { return c >= '0' && c <= '9'; };
  size_t star = s.find('*');
  if (star == std::string::npos || star == 0) return false;
  for (size_t i = 0; i < star; i++) if (!up(s[i]) && !dg(s[i])) return false;
  size_t p = star + 1;
  std::vector<unsigned long> fields;
  for (;;) {
    size_t q = p;
    while (q < s.size() && dg(s[q])) q++;
    if (q == p) return false;
    fields.push_back(std::stoul(s.substr(p, q - p)));
    p = q;
    if (p < s.size() && s[p] == ':') { p++; continue; }
    break;
  }
  if (p < s.size() && up(s[p])) p++;
  if (p != s.size()) return false;
  out = Allele();
  out.gene = s.substr(0, star);
  out.f1 = (uint16_t)fields[0];
  if (fields.size() > 1) out.f2 = (int32_t)fields[1];
  if (fields.size() > 2) out.f3 = (int32_t)fields[2];
  if (fields.size() > 3) out.f4 = (int32_t)fields[3];
  out.name = s;
  return true;
}

std::vector<FastaRecord> read_fasta(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw Error(HLAC_ERR_IO, "cannot open FASTA " + path);
  std::vector<FastaRecord> recs;
  std::string line;
  while (std::getline(f, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '>') {
      FastaRecord r;
      size_t sp = line.find(' ');
      if (sp == std::string::npos) r.id = line.substr(1);
      else { r.id = line.substr(1, sp - 1); r.desc = line.substr(sp + 1); }
      recs.push_back(std::move(r));
    } else {
      if (recs.empty()) throw Error(HLAC_ERR_FORMAT, "FASTA " + path + ": sequence before first header");
      recs.back().seq += line;
    }
  }
  return recs;
}

void write_fasta_record(std::ostream& o, const FastaRecord& r) {
  o << '>' << r.id;
  if (!r.desc.empty()) o << ' ' << r.desc;
  o << '\n' << r.seq << '\n';
}

void read_hla_cds(const std::string& fasta, const std::set<std::string>& allele_set, bool use_filter,
                  std::vector<Bases>& seqs, std::vector<std::string>& tx_ids) {
  static const std::set<std::string> kGenes = {"A", "B", "C", "DRB1", "DPA1", "DPB1", "DQA1", "DQB1"};  // hla.rs:108-117
  struct Entry { Allele al; std::string tx_id, allele_str; Bases seq; };
  std::vector<Entry> hlas;
  for (auto& rec : read_fasta(fasta)) {
    if (rec.desc.empty()) throw Error(HLAC_ERR_FORMAT, "no HLA allele");
    std::string allele_str = rec.desc.substr(0, rec.desc.find(' '));
    if (use_filter && !allele_set.count(allele_str)) continue;
    Allele al;
    if (!parse_allele(allele_str, al)) throw Error(HLAC_ERR_FORMAT, "invalid allele string: " + allele_str);
    if (!kGenes.count(al.gene)) continue;
    Entry e{al, rec.id, allele_str, Bases(rec.seq.size())};
    for (size_t i = 0; i < rec.seq.size(); i++) e.seq[i] = base_code(rec.seq[i]);
    hlas.push_back(std::move(e));
  }
  std::sort(hlas.begin(), hlas.end(), [](const Entry& a, const Entry& b) {
    if (a.al < b.al) return true;
    if (b.al < a.al) return false;
    return std::tie(a.tx_id, a.allele_str, a.seq) < std::tie(b.tx_id, b.allele_str, b.seq);
  });
  std::map<std::tuple<std::string, uint16_t, int32_t>, size_t> two_field_len;
  if (!use_filter)
    for (auto& e : hlas) {
      size_t& m = two_field_len[std::make_tuple(e.al.gene, e.al.f1, e.al.f2)];
      m = std::max(m, e.seq.size());
    }
  size_t i = 0;
  while (i < hlas.size()) {
    size_t j = i + 1;
    while (j < hlas.size() && std::tie(hlas[j].al.gene, hlas[j].al.f1, hlas[j].al.f2, hlas[j].al.f3) ==
                                  std::tie(hlas[i].al.gene, hlas[i].al.f1, hlas[i].al.f2, hlas[i].al.f3)) j++;
    size_t rep = i;   // stable sort by length, then pop(): the last of the longest
    for (size_t q = i + 1; q < j; q++) if (hlas[q].seq.size() >= hlas[rep].seq.size()) rep = q;
    const Entry& e = hlas[rep];
    if (use_filter || e.seq.size() >= two_field_len[std::make_tuple(e.al.gene, e.al.f1, e.al.f2)]) {
      seqs.push_back(e.seq);
      tx_ids.push_back(e.allele_str);
    }
    i = j;
  }
}

std::set<std::string> load_allele_status(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw Error(HLAC_ERR_IO, "cannot open allele status file " + path);
  std::set<std::string> keep;
  std::string line;
  bool saw_header = false;
  while (std::getline(f, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    if (!saw_header) { saw_header = true; continue; }
    std::vector<std::string> col;
    size_t p = 0;
    for (;;) { size_t c = line.find(',', p); col.push_back(line.substr(p, c == std::string::npos ? c : c - p)); if (c == std::string::npos) break; p = c + 1; }
    if (col.size() < 8) throw Error(HLAC_ERR_FORMAT, "allele status file: short row");
    if (col[0].find('N') == std::string::npos && col[3] == "Confirmed" && col[6] == "Full" && col[7] == "gDNA") keep.insert(col[0]);
  }
  return keep;
}

RowLayout build_row_layout(const std::vector<std::string>& tx_names) {
  struct A { std::string gene, name; uint32_t tx; };
  std::vector<A> al;
  for (uint32_t i = 0; i < tx_names.size(); i++) { Allele a; if (parse_allele(tx_names[i], a)) al.push_back({a.gene, a.name, i}); }
  std::stable_sort(al.begin(), al.end(), [](const A& x, const A& y) { return x.gene < y.gene; });
  RowLayout r;
  for (size_t i = 0; i < al.size();) {
    size_t j = i; while (j < al.size() && al[j].gene == al[i].gene) j++;
    r.gene_row0.push_back(r.nrows);
    r.a1.push_back(al[i].tx);
    r.rownames.push_back(al[i].name);
    if (j - i >= 2) {   // two alleles: three rows; a third and later allele of the gene is ignored (mapping.rs:689-696)
      r.a2.push_back(al[i + 1].tx);
      r.rownames.push_back(al[i + 1].name);
      r.rownames.push_back(al[i].gene);
      r.nrows += 3;
    } else { r.a2.push_back(NONE32); r.nrows += 1; }
    r.n_genes++;
    i = j;
  }
  return r;
}

}  // namespace hlac
