// The genotype caller (src/em.rs:361-434) as host logic shared by hlac_call_genotypes (explicit tables) and
// hlac_geno_call (session with the class vectors still on the device). Product code.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "common.hpp"

namespace hlac {

struct Ranked {
  struct W { uint32_t i; double w; uint32_t gene; uint64_t exp; };   // gene = rank of the gene string (lexicographic)
  std::vector<W> wn;                 // sorted by descending weight, then (stably) by gene (em.rs:370-371)
  std::vector<int32_t> cand_of;      // tx -> index into cands, or -1
  std::vector<uint32_t> cands;       // per gene the leading alleles with reads_explained > MIN_READS_CALL (<= WEIGHTS_TO_OUTPUT)
};

inline Ranked rank_alleles(const std::vector<std::string>& names, uint32_t nitems, const uint64_t* reads_explained, const double* theta) {
  Ranked r;
  // gene strings -> ranks in string order (a handful of genes), so that the two stable sorts of em.rs:370-371 (by descending
  // weight, then by gene string) become one stable sort on (gene rank, descending weight) over plain numbers
  std::vector<std::string> gene_of(nitems);
  std::set<std::string> gs;
  for (uint32_t i = 0; i < nitems; i++) { gene_of[i] = names[i].substr(0, names[i].find('*')); gs.insert(gene_of[i]); }
  const std::vector<std::string> sorted_genes(gs.begin(), gs.end());
  r.wn.reserve(nitems);
  for (uint32_t i = 0; i < nitems; i++)
    r.wn.push_back({i, theta[i], (uint32_t)(std::lower_bound(sorted_genes.begin(), sorted_genes.end(), gene_of[i]) - sorted_genes.begin()), reads_explained[i]});
  std::stable_sort(r.wn.begin(), r.wn.end(), [](const Ranked::W& a, const Ranked::W& b) { return a.gene != b.gene ? a.gene < b.gene : -a.w < -b.w; });
  r.cand_of.assign(nitems, -1);
  for (size_t s = 0; s < r.wn.size();) {
    size_t e = s; while (e < r.wn.size() && r.wn[e].gene == r.wn[s].gene) e++;
    size_t k = 0;
    for (size_t q = s; q < e && r.wn[q].exp > MIN_READS_CALL && k < WEIGHTS_TO_OUTPUT; q++, k++) { r.cand_of[r.wn[q].i] = (int32_t)r.cands.size(); r.cands.push_back(r.wn[q].i); }
    s = e;
  }
  return r;
}

// em.rs:373-434 given pair_reads_explained(a, b) for candidate alleles
template <typename PairFn>
inline void call_from_ranked(const Ranked& rk, uint32_t nreads_u32, PairFn pair_explained, hlac_calls* out) {
  const auto& wn = rk.wn;
  const double nreads = (double)nreads_u32;
  std::set<uint32_t> called;
  std::vector<uint32_t> wtx, pa, pb; std::vector<double> wem, wfr, pfr;
  for (size_t s = 0; s < wn.size();) {
    size_t e = s; while (e < wn.size() && wn[e].gene == wn[s].gene) e++;
    size_t written = 0;
    for (size_t q = s; q < e; q++) {
      if (wn[q].exp <= MIN_READS_CALL) break;
      wtx.push_back(wn[q].i); wem.push_back(wn[q].w); wfr.push_back((double)(uint32_t)wn[q].exp / nreads);
      if (++written == WEIGHTS_TO_OUTPUT) break;
    }
    if (written) {
      struct P { uint32_t a, b; double joint, single; };
      std::vector<P> pairs; bool stop = false;
      for (size_t i = 0; i + 1 < written && !stop; i++)
        for (size_t j = i + 1; j < written; j++) {
          const uint32_t ex = pair_explained(wn[s + i].i, wn[s + j].i);
          if (ex > MIN_READS_CALL) pairs.push_back({wn[s + i].i, wn[s + j].i, (double)ex / nreads, (double)(uint32_t)std::max(wn[s + i].exp, wn[s + j].exp) / nreads});
          else { stop = true; break; }
        }
      if (!pairs.empty()) {
        std::stable_sort(pairs.begin(), pairs.end(), [](const P& a, const P& b) { return -a.joint < -b.joint; });
        for (size_t q = 0; q < std::min<size_t>(PAIRS_TO_OUTPUT, pairs.size()); q++) { pa.push_back(pairs[q].a); pb.push_back(pairs[q].b); pfr.push_back(pairs[q].joint); }
        if (pairs[0].joint * HOMOZYGOUS_TH > pairs[0].joint - pairs[0].single) { called.insert(pairs[0].a); called.insert(pairs[0].b); }
        else called.insert(wn[s].i);
      } else called.insert(wn[s].i);
    }
    s = e;
  }
  memset(out, 0, sizeof *out);
  auto dup32 = [](const std::vector<uint32_t>& v) { uint32_t* p = (uint32_t*)malloc(std::max<size_t>(v.size(), 1) * 4); if (!v.empty()) memcpy(p, v.data(), v.size() * 4); return p; };
  auto dupd = [](const std::vector<double>& v) { double* p = (double*)malloc(std::max<size_t>(v.size(), 1) * 8); if (!v.empty()) memcpy(p, v.data(), v.size() * 8); return p; };
  std::vector<uint32_t> cv(called.begin(), called.end());
  out->n_called = (uint32_t)cv.size(); out->called = dup32(cv);
  out->n_weights = (uint32_t)wtx.size(); out->w_tx = dup32(wtx); out->w_em = dupd(wem); out->w_frac = dupd(wfr);
  out->n_pairs = (uint32_t)pa.size(); out->p_a = dup32(pa); out->p_b = dup32(pb); out->p_frac = dupd(pfr);
}

}  // namespace hlac
