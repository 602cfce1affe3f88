// HostIndex: .idx loader (bincode layout), sort-based build_index, derived GPU structures.
#include "host_index.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <unordered_map>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace hlac {

void HostIndex::append_bases(const uint8_t* b, size_t n) {
  for (size_t i = 0; i < n; i++) {
    uint64_t g = n_bases++;
    if ((g >> 4) >= seq32.size()) seq32.push_back(0);
    seq32[g >> 4] |= uint32_t(b[i] & 3) << (30 - 2 * (g & 15));
  }
}

namespace {

struct Reader {
  const uint8_t* p; size_t n; size_t o = 0;
  void need(size_t k) { if (o + k > n) throw Error(HLAC_ERR_FORMAT, "index file truncated"); }
  uint64_t u64() { need(8); uint64_t v; memcpy(&v, p + o, 8); o += 8; return v; }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, p + o, 4); o += 4; return v; }
  uint8_t u8() { need(1); return p[o++]; }
  void skip(uint64_t k) { need(k); o += k; }
  uint64_t count(uint64_t elem) { uint64_t c = u64(); if (c > (n - o) / (elem ? elem : 1) + 1) throw Error(HLAC_ERR_FORMAT, "index file: implausible length"); return c; }
};

// boomphf::Mphf<T> { bitvecs: Vec<BitVector{bits:u64, vector:Vec<u64>}>, ranks: Vec<Vec<u64>> } — not needed:
// every MPHF hit is verified against the node k-mer, so any exact dictionary is equivalent.
void skip_mphf(Reader& r) {
  uint64_t nb = r.count(16);
  for (uint64_t i = 0; i < nb; i++) { r.u64(); uint64_t w = r.count(8); r.skip(w * 8); }
  uint64_t nr = r.count(8);
  for (uint64_t i = 0; i < nr; i++) { uint64_t w = r.count(8); r.skip(w * 8); }
}

}  // namespace

void HostIndex::load_idx(const std::string& path, HostIndex& ix) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(HLAC_ERR_IO, "cannot open index file " + path);
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  Reader r{buf.data(), buf.size()};
  ix = HostIndex();
  uint64_t nw = r.count(8);
  ix.seq32.resize(nw * 2);
  for (uint64_t i = 0; i < nw; i++) { uint64_t w = r.u64(); ix.seq32[2 * i] = uint32_t(w >> 32); ix.seq32[2 * i + 1] = uint32_t(w); }
  ix.n_bases = r.u64();
  if (ix.n_bases > nw * 32 || ix.n_bases >= (1ULL << 32)) throw Error(HLAC_ERR_FORMAT, "index file: bad base count");
  uint64_t ns = r.count(8);
  ix.start.resize(ns);
  for (auto& v : ix.start) { uint64_t s = r.u64(); if (s > ix.n_bases) throw Error(HLAC_ERR_FORMAT, "index file: bad node start"); v = (uint32_t)s; }
  uint64_t nl = r.count(4); ix.len.resize(nl); for (auto& v : ix.len) v = r.u32();
  uint64_t ne = r.count(1); ix.exts.resize(ne); for (auto& v : ix.exts) v = r.u8();
  uint64_t nd = r.count(4); ix.colour.resize(nd); for (auto& v : ix.colour) v = r.u32();
  uint8_t stranded = r.u8();
  for (int pass = 0; pass < 2; pass++) {   // left_order, right_order : BoomHashMap<Kmer24,u32>
    skip_mphf(r);
    uint64_t nk = r.count(8); r.skip(nk * 8);
    uint64_t nv = r.count(4); r.skip(nv * 4);
  }
  uint64_t nc = r.count(8);
  ix.col_off.assign(1, 0);
  for (uint64_t c = 0; c < nc; c++) {
    uint64_t m = r.count(4);
    for (uint64_t i = 0; i < m; i++) ix.col_tx.push_back(r.u32());
    ix.col_off.push_back(ix.col_tx.size());
  }
  skip_mphf(r);                            // dbg_index : NoKeyBoomHashMap<Kmer24,(u32,u32)>
  uint64_t nv = r.count(8); r.skip(nv * 8);
  uint64_t nt = r.count(8);
  for (uint64_t i = 0; i < nt; i++) { uint64_t m = r.count(1); r.need(m); ix.tx_names.emplace_back((const char*)r.p + r.o, m); r.o += m; }
  uint64_t nm = r.count(16);               // tx_gene_mapping : HashMap<String,String>
  for (uint64_t i = 0; i < nm; i++) { uint64_t a = r.count(1); r.skip(a); uint64_t b = r.count(1); r.skip(b); }
  if (r.o != r.n) throw Error(HLAC_ERR_FORMAT, "index file: trailing bytes");
  if (stranded != 1) throw Error(HLAC_ERR_FORMAT, "index file: not a stranded graph");
  if (ns != nl || ns != ne || ns != nd) throw Error(HLAC_ERR_FORMAT, "index file: per-node arrays disagree");
  for (uint64_t i = 0; i < ns; i++) {
    if (ix.len[i] < (uint32_t)K || (uint64_t)ix.start[i] + ix.len[i] > ix.n_bases) throw Error(HLAC_ERR_FORMAT, "index file: bad node extent");
    if (ix.colour[i] >= nc) throw Error(HLAC_ERR_FORMAT, "index file: bad colour id");
  }
  for (uint32_t t : ix.col_tx) if (t >= nt) throw Error(HLAC_ERR_FORMAT, "index file: bad tx id");
  ix.finalize();
}

// build_index::<Kmer24> [EXT] — stranded k=24 coloured compacted de Bruijn graph: a node is a
// maximal path whose interior links have a unique right extension, a unique left extension on the successor, and
// one colour. Sort-based: one record per k-mer occurrence, sorted by (k-mer, tx), then grouped.
void HostIndex::build(const std::vector<Bases>& seqs, const std::vector<std::string>& names, HostIndex& ix, PairSorter sorter) {
  const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr; auto t_last = std::chrono::steady_clock::now();
  auto tick = [&](const char* what) { if (dbg) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[hlac timing] build: %-24s %.3f s\n", what, std::chrono::duration<double>(t - t_last).count()); t_last = t; } };
  ix = HostIndex();
  ix.tx_names = names;
  // one (k-mer, payload) pair per occurrence; payload = tx << 16 | left-ext mask << 8 | right-ext mask
  std::vector<uint64_t> okey, oval;
  size_t total = 0;
  for (auto& s : seqs) if (s.size() >= (size_t)K) total += s.size() - K + 1;
  okey.reserve(total); oval.reserve(total);
  for (uint32_t t = 0; t < seqs.size(); t++) {
    const Bases& s = seqs[t];
    if (s.size() < (size_t)K) continue;
    uint64_t v = 0;
    for (int i = 0; i < K - 1; i++) v = (v << 2) | s[i];
    for (size_t i = 0; i + K <= s.size(); i++) {
      v = ((v << 2) | s[i + K - 1]) & KMER_MASK;
      okey.push_back(v);
      oval.push_back(((uint64_t)t << 16) | (uint64_t)(i > 0 ? 1u << s[i - 1] : 0) << 8 | (uint64_t)(i + K < s.size() ? 1u << s[i + K] : 0));
    }
  }
  tick("enumerate k-mers");
  // stable sort by k-mer: occurrences were generated in ascending tx order, so stability leaves every k-mer's tx list
  // sorted. On a device-resident build the caller supplies a GPU radix sort; otherwise an LSD radix sort on the host.
  if (sorter && total >= (1u << 16)) sorter(okey.data(), oval.data(), total);
  else {
    std::vector<uint64_t> tk(total), tv(total);
    std::vector<size_t> cnt(1 << 16);
    for (int pass = 0; pass < 3; pass++) {
      const int sh = 16 * pass;
      std::fill(cnt.begin(), cnt.end(), 0);
      for (size_t i = 0; i < total; i++) cnt[(okey[i] >> sh) & 0xFFFF]++;
      size_t run = 0;
      for (size_t& c : cnt) { const size_t k = c; c = run; run += k; }
      for (size_t i = 0; i < total; i++) { const size_t d = cnt[(okey[i] >> sh) & 0xFFFF]++; tk[d] = okey[i]; tv[d] = oval[i]; }
      okey.swap(tk); oval.swap(tv);
    }
  }
  tick("sort occurrences");

  // unique k-mers with ext masks and interned colour
  std::vector<uint64_t> uk; std::vector<uint8_t> ul, ur; std::vector<uint32_t> ucol;
  std::unordered_map<uint64_t, std::vector<uint32_t>> by_hash;   // hash(tx list) -> candidate colour ids
  ix.col_off.assign(1, 0);
  std::vector<uint32_t> txl;
  for (size_t i = 0; i < total;) {
    size_t j = i; uint8_t l = 0, r = 0; txl.clear();
    uint64_t h = 0xcbf29ce484222325ULL;
    while (j < total && okey[j] == okey[i]) {
      const uint32_t tx = (uint32_t)(oval[j] >> 16);
      l |= (uint8_t)(oval[j] >> 8); r |= (uint8_t)oval[j];
      if (txl.empty() || txl.back() != tx) { txl.push_back(tx); h = (h ^ tx) * 0x100000001b3ULL; h ^= h >> 29; }
      j++;
    }
    uint32_t cid = NONE32;
    auto& cand = by_hash[h];
    for (uint32_t c : cand) {
      size_t n = ix.col_off[c + 1] - ix.col_off[c];
      if (n == txl.size() && std::equal(txl.begin(), txl.end(), ix.col_tx.begin() + ix.col_off[c])) { cid = c; break; }
    }
    if (cid == NONE32) {
      cid = (uint32_t)ix.col_off.size() - 1;
      ix.col_tx.insert(ix.col_tx.end(), txl.begin(), txl.end());
      ix.col_off.push_back(ix.col_tx.size());
      cand.push_back(cid);
    }
    uk.push_back(okey[i]); ul.push_back(l); ur.push_back(r); ucol.push_back(cid);
    i = j;
  }
  std::vector<uint64_t>().swap(okey); std::vector<uint64_t>().swap(oval);
  tick("group + intern colours");

  const size_t n = uk.size();
  auto find = [&](uint64_t kmer) -> int64_t {
    auto it = std::lower_bound(uk.begin(), uk.end(), kmer);
    return (it != uk.end() && *it == kmer) ? (int64_t)(it - uk.begin()) : -1;
  };
  auto one = [](uint8_t m) { return m != 0 && (m & (m - 1)) == 0; };
  auto bit = [](uint8_t m) { return (uint64_t)__builtin_ctz(m); };
  // link[a] = b if a->b is an interior link of a node, else -1
  std::vector<int64_t> link(n, -1); std::vector<uint8_t> linked_into(n, 0);
  for (size_t a = 0; a < n; a++) {
    if (!one(ur[a])) continue;
    int64_t b = find(((uk[a] << 2) | bit(ur[a])) & KMER_MASK);
    if (b < 0 || (size_t)b == a || !one(ul[b]) || ucol[b] != ucol[a]) continue;
    link[a] = b; linked_into[b] = 1;
  }
  std::vector<uint8_t> done(n, 0);
  std::vector<uint8_t> tmp;
  auto emit = [&](size_t a) {
    tmp.clear();
    for (int i = K - 1; i >= 0; i--) tmp.push_back((uk[a] >> (2 * i)) & 3);
    done[a] = 1;
    size_t cur = a;
    while (link[cur] >= 0 && !done[link[cur]]) { cur = link[cur]; done[cur] = 1; tmp.push_back(uk[cur] & 3); }
    if (ix.n_bases + tmp.size() >= (1ULL << 32)) throw Error(HLAC_ERR_LIMIT, "index exceeds 2^32 bases");
    ix.start.push_back((uint32_t)ix.n_bases);
    ix.len.push_back((uint32_t)tmp.size());
    ix.exts.push_back(uint8_t((ur[cur] << 4) | ul[a]));
    ix.colour.push_back(ucol[a]);
    ix.append_bases(tmp.data(), tmp.size());
  };
  for (size_t a = 0; a < n; a++) if (!linked_into[a]) emit(a);
  for (size_t a = 0; a < n; a++) if (!done[a]) emit(a);   // closed cycles of interior links
  tick("compact nodes");
  ix.finalize();
  tick("finalize (adj/dict/bitmap)");
}

void HostIndex::finalize() {
  const uint32_t nn = n_nodes();
  seq32.resize((n_bases + 15) / 16 + 4, 0);   // padding so that 3-word window reads never run off the end
  // ---- adjacency: resolve each ext bit to the neighbouring node once ----
  std::vector<std::pair<uint64_t, uint32_t>> firsts(nn), lasts(nn);
  n_kmers = 0;
  for (uint32_t i = 0; i < nn; i++) {
    firsts[i] = {kmer_at(start[i]), i};
    lasts[i] = {kmer_at((uint64_t)start[i] + len[i] - K), i};
    n_kmers += len[i] - K + 1;
  }
  std::sort(firsts.begin(), firsts.end()); std::sort(lasts.begin(), lasts.end());
  auto lookup = [](const std::vector<std::pair<uint64_t, uint32_t>>& v, uint64_t k) -> uint32_t {
    auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(k, 0u));
    return (it != v.end() && it->first == k) ? it->second : NONE32;
  };
  radj.assign((size_t)nn * 4, NONE32); ladj.assign((size_t)nn * 4, NONE32);
  for (uint32_t i = 0; i < nn; i++) {
    uint64_t fk = kmer_at(start[i]), lk = kmer_at((uint64_t)start[i] + len[i] - K);
    for (uint64_t b = 0; b < 4; b++) {
      if ((exts[i] >> 4) >> b & 1) radj[(size_t)i * 4 + b] = lookup(firsts, ((lk << 2) | b) & KMER_MASK);
      if ((exts[i] & 0xF) >> b & 1) ladj[(size_t)i * 4 + b] = lookup(lasts, (fk >> 2) | (b << (2 * K - 2)));
    }
  }
  // ---- exact k-mer dictionary: every (node, offset) ----
  dict_bits = 4;
  while ((1ULL << dict_bits) < n_kmers * 2 + 2) dict_bits++;
  if (dict_bits > 31) throw Error(HLAC_ERR_LIMIT, "too many k-mers for the dictionary");
  const uint32_t mask = (1u << dict_bits) - 1;
  dict.assign((size_t)4 << dict_bits, NONE32);
  for (uint32_t i = 0; i < nn; i++) {
    uint64_t v = kmer_at(start[i]);
    for (uint32_t o = 0; o + K <= len[i]; o++) {
      if (o) v = ((v << 2) | base((uint64_t)start[i] + o + K - 1)) & KMER_MASK;
      uint32_t h = dict_hash(v, dict_bits);
      while (dict[(size_t)h * 4 + 1] != NONE32) h = (h + 1) & mask;
      dict[(size_t)h * 4] = (uint32_t)v; dict[(size_t)h * 4 + 1] = (uint32_t)(v >> 32);
      dict[(size_t)h * 4 + 2] = i; dict[(size_t)h * 4 + 3] = o;
    }
  }
  // ---- l-mer presence bitmap for the skip-scan seed filter: smallest l in 8..14 with density <= 1/16 ----
  for (lmer = 8;; lmer++) {
    const uint64_t bits = 1ULL << (2 * lmer);
    bitmap.assign(bits / 32, 0);
    uint64_t distinct = 0;
    const uint32_t lm = (uint32_t)(bits - 1);
    for (uint32_t i = 0; i < nn; i++) {
      uint32_t v = 0;
      for (uint32_t o = 0; o < len[i]; o++) {
        v = ((v << 2) | base((uint64_t)start[i] + o)) & lm;
        if (o + 1 >= lmer) { uint32_t& w = bitmap[v >> 5]; uint32_t b = 1u << (v & 31); if (!(w & b)) { w |= b; distinct++; } }
      }
    }
    if (distinct * 16 <= bits || lmer == 14) break;
  }
}

}  // namespace hlac
