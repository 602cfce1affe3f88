// HostIndex: .idx loader (bincode layout), sort-based build_index, derived GPU structures.
#include "host_index.hpp"
#include "worker_pool.hpp"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <fstream>
#include <memory>
#include <thread>
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
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw Error(HLAC_ERR_IO, "cannot open index file " + path);
  const std::streamoff fsize = f.tellg();
  if (fsize < 0) throw Error(HLAC_ERR_IO, "cannot size index file " + path);
  std::vector<uint8_t> buf((size_t)fsize);
  f.seekg(0);
  if (fsize && !f.read((char*)buf.data(), fsize)) throw Error(HLAC_ERR_IO, "cannot read index file " + path);
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
  for (uint64_t c = 0; c < nc; c++) {   // each list is contiguous little-endian u32: one bulk copy
    const uint64_t m = r.count(4);
    r.need(m * 4);
    const size_t at = ix.col_tx.size();
    ix.col_tx.resize(at + m);
    if (m) memcpy(ix.col_tx.data() + at, r.p + r.o, m * 4);
    r.o += m * 4;
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

// ---- .idx writer -----------------------------------------------------------------------------------------------------
// [EXT] boomphf 0.5.2 (Cargo.lock:286-297) Mphf::new(gamma = 1.7, keys), restated from its published algorithm and
// pinned on the committed fixture test/fake_db/hla_nuc.fasta.idx (all three tables reproduce bit for bit):
//   level i: size = max(255, floor(1.7 * #keys left)); idx(key) = fastmod(fold32(wyhash(key, seed = 1 << 2i)), size);
//   a bit is set where exactly one key lands, keys that collide go to the next level; ranks = cumulative popcount
//   before every 8th word, running across levels. The key is hashed as its u64 storage (derive(Hash) -> write_u64):
//   wyhash 0.3.0 for 8 bytes = wymum(wymum(seed ^ P0, swap32(x) ^ P1), 8 ^ P5).
namespace {

struct Mphf {
  struct Level { uint64_t bits; std::vector<uint64_t> words; std::vector<uint64_t> ranks; };
  std::vector<Level> levels;
  static uint64_t wymum(uint64_t a, uint64_t b) { const unsigned __int128 r = (unsigned __int128)a * b; return (uint64_t)(r >> 64) ^ (uint64_t)r; }
  static uint64_t slot(uint64_t key, uint32_t level, uint64_t size) {
    const uint64_t x = (key << 32) | (key >> 32);
    const uint64_t h = wymum(wymum(((uint64_t)1 << (2 * level)) ^ 0xa0761d6478bd642fULL, x ^ 0xe7037ed1a0b428dbULL), 8 ^ 0xeb44accab455d165ULL);
    if (size < (1ULL << 32)) return ((uint64_t)((uint32_t)h ^ (uint32_t)(h >> 32)) * size) >> 32;
    return h % size;
  }
  explicit Mphf(std::vector<uint64_t> keys) {
    uint64_t pop = 0;
    for (uint32_t level = 0;; level++) {
      if (level > 31) throw Error(HLAC_ERR_LIMIT, "boomphf construction did not converge (duplicate keys?)");
      Level lv;
      lv.bits = std::max<uint64_t>(255, (uint64_t)(1.7 * (double)keys.size()));
      const size_t nw = (lv.bits + 63) / 64;
      lv.words.assign(nw, 0);
      std::vector<uint64_t> collide(nw, 0);
      for (uint64_t k : keys) {
        const uint64_t i = slot(k, level, lv.bits), m = 1ULL << (i & 63);
        if (collide[i >> 6] & m) continue;
        if (lv.words[i >> 6] & m) collide[i >> 6] |= m; else lv.words[i >> 6] |= m;
      }
      std::vector<uint64_t> redo;
      for (uint64_t k : keys) {
        const uint64_t i = slot(k, level, lv.bits), m = 1ULL << (i & 63);
        if (collide[i >> 6] & m) { redo.push_back(k); lv.words[i >> 6] &= ~m; }
      }
      for (size_t w = 0; w < nw; w++) { if (w % 8 == 0) lv.ranks.push_back(pop); pop += (uint64_t)__builtin_popcountll(lv.words[w]); }
      levels.push_back(std::move(lv));
      if (redo.empty()) break;
      keys.swap(redo);
    }
  }
  uint64_t hash(uint64_t key) const {
    for (uint32_t l = 0; l < levels.size(); l++) {
      const Level& lv = levels[l];
      const uint64_t i = slot(key, l, lv.bits);
      if (!((lv.words[i >> 6] >> (i & 63)) & 1)) continue;
      uint64_t r = lv.ranks[i >> 9];
      for (uint64_t w = (i >> 9) * 8; w < (i >> 6); w++) r += (uint64_t)__builtin_popcountll(lv.words[w]);
      return r + (uint64_t)__builtin_popcountll(lv.words[i >> 6] & ((1ULL << (i & 63)) - 1));
    }
    throw Error(HLAC_ERR_STATE, "boomphf: key not in the set");
  }
};

struct Writer {
  std::string o;
  void u8(uint8_t v) { o.push_back((char)v); }
  void u32(uint32_t v) { o.append((const char*)&v, 4); }
  void u64(uint64_t v) { o.append((const char*)&v, 8); }
  void mphf(const Mphf& m) {
    u64(m.levels.size());
    for (auto& l : m.levels) { u64(l.bits); u64(l.words.size()); for (uint64_t w : l.words) u64(w); }
    u64(m.levels.size());
    for (auto& l : m.levels) { u64(l.ranks.size()); for (uint64_t r : l.ranks) u64(r); }
  }
};

}  // namespace

void HostIndex::write_idx(const std::string& path) const {
  const uint32_t nn = n_nodes();
  // every (node, offset) with its k-mer; the graph holds each k-mer once
  std::vector<uint64_t> kmers; std::vector<uint32_t> k_node, k_off;
  kmers.reserve(n_kmers); k_node.reserve(n_kmers); k_off.reserve(n_kmers);
  for (uint32_t n = 0; n < nn; n++) {
    uint64_t v = 0;
    for (uint32_t j = 0; j < len[n]; j++) {
      v = ((v << 2) | base((uint64_t)start[n] + j)) & KMER_MASK;
      if (j + 1 >= (uint32_t)K) { kmers.push_back(v); k_node.push_back(n); k_off.push_back(j + 1 - K); }
    }
  }
  const size_t nk = kmers.size();
  // dbg_index: the reference orders k-mers by this MPHF while it compacts the graph, so node ids are the order in which
  // nodes first appear when the k-mers are walked by hash value
  Mphf all(kmers);
  std::vector<uint32_t> at(nk);   // hash -> k-mer record
  for (size_t i = 0; i < nk; i++) at[all.hash(kmers[i])] = (uint32_t)i;
  std::vector<uint32_t> new_id(nn, NONE32), old_of; old_of.reserve(nn);
  for (size_t h = 0; h < nk; h++) { const uint32_t n = k_node[at[h]]; if (new_id[n] == NONE32) { new_id[n] = (uint32_t)old_of.size(); old_of.push_back(n); } }
  if (old_of.size() != nn) throw Error(HLAC_ERR_STATE, "index has a node without k-mers");
  // colour ids: first appearance walking the k-mers in ascending order (the reference's k-mer collection pass)
  std::vector<uint32_t> by_value(nk);
  for (size_t i = 0; i < nk; i++) by_value[i] = (uint32_t)i;
  std::sort(by_value.begin(), by_value.end(), [&](uint32_t a, uint32_t b) { return kmers[a] < kmers[b]; });
  const uint32_t nc = n_colours();
  std::vector<uint32_t> new_col(nc, NONE32), old_col; old_col.reserve(nc);
  for (uint32_t i : by_value) { const uint32_t c = colour[k_node[i]]; if (new_col[c] == NONE32) { new_col[c] = (uint32_t)old_col.size(); old_col.push_back(c); } }
  for (uint32_t c = 0; c < nc; c++) if (new_col[c] == NONE32) { new_col[c] = (uint32_t)old_col.size(); old_col.push_back(c); }   // unused colours last
  Writer w;
  {   // sequences: node strings concatenated in the new order, 32 bases per word
    std::vector<uint64_t> words((n_bases + 31) / 32, 0);
    uint64_t g = 0;
    for (uint32_t n : old_of) for (uint32_t j = 0; j < len[n]; j++, g++) words[g >> 5] |= (uint64_t)base((uint64_t)start[n] + j) << (62 - 2 * (g & 31));
    w.u64(words.size()); for (uint64_t x : words) w.u64(x);
    w.u64(n_bases);
    w.u64(nn); g = 0; for (uint32_t n : old_of) { w.u64(g); g += len[n]; }
    w.u64(nn); for (uint32_t n : old_of) w.u32(len[n]);
    w.u64(nn); for (uint32_t n : old_of) w.u8(exts[n]);
    w.u64(nn); for (uint32_t n : old_of) w.u32(new_col[colour[n]]);
    w.u8(1);   // stranded
  }
  for (int side = 0; side < 2; side++) {   // left_order / right_order: BoomHashMap<Kmer24, u32>, keys and values in hash order
    std::vector<uint64_t> keys(nn);
    for (uint32_t i = 0; i < nn; i++) { const uint32_t n = old_of[i]; keys[i] = kmer_at((uint64_t)start[n] + (side ? len[n] - K : 0)); }
    Mphf m(keys);
    std::vector<uint64_t> ko(nn); std::vector<uint32_t> vo(nn);
    for (uint32_t i = 0; i < nn; i++) { const uint64_t h = m.hash(keys[i]); ko[h] = keys[i]; vo[h] = i; }
    w.mphf(m);
    w.u64(nn); for (uint64_t k : ko) w.u64(k);
    w.u64(nn); for (uint32_t v : vo) w.u32(v);
  }
  w.u64(nc);
  for (uint32_t c : old_col) { w.u64(col_off[c + 1] - col_off[c]); for (uint64_t q = col_off[c]; q < col_off[c + 1]; q++) w.u32(col_tx[q]); }
  w.mphf(all);
  w.u64(nk); for (size_t h = 0; h < nk; h++) { w.u32(new_id[k_node[at[h]]]); w.u32(k_off[at[h]]); }
  w.u64(tx_names.size()); for (auto& t : tx_names) { w.u64(t.size()); w.o += t; }
  w.u64(0);   // tx_gene_mapping: empty (hla.rs:257)
  std::ofstream f(path, std::ios::binary);
  if (!f) throw Error(HLAC_ERR_IO, "cannot write " + path);
  f.write(w.o.data(), (std::streamsize)w.o.size());
  if (!f) throw Error(HLAC_ERR_IO, "short write to " + path);
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
  // host threads for the embarrassingly parallel passes (enumeration, per-k-mer grouping); HLAC_HOST_THREADS overrides
  unsigned n_threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* e = getenv("HLAC_HOST_THREADS")) n_threads = (unsigned)std::max(1, atoi(e));
  WorkerPool pool(n_threads);
  // occurrence arrays: plain uninitialised storage, so that the first touch (page faults) happens on the worker threads
  struct Raw {
    std::unique_ptr<uint64_t[]> p; size_t n = 0;
    void alloc(size_t k) { p.reset(new uint64_t[k ? k : 1]); n = k; }
    uint64_t* data() { return p.get(); }
    uint64_t& operator[](size_t i) { return p[i]; }
    const uint64_t& operator[](size_t i) const { return p[i]; }
    void swap(Raw& o) { p.swap(o.p); std::swap(n, o.n); }
    void release() { p.reset(); n = 0; }
  } okey, oval;
  std::vector<size_t> first(seqs.size() + 1, 0);   // occurrences are laid out in ascending tx order (the stable sort relies on it)
  for (size_t t = 0; t < seqs.size(); t++) first[t + 1] = first[t] + (seqs[t].size() >= (size_t)K ? seqs[t].size() - K + 1 : 0);
  const size_t total = first[seqs.size()];
  okey.alloc(total); oval.alloc(total);
  pool.for_ranges(seqs.size(), 32, [&](size_t tb, size_t te) {
    for (size_t t = tb; t < te; t++) {
      const Bases& s = seqs[t];
      if (s.size() < (size_t)K) continue;
      uint64_t v = 0;
      for (int i = 0; i < K - 1; i++) v = (v << 2) | s[i];
      size_t at = first[t];
      for (size_t i = 0; i + K <= s.size(); i++, at++) {
        v = ((v << 2) | s[i + K - 1]) & KMER_MASK;
        okey[at] = v;
        oval[at] = ((uint64_t)t << 16) | (uint64_t)(i > 0 ? 1u << s[i - 1] : 0) << 8 | (uint64_t)(i + K < s.size() ? 1u << s[i + K] : 0);
      }
    }
  });
  tick("enumerate k-mers");
  // stable sort by k-mer: occurrences were generated in ascending tx order, so stability leaves every k-mer's tx list
  // sorted. On a device-resident build the caller supplies a GPU radix sort; otherwise an LSD radix sort on the host.
  if (sorter && total >= (1u << 16)) sorter(okey.data(), oval.data(), total);
  else {
    // host LSD radix sort, 3 passes of 16 bits, stable: every thread histograms and then scatters its own contiguous chunk
    // into bucket ranges laid out chunk after chunk
    Raw tk, tv; tk.alloc(total); tv.alloc(total);
    const size_t nch = std::max<size_t>(1, std::min<size_t>(pool.size(), total / (1u << 16) + 1));
    std::vector<std::vector<size_t>> cnt(nch, std::vector<size_t>(1 << 16));
    auto lo = [&](size_t c) { return c * (total / nch) + std::min(c, total % nch); };
    for (int pass = 0; pass < 3; pass++) {
      const int sh = 16 * pass;
      pool.for_ranges(nch, 1, [&](size_t cb, size_t ce) {
        for (size_t c = cb; c < ce; c++) {
          std::fill(cnt[c].begin(), cnt[c].end(), 0);
          for (size_t i = lo(c); i < lo(c + 1); i++) cnt[c][(okey[i] >> sh) & 0xFFFF]++;
        }
      });
      size_t run = 0;
      for (size_t b = 0; b < (1u << 16); b++) for (size_t c = 0; c < nch; c++) { const size_t k = cnt[c][b]; cnt[c][b] = run; run += k; }
      pool.for_ranges(nch, 1, [&](size_t cb, size_t ce) {
        for (size_t c = cb; c < ce; c++)
          for (size_t i = lo(c); i < lo(c + 1); i++) { const size_t d = cnt[c][(okey[i] >> sh) & 0xFFFF]++; tk[d] = okey[i]; tv[d] = oval[i]; }
      });
      okey.swap(tk); oval.swap(tv);
    }
  }
  tick("sort occurrences");

  // unique k-mers with ext masks and interned colour. Parallel part: the sorted occurrences are cut into chunks at k-mer
  // boundaries and every chunk summarises its k-mers (ext masks, occurrence range, length and hash of the de-duplicated tx
  // list). Sequential part: colour ids in first-appearance order over the sorted k-mers, comparing tx lists only on a hash hit.
  struct Uniq { uint64_t kmer, hash; size_t b, e; uint32_t ntx; uint8_t l, r; };
  const size_t n_chunks = std::max<size_t>(1, std::min<size_t>(pool.size() * 8, total / 4096 + 1));
  std::vector<size_t> cut(n_chunks + 1, total);
  cut[0] = 0;
  for (size_t c = 1; c < n_chunks; c++) { size_t b = c * (total / n_chunks); while (b < total && b > 0 && okey[b] == okey[b - 1]) b++; cut[c] = std::max(b, cut[c - 1]); }
  std::vector<std::vector<Uniq>> parts(n_chunks);
  pool.for_ranges(n_chunks, 1, [&](size_t cb, size_t ce) {
    for (size_t c = cb; c < ce; c++) {
      std::vector<Uniq>& out = parts[c];
      for (size_t i = cut[c]; i < cut[c + 1];) {
        Uniq u{okey[i], 0xcbf29ce484222325ULL, i, i, 0, 0, 0};
        uint32_t prev = NONE32;
        size_t j = i;
        while (j < total && okey[j] == u.kmer) {
          const uint32_t tx = (uint32_t)(oval[j] >> 16);
          u.l |= (uint8_t)(oval[j] >> 8); u.r |= (uint8_t)oval[j];
          if (tx != prev) { prev = tx; u.ntx++; u.hash = (u.hash ^ tx) * 0x100000001b3ULL; u.hash ^= u.hash >> 29; }
          j++;
        }
        u.e = j;
        out.push_back(u);
        i = j;
      }
    }
  });
  std::vector<uint64_t> uk; std::vector<uint8_t> ul, ur; std::vector<uint32_t> ucol;
  size_t nu = 0;
  for (auto& pt : parts) nu += pt.size();
  uk.reserve(nu); ul.reserve(nu); ur.reserve(nu); ucol.reserve(nu);
  std::vector<const Uniq*> all; all.reserve(nu);
  for (auto& pt : parts) for (const Uniq& u : pt) all.push_back(&u);
  // sequential: colour id = first appearance of (hash, list length) over the sorted k-mers; rep[c] = the k-mer that defined colour c
  std::unordered_map<uint64_t, std::vector<uint32_t>> by_hash;   // hash(tx list) -> candidate colour ids
  std::vector<uint32_t> rep;
  ix.col_off.assign(1, 0);
  for (size_t q = 0; q < nu; q++) {
    const Uniq& u = *all[q];
    uint32_t cid = NONE32;
    auto& cand = by_hash[u.hash];
    for (uint32_t c : cand) if (all[rep[c]]->ntx == u.ntx) { cid = c; break; }
    if (cid == NONE32) { cid = (uint32_t)rep.size(); rep.push_back((uint32_t)q); cand.push_back(cid); ix.col_off.push_back(ix.col_off.back() + u.ntx); }
    uk.push_back(u.kmer); ul.push_back(u.l); ur.push_back(u.r); ucol.push_back(cid);
  }
  // parallel: materialise every colour from its defining k-mer, then verify every k-mer's tx list against its colour (a
  // 64-bit hash + length collision between different lists would otherwise merge two colours silently)
  ix.col_tx.resize(ix.col_off.back());
  pool.for_ranges(rep.size(), 64, [&](size_t cb, size_t ce) {
    for (size_t c = cb; c < ce; c++) {
      const Uniq& u = *all[rep[c]];
      uint32_t* dst = ix.col_tx.data() + ix.col_off[c]; uint32_t prev = NONE32;
      for (size_t j = u.b; j < u.e; j++) { const uint32_t tx = (uint32_t)(oval[j] >> 16); if (tx != prev) { prev = tx; *dst++ = tx; } }
    }
  });
  std::atomic<bool> collision{false};
  pool.for_ranges(nu, 2048, [&](size_t qb, size_t qe) {
    for (size_t q = qb; q < qe; q++) {
      const Uniq& u = *all[q];
      if (rep[ucol[q]] == q) continue;
      const uint32_t* have = ix.col_tx.data() + ix.col_off[ucol[q]];
      size_t k = 0; uint32_t prev = NONE32;
      for (size_t j = u.b; j < u.e; j++) {
        const uint32_t tx = (uint32_t)(oval[j] >> 16);
        if (tx == prev) continue;
        prev = tx;
        if (have[k++] != tx) { collision = true; return; }
      }
    }
  });
  if (collision) throw Error(HLAC_ERR_STATE, "tx-list hash collision while interning colours");
  okey.release(); oval.release();
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
