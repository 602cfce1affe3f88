// Host driver above the C ABI: the three entry points src/main.rs imports (mapping_wrapper, em_wrapper,
// map_and_count_pseudo), load_barcodes, the output writers and sc_hla_count's _main. BAM/FASTA I/O, barcode and UMI
// interning stay on the host; every map_read / count() / eq_class_counts / squarem call goes through the hlac_* GPU
// sessions — there is no CPU implementation of those here.
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <future>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <unordered_map>

#include "bam_reader.hpp"
#include "worker_pool.hpp"
#include "common.hpp"
#include "hla_fasta.hpp"

using namespace hlac;

namespace {

bool path_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
void make_dirs(const std::string& p) {
  std::string cur;
  for (size_t i = 0; i <= p.size(); i++) {
    if (i == p.size() || p[i] == '/') { if (!cur.empty() && !path_exists(cur)) { if (mkdir(cur.c_str(), 0777) != 0) throw Error(HLAC_ERR_IO, "cannot create directory " + cur); } }
    if (i < p.size()) cur += p[i];
  }
}
std::string join(const std::string& a, const std::string& b) { return a.empty() ? b : (a.back() == '/' ? a + b : a + "/" + b); }
void info(const std::string& m) { std::cerr << "[INFO] " << m << std::endl; }

void check(int rc) { if (rc != HLAC_OK) throw Error(rc, hlac_last_error()); }

// src/locus.rs:29-48 — regex ^(.*):([0-9,]+)(-|..)([0-9,]+)$ : greedy contig up to the last colon that lets the rest
// match, start and end may contain commas, the separator is '-' or ANY two characters (the reference's unescaped "..").
struct Locus { std::string chrom; uint64_t start = 0, end = 0; };
Locus parse_locus(const std::string& s) {
  auto numeric = [](const std::string& t) { if (t.empty()) return false; for (char c : t) if (!((c >= '0' && c <= '9') || c == ',')) return false; return true; };
  auto value = [&](std::string t) {
    t.erase(std::remove(t.begin(), t.end(), ','), t.end());
    if (t.empty() || t.size() > 10 || std::stoull(t) > 0xFFFFFFFFull) throw Error(HLAC_ERR_ARG, "invalid locus: " + s);   // u32::from_str(..).unwrap()
    return (uint64_t)std::stoull(t);
  };
  for (size_t colon = s.rfind(':'); colon != std::string::npos; colon = colon ? s.rfind(':', colon - 1) : std::string::npos) {
    const std::string rest = s.substr(colon + 1);
    size_t run = 0;
    while (run < rest.size() && ((rest[run] >= '0' && rest[run] <= '9') || rest[run] == ',')) run++;
    for (size_t n2 = run; n2 >= 1; n2--) {   // group 2 is greedy, then backtracks
      for (size_t sep : {size_t(1), size_t(2)}) {
        if (sep == 1 && (n2 >= rest.size() || rest[n2] != '-')) continue;
        if (n2 + sep >= rest.size() || !numeric(rest.substr(n2 + sep))) continue;
        Locus l; l.chrom = s.substr(0, colon); l.start = value(rest.substr(0, n2)); l.end = value(rest.substr(n2 + sep));
        return l;
      }
    }
    if (colon == 0) break;
  }
  throw Error(HLAC_ERR_ARG, "invalid locus: " + s);
}

// src/main.rs:406-424 + src/io.rs: column = line index, last occurrence wins, gzip sniffed
std::unordered_map<std::string, uint32_t> load_barcodes(const std::string& path) {
  { struct stat st; if (stat(path.c_str(), &st) != 0) throw Error(HLAC_ERR_IO, "Error opening file: \"" + path + "\"");
    if (st.st_size < 4) throw Error(HLAC_ERR_IO, "failed to fill whole buffer"); }   // io.rs:20 read_exact of the 4 sniff bytes
  gzFile g = gzopen(path.c_str(), "rb");   // gzopen reads plain files transparently
  if (!g) throw Error(HLAC_ERR_IO, "cannot open barcode file " + path);
  std::string data; char buf[1 << 16]; int k;
  while ((k = gzread(g, buf, sizeof buf)) > 0) data.append(buf, k);
  gzclose(g);
  std::unordered_map<std::string, uint32_t> m;
  uint32_t line = 0; size_t p = 0;
  while (p < data.size()) {
    size_t e = data.find('\n', p);
    std::string l = data.substr(p, e == std::string::npos ? std::string::npos : e - p);
    if (!l.empty() && l.back() == '\r') l.pop_back();
    m[l] = line++;
    if (e == std::string::npos) break;
    p = e + 1;
  }
  if (m.empty()) throw Error(HLAC_ERR_FORMAT, "Loaded 0 barcodes. Is your barcode file gzipped or empty?");
  return m;
}

// Rust `{}` for f64: shortest round-trip digits, never scientific
std::string fmt_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  char buf[512];
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  return std::string(buf, r.ptr);
}

// pinned batch assembly for the GPU sessions
struct BatchBuilder {
  uint32_t wpr = 0; size_t cap; uint64_t n = 0;
  uint64_t* seq = nullptr; uint16_t* len = nullptr; uint32_t* cell = nullptr; uint32_t* umi = nullptr; uint8_t* flags = nullptr;
  uint64_t* ordinal = nullptr;   // optional (with_ordinal): stream position of every record, for sessions that see a shard
  uint64_t ordinal_base = 0;     // added to the positions (a second pass continues the numbering of the first)
  std::unordered_map<std::string, uint32_t> umi_ids;   // UMIs that are not <=15 ACGT bases
  bool pinned;   // page-locked buffers for the H2D copies; pageable for the host-only self-check
  bool with_ordinal;
  explicit BatchBuilder(size_t capacity, bool pinned_ = true, bool with_ordinal_ = false) : cap(capacity), pinned(pinned_), with_ordinal(with_ordinal_) {}
  ~BatchBuilder() { release(); }
  void release() {
    for (void* p : {(void*)seq, (void*)len, (void*)cell, (void*)umi, (void*)flags, (void*)ordinal}) if (p) { if (pinned) hlac_host_free(p); else free(p); }
    seq = nullptr; len = nullptr; cell = nullptr; umi = nullptr; flags = nullptr; ordinal = nullptr;
  }
  void alloc(uint32_t words) {
    release(); wpr = words;
    auto get = [&](size_t bytes, void** out) { if (pinned) check(hlac_host_alloc(bytes, out)); else if (!(*out = malloc(bytes))) throw Error(HLAC_ERR_IO, "out of memory"); };
    get(cap * wpr * 8, (void**)&seq); get(cap * 2, (void**)&len); get(cap * 4, (void**)&cell); get(cap * 4, (void**)&umi); get(cap, (void**)&flags);
    if (with_ordinal) get(cap * 8, (void**)&ordinal);
  }
  uint32_t umi_key(const std::string& u) {
    uint32_t k;
    if (hlac_umi_key(u.data(), (uint32_t)u.size(), &k) == HLAC_OK) return k;
    auto it = umi_ids.find(u);
    if (it == umi_ids.end()) it = umi_ids.emplace(u, 0x80000000u | (uint32_t)umi_ids.size()).first;
    return it->second;
  }
  bool fits(size_t bases) const { return wpr && bases <= (size_t)wpr * 32 && n < cap; }
  void add(const BamRecord& r, uint32_t cell_id) {
    uint64_t* w = seq + n * wpr;
    for (uint32_t i = 0; i < wpr; i++) w[i] = 0;
    for (size_t i = 0; i < r.seq.size(); i++) w[i >> 5] |= (uint64_t)r.seq[i] << (62 - 2 * (i & 31));
    len[n] = (uint16_t)r.seq.size(); cell[n] = cell_id; umi[n] = umi_key(r.umi); flags[n] = r.primary ? HLAC_FLAG_PRIMARY : 0;
    n++;
  }
  hlac_batch batch() const { hlac_batch b{}; b.seq = seq; b.len = len; b.cell = cell; b.umi = umi; b.flags = flags; b.n = n; b.words_per_read = wpr; b.ordinal = ordinal; return b; }
};

// Streams the selected BAM records into a session through pinned batches. `flush` pushes one batch and syncs.
// The reader frames raw records sequentially (a copy), then a pool of host threads does the expensive part per record
// - overlap test, CB / UB aux scan, UMI key, 4-bit -> 2-bit packing, and the barcode lookup when `cell_of` is a pure
// function (the counting whitelist) - and one thread appends the survivors to the batch in file order. cell_of(cb, ub)
// with side effects (first-seen barcode ids in mapping_wrapper) runs in that ordered append instead.
struct StrRef { const char* p; uint32_t n; std::string str() const { return std::string(p, n); } };

WorkerPool& host_pool() {
  static WorkerPool pool([] {
    unsigned n = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("HLAC_BAM_THREADS")) n = (unsigned)std::max(1, atoi(e));
    return n; }());
  return pool;
}

// cell_try(cb): optional read-only pre-lookup for a cell_of with side effects - it runs on the parser threads while nothing
// is being inserted and returns CELL_UNKNOWN when the ordered append has to call cell_of (a barcode not seen before).
constexpr uint32_t CELL_UNKNOWN = 0xFFFFFFFEu;
struct NoCellTry { uint32_t operator()(StrRef) const { return CELL_UNKNOWN; } };

// Counting only: reads whose barcode is not in the whitelist are never scored (mapping.rs:760-763), so the driver keeps them
// off the PCIe link and the GPU and only accounts for them (mapping.rs:752-763), see hlac_count_add_filtered.
struct HostFilter { bool primary_only = false; uint64_t n_reads = 0, n_non_primary = 0, n_not_cell = 0; };

// route(cell) picks the batch (= device lane) a record goes to; flush(lane) pushes that lane's batch.
template <typename CellFn, typename Flush, typename Route, typename CellTry = NoCellTry>
uint64_t stream_records_routed(BamReader& bam, std::vector<BatchBuilder*>& lanes, Route route, CellFn cell_of, Flush flush, bool cell_of_is_pure,
                               CellTry cell_try = CellTry(), HostFilter* filter = nullptr) {
  if (filter && !cell_of_is_pure) throw Error(HLAC_ERR_STATE, "the host filter needs a pure barcode lookup");
  constexpr size_t CHUNK = 1 << 16, GRAIN = 1024;
  constexpr uint32_t UMI_INTERN = 0xFFFFFFFFu;
  WorkerPool& pool = host_pool();
  std::vector<uint8_t> ok; std::vector<uint64_t> words; std::vector<BamReader::RecView> views; std::vector<uint32_t> ukey, cells;
  uint64_t total = 0;
  const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr;
  double t_frame = 0, t_parse = 0, t_pack = 0, t_append = 0, t_work = 0;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
  struct Report { bool on; double &a, &b, &c, &d, &e; ~Report() { if (on) fprintf(stderr, "[hlac timing] ingest: wait for frames %.3f s (framing itself %.3f s on the helper thread), parse %.3f s, pack %.3f s, append %.3f s\n", a, e, b, c, d); } } report{dbg, t_frame, t_parse, t_pack, t_append, t_work};
  // two chunks in flight: while this thread and the pool parse / append chunk k (and the session maps the batch it flushes),
  // a helper thread frames chunk k + 1 (which is also what keeps the inflate workers fed)
  BamReader::Chunk chunks[2];
  int cur = 0;
  double t0 = now();
  size_t n = bam.frame(chunks[0], CHUNK);
  t_frame += now() - t0;
  for (; n > 0; cur ^= 1) {
    const std::vector<BamReader::RawRec>& recs = chunks[cur].recs;
    std::future<size_t> next = std::async(std::launch::async, [&bam, &chunks, cur, &t_work, &now] { const double a = now(); const size_t k = bam.frame(chunks[cur ^ 1], CHUNK); t_work += now() - a; return k; });
    struct Join { std::future<size_t>& f; ~Join() { if (f.valid()) f.wait(); } } join{next};   // never leave the helper running behind an exception
    t0 = now();
    views.resize(n); ok.assign(n, 0); ukey.resize(n); cells.resize(n);
    pool.for_ranges(n, GRAIN, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
        BamReader::RecView& v = views[i];
        if (!bam.parse(recs[i].p, recs[i].n, v)) continue;
        if (v.l_seq > 65535) throw Error(HLAC_ERR_LIMIT, "read longer than 65535 bases");
        if (v.l_seq > 512) throw Error(HLAC_ERR_LIMIT, "read longer than 512 bases");
        ok[i] = 1;
        uint32_t k = UMI_INTERN;   // hlac_umi_key: <= 15 ACGT bases -> (1 << 2n) | 2-bit value, anything else is interned
        if (v.ub_len <= 15) {
          k = 1;
          for (uint32_t j = 0; j < v.ub_len; j++) {
            const char c = v.ub[j];
            const uint32_t code = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
            if (code > 3) { k = UMI_INTERN; break; }
            k = (k << 2) | code;
          }
        }
        ukey[i] = k;
        cells[i] = cell_of_is_pure ? cell_of(StrRef{v.cb, v.cb_len}, StrRef{v.ub, v.ub_len}) : cell_try(StrRef{v.cb, v.cb_len});
      }
    });
    t_parse += now() - t0; t0 = now();
    uint32_t need = 3;
    for (size_t i = 0; i < n; i++) if (ok[i]) need = std::max<uint32_t>(need, (uint32_t)((views[i].l_seq + 31) / 32));
    words.resize(n * (size_t)need);
    pool.for_ranges(n, GRAIN, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++)
        if (ok[i] && !(filter && cells[i] == HLAC_NOT_A_CELL)) BamReader::pack_2bit(views[i].seq4, views[i].l_seq, words.data() + i * need, need);
    });
    t_pack += now() - t0; t0 = now();
    for (size_t i = 0; i < n; i++) {
      if (!ok[i]) continue;
      const BamReader::RecView& v = views[i];
      if (filter && cells[i] == HLAC_NOT_A_CELL) {
        filter->n_reads++; total++;
        if (!v.primary) filter->n_non_primary++;
        if (v.primary || !filter->primary_only) filter->n_not_cell++;
        continue;
      }
      const uint32_t cell = (cell_of_is_pure || cells[i] != CELL_UNKNOWN) ? cells[i] : cell_of(StrRef{v.cb, v.cb_len}, StrRef{v.ub, v.ub_len});
      const size_t lane = route(cell, total);
      BatchBuilder& bb = *lanes[lane];
      if (!bb.fits((size_t)v.l_seq)) {
        if (bb.n) { flush(lane); bb.n = 0; }
        const uint32_t want = std::max<uint32_t>(3, (uint32_t)((v.l_seq + 31) / 32));
        if (want > bb.wpr) bb.alloc(want);
      }
      uint64_t* w = bb.seq + bb.n * bb.wpr;
      const uint32_t have = (uint32_t)((v.l_seq + 31) / 32);   // words carrying bases; the rest of both rows is zero
      for (uint32_t j = 0; j < bb.wpr; j++) w[j] = j < have ? words[i * need + j] : 0;
      bb.len[bb.n] = (uint16_t)v.l_seq;
      bb.cell[bb.n] = cell;
      // interned UMI ids (UMIs that are not <= 15 ACGT bases) come from lane 0's table so that they are unique over all lanes
      bb.umi[bb.n] = ukey[i] != UMI_INTERN ? ukey[i] : lanes[0]->umi_key(std::string(v.ub, v.ub_len));
      bb.flags[bb.n] = v.primary ? HLAC_FLAG_PRIMARY : 0;
      if (bb.ordinal) bb.ordinal[bb.n] = bb.ordinal_base + total;   // position in the stream (first-seen class ids are global, mapping.rs:147-154)
      bb.n++;
      total++;
    }
    t_append += now() - t0; t0 = now();
    n = next.get();
    t_frame += now() - t0;   // time spent WAITING for the next chunk
  }
  for (size_t lane = 0; lane < lanes.size(); lane++) if (lanes[lane]->n) { flush(lane); lanes[lane]->n = 0; }
  return total;
}

template <typename CellFn, typename Flush, typename CellTry = NoCellTry>
uint64_t stream_records(BamReader& bam, BatchBuilder& bb, CellFn cell_of, Flush flush, bool cell_of_is_pure, CellTry cell_try = CellTry(),
                        HostFilter* filter = nullptr) {
  std::vector<BatchBuilder*> lanes{&bb};
  return stream_records_routed(bam, lanes, [](uint32_t, uint64_t) { return size_t(0); }, cell_of, [&](size_t) { flush(); }, cell_of_is_pure, cell_try, filter);
}

constexpr const char* COUNTS_MAGIC = "HLACCNT1";

void write_tables(const std::string& path, const hlac_class_tables& t) {
  std::ofstream o(path, std::ios::binary);
  if (!o) throw Error(HLAC_ERR_IO, "cannot write " + path);
  auto w = [&](const void* p, size_t n) { o.write((const char*)p, n); };
  w(COUNTS_MAGIC, 8); w(&t.nitems, 4); w(&t.nreads, 4);
  w(&t.n_read_classes, 8); w(t.reads_off, (t.n_read_classes + 1) * 8);
  const uint64_t nr = t.reads_off[t.n_read_classes]; w(t.reads_tx, nr * 4); w(t.reads_cnt, t.n_read_classes * 4);
  w(&t.n_umi_classes, 8); w(t.umi_off, (t.n_umi_classes + 1) * 8);
  const uint64_t nu = t.umi_off[t.n_umi_classes]; w(t.umi_tx, nu * 4); w(t.umi_cnt, t.n_umi_classes * 4);
  w(t.reads_explained, (size_t)t.nitems * 8);
}

bool is_own_counts_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(HLAC_ERR_IO, "cannot open " + path);
  char magic[8] = {0}; f.read(magic, 8);
  return f.gcount() == 8 && memcmp(magic, COUNTS_MAGIC, 8) == 0;
}

struct OwnedTables {
  hlac_class_tables t{};
  std::vector<uint64_t> roff, uoff, rex; std::vector<uint32_t> rtx, rcnt, utx, ucnt;
};
void read_tables(const std::string& path, OwnedTables& ot) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(HLAC_ERR_IO, "cannot open " + path);
  auto r = [&](void* p, size_t n) { f.read((char*)p, n); if ((size_t)f.gcount() != n) throw Error(HLAC_ERR_FORMAT, path + ": truncated counts file"); };
  char magic[8]; r(magic, 8);
  if (memcmp(magic, COUNTS_MAGIC, 8) != 0) throw Error(HLAC_ERR_FORMAT, path + ": not a counts file in this library's table format");
  auto& t = ot.t;
  r(&t.nitems, 4); r(&t.nreads, 4);
  r(&t.n_read_classes, 8); ot.roff.resize(t.n_read_classes + 1); r(ot.roff.data(), ot.roff.size() * 8);
  ot.rtx.resize(std::max<uint64_t>(ot.roff.back(), 1)); r(ot.rtx.data(), ot.roff.back() * 4); ot.rcnt.resize(std::max<uint64_t>(t.n_read_classes, 1)); r(ot.rcnt.data(), t.n_read_classes * 4);
  r(&t.n_umi_classes, 8); ot.uoff.resize(t.n_umi_classes + 1); r(ot.uoff.data(), ot.uoff.size() * 8);
  ot.utx.resize(std::max<uint64_t>(ot.uoff.back(), 1)); r(ot.utx.data(), ot.uoff.back() * 4); ot.ucnt.resize(std::max<uint64_t>(t.n_umi_classes, 1)); r(ot.ucnt.data(), t.n_umi_classes * 4);
  ot.rex.resize(std::max<uint32_t>(t.nitems, 1)); r(ot.rex.data(), (size_t)t.nitems * 8);
  t.reads_off = ot.roff.data(); t.reads_tx = ot.rtx.data(); t.reads_cnt = ot.rcnt.data();
  t.umi_off = ot.uoff.data(); t.umi_tx = ot.utx.data(); t.umi_cnt = ot.ucnt.data(); t.reads_explained = ot.rex.data();
}

void copy_out(const std::string& s, char* out, size_t cap) {
  if (!out || cap == 0) return;
  if (s.size() + 1 > cap) throw Error(HLAC_ERR_ARG, "output path buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
}

struct IndexHandle { hlac_index* p = nullptr; ~IndexHandle() { if (p) hlac_index_free(p); } };
struct GenoHandle { hlac_geno* p = nullptr; ~GenoHandle() { if (p) hlac_geno_free(p); } };
struct CountHandle { hlac_count* p = nullptr; ~CountHandle() { if (p) hlac_count_free(p); } };

constexpr size_t BATCH_READS = 1 << 20;

}  // namespace

#define HLAC_API_TRY try {
#define HLAC_API_CATCH                                                                       \
  }                                                                                          \
  catch (const hlac::Error& e) { hlac::set_last_error(e.what()); return e.code; }           \
  catch (const std::exception& e) { hlac::set_last_error(e.what()); return HLAC_ERR_STATE; } \
  return HLAC_OK;

namespace {

// mapping_wrapper's body (mapping.rs:320-404) on an index that is already resident on the device(s). With several devices
// (index[d] resident on device d) reads are routed by barcode (first-seen barcode id % n) with their stream ordinals, every
// device runs its own session, and the sessions merge inside hlac_geno_finish through a communicator.
std::string mapping_impl(const std::vector<hlac_index*>& index, const char* outdir, const char* bam, const char* region, int use_unmapped) {
  info("Mapping reads from BAM to database");
  const size_t nd = index.size();
  std::vector<GenoHandle> g(nd);
  struct CommHandle { hlac_comm* p = nullptr; ~CommHandle() { if (p) hlac_comm_free(p); } };
  std::vector<CommHandle> comms(nd);
  for (size_t d = 0; d < nd; d++) check(hlac_geno_new(index[d], &g[d].p));
  if (nd > 1) {
    std::vector<int> devs(nd); std::vector<hlac_comm*> raw(nd, nullptr);
    for (size_t d = 0; d < nd; d++) { hlac_index_info unused; (void)unused; devs[d] = hlac_index_device(index[d]); }
    check(hlac_comm_init_all(devs.data(), (int)nd, raw.data()));
    for (size_t d = 0; d < nd; d++) { comms[d].p = raw[d]; check(hlac_geno_attach_comm(g[d].p, raw[d])); }
    info("Genotyping on " + std::to_string(nd) + " GPUs: reads routed by barcode, paths merged over NCCL");
  }
  // The hand-over file is the reference's own: bincode of EqClassDb (mapping.rs:84-90,403), so that either side of the pair
  // mapping_wrapper / em_wrapper can be the reference's. It needs the per-read class ids from the session(s) and the barcode /
  // UMI strings of every read from here. HLAC_COUNTS_FORMAT=tables writes this library's compact table format instead
  // (the three tables, no per-read triples), which only hlac_em_wrapper reads.
  const char* fmt = getenv("HLAC_COUNTS_FORMAT");
  const bool bincode = !fmt || std::string(fmt) == "bincode";
  if (fmt && !bincode && std::string(fmt) != "tables") throw Error(HLAC_ERR_ARG, "HLAC_COUNTS_FORMAT must be 'tables' or 'bincode'");
  if (bincode) for (size_t d = 0; d < nd; d++) check(hlac_geno_record_reads(g[d].p, 1));
  std::vector<uint32_t> read_bc, read_umi;             // bincode only: per read, index into bc_names / umi_names
  std::vector<std::string> bc_names, umi_names;
  std::unordered_map<std::string, uint32_t> umi_idx;
  BamReader reader(bam);
  std::unordered_map<std::string, uint32_t> bc_ids;   // mapping_wrapper uses every barcode (no whitelist)
  std::string key;   // reused buffer: one hash lookup per record, no allocation
  auto cell_of = [&](StrRef cb, StrRef ub) {
    key.assign(cb.p, cb.n);
    const uint32_t id = bc_ids.emplace(key, (uint32_t)bc_ids.size()).first->second;
    if (bincode) {
      if (id == bc_names.size()) bc_names.push_back(key);
      key.assign(ub.p, ub.n);
      auto it = umi_idx.find(key);
      if (it == umi_idx.end()) { it = umi_idx.emplace(key, (uint32_t)umi_names.size()).first; umi_names.push_back(key); }
      read_bc.push_back(id); read_umi.push_back(it->second);
    }
    return id;
  };
  std::vector<std::unique_ptr<BatchBuilder>> bbs;
  std::vector<BatchBuilder*> lanes;
  for (size_t d = 0; d < nd; d++) { bbs.emplace_back(new BatchBuilder(BATCH_READS, true, nd > 1)); lanes.push_back(bbs.back().get()); }
  std::string counts = outdir;   // PathBuf::set_extension("counts.bin") (mapping.rs:328-329)
  { size_t slash = counts.rfind('/'), dot = counts.rfind('.');
    if (dot != std::string::npos && (slash == std::string::npos || dot > slash) && dot != slash + 1) counts.erase(dot);
    counts += ".counts.bin"; }
  auto report = [&](const char* what, uint64_t rid) {
    uint64_t some = 0, lng = 0;
    for (size_t d = 0; d < nd; d++) { uint64_t a = 0, b = 0; check(hlac_geno_stats(g[d].p, &a, &b, nullptr, nullptr, nullptr)); some += a; lng += b; }
    info("analyzed " + std::to_string(rid) + what + " reads. Mapped " + std::to_string(some) + ", used " + std::to_string(lng) +
         " with score at least " + std::to_string(MIN_SCORE_CALL));
  };
  if (region) { Locus l = parse_locus(region); reader.fetch(l.chrom, l.start, l.end); }
  else reader.fetch_unmapped();   // locus == None never happens from main.rs; keep the call total
  // barcodes seen in earlier chunks are resolved on the parser threads (read-only lookups; inserts only happen in the ordered
  // append, between two parallel stages); the bincode hand-over needs every record to pass through cell_of
  auto cell_try = [&](StrRef cb) { if (bincode) return CELL_UNKNOWN; auto it = bc_ids.find(cb.str()); return it == bc_ids.end() ? CELL_UNKNOWN : it->second; };
  auto route = [nd](uint32_t cell, uint64_t) { return (size_t)(cell % nd); };
  uint64_t rid = stream_records_routed(reader, lanes, route, cell_of, [&](size_t d) { hlac_batch b = lanes[d]->batch(); check(hlac_geno_push(g[d].p, &b, 0)); }, false, cell_try);
  report("", rid);
  if (use_unmapped) {
    reader.fetch_unmapped();
    for (auto* l : lanes) l->ordinal_base = rid;   // the unmapped pass continues the stream
    uint64_t rid2 = stream_records_routed(reader, lanes, route, cell_of, [&](size_t d) { hlac_batch b = lanes[d]->batch(); check(hlac_geno_push(g[d].p, &b, 1)); }, false, cell_try);
    report(" unaligned", rid2);
  }
  hlac_class_tables t{};
  struct Free { hlac_class_tables* t; ~Free() { hlac_class_tables_free(t); } } fr{&t};
  if (nd == 1) check(hlac_geno_finish(g[0].p, &t));
  else {   // the finish is a collective: one host thread per device; device 0's tables are the result
    std::vector<std::future<void>> fut;
    std::vector<std::string> errs(nd);
    std::vector<int> rcs(nd, HLAC_OK);
    for (size_t d = 1; d < nd; d++)
      fut.push_back(std::async(std::launch::async, [&, d] {
        hlac_class_tables td{}; rcs[d] = hlac_geno_finish(g[d].p, &td); if (rcs[d] != HLAC_OK) errs[d] = hlac_last_error(); hlac_class_tables_free(&td); }));
    rcs[0] = hlac_geno_finish(g[0].p, &t);
    if (rcs[0] != HLAC_OK) errs[0] = hlac_last_error();
    for (auto& f : fut) f.get();
    for (size_t d = 0; d < nd; d++) if (rcs[d] != HLAC_OK) throw Error(rcs[d], "device " + std::to_string(d) + ": " + errs[d]);
  }
  if (bincode) {
    auto csr = [](const std::vector<std::string>& v, std::vector<uint64_t>& off, std::string& bytes) {
      off.assign(1, 0); for (auto& x : v) { bytes += x; off.push_back(bytes.size()); } };
    std::vector<uint64_t> boff, uoff; std::string bbytes, ubytes;
    csr(bc_names, boff, bbytes); csr(umi_names, uoff, ubytes);
    // per-read class ids in stream order: every session knows its own reads in its own push order (= stream order within a lane)
    std::vector<uint32_t> ids(std::max<size_t>(read_bc.size(), 1));
    if (nd == 1) { if (!read_bc.empty()) check(hlac_geno_read_class_ids(g[0].p, ids.data(), read_bc.size())); }
    else {
      std::vector<std::vector<uint32_t>> per(nd);
      std::vector<uint64_t> cnt(nd, 0);
      for (uint32_t b : read_bc) cnt[b % nd]++;
      for (size_t d = 0; d < nd; d++) { per[d].resize(std::max<uint64_t>(cnt[d], 1)); if (cnt[d]) check(hlac_geno_read_class_ids(g[d].p, per[d].data(), cnt[d])); }
      std::vector<uint64_t> cur(nd, 0);
      for (size_t i = 0; i < read_bc.size(); i++) { const size_t d = read_bc[i] % nd; ids[i] = per[d][cur[d]++]; }
    }
    hlac_eqclassdb db{};
    check(hlac_eqclassdb_from_ids(&t, ids.data(), read_bc.data(), read_umi.data(), read_bc.size(), boff.data(), (const uint8_t*)bbytes.data(), (uint32_t)bc_names.size(),
                                  uoff.data(), (const uint8_t*)ubytes.data(), (uint32_t)umi_names.size(), &db));
    struct FreeDb { hlac_eqclassdb* d; ~FreeDb() { hlac_eqclassdb_free(d); } } fdb{&db};
    check(hlac_eqclassdb_write(counts.c_str(), &db));
  } else write_tables(counts, t);
  return counts;
}

// em_wrapper's body (em.rs:346-475) given the index handle (names only are needed from it)
void em_impl(hlac_index* index, const char* hla_counts, const char* cds_db, const char* gen_db, const char* outdir, int device,
             std::string& gen_name, std::string& cds_name) {
  struct { hlac_index* p; } ix{index};
  hlac_index_info inf; check(hlac_index_get_info(ix.p, &inf));
  OwnedTables ot;
  struct FreeT { hlac_class_tables* t = nullptr; ~FreeT() { if (t) hlac_class_tables_free(t); } } free_t;
  if (is_own_counts_file(hla_counts)) read_tables(hla_counts, ot);
  else {   // the reference's bincode EqClassDb (em.rs:347): eq_class_counts runs on the device (em.rs:358)
    hlac_eqclassdb db{}; check(hlac_eqclassdb_read(hla_counts, &db));
    struct FreeDb { hlac_eqclassdb* d; ~FreeDb() { hlac_eqclassdb_free(d); } } fdb{&db};
    check(hlac_eqclassdb_counts(&db, inf.n_tx, device, &ot.t));
    free_t.t = &ot.t;
  }
  if (inf.n_tx != ot.t.nitems) throw Error(HLAC_ERR_ARG, "counts file does not match the index");
  std::vector<double> theta(std::max<uint32_t>(ot.t.nitems, 1)); uint32_t iters = 0;
  check(hlac_squarem(&ot.t, device, theta.data(), &iters));
  hlac_calls calls{}; check(hlac_call_genotypes(ix.p, &ot.t, theta.data(), &calls));
  struct Free { hlac_calls* c; ~Free() { hlac_calls_free(c); } } fr{&calls};
  {
    std::ofstream w(join(outdir, "weights.tsv")); if (!w) throw Error(HLAC_ERR_IO, "cannot write weights.tsv");
    w << "allele_name\tem_weight\treads_explained\n";
    for (uint32_t i = 0; i < calls.n_weights; i++) w << hlac_index_tx_name(ix.p, calls.w_tx[i]) << '\t' << fmt_f64(calls.w_em[i]) << '\t' << fmt_f64(calls.w_frac[i]) << '\n';
    std::ofstream p(join(outdir, "pairs.tsv")); if (!p) throw Error(HLAC_ERR_IO, "cannot write pairs.tsv");
    p << "allele_1\tallele_2\treads_explained\n";
    for (uint32_t i = 0; i < calls.n_pairs; i++) p << hlac_index_tx_name(ix.p, calls.p_a[i]) << '\t' << hlac_index_tx_name(ix.p, calls.p_b[i]) << '\t' << fmt_f64(calls.p_frac[i]) << '\n';
  }
  std::set<std::string> called;
  for (uint32_t i = 0; i < calls.n_called; i++) called.insert(hlac_index_tx_name(ix.p, calls.called[i]));
  info("Writing CDS of " + std::to_string(called.size()) + " alleles to file");
  gen_name = join(outdir, "gen_pseudoaln.fasta"); cds_name = join(outdir, "cds_pseudoaln.fasta");
  auto filter = [&](const std::string& src, const std::string& dst) {   // em.rs:445-473
    std::ofstream o(dst); if (!o) throw Error(HLAC_ERR_IO, "cannot write " + dst);
    for (auto& rec : read_fasta(src)) {
      if (rec.desc.empty()) throw Error(HLAC_ERR_FORMAT, "no HLA allele");
      if (called.count(rec.desc.substr(0, rec.desc.find(' ')))) write_fasta_record(o, rec);
    }
  };
  filter(cds_db, cds_name); filter(gen_db, gen_name);
}

}  // namespace

extern "C" {

// src/mapping.rs:310-405
int hlac_mapping_wrapper(const char* hla_index, const char* outdir, const char* bam, const char* region, int use_unmapped,
                         int device, char* counts_path_out, size_t cap) {
  HLAC_API_TRY
  if (!hla_index || !outdir || !bam) throw Error(HLAC_ERR_ARG, "null argument");
  info("Reading database index from disk");
  IndexHandle ix; check(hlac_index_from_idx_file(hla_index, device, &ix.p));
  info("Finished reading index!");
  copy_out(mapping_impl({ix.p}, outdir, bam, region, use_unmapped), counts_path_out, cap);
  HLAC_API_CATCH
}

int hlac_mapping_wrapper_multi(const char* hla_index, const char* outdir, const char* bam, const char* region, int use_unmapped,
                               const int* devices, int n_devices, char* counts_path_out, size_t cap) {
  HLAC_API_TRY
  if (!hla_index || !outdir || !bam || !devices || n_devices < 1) throw Error(HLAC_ERR_ARG, "null argument");
  info("Reading database index from disk");
  std::vector<IndexHandle> ix(n_devices);
  std::vector<hlac_index*> raw(n_devices);
  check(hlac_index_from_idx_file(hla_index, devices[0], &ix[0].p));
  for (int d = 1; d < n_devices; d++) check(hlac_index_clone(ix[0].p, devices[d], &ix[d].p));
  for (int d = 0; d < n_devices; d++) raw[d] = ix[d].p;
  info("Finished reading index!");
  copy_out(mapping_impl(raw, outdir, bam, region, use_unmapped), counts_path_out, cap);
  HLAC_API_CATCH
}

// src/em.rs:339-476
int hlac_em_wrapper(const char* hla_index, const char* hla_counts, const char* cds_db, const char* gen_db, const char* outdir,
                    int device, char* gen_path_out, char* cds_path_out, size_t cap) {
  HLAC_API_TRY
  if (!hla_index || !hla_counts || !cds_db || !gen_db || !outdir) throw Error(HLAC_ERR_ARG, "null argument");
  IndexHandle ix; check(hlac_index_from_idx_file(hla_index, -1, &ix.p));   // names only; the EM itself runs on `device`
  std::string g, c;
  em_impl(ix.p, hla_counts, cds_db, gen_db, outdir, device, g, c);
  copy_out(g, gen_path_out, cap); copy_out(c, cds_path_out, cap);
  HLAC_API_CATCH
}

// src/mapping.rs:640-821 on one or several devices. With several, reads are routed by barcode column (column % n) so that
// every (cell, UMI) molecule is complete on one device; every device holds a session over ALL columns, and
// hlac_count_finish sums the matrices onto devices[0] through the communicator the sessions share.
static void map_and_count_impl(const char* bam, const char* out_dir, const char* barcodes_path, const char* region, const char* cds,
                               const char* genomic, int primary_only, const std::vector<int>& devices, hlac_coo* entries,
                               hlac_metrics* metrics, uint32_t* n_cells_out) {
  const size_t nd = devices.size();
  if (nd == 0) throw Error(HLAC_ERR_ARG, "no device given");
  for (size_t i = 0; i < nd; i++) for (size_t j = 0; j < i; j++) if (devices[i] == devices[j]) throw Error(HLAC_ERR_ARG, "a device is listed twice");
  std::vector<IndexHandle> gix(nd), cix(nd);
  check(hlac_index_from_fasta(genomic, nullptr, devices[0], &gix[0].p)); info("Built genome index for pseudoalignment");
  check(hlac_index_from_fasta(cds, nullptr, devices[0], &cix[0].p)); info("Built CDS index for pseudoalignment");
  for (size_t d = 1; d < nd; d++) { check(hlac_index_clone(gix[0].p, devices[d], &gix[d].p)); check(hlac_index_clone(cix[0].p, devices[d], &cix[d].p)); }
  auto barcodes = load_barcodes(barcodes_path);
  const uint32_t n_cells = (uint32_t)barcodes.size();
  uint32_t max_col = 0; for (auto& kv : barcodes) max_col = std::max(max_col, kv.second);
  // The reference keeps the LAST line index of a duplicated barcode (main.rs:412-416) and sizes the matrix by the number of
  // distinct barcodes, so a count landing in a column >= that size panics in TriMat::add_triplet (main.rs:279) - but only if
  // such a count occurs. Here the file is rejected up front: a deliberate, stricter check (README.md).
  if (max_col >= n_cells) {
    throw Error(HLAC_ERR_ARG, "duplicate lines in the barcode file " + std::string(barcodes_path) + ": " + std::to_string(max_col + 1) + " lines but " + std::to_string(n_cells) +
                              " distinct barcodes (the reference would index past its matrix for the later ones, main.rs:279,412-416); remove the duplicates");
  }
  struct CommHandle { hlac_comm* p = nullptr; ~CommHandle() { if (p) hlac_comm_free(p); } };
  std::vector<CountHandle> sess(nd);
  std::vector<CommHandle> comms(nd);
  for (size_t d = 0; d < nd; d++) check(hlac_count_new(cix[d].p, gix[d].p, n_cells, primary_only, &sess[d].p));
  if (nd > 1) {
    std::vector<hlac_comm*> raw(nd, nullptr);
    check(hlac_comm_init_all(devices.data(), (int)nd, raw.data()));
    for (size_t d = 0; d < nd; d++) { comms[d].p = raw[d]; check(hlac_count_attach_comm(sess[d].p, raw[d], 0)); }
    info("Counting on " + std::to_string(nd) + " GPUs: reads routed by barcode, matrices summed onto device " + std::to_string(devices[0]));
  }
  { std::ofstream lab(join(out_dir, "labels.tsv")); if (!lab) throw Error(HLAC_ERR_IO, "cannot write labels.tsv");
    for (uint32_t i = 0; i < hlac_count_nrows(sess[0].p); i++) lab << hlac_count_row_name(sess[0].p, i) << '\n'; }
  BamReader reader(bam);
  Locus l = parse_locus(region); reader.fetch(l.chrom, l.start, l.end);
  std::vector<std::unique_ptr<BatchBuilder>> bbs;
  std::vector<BatchBuilder*> lanes;
  for (size_t d = 0; d < nd; d++) { bbs.emplace_back(new BatchBuilder(BATCH_READS)); lanes.push_back(bbs.back().get()); }
  // the whitelist is read-only here: the lookup is a pure function and runs on the parser threads
  auto cell_of = [&](StrRef cb, StrRef) { auto it = barcodes.find(cb.str()); return it == barcodes.end() ? HLAC_NOT_A_CELL : it->second; };
  HostFilter filt; filt.primary_only = primary_only != 0;
  stream_records_routed(reader, lanes, [nd](uint32_t cell, uint64_t) { return (size_t)(cell % nd); }, cell_of,
                        [&](size_t d) { hlac_batch b = lanes[d]->batch(); check(hlac_count_push(sess[d].p, &b)); check(hlac_count_sync(sess[d].p)); },
                        true, NoCellTry(), &filt);
  check(hlac_count_add_filtered(sess[0].p, filt.n_reads, filt.n_non_primary, filt.n_not_cell));
  hlac_metrics m{}; hlac_coo tmp{};
  if (nd == 1) check(hlac_count_finish(sess[0].p, &tmp, &m));
  else {   // the finish is a collective: one host thread per device
    std::vector<std::future<void>> fut;
    std::vector<std::string> errs(nd);
    std::vector<int> rcs(nd, HLAC_OK);
    for (size_t d = 1; d < nd; d++)
      fut.push_back(std::async(std::launch::async, [&, d] { hlac_coo c{}; hlac_metrics mm{}; rcs[d] = hlac_count_finish(sess[d].p, &c, &mm); if (rcs[d] != HLAC_OK) errs[d] = hlac_last_error(); }));
    rcs[0] = hlac_count_finish(sess[0].p, &tmp, &m);
    if (rcs[0] != HLAC_OK) errs[0] = hlac_last_error();
    for (auto& f : fut) f.get();
    for (size_t d = 0; d < nd; d++) if (rcs[d] != HLAC_OK) throw Error(rcs[d], "device " + std::to_string(devices[d]) + ": " + errs[d]);
  }
  if (entries) {   // the session (and its pinned result buffer) ends with this call: hand out owned copies
    *entries = tmp; entries->borrowed = 0;
    const size_t bytes = std::max<uint64_t>(tmp.n, 1) * 4;
    entries->row = (uint32_t*)malloc(bytes); entries->col = (uint32_t*)malloc(bytes); entries->val = (uint32_t*)malloc(bytes);
    if (tmp.n) { memcpy(entries->row, tmp.row, tmp.n * 4); memcpy(entries->col, tmp.col, tmp.n * 4); memcpy(entries->val, tmp.val, tmp.n * 4); }
  }
  if (metrics) *metrics = m;
  if (n_cells_out) *n_cells_out = n_cells;
  info("analyzed " + std::to_string(m.num_reads) + " reads. Mapped " + std::to_string(m.num_gen_align + m.num_cds_align) + " with score at least " + std::to_string(MIN_SCORE_COUNT_PSEUDO));
}

int hlac_map_and_count_pseudo(const char* bam, const char* out_dir, const char* barcodes_path, const char* region, const char* cds,
                              const char* genomic, int primary_only, int device, hlac_coo* entries, hlac_metrics* metrics,
                              uint32_t* n_cells_out) {
  HLAC_API_TRY
  if (!bam || !out_dir || !barcodes_path || !region || !cds || !genomic) throw Error(HLAC_ERR_ARG, "null argument");
  map_and_count_impl(bam, out_dir, barcodes_path, region, cds, genomic, primary_only, {device}, entries, metrics, n_cells_out);
  HLAC_API_CATCH
}

int hlac_map_and_count_pseudo_multi(const char* bam, const char* out_dir, const char* barcodes_path, const char* region, const char* cds,
                                    const char* genomic, int primary_only, const int* devices, int n_devices, hlac_coo* entries,
                                    hlac_metrics* metrics, uint32_t* n_cells_out) {
  HLAC_API_TRY
  if (!bam || !out_dir || !barcodes_path || !region || !cds || !genomic || !devices || n_devices < 1) throw Error(HLAC_ERR_ARG, "null argument");
  map_and_count_impl(bam, out_dir, barcodes_path, region, cds, genomic, primary_only, std::vector<int>(devices, devices + n_devices), entries, metrics, n_cells_out);
  HLAC_API_CATCH
}

// Host-reader self check: number of CB+UB records the selection yields and an FNV-1a checksum over
// (2-bit sequence, CB, UB, primary) in file order. region == NULL selects the unmapped tail ("*").
int hlac_bam_scan(const char* bam, const char* region, uint64_t* n_records, uint64_t* checksum) {
  HLAC_API_TRY
  if (!bam) throw Error(HLAC_ERR_ARG, "null argument");
  BamReader reader(bam);
  if (region) { Locus l = parse_locus(region); reader.fetch(l.chrom, l.start, l.end); } else reader.fetch_unmapped();
  uint64_t n = 0, h = 0xcbf29ce484222325ULL;
  auto mixb = [&](uint8_t b) { h = (h ^ b) * 0x100000001b3ULL; };
  BamRecord r;
  while (reader.next(r)) {
    for (uint8_t b : r.seq) mixb(b);
    mixb(0xFF); for (char c : r.barcode) mixb((uint8_t)c);
    mixb(0xFE); for (char c : r.umi) mixb((uint8_t)c);
    mixb(r.primary ? 1 : 0);
    n++;
  }
  if (n_records) *n_records = n;
  if (checksum) *checksum = h;
  HLAC_API_CATCH
}

// Self-check of the batch path the wrappers use (frame -> parallel parse / pack -> ordered append): FNV-1a over what the
// sessions receive per record - the 2-bit words that carry bases, length, cell (first-seen barcode id, as mapping_wrapper
// numbers them), UMI key, flags - in file order. Runs the selection twice, with the barcode lookup in the ordered append
// and (using the ids of the first run) on the parser threads; both must agree.
int hlac_bam_batch_checksum(const char* bam, const char* region, uint64_t* n_records, uint64_t* checksum) {
  HLAC_API_TRY
  if (!bam) throw Error(HLAC_ERR_ARG, "null argument");
  std::unordered_map<std::string, uint32_t> ids;
  uint64_t sums[2] = {0, 0}, counts[2] = {0, 0};
  for (int pass = 0; pass < (checksum ? 2 : 1); pass++) {   // checksum == NULL: one pass, nothing hashed (ingest timing)
    BamReader reader(bam);
    if (region) { Locus l = parse_locus(region); reader.fetch(l.chrom, l.start, l.end); } else reader.fetch_unmapped();
    BatchBuilder bb(1 << 10, false);   // small pageable batches: several flushes even on the fixture, no device needed
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mix = [&](const void* p, size_t k) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < k; i++) h = (h ^ b[i]) * 0x100000001b3ULL; };
    auto flush = [&] {
      if (!checksum) return;
      for (uint64_t i = 0; i < bb.n; i++) {
        mix(bb.seq + i * bb.wpr, (size_t)((bb.len[i] + 31) / 32) * 8);
        for (uint32_t j = (bb.len[i] + 31) / 32; j < bb.wpr; j++) if (bb.seq[i * bb.wpr + j]) throw Error(HLAC_ERR_STATE, "padding words are not zero");
        mix(bb.len + i, 2); mix(bb.cell + i, 4); mix(bb.umi + i, 4); mix(bb.flags + i, 1);
      }
    };
    std::string key;
    if (pass == 0) counts[0] = stream_records(reader, bb, [&](StrRef cb, StrRef) { key.assign(cb.p, cb.n); return ids.emplace(key, (uint32_t)ids.size()).first->second; }, flush, false,
                                              [&](StrRef cb) { auto it = ids.find(cb.str()); return it == ids.end() ? CELL_UNKNOWN : it->second; });
    else counts[1] = stream_records(reader, bb, [&](StrRef cb, StrRef) { return ids.at(cb.str()); }, flush, true);
    sums[pass] = h;
  }
  if (checksum && (sums[0] != sums[1] || counts[0] != counts[1])) throw Error(HLAC_ERR_STATE, "ordered and parallel barcode lookups disagree");
  if (n_records) *n_records = counts[0];
  if (checksum) *checksum = sums[0];
  HLAC_API_CATCH
}

// src/main.rs:151-298
int hlac_main(int argc, const char* const* argv) {
  HLAC_API_TRY
  std::map<std::string, std::string> opt = {{"out_dir", "hla-typer-results"}, {"pl_tmp", "pseudoaligner_tmp"}, {"region", "6:28510120-33480577"},
                                            {"log_level", "info"}, {"device", "0"}};
  std::map<std::string, bool> flag;
  static const std::map<std::string, std::string> names = {
      {"-b", "bam"}, {"--bam", "bam"}, {"-c", "cell_barcodes"}, {"--cell-barcodes", "cell_barcodes"}, {"-o", "out_dir"}, {"--out-dir", "out_dir"},
      {"--pl-tmp", "pl_tmp"}, {"-g", "fastagenomic"}, {"--fasta-genomic", "fastagenomic"}, {"-f", "fastacds"}, {"--fasta-cds", "fastacds"},
      {"-d", "hladbdir"}, {"--hladb-dir", "hladbdir"}, {"-i", "hlaindex"}, {"--hla-index", "hlaindex"}, {"-r", "region"}, {"--region", "region"},
      {"--log-level", "log_level"}, {"--device", "device"}, {"--devices", "devices"}};
  static const std::map<std::string, std::string> flags = {{"--primary-alignments", "primary"}, {"--use-exact-count", "exact"}, {"--unmapped", "unmapped"}};
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (flags.count(a)) { flag[flags.at(a)] = true; continue; }
    std::string val; size_t eq = a.find('=');
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); }
    else if (names.count(a)) { if (i + 1 >= argc) throw Error(HLAC_ERR_ARG, "missing value for " + a); val = argv[++i]; }
    if (!names.count(a)) throw Error(HLAC_ERR_ARG, "unknown argument " + a);
    opt[names.at(a)] = val;
  }
  if (!opt.count("bam")) throw Error(HLAC_ERR_ARG, "You must provide a BAM file");
  if (!opt.count("cell_barcodes")) throw Error(HLAC_ERR_ARG, "You must provide a cell barcodes file");
  if (opt["log_level"] != "info" && opt["log_level"] != "debug" && opt["log_level"] != "error") throw Error(HLAC_ERR_ARG, "Log level must be 'info', 'debug', or 'error'");
  if (flag["exact"]) throw Error(HLAC_ERR_ARG, "--use-exact-count (banded Smith-Waterman, mapping.rs:407-637) is outside this build's hot path");
  // --devices 0,1,2,3 (or 0-3): the counting pass runs on all of them; everything else on the first
  std::vector<int> devices;
  if (opt.count("devices")) {
    std::stringstream ss(opt["devices"]); std::string tok;
    while (std::getline(ss, tok, ',')) {
      const size_t dash = tok.find('-');
      if (tok.empty()) continue;
      if (dash != std::string::npos && dash > 0) { for (int d = std::stoi(tok.substr(0, dash)); d <= std::stoi(tok.substr(dash + 1)); d++) devices.push_back(d); }
      else devices.push_back(std::stoi(tok));
    }
    if (devices.empty()) throw Error(HLAC_ERR_ARG, "--devices needs a list like 0,1,2,3 or 0-3");
  } else devices.push_back(std::stoi(opt["device"]));
  const int device = devices[0];
  const std::string bam = opt["bam"], bcs = opt["cell_barcodes"], out_dir = opt["out_dir"], pl = opt["pl_tmp"];
  // check_inputs_exist (main.rs:302-348)
  for (auto& p : {bam, bcs}) if (!path_exists(p)) throw Error(HLAC_ERR_IO, "Input \"" + p + "\" does not exist");
  if (bam.size() < 4 || bam.substr(bam.size() - 4) != ".bam") throw Error(HLAC_ERR_ARG, "BAM file did not end in .bam (CRAM is not supported by this build). Unable to validate");
  if (!path_exists(bam + ".bai")) throw Error(HLAC_ERR_IO, "BAM index " + bam + ".bai does not exist");
  if (path_exists(out_dir)) throw Error(HLAC_ERR_IO, "Specified output directory " + out_dir + " already exists");
  make_dirs(out_dir);
  Locus region = parse_locus(opt["region"]); (void)region;
  std::string genomic = opt["fastagenomic"], cds = opt["fastacds"];
  if (genomic.empty() || cds.empty()) {
    if (path_exists(pl)) throw Error(HLAC_ERR_IO, "Specified tempdir " + pl + " already exists");
    make_dirs(pl);
    const std::string db = opt["hladbdir"];
    if (db.empty()) throw Error(HLAC_ERR_ARG, "Must provide either -d (database directory) or both -g and -f (FASTA sequences).");
    if (!path_exists(db)) throw Error(HLAC_ERR_IO, "IMGT-HLA database directory " + db + " does not exist");
    for (auto f : {"hla_gen.fasta", "hla_nuc.fasta", "Allele_status.txt"}) if (!path_exists(join(db, f))) throw Error(HLAC_ERR_IO, std::string("IMGT-HLA database file ") + f + " does not exist");
    const std::string index = opt["hlaindex"];
    IndexHandle ix;
    if (index.empty()) {
      // make_hla_index (hla.rs:230-264; main.rs:198-201): build, then write <pl-tmp>/hla_nuc.fasta.idx for reuse with -i
      info("Building index from fasta");
      check(hlac_index_from_fasta(join(db, "hla_nuc.fasta").c_str(), join(db, "Allele_status.txt").c_str(), device, &ix.p));
      info("Finished building index!");
      info("Writing index to disk");
      check(hlac_index_write_idx(ix.p, join(pl, "hla_nuc.fasta.idx").c_str()));
      info("Finished writing index!");
    } else {
      if (!path_exists(index)) throw Error(HLAC_ERR_IO, "Pseudoalignment index " + index + " does not exist. Omit parameter -i to generate automatically.");
      info("Reading database index from disk");
      check(hlac_index_from_idx_file(index.c_str(), device, &ix.p));
    }
    std::vector<IndexHandle> more(devices.size());
    std::vector<hlac_index*> ixs{ix.p};
    for (size_t d = 1; d < devices.size(); d++) { check(hlac_index_clone(ix.p, devices[d], &more[d].p)); ixs.push_back(more[d].p); }
    const std::string counts = mapping_impl(ixs, join(pl, "pseudoaligner").c_str(), bam.c_str(), opt["region"].c_str(), flag["unmapped"]);
    em_impl(ix.p, counts.c_str(), join(db, "hla_nuc.fasta").c_str(), join(db, "hla_gen.fasta").c_str(), pl.c_str(), device, genomic, cds);
  }
  for (auto& p : {cds, genomic}) if (!path_exists(p)) throw Error(HLAC_ERR_IO, "Input file \"" + p + "\" does not exist");
  hlac_coo coo{}; hlac_metrics m{}; uint32_t n_cells = 0;
  check(hlac_map_and_count_pseudo_multi(bam.c_str(), out_dir.c_str(), bcs.c_str(), opt["region"].c_str(), cds.c_str(), genomic.c_str(), flag["primary"],
                                        devices.data(), (int)devices.size(), &coo, &m, &n_cells));
  struct Free { hlac_coo* c; ~Free() { hlac_coo_free(c); } } fr{&coo};
  info("Initialized a " + std::to_string(coo.nrows) + " features x " + std::to_string(n_cells) + " cell barcodes matrix.");
  info("Number of alignments evaluated (with 'CB' and 'UB' tags): " + std::to_string(m.num_reads));
  info(std::string(flag["primary"] ? "Number of alignments skipped due to not being primary: " : "Number of alignments that were not primary: ") + std::to_string(m.num_non_primary));
  info("Number of alignments skipped due to not being associated with a cell barcode in the list provided: " + std::to_string(m.num_not_cell_bc));
  info("Number of reads with no alignment score above threshold: " + std::to_string(m.num_not_aligned));
  info("Number of alignments to CDS sequence: " + std::to_string(m.num_cds_align));
  info("Number of alignments to genomic sequence: " + std::to_string(m.num_gen_align));
  // sprs::io::write_matrix_market of the TriMat (main.rs:278-283): 1-based triplets in insertion order
  { std::ofstream mm(join(out_dir, "count_matrix.mtx")); if (!mm) throw Error(HLAC_ERR_IO, "cannot write count_matrix.mtx");
    mm << "%%MatrixMarket matrix coordinate integer general\n% written by sprs\n" << coo.nrows << ' ' << n_cells << ' ' << coo.n << '\n';
    for (uint64_t i = 0; i < coo.n; i++) mm << coo.row[i] + 1 << ' ' << coo.col[i] + 1 << ' ' << coo.val[i] << '\n'; }
  // summary.tsv: every row with its molecule total (main.rs:286-296); row names = labels.tsv
  { std::vector<uint64_t> sums(coo.nrows, 0);
    for (uint64_t i = 0; i < coo.n; i++) sums[coo.row[i]] += coo.val[i];
    std::ifstream lab(join(out_dir, "labels.tsv")); std::ofstream sm(join(out_dir, "summary.tsv"));
    std::string name; for (uint32_t r = 0; r < coo.nrows && std::getline(lab, name); r++) sm << name << '\t' << sums[r] << '\n'; }
  HLAC_API_CATCH
}

}  // extern "C"
