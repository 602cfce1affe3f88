#include "bam_reader.hpp"

#include "inflate_fast.hpp"

#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>

namespace hlac {

namespace {
template <typename T>
T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }

void zlib_inflate(const std::vector<uint8_t>& cdata, std::vector<uint8_t>& out) {
  z_stream zs{};
  zs.next_in = const_cast<Bytef*>(cdata.data()); zs.avail_in = (uInt)cdata.size(); zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
  if (inflateInit2(&zs, -15) != Z_OK) throw Error(HLAC_ERR_FORMAT, "inflateInit2 failed");
  const int rc = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  if (rc != Z_STREAM_END) throw Error(HLAC_ERR_FORMAT, "BGZF inflate failed");
}
// One BGZF block. The library's own decoder (inflate_fast.cpp, ~1.5 x zlib's rate on BAM data) goes first; whatever it
// declines - and any block whose CRC32 does not come out right after it - is inflated again by zlib, and the CRC32 decides.
std::vector<uint8_t> inflate_block(const std::vector<uint8_t>& cdata, uint32_t isize, uint32_t crc) {
  static const bool use_fast = !getenv("HLAC_ZLIB_INFLATE");
  std::vector<uint8_t> out(isize);
  auto crc_ok = [&] { return (uint32_t)crc32(crc32(0L, Z_NULL, 0), out.data(), (uInt)out.size()) == crc; };
  if (isize) {
    if (use_fast && inflate_raw_fast(cdata.data(), cdata.size(), out.data(), isize) && crc_ok()) return out;
    zlib_inflate(cdata, out);
  }
  if (!crc_ok()) throw Error(HLAC_ERR_FORMAT, "BGZF block fails its CRC32");
  return out;
}
}  // namespace

// Fixed pool of inflate workers (BAM ingest off the critical path). HLAC_BAM_THREADS overrides the
// worker count; 0 or 1 inflates inline.
class InflatePool {
 public:
  InflatePool() {
    unsigned n = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("HLAC_BAM_THREADS")) n = (unsigned)std::max(0, atoi(e));
    if (n > 1) for (unsigned i = 0; i < n; i++) workers_.emplace_back([this] { run(); });
  }
  ~InflatePool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  // blocks inflated ahead of the consumer: enough (32 MB inflated) to keep the workers busy while the host driver parses
  // and appends the previous chunk of records
  size_t depth() const { return workers_.empty() ? 1 : std::max<size_t>(workers_.size() * 4, 512); }
  std::future<std::vector<uint8_t>> submit(std::vector<uint8_t> cdata, uint32_t isize, uint32_t crc) {
    auto task = std::make_shared<std::packaged_task<std::vector<uint8_t>()>>([c = std::move(cdata), isize, crc] { return inflate_block(c, isize, crc); });
    auto fut = task->get_future();
    if (workers_.empty()) (*task)();
    else { { std::lock_guard<std::mutex> l(m_); q_.push([task] { (*task)(); }); } cv_.notify_one(); }
    return fut;
  }

 private:
  void run() {
    for (;;) {
      std::function<void()> job;
      { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [this] { return stop_ || !q_.empty(); }); if (stop_ && q_.empty()) return; job = std::move(q_.front()); q_.pop(); }
      job();
    }
  }
  std::vector<std::thread> workers_;
  std::queue<std::function<void()>> q_;
  std::mutex m_;
  std::condition_variable cv_;
  bool stop_ = false;
};

BamReader::BamReader(const std::string& path) {
  f_ = fopen(path.c_str(), "rb");
  if (!f_) throw Error(HLAC_ERR_IO, "cannot open BAM " + path);
  pool_ = std::make_unique<InflatePool>();
  next_coff_ = 0; read_coff_ = 0;
  block_.clear(); block_pos_ = 0;
  uint8_t magic[4];
  if (!read_bytes(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) throw Error(HLAC_ERR_FORMAT, path + " is not a BAM file");
  int32_t l_text; if (!read_bytes(&l_text, 4)) throw Error(HLAC_ERR_FORMAT, "BAM header truncated");
  std::vector<uint8_t> text(l_text); if (l_text && !read_bytes(text.data(), l_text)) throw Error(HLAC_ERR_FORMAT, "BAM header truncated");
  int32_t n_ref; if (!read_bytes(&n_ref, 4)) throw Error(HLAC_ERR_FORMAT, "BAM header truncated");
  for (int32_t i = 0; i < n_ref; i++) {
    int32_t l_name; if (!read_bytes(&l_name, 4)) throw Error(HLAC_ERR_FORMAT, "BAM header truncated");
    std::string name(l_name, '\0'); if (!read_bytes(&name[0], l_name)) throw Error(HLAC_ERR_FORMAT, "BAM header truncated");
    if (!name.empty() && name.back() == '\0') name.pop_back();
    int32_t l_ref; if (!read_bytes(&l_ref, 4)) throw Error(HLAC_ERR_FORMAT, "BAM header truncated");
    ref_names_.push_back(name);
  }
  first_record_voff_ = (block_coff_ << 16) | block_pos_;
  if (block_pos_ == block_.size()) first_record_voff_ = next_coff_ << 16;
  // ---- .bai: linear index per reference, and the largest chunk end (where the unmapped tail starts) ----
  FILE* b = fopen((path + ".bai").c_str(), "rb");
  if (!b) throw Error(HLAC_ERR_IO, "BAM index " + path + ".bai does not exist");
  std::vector<uint8_t> bai;
  { uint8_t tmp[1 << 16]; size_t k; while ((k = fread(tmp, 1, sizeof tmp, b)) > 0) bai.insert(bai.end(), tmp, tmp + k); }
  fclose(b);
  size_t o = 0;
  auto need = [&](size_t k) { if (o + k > bai.size()) throw Error(HLAC_ERR_FORMAT, "BAM index truncated"); };
  need(8);
  if (memcmp(bai.data(), "BAI\1", 4) != 0) throw Error(HLAC_ERR_FORMAT, "not a BAI index");
  int32_t nr = rd<int32_t>(&bai[4]); o = 8;
  linear_.resize(nr);
  unmapped_voff_ = first_record_voff_;
  for (int32_t r = 0; r < nr; r++) {
    need(4); int32_t n_bin = rd<int32_t>(&bai[o]); o += 4;
    for (int32_t i = 0; i < n_bin; i++) {
      need(8); uint32_t bin = rd<uint32_t>(&bai[o]); int32_t n_chunk = rd<int32_t>(&bai[o + 4]); o += 8;
      need((size_t)n_chunk * 16);
      for (int32_t c = 0; c < n_chunk; c++) {
        uint64_t end = rd<uint64_t>(&bai[o + 16 * c + 8]);
        if (bin != 37450) unmapped_voff_ = std::max(unmapped_voff_, end);   // pseudo-bin holds counts, not offsets
      }
      o += (size_t)n_chunk * 16;
    }
    need(4); int32_t n_intv = rd<int32_t>(&bai[o]); o += 4;
    need((size_t)n_intv * 8);
    linear_[r].resize(n_intv);
    for (int32_t i = 0; i < n_intv; i++) linear_[r][i] = rd<uint64_t>(&bai[o + 8 * i]);
    o += (size_t)n_intv * 8;
  }
}

BamReader::~BamReader() { pending_.clear(); pool_.reset(); if (f_) fclose(f_); }

bool BamReader::enqueue_block() {
  if (raw_eof_) return false;
  uint8_t hdr[18];
  if (file_pos_ != read_coff_) {   // sequential reads need no seek (a seek drops the stdio buffer)
    if (fseeko(f_, (off_t)read_coff_, SEEK_SET) != 0) { raw_eof_ = true; return false; }
    file_pos_ = read_coff_;
  }
  size_t k = fread(hdr, 1, 18, f_);
  if (k == 0) { raw_eof_ = true; return false; }
  if (k < 18 || hdr[0] != 31 || hdr[1] != 139 || !(hdr[3] & 4)) throw Error(HLAC_ERR_FORMAT, "bad BGZF block header");
  uint16_t xlen = rd<uint16_t>(&hdr[10]);
  // the BC subfield is normally first (SI1=66, SI2=67, SLEN=2); handle the general case
  std::vector<uint8_t> extra(xlen);
  memcpy(extra.data(), hdr + 12, std::min<size_t>(6, xlen));
  if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f_) != (size_t)(xlen - 6)) throw Error(HLAC_ERR_FORMAT, "BGZF truncated");
  int bsize = -1;
  for (size_t p = 0; p + 4 <= extra.size();) {
    uint16_t slen = rd<uint16_t>(&extra[p + 2]);
    if (extra[p] == 66 && extra[p + 1] == 67 && slen == 2) bsize = rd<uint16_t>(&extra[p + 4]) + 1;
    p += 4 + slen;
  }
  if (bsize < 0 || bsize < 12 + xlen + 8) throw Error(HLAC_ERR_FORMAT, "BGZF block without a valid BSIZE");
  const size_t clen = bsize - 12 - xlen - 8;
  std::vector<uint8_t> cdata(clen + 8);
  if (fread(cdata.data(), 1, clen + 8, f_) != clen + 8) throw Error(HLAC_ERR_FORMAT, "BGZF truncated");
  const uint32_t crc = rd<uint32_t>(&cdata[clen]), isize = rd<uint32_t>(&cdata[clen + 4]);
  cdata.resize(clen);
  Pending pd; pd.coff = read_coff_; pd.bsize = (uint64_t)bsize;
  pd.data = pool_->submit(std::move(cdata), isize, crc);
  pending_.push_back(std::move(pd));
  read_coff_ += bsize;
  file_pos_ = read_coff_;
  return true;
}

bool BamReader::read_block() {
  while (pending_.size() < pool_->depth() && enqueue_block()) {}
  if (pending_.empty()) return false;
  Pending pd = std::move(pending_.front());
  pending_.pop_front();
  block_ = pd.data.get();   // rethrows an inflate error
  block_coff_ = pd.coff;
  next_coff_ = pd.coff + pd.bsize;
  block_pos_ = 0;
  return true;
}

bool BamReader::read_bytes(void* dst, size_t n) {
  uint8_t* d = (uint8_t*)dst;
  while (n) {
    if (block_pos_ == block_.size()) { if (!read_block()) return false; continue; }
    size_t k = std::min(n, block_.size() - block_pos_);
    memcpy(d, block_.data() + block_pos_, k);
    d += k; n -= k; block_pos_ += k;
  }
  return true;
}

void BamReader::seek_virtual(uint64_t voff) {
  pending_.clear();   // drop the blocks inflated ahead of the old position
  next_coff_ = read_coff_ = voff >> 16; raw_eof_ = false;
  block_.clear(); block_pos_ = 0;
  carry_.clear(); one_.recs.clear(); one_pos_ = 0;
  if ((voff & 0xFFFF) != 0) {
    if (!read_block()) throw Error(HLAC_ERR_FORMAT, "bad virtual offset");
    block_pos_ = std::min<size_t>(voff & 0xFFFF, block_.size());
  }
}

void BamReader::fetch(const std::string& chrom, uint64_t start, uint64_t end) {
  auto it = std::find(ref_names_.begin(), ref_names_.end(), chrom);
  if (it == ref_names_.end())
    throw Error(HLAC_ERR_ARG, "Error fetching locus " + chrom + " from BAM file. Are you using a valid contig name? Check for '6' vs 'chr6'.");
  sel_tid_ = (int)(it - ref_names_.begin()); sel_beg_ = (int64_t)start; sel_end_ = (int64_t)end; done_ = false;
  uint64_t voff = first_record_voff_;
  if ((size_t)sel_tid_ < linear_.size() && !linear_[sel_tid_].empty()) {
    size_t w = std::min<size_t>(start >> 14, linear_[sel_tid_].size() - 1);
    voff = linear_[sel_tid_][w];
    if (voff == 0) {   // empty leading windows: fall back to the first non-zero one before w, else the file start
      voff = first_record_voff_;
      for (size_t q = w; q-- > 0;) if (linear_[sel_tid_][q]) { voff = linear_[sel_tid_][q]; break; }
    }
  } else done_ = true;   // no alignments on this reference
  if (!done_) seek_virtual(voff);
}

void BamReader::fetch_unmapped() {
  sel_tid_ = -1; done_ = false;
  seek_virtual(unmapped_voff_);
}

size_t BamReader::frame(Chunk& chunk, size_t max_records) {
  std::vector<uint8_t>& buf = chunk.buf;
  chunk.recs.clear();
  buf.clear();
  buf.swap(carry_);                               // the tail the previous chunk could not use (possibly a partial record)
  std::vector<std::pair<size_t, size_t>> found;   // offsets: the buffer may still grow (and move) while framing
  size_t at = 0;
  auto fill = [&](size_t need_to) {               // make buf hold at least need_to bytes; false at end of file
    while (buf.size() < need_to) {
      if (block_pos_ == block_.size()) { if (!read_block()) return false; if (block_.empty()) continue; }
      buf.insert(buf.end(), block_.begin() + (ptrdiff_t)block_pos_, block_.end());
      block_pos_ = block_.size();
    }
    return true;
  };
  while (!done_ && found.size() < max_records) {
    if (!fill(at + 4)) { if (buf.size() > at) throw Error(HLAC_ERR_FORMAT, "BAM record truncated"); done_ = true; break; }
    const int32_t bs = rd<int32_t>(buf.data() + at);
    if (bs < 32) throw Error(HLAC_ERR_FORMAT, "bad BAM record size");
    if (!fill(at + 4 + (size_t)bs)) throw Error(HLAC_ERR_FORMAT, "BAM record truncated");
    const uint8_t* r = buf.data() + at + 4;
    const int32_t tid = rd<int32_t>(r), pos = rd<int32_t>(r + 4);
    at += 4 + (size_t)bs;
    bool keep = true;
    if (sel_tid_ >= 0) {
      if (tid > sel_tid_ || tid < 0 || (tid == sel_tid_ && pos >= sel_end_)) { done_ = true; break; }   // coordinate sorted
      if (tid < sel_tid_) keep = false;
    } else if (tid != -1) keep = false;
    if (keep) found.emplace_back(at - (size_t)bs, (size_t)bs);
    else if (found.empty() && at >= (1u << 20)) { buf.erase(buf.begin(), buf.begin() + (ptrdiff_t)at); at = 0; }   // a long run of records before the selection: do not hoard it
  }
  if (!done_ && at < buf.size()) carry_.assign(buf.begin() + (ptrdiff_t)at, buf.end());
  chunk.recs.reserve(found.size());
  for (auto& f : found) chunk.recs.push_back({buf.data() + f.first, f.second});
  return chunk.recs.size();
}

bool BamReader::parse(const uint8_t* r, size_t bs, RecView& out) const {
  const int32_t pos = rd<int32_t>(r + 4);
  const uint8_t l_name = r[8];
  const uint16_t n_cigar = rd<uint16_t>(r + 12), flag = rd<uint16_t>(r + 14);
  const int32_t l_seq = rd<int32_t>(r + 16);
  size_t p = 32 + (size_t)l_name;
  if (l_seq < 0 || p + 4ull * n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq > bs) throw Error(HLAC_ERR_FORMAT, "BAM record fields overrun");
  if (sel_tid_ >= 0) {
    int64_t ref_len = 0;
    for (uint16_t c = 0; c < n_cigar; c++) {
      const uint32_t op = rd<uint32_t>(r + p + 4 * c), k = op & 15;
      if (k == 0 || k == 2 || k == 3 || k == 7 || k == 8) ref_len += op >> 4;
    }
    const int64_t endpos = pos + (ref_len > 0 ? ref_len : 1);
    if (!(pos < sel_end_ && endpos > sel_beg_)) return false;
  }
  p += 4ull * n_cigar;
  out.seq4 = r + p; out.l_seq = l_seq;
  p += ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
  // aux: only CB:Z and UB:Z matter (config.rs:17-18)
  bool has_cb = false, has_ub = false;
  while (p + 3 <= bs) {
    const char t0 = (char)r[p], t1 = (char)r[p + 1], ty = (char)r[p + 2];
    p += 3;
    if (ty == 'Z' || ty == 'H') {
      const void* z = memchr(r + p, 0, bs - p);
      const size_t e = z ? (size_t)((const uint8_t*)z - r) : bs;
      if (ty == 'Z' && t0 == 'C' && t1 == 'B') { out.cb = (const char*)r + p; out.cb_len = (uint32_t)(e - p); has_cb = true; }
      if (ty == 'Z' && t0 == 'U' && t1 == 'B') { out.ub = (const char*)r + p; out.ub_len = (uint32_t)(e - p); has_ub = true; }
      p = e + 1;
    } else {
      size_t w;   // fixed-width values (SAM spec 4.2.4; 'd' is htslib's historical double)
      if (ty == 'A' || ty == 'c' || ty == 'C') w = 1;
      else if (ty == 's' || ty == 'S') w = 2;
      else if (ty == 'i' || ty == 'I' || ty == 'f') w = 4;
      else if (ty == 'd') w = 8;
      else if (ty == 'B') {
        if (p + 5 > bs) throw Error(HLAC_ERR_FORMAT, "BAM aux array header overruns the record");
        const char sub = (char)r[p]; const int32_t cnt = rd<int32_t>(r + p + 1);
        const size_t ew = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
        if (ew == 0) throw Error(HLAC_ERR_FORMAT, "bad BAM aux array element type");
        if (cnt < 0 || (uint64_t)cnt * ew > bs - p - 5) throw Error(HLAC_ERR_FORMAT, "BAM aux array overruns the record");
        w = 5 + (size_t)cnt * ew;
      } else throw Error(HLAC_ERR_FORMAT, "bad BAM aux type");
      if (w > bs - p) throw Error(HLAC_ERR_FORMAT, "BAM aux value overruns the record");
      p += w;
    }
  }
  if (!has_cb || !has_ub) return false;   // mapping.rs:254-270
  out.primary = !((flag & 0x100) || (flag & 0x800));
  return true;
}

void BamReader::pack_2bit(const uint8_t* seq4, int32_t l_seq, uint64_t* words, uint32_t nwords) {
  // one table lookup per byte = two bases: "=ACMGRSVTWYHKDBN" -> A C G T = 0 1 2 3, everything else 0
  static const struct Lut { uint8_t v[256]; Lut() { static const uint8_t n2[16] = {0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0};
                                                     for (int b = 0; b < 256; b++) v[b] = (uint8_t)((n2[b >> 4] << 2) | n2[b & 15]); } } lut;
  for (uint32_t w = 0; w < nwords; w++) words[w] = 0;
  const int32_t full = l_seq >> 1;   // bytes holding two bases
  for (int32_t b = 0; b < full; b++) {
    const int32_t i = 2 * b;
    if ((uint32_t)(i >> 5) >= nwords) return;
    words[i >> 5] |= (uint64_t)lut.v[seq4[b]] << (60 - 2 * (i & 31));
  }
  if ((l_seq & 1) && (uint32_t)((l_seq - 1) >> 5) < nwords) {
    const int32_t i = l_seq - 1;
    words[i >> 5] |= (uint64_t)(lut.v[seq4[full]] >> 2) << (62 - 2 * (i & 31));
  }
}

bool BamReader::next(BamRecord& out) {
  static const uint8_t nt16_to_2bit[16] = {0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0};   // "=ACMGRSVTWYHKDBN"
  for (;;) {
    if (one_pos_ == one_.recs.size()) { one_pos_ = 0; if (frame(one_, 4096) == 0) { one_.recs.clear(); return false; } }
    const RawRec rr = one_.recs[one_pos_++];
    RecView v;
    if (!parse(rr.p, rr.n, v)) continue;
    out.barcode.assign(v.cb, v.cb_len); out.umi.assign(v.ub, v.ub_len);
    out.seq.resize(v.l_seq);
    for (int32_t i = 0; i < v.l_seq; i++) out.seq[i] = nt16_to_2bit[(v.seq4[i >> 1] >> ((i & 1) ? 0 : 4)) & 15];
    out.primary = v.primary;
    return true;
  }
}

}  // namespace hlac
