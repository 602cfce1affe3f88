#include "bam_reader.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstring>

namespace hlac {

namespace {
template <typename T>
T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }
}  // namespace

BamReader::BamReader(const std::string& path) {
  f_ = fopen(path.c_str(), "rb");
  if (!f_) throw Error(HLAC_ERR_IO, "cannot open BAM " + path);
  next_coff_ = 0;
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

BamReader::~BamReader() { if (f_) fclose(f_); }

bool BamReader::read_block() {
  uint8_t hdr[18];
  if (fseeko(f_, (off_t)next_coff_, SEEK_SET) != 0) return false;
  size_t k = fread(hdr, 1, 18, f_);
  if (k == 0) return false;
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
  if (bsize < 0) throw Error(HLAC_ERR_FORMAT, "BGZF block without BSIZE");
  size_t clen = bsize - 12 - xlen - 8;
  cbuf_.resize(clen + 8);
  if (fread(cbuf_.data(), 1, clen + 8, f_) != clen + 8) throw Error(HLAC_ERR_FORMAT, "BGZF truncated");
  uint32_t isize = rd<uint32_t>(&cbuf_[clen + 4]);
  block_.resize(isize);
  if (isize) {
    z_stream zs{}; zs.next_in = cbuf_.data(); zs.avail_in = (uInt)clen; zs.next_out = block_.data(); zs.avail_out = isize;
    if (inflateInit2(&zs, -15) != Z_OK) throw Error(HLAC_ERR_FORMAT, "inflateInit2 failed");
    int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) throw Error(HLAC_ERR_FORMAT, "BGZF inflate failed");
  }
  block_coff_ = next_coff_;
  next_coff_ += bsize;
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
  next_coff_ = voff >> 16;
  block_.clear(); block_pos_ = 0;
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

bool BamReader::next(BamRecord& out) {
  static const uint8_t nt16_to_2bit[16] = {0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0};   // "=ACMGRSVTWYHKDBN"
  while (!done_) {
    int32_t bs;
    if (!read_bytes(&bs, 4)) { done_ = true; break; }
    if (bs < 32) throw Error(HLAC_ERR_FORMAT, "bad BAM record size");
    rec_.resize(bs);
    if (!read_bytes(rec_.data(), bs)) throw Error(HLAC_ERR_FORMAT, "BAM record truncated");
    const uint8_t* r = rec_.data();
    const int32_t tid = rd<int32_t>(r), pos = rd<int32_t>(r + 4);
    const uint8_t l_name = r[8];
    const uint16_t n_cigar = rd<uint16_t>(r + 12), flag = rd<uint16_t>(r + 14);
    const int32_t l_seq = rd<int32_t>(r + 16);
    if (sel_tid_ >= 0) {
      if (tid > sel_tid_ || tid < 0 || (tid == sel_tid_ && pos >= sel_end_)) { done_ = true; break; }   // coordinate sorted
      if (tid < sel_tid_) continue;
    } else if (tid != -1) continue;
    size_t p = 32 + l_name;
    if (p + 4ull * n_cigar + (l_seq + 1) / 2 + l_seq > (size_t)bs) throw Error(HLAC_ERR_FORMAT, "BAM record fields overrun");
    if (sel_tid_ >= 0) {
      int64_t ref_len = 0;
      for (uint16_t c = 0; c < n_cigar; c++) {
        uint32_t op = rd<uint32_t>(r + p + 4 * c);
        uint32_t k = op & 15;
        if (k == 0 || k == 2 || k == 3 || k == 7 || k == 8) ref_len += op >> 4;
      }
      const int64_t endpos = pos + (ref_len > 0 ? ref_len : 1);
      if (!(pos < sel_end_ && endpos > sel_beg_)) continue;
    }
    p += 4ull * n_cigar;
    const uint8_t* sq = r + p;
    p += (l_seq + 1) / 2 + l_seq;
    // aux: only CB:Z and UB:Z matter (config.rs:17-18)
    bool has_cb = false, has_ub = false;
    while (p + 3 <= (size_t)bs) {
      const char t0 = r[p], t1 = r[p + 1], ty = r[p + 2];
      p += 3;
      if (ty == 'Z' || ty == 'H') {
        size_t e = p; while (e < (size_t)bs && r[e]) e++;
        if (ty == 'Z' && t0 == 'C' && t1 == 'B') { out.barcode.assign((const char*)r + p, e - p); has_cb = true; }
        if (ty == 'Z' && t0 == 'U' && t1 == 'B') { out.umi.assign((const char*)r + p, e - p); has_ub = true; }
        p = e + 1;
      } else if (ty == 'A' || ty == 'c' || ty == 'C') p += 1;
      else if (ty == 's' || ty == 'S') p += 2;
      else if (ty == 'i' || ty == 'I' || ty == 'f') p += 4;
      else if (ty == 'B') {
        if (p + 5 > (size_t)bs) break;
        const char sub = r[p]; const int32_t cnt = rd<int32_t>(r + p + 1);
        const int w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        p += 5 + (size_t)cnt * w;
      } else throw Error(HLAC_ERR_FORMAT, "bad BAM aux type");
    }
    if (!has_cb || !has_ub) continue;   // mapping.rs:254-270
    out.seq.resize(l_seq);
    for (int32_t i = 0; i < l_seq; i++) out.seq[i] = nt16_to_2bit[(sq[i >> 1] >> ((i & 1) ? 0 : 4)) & 15];
    out.primary = !((flag & 0x100) || (flag & 0x800));
    return true;
  }
  return false;
}

}  // namespace hlac
