// Test driver for schlacount_b200/csrc/inflate_fast.cpp (built and run by tests/test_host_cpu.py): raw deflate streams made
// by zlib at every level / strategy over several kinds of data are inflated with the library's decoder and compared byte by
// byte; truncated and bit-flipped streams must be rejected or decode to something (never crash / overrun - the buffers
// carry guard bytes that are checked), and a BAM file's BGZF blocks can be given on the command line.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../schlacount_b200/csrc/inflate_fast.hpp"

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& data, int level, int strategy, int mem_level) {
  z_stream zs{};
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, mem_level, strategy) != Z_OK) { fprintf(stderr, "deflateInit2 failed\n"); exit(2); }
  std::vector<uint8_t> out(deflateBound(&zs, (uLong)data.size()) + 64);
  zs.next_in = const_cast<Bytef*>(data.data()); zs.avail_in = (uInt)data.size();
  zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "deflate failed\n"); exit(2); }
  out.resize(zs.total_out);
  deflateEnd(&zs);
  return out;
}

static const size_t GUARD = 64;
struct Guarded {
  std::vector<uint8_t> buf; size_t n;
  explicit Guarded(size_t n_) : buf(n_ + 2 * GUARD, 0xCD), n(n_) {}
  uint8_t* p() { return buf.data() + GUARD; }
  bool intact() const {
    for (size_t i = 0; i < GUARD; i++) if (buf[i] != 0xCD || buf[GUARD + n + i] != 0xCD) return false;
    return true;
  }
};

static uint64_t n_ok = 0, n_rejected = 0, n_fallback = 0;

static void check_stream(const std::vector<uint8_t>& data, const std::vector<uint8_t>& comp, const char* what) {
  // exact-size copy of the input so that an over-read would leave the allocation (ASan / valgrind) - and guard the output
  std::vector<uint8_t> in(comp);
  Guarded out(data.size());
  const bool ok = hlac::inflate_raw_fast(in.data(), in.size(), out.p(), data.size());
  if (!out.intact()) { fprintf(stderr, "FAIL %s: output guard bytes overwritten\n", what); exit(1); }
  if (!ok) { n_fallback++; fprintf(stderr, "note: %s (%zu -> %zu bytes) was declined\n", what, comp.size(), data.size()); return; }
  if (!data.empty() && memcmp(out.p(), data.data(), data.size()) != 0) { fprintf(stderr, "FAIL %s: wrong output\n", what); exit(1); }
  n_ok++;
}

static void check_damaged(const std::vector<uint8_t>& data, const std::vector<uint8_t>& comp, std::mt19937& rng) {
  for (int trial = 0; trial < 24; trial++) {
    std::vector<uint8_t> in(comp);
    if (trial % 3 == 0 && in.size() > 2) in.resize(rng() % (in.size() - 1) + 1);                 // truncated
    else if (trial % 3 == 1 && !in.empty()) in[rng() % in.size()] ^= (uint8_t)(1u << (rng() % 8));   // one bit flipped
    else for (int k = 0; k < 4 && !in.empty(); k++) in[rng() % in.size()] = (uint8_t)rng();      // a few bytes replaced
    const size_t claim = (trial % 5 == 0) ? data.size() / 2 : data.size();                       // also: a wrong size claim
    Guarded out(claim);
    const bool ok = hlac::inflate_raw_fast(in.data(), in.size(), out.p(), claim);
    if (!out.intact()) { fprintf(stderr, "FAIL: damaged stream wrote outside its buffer\n"); exit(1); }
    if (!ok) n_rejected++;
  }
}

int main(int argc, char** argv) {
  std::mt19937 rng(20191101);
  std::vector<std::vector<uint8_t>> datasets;
  datasets.push_back({});                                  // empty
  datasets.push_back({'A'});
  { std::vector<uint8_t> v(70000, 'A'); datasets.push_back(v); }                                  // one long run (distance 1)
  { std::vector<uint8_t> v(65280); for (auto& b : v) b = (uint8_t)rng(); datasets.push_back(v); }  // incompressible: stored blocks
  { std::vector<uint8_t> v(65280); for (size_t i = 0; i < v.size(); i++) v[i] = "ACGT"[rng() & 3]; datasets.push_back(v); }
  { std::vector<uint8_t> v; const char* rec = "NB501047:228:HWLCKBGX5:1:11101:10000:1234\tCB:Z:ACGTACGTACGTACGT-1\tUB:Z:ACGTACGTAC\t";
    while (v.size() < 65000) { for (const char* p = rec; *p; p++) v.push_back((uint8_t)*p); for (int k = 0; k < 46; k++) v.push_back((uint8_t)(rng() % 7 ? 0x11 * (1 + rng() % 8) : rng())); }
    datasets.push_back(v); }                                                                       // BAM-record-like
  { std::vector<uint8_t> v(40000); for (size_t i = 0; i < v.size(); i++) v[i] = (uint8_t)((i * i) >> 3); datasets.push_back(v); }
  { std::vector<uint8_t> v(300); for (size_t i = 0; i < v.size(); i++) v[i] = (uint8_t)(i % 7 + 'a'); datasets.push_back(v); }   // short periods: distances 2..7
  { std::vector<uint8_t> v(3000); for (size_t i = 0; i < v.size(); i++) v[i] = (uint8_t)(rng() % 3 ? 'x' : rng()); datasets.push_back(v); }
  for (int len : {2, 3, 7, 8, 9, 257, 258, 259, 297, 298, 299, 1000}) { std::vector<uint8_t> v(len); for (auto& b : v) b = (uint8_t)(rng() % 4 + 'A'); datasets.push_back(v); }
  const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
  char what[128];
  for (size_t d = 0; d < datasets.size(); d++)
    for (int level = 0; level <= 9; level++)
      for (int strategy : strategies)
        for (int mem_level : {1, 8, 9}) {
          const std::vector<uint8_t> comp = deflate_raw(datasets[d], level, strategy, mem_level);
          snprintf(what, sizeof what, "dataset %zu level %d strategy %d memlevel %d", d, level, strategy, mem_level);
          check_stream(datasets[d], comp, what);
          if (mem_level == 8 && (level == 1 || level == 6)) check_damaged(datasets[d], comp, rng);
        }
  // multi-block streams with full flushes in the middle (what a streaming compressor emits)
  for (size_t d = 0; d < datasets.size(); d++) {
    const std::vector<uint8_t>& data = datasets[d];
    z_stream zs{};
    deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    std::vector<uint8_t> out(deflateBound(&zs, (uLong)data.size()) + 1024);
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    size_t at = 0;
    while (at < data.size()) {
      const size_t k = std::min<size_t>(data.size() - at, 1 + rng() % 5000);
      zs.next_in = const_cast<Bytef*>(data.data() + at); zs.avail_in = (uInt)k;
      deflate(&zs, (rng() & 1) ? Z_FULL_FLUSH : Z_SYNC_FLUSH);
      at += k;
    }
    zs.next_in = nullptr; zs.avail_in = 0;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "deflate finish failed\n"); return 2; }
    out.resize(zs.total_out);
    deflateEnd(&zs);
    snprintf(what, sizeof what, "dataset %zu flushed", d);
    check_stream(data, out, what);
  }
  // BGZF blocks of real files
  for (int a = 1; a < argc; a++) {
    FILE* f = fopen(argv[a], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
    std::vector<uint8_t> raw; uint8_t tmp[1 << 16]; size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) raw.insert(raw.end(), tmp, tmp + k);
    fclose(f);
    size_t p = 0, blocks = 0;
    while (p + 18 <= raw.size()) {
      const size_t xlen = raw[p + 10] | (raw[p + 11] << 8), bsize = (raw[p + 16] | (raw[p + 17] << 8)) + 1;
      const size_t clen = bsize - 12 - xlen - 8;
      const uint8_t* c = raw.data() + p + 12 + xlen;
      uint32_t isize; memcpy(&isize, c + clen + 4, 4);
      std::vector<uint8_t> ref(isize);
      z_stream zs{}; inflateInit2(&zs, -15);
      zs.next_in = const_cast<Bytef*>(c); zs.avail_in = (uInt)clen; zs.next_out = ref.data(); zs.avail_out = isize;
      if (isize && inflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "zlib cannot inflate block %zu of %s\n", blocks, argv[a]); return 2; }
      inflateEnd(&zs);
      snprintf(what, sizeof what, "%s block %zu", argv[a], blocks);
      check_stream(ref, std::vector<uint8_t>(c, c + clen), what);
      p += bsize; blocks++;
    }
  }
  printf("ok: %llu streams decoded identically, %llu declined, %llu of the damaged streams rejected\n", (unsigned long long)n_ok,
         (unsigned long long)n_fallback, (unsigned long long)n_rejected);
  return n_fallback * 20 > n_ok ? 1 : 0;   // declining is allowed (zlib takes over) but must stay rare
}
