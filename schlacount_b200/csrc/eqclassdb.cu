// The reference's `pseudoaligner.counts.bin` (mapping.rs:403 write_obj / em.rs:347 read_obj): bincode 1.x, default
// config (little-endian, fixed-width integers, u64 length prefixes) of
//   struct EqClassDb { barcodes: HashMap<SmallVec<[u8;24]>, u32>, umis: HashMap<SmallVec<[u8;16]>, u32>,
//                      eq_classes: HashMap<SmallVec<[u32;4]>, u32>, counts: Vec<BusCount{barcode_id, umi_id, eq_class_id}> }
// (mapping.rs:77-90, config.rs:23-26). A HashMap serialises as len + (key, value) pairs in iteration order, which is
// arbitrary in the reference (RandomState), so any order is a valid file; this writer emits id order.
//
// hlac_eqclassdb_counts is EqClassDb::eq_class_counts (mapping.rs:169-228) on the device for such a triple list: it
// lets the GPU em_wrapper consume a counts file written by the reference's mapping_wrapper. The grouping by
// (barcode, umi), the per-UMI lowest class id and the per-class histograms run on the GPU; only the
// sort + dedup of the (few) distinct class vectors into counts_umi keys is host code.

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <vector>

#include "common.hpp"
#include "cuda_util.hpp"
#include "radix_sort.cuh"

using namespace hlac;

namespace {

struct Cursor {
  const uint8_t* p; const uint8_t* e; const std::string& path;
  void need(uint64_t n) const { if ((uint64_t)(e - p) < n) throw Error(HLAC_ERR_FORMAT, path + ": truncated EqClassDb (bincode)"); }
  uint64_t u64() { need(8); uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
  // a length prefix for entries of at least `each` bytes
  uint64_t count(uint64_t each) { const uint64_t n = u64(); if (n > (uint64_t)(e - p) / each) throw Error(HLAC_ERR_FORMAT, path + ": truncated EqClassDb (bincode)"); return n; }
};

// HashMap<SmallVec<[u8; N]>, u32> -> strings by id
void read_string_map(Cursor& c, uint32_t& n_out, uint64_t*& off_out, uint8_t*& bytes_out) {
  const uint64_t n = c.count(12);   // every entry has at least a length and an id
  if (n >= NONE32) throw Error(HLAC_ERR_LIMIT, c.path + ": too many map entries");
  std::vector<std::pair<const uint8_t*, uint64_t>> by_id(n, {nullptr, 0});
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t len = c.u64(); if (len > (uint64_t)(c.e - c.p)) c.need(len); c.need(len + 4);
    const uint8_t* s = c.p; c.p += len;
    const uint32_t id = c.u32();
    if (id >= n || by_id[id].first) throw Error(HLAC_ERR_FORMAT, c.path + ": ids of a HashMap are not a permutation of 0..n");
    by_id[id] = {s, len}; total += len;
  }
  off_out = (uint64_t*)calloc(n + 1, 8); bytes_out = (uint8_t*)malloc(std::max<uint64_t>(total, 1));
  uint64_t q = 0;
  for (uint64_t i = 0; i < n; i++) { memcpy(bytes_out + q, by_id[i].first, by_id[i].second); q += by_id[i].second; off_out[i + 1] = q; }
  n_out = (uint32_t)n;
}

void put64(std::string& o, uint64_t v) { o.append((const char*)&v, 8); }
void put32(std::string& o, uint32_t v) { o.append((const char*)&v, 4); }

__global__ void k_make_keys(const uint32_t* __restrict__ bc, const uint32_t* __restrict__ umi, uint64_t n, uint64_t* keys) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((uint64_t)bc[i] << 32) | umi[i];
}
// counts_reads[class]++ per read (mapping.rs:199-201)
__global__ void k_class_hist(const uint32_t* __restrict__ cls, uint64_t n, uint32_t* hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(hist + cls[i], 1u);
}
// one thread per (barcode, umi) run of the sorted triples: the UMI's class is its lowest class id (mapping.rs:184-193:
// buf starts as the first hit's class after the (bc, umi, class id) sort; later hits only permute it)
__global__ void k_umi_lowest(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cls, uint64_t n, uint32_t* umi_hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  if (i > 0 && keys[i - 1] == k) return;
  uint32_t best = cls[i];
  for (uint64_t j = i + 1; j < n && keys[j] == k; j++) best = min(best, cls[j]);
  atomicAdd(umi_hist + best, 1u);
}
// reads_explained[t] += reads of the class, for every member t (mapping.rs:196-198); one warp per class
__global__ void k_explained(const uint64_t* __restrict__ off, const uint32_t* __restrict__ tx, const uint32_t* __restrict__ reads, uint32_t nc,
                            unsigned long long* expl) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (c >= nc) return;
  const unsigned long long t = reads[c];
  if (t == 0) return;
  for (uint64_t q = off[c] + lane; q < off[c + 1]; q += 32) atomicAdd(expl + tx[q], t);
}

}  // namespace

#define HLAC_API_TRY try {
#define HLAC_API_CATCH                                                                       \
  }                                                                                          \
  catch (const hlac::Error& e) { hlac::set_last_error(e.what()); return e.code; }           \
  catch (const std::exception& e) { hlac::set_last_error(e.what()); return HLAC_ERR_STATE; } \
  return HLAC_OK;

extern "C" {

int hlac_eqclassdb_free(hlac_eqclassdb* db) {
  if (!db) return HLAC_OK;
  free(db->barcode_off); free(db->barcode_bytes); free(db->umi_off); free(db->umi_bytes); free(db->class_off); free(db->class_tx);
  free(db->bc); free(db->umi); free(db->cls);
  memset(db, 0, sizeof *db);
  return HLAC_OK;
}

int hlac_eqclassdb_read(const char* path, hlac_eqclassdb* out) {
  HLAC_API_TRY
  if (!path || !out) throw Error(HLAC_ERR_ARG, "null argument");
  memset(out, 0, sizeof *out);
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw Error(HLAC_ERR_IO, std::string("cannot open ") + path);
  const std::streamoff fsize = f.tellg();
  if (fsize < 0) throw Error(HLAC_ERR_IO, std::string("cannot size ") + path);
  std::string data((size_t)fsize, '\0');
  f.seekg(0);
  if (fsize && !f.read(&data[0], fsize)) throw Error(HLAC_ERR_IO, std::string("cannot read ") + path);
  const std::string spath(path);
  Cursor c{(const uint8_t*)data.data(), (const uint8_t*)data.data() + data.size(), spath};
  struct Guard { hlac_eqclassdb* d; bool ok = false; ~Guard() { if (!ok) hlac_eqclassdb_free(d); } } guard{out};
  read_string_map(c, out->n_barcodes, out->barcode_off, out->barcode_bytes);
  read_string_map(c, out->n_umis, out->umi_off, out->umi_bytes);
  {   // eq_classes: HashMap<SmallVec<[u32;4]>, u32>
    const uint64_t n = c.count(12);
    if (n >= NONE32) throw Error(HLAC_ERR_LIMIT, spath + ": too many classes");
    std::vector<std::pair<const uint8_t*, uint64_t>> by_id(n, {nullptr, 0});
    std::vector<uint8_t> seen(n, 0);
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; i++) {
      const uint64_t len = c.u64();
      if (len > (uint64_t)(c.e - c.p) / 4) throw Error(HLAC_ERR_FORMAT, spath + ": truncated EqClassDb (bincode)");
      const uint8_t* s = c.p; c.p += len * 4;
      const uint32_t id = c.u32();
      if (id >= n || seen[id]) throw Error(HLAC_ERR_FORMAT, spath + ": eq_class ids are not a permutation of 0..n");
      seen[id] = 1; by_id[id] = {s, len}; total += len;
    }
    out->class_off = (uint64_t*)calloc(n + 1, 8); out->class_tx = (uint32_t*)malloc(std::max<uint64_t>(total, 1) * 4);
    uint64_t q = 0;
    for (uint64_t i = 0; i < n; i++) { memcpy(out->class_tx + q, by_id[i].first, by_id[i].second * 4); q += by_id[i].second; out->class_off[i + 1] = q; }
    out->n_classes = (uint32_t)n;
  }
  {   // counts: Vec<BusCount>
    const uint64_t n = c.count(12);
    out->bc = (uint32_t*)malloc(std::max<uint64_t>(n, 1) * 4); out->umi = (uint32_t*)malloc(std::max<uint64_t>(n, 1) * 4);
    out->cls = (uint32_t*)malloc(std::max<uint64_t>(n, 1) * 4);
    for (uint64_t i = 0; i < n; i++) {
      uint32_t t[3]; memcpy(t, c.p + i * 12, 12);
      if (t[0] >= out->n_barcodes || t[1] >= out->n_umis || t[2] >= out->n_classes) throw Error(HLAC_ERR_FORMAT, spath + ": BusCount id out of range");
      out->bc[i] = t[0]; out->umi[i] = t[1]; out->cls[i] = t[2];
    }
    c.p += n * 12;
    out->n_counts = n;
  }
  if (c.p != c.e) throw Error(HLAC_ERR_FORMAT, spath + ": trailing bytes after EqClassDb");
  guard.ok = true;
  HLAC_API_CATCH
}

int hlac_eqclassdb_write(const char* path, const hlac_eqclassdb* db) {
  HLAC_API_TRY
  if (!path || !db) throw Error(HLAC_ERR_ARG, "null argument");
  std::string o;
  auto strings = [&](uint32_t n, const uint64_t* off, const uint8_t* bytes) {
    put64(o, n);
    for (uint32_t i = 0; i < n; i++) { put64(o, off[i + 1] - off[i]); o.append((const char*)bytes + off[i], off[i + 1] - off[i]); put32(o, i); }
  };
  strings(db->n_barcodes, db->barcode_off, db->barcode_bytes);
  strings(db->n_umis, db->umi_off, db->umi_bytes);
  put64(o, db->n_classes);
  for (uint32_t i = 0; i < db->n_classes; i++) {
    put64(o, db->class_off[i + 1] - db->class_off[i]);
    o.append((const char*)(db->class_tx + db->class_off[i]), (db->class_off[i + 1] - db->class_off[i]) * 4);
    put32(o, i);
  }
  put64(o, db->n_counts);
  for (uint64_t i = 0; i < db->n_counts; i++) { put32(o, db->bc[i]); put32(o, db->umi[i]); put32(o, db->cls[i]); }
  std::ofstream f(path, std::ios::binary);
  if (!f) throw Error(HLAC_ERR_IO, std::string("cannot write ") + path);
  f.write(o.data(), (std::streamsize)o.size());
  if (!f) throw Error(HLAC_ERR_IO, std::string("short write to ") + path);
  HLAC_API_CATCH
}

int hlac_eqclassdb_counts(const hlac_eqclassdb* db, uint32_t nitems, int device, hlac_class_tables* out) {
  HLAC_API_TRY
  if (!db || !out) throw Error(HLAC_ERR_ARG, "null argument");
  if (device < 0) throw Error(HLAC_ERR_CUDA, "eq_class_counts needs a CUDA device, there is no CPU fallback");
  const uint32_t nc = db->n_classes;
  const uint64_t n = db->n_counts;
  if (n >= NONE32) throw Error(HLAC_ERR_LIMIT, "nreads is a u32 in the reference (mapping.rs:174)");
  const uint64_t ntx = nc ? db->class_off[nc] : 0;
  for (uint64_t q = 0; q < ntx; q++) if (db->class_tx[q] >= nitems) throw Error(HLAC_ERR_ARG, "counts file does not match the index (class member out of range)");
  for (uint64_t i = 0; i < n; i++) if (db->cls[i] >= nc) throw Error(HLAC_ERR_ARG, "BusCount eq_class_id out of range");   // ids index device arrays
  DeviceGuard g(device);
  cudaStream_t stream; HLAC_CUDA(cudaStreamCreate(&stream));
  struct SG { cudaStream_t s; ~SG() { cudaStreamDestroy(s); } } sg{stream};
  std::vector<uint32_t> reads_hist(std::max<uint32_t>(nc, 1), 0), umi_hist(std::max<uint32_t>(nc, 1), 0);
  std::vector<unsigned long long> expl(std::max<uint32_t>(nitems, 1), 0);
  if (n && nc) {
    DevBuf<uint32_t> d_bc, d_umi, d_cls, d_cls_sorted, d_rh, d_uh, d_tx; DevBuf<uint64_t> d_keys, d_keys_sorted, d_off; DevBuf<uint8_t> tmp;
    DevBuf<unsigned long long> d_expl;
    d_bc.upload(db->bc, n, stream); d_umi.upload(db->umi, n, stream); d_cls.upload(db->cls, n, stream);
    d_keys.reserve(n, stream); d_keys_sorted.reserve(n, stream); d_cls_sorted.reserve(n, stream);
    d_rh.reserve(nc, stream); d_uh.reserve(nc, stream); d_expl.reserve(std::max<uint32_t>(nitems, 1), stream);
    HLAC_CUDA(cudaMemsetAsync(d_rh.p, 0, (size_t)nc * 4, stream)); HLAC_CUDA(cudaMemsetAsync(d_uh.p, 0, (size_t)nc * 4, stream));
    HLAC_CUDA(cudaMemsetAsync(d_expl.p, 0, (size_t)std::max<uint32_t>(nitems, 1) * 8, stream));
    d_off.upload(db->class_off, (size_t)nc + 1, stream); d_tx.upload(db->class_tx, std::max<uint64_t>(ntx, 1), stream);
    const unsigned grid = (unsigned)((n + 255) / 256);
    k_make_keys<<<grid, 256, 0, stream>>>(d_bc.p, d_umi.p, n, d_keys.p);
    k_class_hist<<<grid, 256, 0, stream>>>(d_cls.p, n, d_rh.p);
    HLAC_CUDA(cudaGetLastError());
    RadixSortScratch rs;
    const bool second = radix_sort_pairs(d_keys.p, d_keys_sorted.p, d_cls.p, d_cls_sorted.p, (uint64_t)n, 0, 64, rs, stream);
    k_umi_lowest<<<grid, 256, 0, stream>>>(second ? d_keys_sorted.p : d_keys.p, second ? d_cls_sorted.p : d_cls.p, n, d_uh.p);
    k_explained<<<(unsigned)(((uint64_t)nc * 32 + 255) / 256), 256, 0, stream>>>(d_off.p, d_tx.p, d_rh.p, nc, d_expl.p);
    HLAC_CUDA(cudaGetLastError());
    HLAC_CUDA(cudaMemcpyAsync(reads_hist.data(), d_rh.p, (size_t)nc * 4, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaMemcpyAsync(umi_hist.data(), d_uh.p, (size_t)nc * 4, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaMemcpyAsync(expl.data(), d_expl.p, (size_t)nitems * 8, cudaMemcpyDeviceToHost, stream));
    HLAC_CUDA(cudaStreamSynchronize(stream));
  }
  memset(out, 0, sizeof *out);
  out->nitems = nitems; out->nreads = (uint32_t)n;
  out->reads_explained = (uint64_t*)calloc(std::max<uint32_t>(nitems, 1), 8);
  for (uint32_t t = 0; t < nitems; t++) out->reads_explained[t] = expl[t];
  // counts_reads: the classes that were hit, in id (= first-seen) order
  uint64_t nr = 0, nrtx = 0;
  for (uint32_t c = 0; c < nc; c++) if (reads_hist[c]) { nr++; nrtx += db->class_off[c + 1] - db->class_off[c]; }
  out->n_read_classes = nr;
  out->reads_off = (uint64_t*)calloc(nr + 1, 8); out->reads_tx = (uint32_t*)malloc(std::max<uint64_t>(nrtx, 1) * 4); out->reads_cnt = (uint32_t*)calloc(std::max<uint64_t>(nr, 1), 4);
  uint64_t k = 0, q = 0;
  for (uint32_t c = 0; c < nc; c++) if (reads_hist[c]) {
    const uint64_t len = db->class_off[c + 1] - db->class_off[c];
    memcpy(out->reads_tx + q, db->class_tx + db->class_off[c], len * 4);
    q += len; out->reads_cnt[k] = reads_hist[c]; out->reads_off[++k] = q;
  }
  // counts_umi: key = sort + dedup of the UMI's class (mapping.rs:205-207); emitted in lexicographic key order
  std::map<std::vector<uint32_t>, uint64_t> by_key;
  for (uint32_t c = 0; c < nc; c++) if (umi_hist[c]) {
    std::vector<uint32_t> v(db->class_tx + db->class_off[c], db->class_tx + db->class_off[c + 1]);
    std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
    by_key[v] += umi_hist[c];
  }
  uint64_t utx = 0; for (auto& kv : by_key) utx += kv.first.size();
  out->n_umi_classes = by_key.size();
  out->umi_off = (uint64_t*)calloc(by_key.size() + 1, 8); out->umi_tx = (uint32_t*)malloc(std::max<uint64_t>(utx, 1) * 4); out->umi_cnt = (uint32_t*)calloc(std::max<size_t>(by_key.size(), 1), 4);
  k = 0; q = 0;
  for (auto& kv : by_key) {
    memcpy(out->umi_tx + q, kv.first.data(), kv.first.size() * 4);
    q += kv.first.size(); out->umi_cnt[k] = (uint32_t)kv.second; out->umi_off[++k] = q;
  }
  HLAC_API_CATCH
}

int hlac_geno_eqclassdb(hlac_geno* s, const hlac_class_tables* t, const uint32_t* read_bc, const uint32_t* read_umi, uint64_t n_reads,
                        const uint64_t* bc_off, const uint8_t* bc_bytes, uint32_t n_bc, const uint64_t* umi_off, const uint8_t* umi_bytes,
                        uint32_t n_umi, hlac_eqclassdb* out) {
  if (!s) { set_last_error("null argument"); return HLAC_ERR_ARG; }
  std::vector<uint32_t> ids(std::max<uint64_t>(n_reads, 1));
  { const int rc = hlac_geno_read_class_ids(s, ids.data(), n_reads); if (rc != HLAC_OK) return rc; }
  return hlac_eqclassdb_from_ids(t, ids.data(), read_bc, read_umi, n_reads, bc_off, bc_bytes, n_bc, umi_off, umi_bytes, n_umi, out);
}

// the same from explicit per-read class ids (first-seen eq_class ids, 0xFFFFFFFF = read not counted): what a driver that
// spread the reads over several sessions assembles from their hlac_geno_read_class_ids
int hlac_eqclassdb_from_ids(const hlac_class_tables* t, const uint32_t* class_ids, const uint32_t* read_bc, const uint32_t* read_umi, uint64_t n_reads,
                            const uint64_t* bc_off, const uint8_t* bc_bytes, uint32_t n_bc, const uint64_t* umi_off, const uint8_t* umi_bytes,
                            uint32_t n_umi, hlac_eqclassdb* out) {
  HLAC_API_TRY
  if (!t || !out || (n_reads && (!read_bc || !read_umi || !class_ids)) || !bc_off || !umi_off) throw Error(HLAC_ERR_ARG, "null argument");
  if (!t->reads_off || !t->reads_tx) throw Error(HLAC_ERR_STATE, "the tables are lean: counts_reads (the class vectors) is not on the host");
  if (t->n_read_classes >= NONE32) throw Error(HLAC_ERR_LIMIT, "too many classes");
  const uint32_t* ids = class_ids;
  memset(out, 0, sizeof *out);
  struct Guard { hlac_eqclassdb* d; bool ok = false; ~Guard() { if (!ok) hlac_eqclassdb_free(d); } } guard{out};
  // ids in first-counted order, like EqClassDb::count (mapping.rs:128-163)
  std::vector<uint32_t> bc_new(n_bc, NONE32), umi_new(n_umi, NONE32), bc_old, umi_old;
  uint64_t kept = 0;
  for (uint64_t i = 0; i < n_reads; i++) if (ids[i] != NONE32) kept++;
  out->bc = (uint32_t*)malloc(std::max<uint64_t>(kept, 1) * 4); out->umi = (uint32_t*)malloc(std::max<uint64_t>(kept, 1) * 4);
  out->cls = (uint32_t*)malloc(std::max<uint64_t>(kept, 1) * 4);
  uint64_t k = 0;
  for (uint64_t i = 0; i < n_reads; i++) {
    if (ids[i] == NONE32) continue;
    if (ids[i] >= t->n_read_classes || read_bc[i] >= n_bc || read_umi[i] >= n_umi) throw Error(HLAC_ERR_ARG, "read table index out of range");
    uint32_t& b = bc_new[read_bc[i]]; if (b == NONE32) { b = (uint32_t)bc_old.size(); bc_old.push_back(read_bc[i]); }
    uint32_t& u = umi_new[read_umi[i]]; if (u == NONE32) { u = (uint32_t)umi_old.size(); umi_old.push_back(read_umi[i]); }
    out->bc[k] = b; out->umi[k] = u; out->cls[k] = ids[i]; k++;
  }
  out->n_counts = kept;
  auto strings = [](const std::vector<uint32_t>& old, const uint64_t* off, const uint8_t* bytes, uint32_t& n, uint64_t*& ooff, uint8_t*& obytes) {
    uint64_t total = 0; for (uint32_t o : old) total += off[o + 1] - off[o];
    n = (uint32_t)old.size(); ooff = (uint64_t*)calloc(old.size() + 1, 8); obytes = (uint8_t*)malloc(std::max<uint64_t>(total, 1));
    uint64_t q = 0;
    for (size_t i = 0; i < old.size(); i++) { const uint64_t len = off[old[i] + 1] - off[old[i]]; if (len) memcpy(obytes + q, bytes + off[old[i]], len); q += len; ooff[i + 1] = q; }
  };
  strings(bc_old, bc_off, bc_bytes, out->n_barcodes, out->barcode_off, out->barcode_bytes);
  strings(umi_old, umi_off, umi_bytes, out->n_umis, out->umi_off, out->umi_bytes);
  const uint64_t nc = t->n_read_classes, ntx = t->reads_off[nc];
  out->n_classes = (uint32_t)nc;
  out->class_off = (uint64_t*)malloc((nc + 1) * 8); memcpy(out->class_off, t->reads_off, (nc + 1) * 8);
  out->class_tx = (uint32_t*)malloc(std::max<uint64_t>(ntx, 1) * 4); if (ntx) memcpy(out->class_tx, t->reads_tx, ntx * 4);
  guard.ok = true;
  HLAC_API_CATCH
}

}  // extern "C"
