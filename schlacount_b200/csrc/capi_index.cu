// C ABI: error state, index construction, host helpers (include/hlacount.h). Product code.
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>

#include "device_index.hpp"
#include "hla_fasta.hpp"
#include "radix_sort.cuh"

using namespace hlac;

namespace hlac {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
static int g_seed_stride = 0;   // 0 = not set yet: HLAC_SEED_STRIDE from the environment, else the default
int seed_stride() {
  if (g_seed_stride == 0) {
    const char* e = getenv("HLAC_SEED_STRIDE");
    const int v = e ? atoi(e) : DEFAULT_SEED_STRIDE;
    g_seed_stride = (v == 1 || v == 3) ? v : DEFAULT_SEED_STRIDE;
  }
  return g_seed_stride;
}
}  // namespace hlac

#define HLAC_API_BEGIN try {
#define HLAC_API_END                                                                         \
  }                                                                                          \
  catch (const hlac::Error& e) { hlac::set_last_error(e.what()); return e.code; }           \
  catch (const std::exception& e) { hlac::set_last_error(e.what()); return HLAC_ERR_STATE; } \
  return HLAC_OK;

namespace {
// build_index's dominant step on the device: stable radix sort of the (k-mer, payload) occurrence pairs
void device_sort_pairs(uint64_t* keys, uint64_t* vals, size_t n) {
  DevBuf<uint64_t> k0, k1, v0, v1; RadixSortScratch rs;
  k0.upload(keys, n); v0.upload(vals, n); k1.reserve(n); v1.reserve(n);
  const bool second = radix_sort_pairs(k0.p, k1.p, v0.p, v1.p, (uint64_t)n, 0, 2 * K, rs, (cudaStream_t)0);
  HLAC_CUDA(cudaMemcpy(keys, second ? k1.p : k0.p, n * 8, cudaMemcpyDeviceToHost));
  HLAC_CUDA(cudaMemcpy(vals, second ? v1.p : v0.p, n * 8, cudaMemcpyDeviceToHost));
}
HostIndex::PairSorter sorter_for(int device) {
  int ndev = 0;
  if (device < 0 || cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return nullptr;
  return device_sort_pairs;
}
// build_index: on the device when there is one (HLAC_INDEX_BUILD=host forces the host threads + device sort of round 1),
// on the host for device-less handles, tiny inputs and the rare graphs the device build declines
void build_index_for(const std::vector<Bases>& seqs, const std::vector<std::string>& names, HostIndex& host, int device, DeviceIndexParts* keep) {
  const char* how = getenv("HLAC_INDEX_BUILD");
  const bool want_host = how && std::string(how) == "host";
  if (sorter_for(device) && !want_host && device_build_index(seqs, names, host, device, keep)) return;
  std::unique_ptr<DeviceGuard> g; if (sorter_for(device)) g = std::make_unique<DeviceGuard>(device);
  HostIndex::build(seqs, names, host, sorter_for(device));
}
}  // namespace

extern "C" {

const char* hlac_last_error(void) { return g_last_error.c_str(); }

// sizeof of every struct that crosses the ABI, in header order (bindings check their own layouts against this)
int hlac_abi_sizes(uint32_t* out, uint32_t cap) {
  const uint32_t v[] = {(uint32_t)sizeof(hlac_batch), (uint32_t)sizeof(hlac_metrics), (uint32_t)sizeof(hlac_coo), (uint32_t)sizeof(hlac_index_info),
                        (uint32_t)sizeof(hlac_class_tables), (uint32_t)sizeof(hlac_paths), (uint32_t)sizeof(hlac_calls), (uint32_t)sizeof(hlac_eqclassdb), (uint32_t)sizeof(hlac_packed_batch), (uint32_t)sizeof(hlac_wire_batch)};
  if (!out || cap < 7) { set_last_error("hlac_abi_sizes needs room for 7 values"); return HLAC_ERR_ARG; }
  for (uint32_t i = 0; i < 10 && i < cap; i++) out[i] = v[i];   // the 8th (hlac_eqclassdb) and 9th (hlac_packed_batch) only if the caller made room
  return HLAC_OK;
}

int hlac_device_cache_trim(uint64_t* bytes_released, uint64_t* bytes_live) try {
  hlac::AllocCache& c = hlac::alloc_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  const uint64_t before = c.cached_bytes;
  hlac::alloc_cache_trim_locked(c, 0);
  if (bytes_released) *bytes_released = before - c.cached_bytes;
  if (bytes_live) { uint64_t live = 0; for (const auto& kv : c.live) live += kv.second.second; *bytes_live = live; }
  return HLAC_OK;
} catch (const std::exception& e) { set_last_error(e.what()); return HLAC_ERR_STATE; }

int hlac_set_seed_stride(int stride) {
  if (stride != 1 && stride != 3) { set_last_error("seed stride must be 1 or 3 (the kernels are instantiated for these two)"); return HLAC_ERR_ARG; }
  hlac::g_seed_stride = stride;
  return HLAC_OK;
}
int hlac_get_seed_stride(void) { return hlac::seed_stride(); }

int hlac_device_count(int* n) {
  HLAC_API_BEGIN
  if (!n) throw Error(HLAC_ERR_ARG, "null argument");
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) c = 0;
  *n = c;
  HLAC_API_END
}

int hlac_index_from_idx_file(const char* path, int device, hlac_index** out) {
  HLAC_API_BEGIN
  if (!path || !out) throw Error(HLAC_ERR_ARG, "null argument");
  auto ix = std::make_unique<hlac_index>();
  HostIndex::load_idx(path, ix->host);
  ix->upload(device);
  *out = ix.release();
  HLAC_API_END
}

int hlac_index_from_fasta(const char* fasta_path, const char* allele_status, int device, hlac_index** out) {
  HLAC_API_BEGIN
  if (!fasta_path || !out) throw Error(HLAC_ERR_ARG, "null argument");
  std::set<std::string> keep;
  if (allele_status) keep = load_allele_status(allele_status);
  std::vector<Bases> seqs; std::vector<std::string> names;
  const bool dbg = getenv("HLAC_DEBUG_TIMING") != nullptr;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
  double t0 = now();
  read_hla_cds(fasta_path, keep, allele_status != nullptr, seqs, names);
  if (dbg) { fprintf(stderr, "[hlac timing] index: read_hla_cds                 %.3f s\n", now() - t0); t0 = now(); }
  auto ix = std::make_unique<hlac_index>();
  DeviceIndexParts parts;
  build_index_for(seqs, names, ix->host, device, &parts);
  if (dbg) { fprintf(stderr, "[hlac timing] index: build                        %.3f s\n", now() - t0); t0 = now(); }
  ix->upload(device, &parts);
  if (dbg) { fprintf(stderr, "[hlac timing] index: make resident                %.3f s\n", now() - t0); t0 = now(); }
  *out = ix.release();
  HLAC_API_END
}

int hlac_index_from_sequences(const uint64_t* packed, const uint64_t* seq_word_start, const uint32_t* seq_len,
                              const char* const* names, uint32_t n_seq, int device, hlac_index** out) {
  HLAC_API_BEGIN
  if (!packed || !seq_word_start || !seq_len || !names || !out) throw Error(HLAC_ERR_ARG, "null argument");
  std::vector<Bases> seqs(n_seq); std::vector<std::string> nm(n_seq);
  for (uint32_t i = 0; i < n_seq; i++) {
    seqs[i].resize(seq_len[i]);
    const uint64_t* w = packed + seq_word_start[i];
    for (uint32_t j = 0; j < seq_len[i]; j++) seqs[i][j] = (w[j >> 5] >> (62 - 2 * (j & 31))) & 3;
    nm[i] = names[i];
  }
  auto ix = std::make_unique<hlac_index>();
  DeviceIndexParts parts;
  build_index_for(seqs, nm, ix->host, device, &parts);
  ix->upload(device, &parts);
  *out = ix.release();
  HLAC_API_END
}

// a second resident copy of an index on another device (multi-GPU runs replicate the index): the host
// structures are copied, nothing is rebuilt
int hlac_index_clone(const hlac_index* src, int device, hlac_index** out) {
  HLAC_API_BEGIN
  if (!src || !out) throw Error(HLAC_ERR_ARG, "null argument");
  auto ix = std::make_unique<hlac_index>();
  ix->host = src->host;
  ix->upload(device);
  if (src->dict_device_only && device >= 0) {   // the dictionary exists on the source device only
    DeviceGuard g(device);
    ix->dict.reserve(src->dict.cap);
    HLAC_CUDA(cudaMemcpyPeer(ix->dict.p, device, src->dict.p, src->device, ((size_t)16 << src->host.dict_bits)));
    ix->view.dict = reinterpret_cast<const uint4*>(ix->dict.p);
    ix->dict_device_only = true;
  }
  *out = ix.release();
  HLAC_API_END
}

int hlac_index_device(const hlac_index* ix) { return ix ? ix->device : -1; }

int hlac_index_free(hlac_index* ix) {
  if (ix) { int prev = 0; cudaGetDevice(&prev); if (ix->device >= 0) cudaSetDevice(ix->device); delete ix; cudaSetDevice(prev); }   // the caller's current device is left as it was
  return HLAC_OK;
}

int hlac_index_get_info(const hlac_index* ix, hlac_index_info* o) {
  HLAC_API_BEGIN
  if (!ix || !o) throw Error(HLAC_ERR_ARG, "null argument");
  o->n_nodes = ix->host.n_nodes(); o->n_tx = (uint32_t)ix->host.tx_names.size(); o->n_colours = ix->host.n_colours();
  o->lmer = ix->host.lmer; o->n_bases = ix->host.n_bases; o->n_kmers = ix->host.n_kmers;
  o->dict_slots = 1ull << ix->host.dict_bits; o->bitmap_bytes = ix->host.bitmap.size() * 4;
  HLAC_API_END
}

int hlac_index_write_idx(const hlac_index* ix, const char* path) {
  HLAC_API_BEGIN
  if (!ix || !path) throw Error(HLAC_ERR_ARG, "null argument");
  ix->host.write_idx(path);
  HLAC_API_END
}

const char* hlac_index_tx_name(const hlac_index* ix, uint32_t i) {
  return (ix && i < ix->host.tx_names.size()) ? ix->host.tx_names[i].c_str() : nullptr;
}

int hlac_index_node(const hlac_index* ix, uint32_t node, char* seq, uint32_t seq_cap, uint32_t* len, uint32_t* exts,
                    uint32_t* colours, uint32_t col_cap, uint32_t* n_col) {
  HLAC_API_BEGIN
  if (!ix || node >= ix->host.n_nodes()) throw Error(HLAC_ERR_ARG, "bad node");
  const HostIndex& h = ix->host;
  if (len) *len = h.len[node];
  if (exts) *exts = h.exts[node];
  if (seq) for (uint32_t i = 0; i < h.len[node] && i < seq_cap; i++) seq[i] = "ACGT"[h.base((uint64_t)h.start[node] + i)];
  const uint64_t a = h.col_off[h.colour[node]], b = h.col_off[h.colour[node] + 1];
  if (n_col) *n_col = (uint32_t)(b - a);
  if (colours) for (uint64_t i = a; i < b && i - a < col_cap; i++) colours[i - a] = h.col_tx[i];
  HLAC_API_END
}

int hlac_host_alloc(size_t bytes, void** out) {
  HLAC_API_BEGIN
  if (!out) throw Error(HLAC_ERR_ARG, "null argument");
  HLAC_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  HLAC_API_END
}
int hlac_host_free(void* p) {
  HLAC_API_BEGIN
  if (p) HLAC_CUDA(cudaFreeHost(p));
  HLAC_API_END
}

int hlac_pack_read(const char* ascii, uint32_t len, uint64_t* words, uint32_t nwords) {
  HLAC_API_BEGIN
  if (!ascii || !words) throw Error(HLAC_ERR_ARG, "null argument");
  if ((uint64_t)nwords * 32 < len) throw Error(HLAC_ERR_ARG, "read longer than 32*nwords");
  memset(words, 0, (size_t)nwords * 8);
  for (uint32_t i = 0; i < len; i++) words[i >> 5] |= (uint64_t)base_code(ascii[i]) << (62 - 2 * (i & 31));
  HLAC_API_END
}

int hlac_umi_key(const char* umi, uint32_t len, uint32_t* key) {
  HLAC_API_BEGIN
  if (!umi || !key) throw Error(HLAC_ERR_ARG, "null argument");
  if (len > 15) throw Error(HLAC_ERR_LIMIT, "UMI longer than 15 bases needs interning");
  uint32_t v = 1;
  for (uint32_t i = 0; i < len; i++) {
    uint32_t c;
    switch (umi[i]) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break;
      default: throw Error(HLAC_ERR_LIMIT, "non-ACGT UMI needs interning"); }
    v = (v << 2) | c;
  }
  *key = v;
  HLAC_API_END
}

}  // extern "C"
