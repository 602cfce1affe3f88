#include "device_index.hpp"

#include <vector>

using namespace hlac;

void hlac_index::upload(int dev, DeviceIndexParts* parts) {
  if (dev == -1) { device = -1; return; }   // host-only handle: inspection (hlac_index_node/info) but no compute
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw Error(HLAC_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (dev < 0 || dev >= ndev) throw Error(HLAC_ERR_ARG, "bad device ordinal");
  device = dev;
  DeviceGuard g(dev);
  const uint32_t nn = host.n_nodes();
  // the kernels carry (offset in node, read position, orientation) of a seed in one word: 20 bits of offset
  for (uint32_t i = 0; i < nn; i++)
    if (host.len[i] >= (1u << 20) + K) throw Error(HLAC_ERR_LIMIT, "a graph node is longer than 2^20 bases (unbranched sequence of " + std::to_string(host.len[i]) + " bases): the seed record holds a 20-bit node offset");
  std::vector<uint32_t> rec((size_t)nn * 12);
  for (uint32_t i = 0; i < nn; i++) {
    uint32_t* r = &rec[(size_t)i * 12];
    r[0] = host.start[i]; r[1] = host.len[i]; r[2] = host.exts[i]; r[3] = host.colour[i];
    for (int b = 0; b < 4; b++) { r[4 + b] = host.radj[(size_t)i * 4 + b]; r[8 + b] = host.ladj[(size_t)i * 4 + b]; }
  }
  auto take = [](auto& dst, auto& src) { std::swap(dst.p, src.p); std::swap(dst.cap, src.cap); };
  if (parts && parts->valid && parts->device == dev) {   // built on this device: adopt the buffers instead of copying them back up
    take(seq32, parts->seq32); take(dict, parts->dict); take(bitmap, parts->bitmap); take(col_tx, parts->col_tx); take(col_off, parts->col_off);
    dict_device_only = host.dict.empty();
    parts->valid = false;
  } else {
    seq32.upload(host.seq32.data(), host.seq32.size());
    if (!host.dict.empty()) dict.upload(host.dict.data(), host.dict.size());
    bitmap.upload(host.bitmap.data(), host.bitmap.size());
    col_tx.upload(host.col_tx.data(), host.col_tx.size());
    col_off.upload(host.col_off.data(), host.col_off.size());
  }
  nodes.upload(rec.data(), rec.size());
  HLAC_CUDA(cudaDeviceSynchronize());
  view.seq32 = seq32.p;
  view.nodes = reinterpret_cast<const uint4*>(nodes.p);
  view.dict = reinterpret_cast<const uint4*>(dict.p);
  view.bitmap = bitmap.p;
  view.dict_bits = host.dict_bits;
  view.dict_mask = (1u << host.dict_bits) - 1;
  view.lmer = host.lmer;
  view.bitmap_words = (uint32_t)host.bitmap.size();
  view.n_nodes = nn;
  view.seq_words = (uint32_t)host.seq32.size();
}
