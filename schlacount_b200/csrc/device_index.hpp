// An index resident on one device: HostIndex + its device buffers (the `hlac_index` handle). Product code.
#pragma once
#include "cuda_util.hpp"
#include "device_index.cuh"
#include "host_index.hpp"

namespace hlac {
struct DeviceIndexParts {   // device-resident parts a device-side index build hands over
  DevBuf<uint32_t> seq32, dict, bitmap, col_tx;
  DevBuf<uint64_t> col_off;
  int device = -1;
  bool valid = false;
};
}  // namespace hlac

struct hlac_index {
  hlac::HostIndex host;
  int device = 0;
  hlac::DevBuf<uint32_t> seq32, dict, nodes, bitmap;
  hlac::DevBuf<uint32_t> col_tx;
  hlac::DevBuf<uint64_t> col_off;
  hlac::DevBuf<uint32_t> col_bits;   // per colour: membership bitset over tx ids (built on first use by the genotyping finish)
  uint32_t col_bits_words = 0;
  hlac::DevIndexView view{};
  bool dict_device_only = false;   // the dictionary was built on the device and never downloaded (host.dict is empty)
  void upload(int dev, hlac::DeviceIndexParts* parts = nullptr);
};
