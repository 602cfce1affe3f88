// hlac_comm over NCCL (see comm.hpp). Product code.
#include "comm.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "cuda_util.hpp"

using namespace hlac;

namespace {
struct Nccl {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char* names[] = {getenv("HLAC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.so) break;
      err = dlerror();
    }
    if (!n.so) return;
#define HLAC_SYM(F) n.F = reinterpret_cast<decltype(n.F)>(dlsym(n.so, "nccl" #F)); if (!n.F) { err = "symbol nccl" #F " missing"; n.so = nullptr; return; }
    HLAC_SYM(GetUniqueId) HLAC_SYM(CommInitRank) HLAC_SYM(CommInitAll) HLAC_SYM(CommDestroy) HLAC_SYM(AllReduce) HLAC_SYM(Reduce)
    HLAC_SYM(AllGather) HLAC_SYM(Broadcast) HLAC_SYM(GroupStart) HLAC_SYM(GroupEnd) HLAC_SYM(GetErrorString) HLAC_SYM(GetVersion)
#undef HLAC_SYM
  });
  if (!n.so) throw Error(HLAC_ERR_STATE, "NCCL is not available (" + err + "): multi-GPU merges need libnccl.so.2 (HLAC_NCCL_LIB overrides the name)");
  return n;
}
void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw Error(HLAC_ERR_CUDA, std::string(what) + " failed: " + nccl().GetErrorString(r));
}
#define HLAC_NCCL(x) nccl_check((x), #x)
ncclDataType_t dt(int d) { return d == HLAC_COMM_U32 ? ncclUint32 : d == HLAC_COMM_U64 ? ncclUint64 : ncclUint8; }
}  // namespace

struct hlac_comm {
  ncclComm_t c = nullptr;
  int rank = 0, nranks = 1, device = 0;
};

namespace hlac {
int comm_rank(const hlac_comm* c) { return c->rank; }
int comm_nranks(const hlac_comm* c) { return c->nranks; }
int comm_device(const hlac_comm* c) { return c->device; }
void comm_group_begin(hlac_comm*) { HLAC_NCCL(nccl().GroupStart()); }
void comm_group_end(hlac_comm*) { HLAC_NCCL(nccl().GroupEnd()); }
void comm_allreduce_sum(hlac_comm* c, const void* send, void* recv, size_t count, int dtype, cudaStream_t st) {
  HLAC_NCCL(nccl().AllReduce(send, recv, count, dt(dtype), ncclSum, c->c, st));
}
void comm_allreduce_min(hlac_comm* c, const void* send, void* recv, size_t count, int dtype, cudaStream_t st) {
  HLAC_NCCL(nccl().AllReduce(send, recv, count, dt(dtype), ncclMin, c->c, st));
}
void comm_reduce_sum(hlac_comm* c, const void* send, void* recv, size_t count, int dtype, int root, cudaStream_t st) {
  HLAC_NCCL(nccl().Reduce(send, recv, count, dt(dtype), ncclSum, root, c->c, st));
}
void comm_allgather(hlac_comm* c, const void* send, void* recv, size_t count_per_rank, int dtype, cudaStream_t st) {
  HLAC_NCCL(nccl().AllGather(send, recv, count_per_rank, dt(dtype), c->c, st));
}
void comm_broadcast(hlac_comm* c, void* buf, size_t count, int dtype, int root, cudaStream_t st) {
  HLAC_NCCL(nccl().Broadcast(buf, buf, count, dt(dtype), root, c->c, st));
}
void comm_allgatherv(hlac_comm* c, const void* send, void* recv, const size_t* counts, const size_t* displs, int dtype, cudaStream_t st) {
  const size_t esz = dtype == HLAC_COMM_U32 ? 4 : dtype == HLAC_COMM_U64 ? 8 : 1;
  HLAC_NCCL(nccl().GroupStart());
  for (int r = 0; r < c->nranks; r++) {
    if (counts[r] == 0) continue;
    char* slot = (char*)recv + displs[r] * esz;
    HLAC_NCCL(nccl().Broadcast(r == c->rank ? send : slot, slot, counts[r], dt(dtype), r, c->c, st));
  }
  HLAC_NCCL(nccl().GroupEnd());
}
}  // namespace hlac

#define HLAC_API_TRY try {
#define HLAC_API_CATCH                                                                       \
  }                                                                                          \
  catch (const hlac::Error& e) { hlac::set_last_error(e.what()); return e.code; }           \
  catch (const std::exception& e) { hlac::set_last_error(e.what()); return HLAC_ERR_STATE; } \
  return HLAC_OK;

extern "C" {

int hlac_comm_unique_id(uint8_t id[HLAC_COMM_ID_BYTES]) {
  HLAC_API_TRY
  if (!id) throw Error(HLAC_ERR_ARG, "null argument");
  static_assert(sizeof(ncclUniqueId) == HLAC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  HLAC_NCCL(nccl().GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  HLAC_API_CATCH
}

int hlac_comm_init_rank(const uint8_t id[HLAC_COMM_ID_BYTES], int rank, int nranks, int device, hlac_comm** out) {
  HLAC_API_TRY
  if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) throw Error(HLAC_ERR_ARG, "bad argument");
  DeviceGuard g(device);
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  auto c = std::make_unique<hlac_comm>();
  c->rank = rank; c->nranks = nranks; c->device = device;
  HLAC_NCCL(nccl().CommInitRank(&c->c, nranks, u, rank));
  {   // NCCL connects its transports lazily, at the first collective (seconds): pay that here, not inside the first merge
    DevBuf<uint32_t> one; one.reserve(1);
    HLAC_CUDA(cudaMemset(one.p, 0, 4));
    HLAC_NCCL(nccl().AllReduce(one.p, one.p, 1, ncclUint32, ncclSum, c->c, (cudaStream_t)0));
    HLAC_CUDA(cudaStreamSynchronize(0));
  }
  *out = c.release();
  HLAC_API_CATCH
}

// one process driving several devices (sc_hla_count --devices): comms[i] is rank i on devices[i]
int hlac_comm_init_all(const int* devices, int n, hlac_comm** comms) {
  HLAC_API_TRY
  if (!devices || !comms || n < 1) throw Error(HLAC_ERR_ARG, "bad argument");
  std::vector<ncclComm_t> cs(n);
  HLAC_NCCL(nccl().CommInitAll(cs.data(), n, devices));
  for (int i = 0; i < n; i++) {
    auto c = new hlac_comm();
    c->c = cs[i]; c->rank = i; c->nranks = n; c->device = devices[i];
    comms[i] = c;
  }
  {   // first collective (lazy transport setup) now: all ranks from this thread, in one group
    std::vector<uint32_t*> bufs(n, nullptr);
    int prev = 0; HLAC_CUDA(cudaGetDevice(&prev));
    for (int i = 0; i < n; i++) { HLAC_CUDA(cudaSetDevice(devices[i])); HLAC_CUDA(cudaMalloc(&bufs[i], 4)); HLAC_CUDA(cudaMemset(bufs[i], 0, 4)); }
    HLAC_NCCL(nccl().GroupStart());
    for (int i = 0; i < n; i++) { HLAC_CUDA(cudaSetDevice(devices[i])); HLAC_NCCL(nccl().AllReduce(bufs[i], bufs[i], 1, ncclUint32, ncclSum, cs[i], (cudaStream_t)0)); }
    HLAC_NCCL(nccl().GroupEnd());
    for (int i = 0; i < n; i++) { HLAC_CUDA(cudaSetDevice(devices[i])); HLAC_CUDA(cudaStreamSynchronize(0)); cudaFree(bufs[i]); }
    cudaSetDevice(prev);
  }
  HLAC_API_CATCH
}

int hlac_comm_free(hlac_comm* c) {
  if (c) {
    if (c->c) { int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device); nccl().CommDestroy(c->c); cudaSetDevice(prev); }
    delete c;
  }
  return HLAC_OK;
}

int hlac_comm_info(const hlac_comm* c, int* rank, int* nranks, int* device, int* nccl_version) {
  HLAC_API_TRY
  if (!c) throw Error(HLAC_ERR_ARG, "null argument");
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  if (device) *device = c->device;
  if (nccl_version) HLAC_NCCL(nccl().GetVersion(nccl_version));
  HLAC_API_CATCH
}

}  // extern "C"
