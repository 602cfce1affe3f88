// hlac_comm: the communicator a multi-GPU run attaches to its sessions (include/hlacount.h). One rank = one device; the
// collectives are NCCL's, loaded at run time (dlopen of libnccl.so.2: inside a process that already carries an NCCL —
// torch's bundled one — that copy is reused, a stand-alone sc_hla_count picks up the system library), so the library has
// no link-time dependency on NCCL and single-GPU users never load it. Product code.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "../../include/hlacount.h"

enum { HLAC_COMM_U32 = 0, HLAC_COMM_U64 = 1, HLAC_COMM_U8 = 2 };

namespace hlac {
int comm_rank(const hlac_comm*);
int comm_nranks(const hlac_comm*);
int comm_device(const hlac_comm*);
void comm_group_begin(hlac_comm*);
void comm_group_end(hlac_comm*);
void comm_allreduce_sum(hlac_comm*, const void* send, void* recv, size_t count, int dtype, cudaStream_t);
void comm_allreduce_min(hlac_comm*, const void* send, void* recv, size_t count, int dtype, cudaStream_t);
void comm_reduce_sum(hlac_comm*, const void* send, void* recv, size_t count, int dtype, int root, cudaStream_t);
void comm_allgather(hlac_comm*, const void* send, void* recv, size_t count_per_rank, int dtype, cudaStream_t);
void comm_broadcast(hlac_comm*, void* buf, size_t count, int dtype, int root, cudaStream_t);
// variable-size all-gather: rank r contributes counts[r] elements, which land at recv + displs[r] elements on every rank
// (one grouped launch of per-rank broadcasts). `send` is this rank's contribution (may alias its slot in recv).
void comm_allgatherv(hlac_comm*, const void* send, void* recv, const size_t* counts, const size_t* displs, int dtype, cudaStream_t);
}  // namespace hlac
