// NCCL plumbing of the hypothesis-sharded verification: rank r verifies the hypotheses h with
// h % world == r on replicated down-sampled clouds, the winner is agreed with ONE ncclAllReduce(ncclUint64, ncclMax)
// of a packed key {score bits | ~index} on the context's stream.  libnccl.so.2 is opened at run time (dlopen), so the
// library has no link-time dependency on NCCL and single-GPU use never touches it; inside a torch process the call
// resolves to the NCCL that torch has already loaded.
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace plade {

struct ShardComm;     // opaque: ncclComm_t + rank / world

// all throw std::runtime_error with NCCL's message on failure
void nccl_unique_id(char out128[128]);
ShardComm *nccl_comm_init_rank(const char id128[128], int rank, int world);        // the current device
void nccl_comm_init_all(ShardComm **out, const int *devices, int n);              // one process, n devices
void nccl_comm_destroy(ShardComm *c);
int nccl_rank(const ShardComm *c);
int nccl_world(const ShardComm *c);
// in-place all-reduce(MAX) of n u64 values resident on the device, on `stream`
void nccl_allreduce_max_u64(ShardComm *c, unsigned long long *d_values, int n, cudaStream_t stream);

// in-place broadcast of nbytes resident on the device from rank `root`, on `stream`
void nccl_broadcast_bytes(ShardComm *c, void *d_buf, size_t nbytes, int root, cudaStream_t stream);

}  // namespace plade
