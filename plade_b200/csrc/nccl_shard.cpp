// see nccl_shard.h
#include "nccl_shard.h"
#include <nccl.h>          // types and prototypes only; the symbols are resolved with dlsym
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace plade {

namespace {
struct Api {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
};
Api &api() {
  static Api a;
  static std::once_flag once;
  static std::string why;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) { why = std::string("cannot open libnccl.so.2: ") + dlerror(); return; }
#define PLADE_NCCL_SYM(f) a.f = reinterpret_cast<decltype(a.f)>(dlsym(a.handle, "nccl" #f)); if (!a.f) { why = "libnccl lacks nccl" #f; a.handle = nullptr; return; }
    PLADE_NCCL_SYM(GetUniqueId) PLADE_NCCL_SYM(CommInitRank) PLADE_NCCL_SYM(CommInitAll) PLADE_NCCL_SYM(CommDestroy)
    PLADE_NCCL_SYM(AllReduce) PLADE_NCCL_SYM(Broadcast) PLADE_NCCL_SYM(GetErrorString) PLADE_NCCL_SYM(GroupStart) PLADE_NCCL_SYM(GroupEnd)
#undef PLADE_NCCL_SYM
  });
  if (!a.handle) throw std::runtime_error("NCCL is not available: " + why);
  return a;
}
void check(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + api().GetErrorString(r));
}
}  // namespace

struct ShardComm { ncclComm_t comm = nullptr; int rank = 0, world = 1; };

void nccl_unique_id(char out128[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  check(api().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, &id, 128);
}

ShardComm *nccl_comm_init_rank(const char id128[128], int rank, int world) {
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ShardComm *c = new ShardComm;
  c->rank = rank; c->world = world;
  ncclResult_t r = api().CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) { delete c; check(r, "ncclCommInitRank"); }
  return c;
}

void nccl_comm_init_all(ShardComm **out, const int *devices, int n) {
  std::vector<ncclComm_t> comms(n);
  check(api().CommInitAll(comms.data(), n, devices), "ncclCommInitAll");
  for (int i = 0; i < n; ++i) { out[i] = new ShardComm; out[i]->comm = comms[i]; out[i]->rank = i; out[i]->world = n; }
}

void nccl_comm_destroy(ShardComm *c) {
  if (!c) return;
  if (c->comm) api().CommDestroy(c->comm);
  delete c;
}
int nccl_rank(const ShardComm *c) { return c->rank; }
int nccl_world(const ShardComm *c) { return c->world; }

void nccl_allreduce_max_u64(ShardComm *c, unsigned long long *d_values, int n, cudaStream_t stream) {
  check(api().AllReduce(d_values, d_values, (size_t) n, ncclUint64, ncclMax, c->comm, stream), "ncclAllReduce");
}

void nccl_broadcast_bytes(ShardComm *c, void *d_buf, size_t nbytes, int root, cudaStream_t stream) {
  check(api().Broadcast(d_buf, d_buf, nbytes, ncclChar, root, c->comm, stream), "ncclBroadcast");
}

}  // namespace plade
