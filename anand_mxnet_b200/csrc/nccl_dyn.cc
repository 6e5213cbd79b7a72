// nccl_dyn.cc -- see nccl_dyn.h.
#include "nccl_dyn.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.h"

namespace b200kv {

Nccl* Nccl::Get() {
  static Nccl* inst = new Nccl();
  inst->Load(true);
  return inst;
}

bool Nccl::Available() {
  static Nccl* probe = new Nccl();
  return probe->Load(false);
}

bool Nccl::Load(bool fatal) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (lib_ != nullptr) return true;
  const char* names[] = {std::getenv("B200KV_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (n == nullptr || n[0] == '\0') continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h != nullptr) break;
  }
  if (h == nullptr) {
    if (fatal) {
      KV_FATAL << "kvstore 'nccl': libnccl.so.2 could not be loaded (" << dlerror()
               << "); set B200KV_NCCL_LIB to its path, or use kvstore 'device' (NVLink peer memory)";
    }
    return false;
  }
  auto sym = [&](const char* name) {
    void* p = dlsym(h, name);
    if (p == nullptr && fatal) KV_FATAL << "libnccl: symbol " << name << " is missing";
    return p;
  };
  get_unique_id_ = reinterpret_cast<decltype(get_unique_id_)>(sym("ncclGetUniqueId"));
  comm_init_rank_ = reinterpret_cast<decltype(comm_init_rank_)>(sym("ncclCommInitRank"));
  comm_destroy_ = reinterpret_cast<decltype(comm_destroy_)>(sym("ncclCommDestroy"));
  all_reduce_ = reinterpret_cast<decltype(all_reduce_)>(sym("ncclAllReduce"));
  broadcast_ = reinterpret_cast<decltype(broadcast_)>(sym("ncclBroadcast"));
  get_version_ = reinterpret_cast<decltype(get_version_)>(sym("ncclGetVersion"));
  get_error_string_ = reinterpret_cast<decltype(get_error_string_)>(sym("ncclGetErrorString"));
  if (!get_unique_id_ || !comm_init_rank_ || !comm_destroy_ || !all_reduce_ || !broadcast_) return false;
  int v = 0;
  if (get_version_ != nullptr && get_version_(&v) == 0) {
    std::snprintf(version_, sizeof(version_), "%d.%d.%d", v / 10000, (v / 100) % 100, v % 100);
  }
  lib_ = h;
  return true;
}

void Nccl::Check(int rc, const char* what) {
  if (rc == 0) return;
  KV_FATAL << "NCCL " << what << " failed: "
           << (get_error_string_ ? get_error_string_(rc) : "unknown error") << " (code " << rc << ")";
}

// nccl.h: ncclDataType_t / ncclRedOp_t values (stable across NCCL 2.x)
static int NcclType(int dtype) {
  switch (dtype) {
    case kFloat32: return 7;
    case kFloat64: return 8;
    case kFloat16: return 6;
    case kBfloat16: return 9;
    case kUint8: return 1;
    case kInt32: return 2;
    case kInt8: return 0;
    case kInt64: return 4;
  }
  KV_FATAL << "NCCL collective: unsupported dtype " << DTypeName(dtype);
  return -1;
}

void Nccl::GetUniqueId(NcclUniqueId* id) { Check(get_unique_id_(id), "ncclGetUniqueId"); }

NcclComm Nccl::CommInitRank(int nranks, const NcclUniqueId& id, int rank) {
  NcclComm c = nullptr;
  Check(comm_init_rank_(&c, nranks, id, rank), "ncclCommInitRank");
  return c;
}

void Nccl::CommDestroy(NcclComm comm) {
  if (comm != nullptr) comm_destroy_(comm);
}

void Nccl::AllReduceSum(const void* send, void* recv, size_t count, int dtype, NcclComm comm,
                        cudaStream_t s) {
  Check(all_reduce_(send, recv, count, NcclType(dtype), /*ncclSum=*/0, comm, s), "ncclAllReduce");
}

void Nccl::Broadcast(const void* send, void* recv, size_t count, int dtype, int root, NcclComm comm,
                     cudaStream_t s) {
  Check(broadcast_(send, recv, count, NcclType(dtype), root, comm, s), "ncclBroadcast");
}

}  // namespace b200kv
