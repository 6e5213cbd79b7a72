// nccl_dyn.h -- NCCL as the FALLBACK collective (north_star: "NCCL only as the fallback collective
// where the key shards naturally"; reference: src/kvstore/kvstore_nccl.h:62-551), bound at run time.
//
// The library does not link against NCCL: the handful of entry points it needs are resolved with
// dlopen/dlsym from libnccl.so.2 (B200KV_NCCL_LIB, else the loader's search path -- the copy a
// framework such as torch has already loaded is found first). Without the library a store of type
// 'nccl' fails at creation with a clear message; the peer-memory stores never touch it.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace b200kv {

struct NcclUniqueId { char internal[128]; };   // nccl.h: NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;

class Nccl {
 public:
  static Nccl* Get();        // loads libnccl.so.2 on first use; KV_FATAL when it cannot
  static bool Available();   // true when the library can be loaded (no error raised)

  void GetUniqueId(NcclUniqueId* id);
  NcclComm CommInitRank(int nranks, const NcclUniqueId& id, int rank);
  void CommDestroy(NcclComm comm);
  // dtype: DType of common.h (kFloat32 / kFloat16 / kBfloat16 / ...); in place allowed
  void AllReduceSum(const void* send, void* recv, size_t count, int dtype, NcclComm comm, cudaStream_t s);
  void Broadcast(const void* send, void* recv, size_t count, int dtype, int root, NcclComm comm,
                 cudaStream_t s);
  const char* version() const { return version_; }

 private:
  Nccl() {}
  bool Load(bool fatal);
  void Check(int rc, const char* what);
  void* lib_ = nullptr;
  char version_[32] = "?";
  int (*get_unique_id_)(NcclUniqueId*) = nullptr;
  int (*comm_init_rank_)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*comm_destroy_)(NcclComm) = nullptr;
  int (*all_reduce_)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*broadcast_)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*get_version_)(int*) = nullptr;
  const char* (*get_error_string_)(int) = nullptr;
};

}  // namespace b200kv
