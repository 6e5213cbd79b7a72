// group.h -- one-process-per-GPU peer group: NVLink peer memory across processes through CUDA IPC.
//
// The reference is single-process (one process drives all GPUs, peer access by
// cudaDeviceEnablePeerAccess, src/kvstore/comm.h:715-757). The B200 deployment model is one rank
// per GPU (torchrun); the same fused kernel then reads the peers' gradients and writes the peers'
// weights through IPC-mapped pointers:
//   * every rank owns one IPC ARENA (a single cudaMalloc, exported once with cudaIpcGetMemHandle and
//     mapped by every peer at group creation); the library's GPU allocations of that rank come out
//     of the arena, so any NDArray created through the C ABI is addressable by every peer as
//     peer_base[rank] + offset;
//   * plan building is collective: ranks all-gather the arena offsets of their operands once per
//     distinct call signature (host callback supplied by the launcher -- torch.distributed -- the
//     library itself never opens a socket);
//   * steady state has NO host-side communication: cross-rank ordering is an in-kernel barrier on
//     IPC-mapped signal pads (st.release.sys / ld.acquire.sys), one at kernel start ("my gradients
//     are ready") and one at the end ("I have finished reading yours and writing your weights").
#pragma once
#include <cstdint>
#include <vector>

#include "common.h"

extern "C" typedef int (*B200KVAllGatherFnC)(const void* send, void* recv, size_t nbytes, void* ctx);

namespace b200kv {

constexpr int kPadStride = 32;                         // uint32 slots per flag: one 128-byte line
constexpr int kPadWords = 2 * kMaxDevices * kPadStride;  // start flags, then end flags

class PeerGroup {
 public:
  static PeerGroup* Get();  // nullptr until Init()
  // ext_arena == nullptr: the library allocates and IPC-exports its own arena. Otherwise the
  // launcher supplies an arena it has already peer-mapped (ext_peers[r] = rank r's arena as seen
  // from this rank) and, optionally, an NVSwitch MULTICAST mapping of all ranks' arenas
  // (multicast_base: a store to it lands in every rank's arena at the same offset, a
  // multimem.ld_reduce from it returns the sum over all ranks -- NVLS).
  static void Init(int rank, int world, int dev, B200KVAllGatherFnC fn, void* ctx,
                   void* ext_arena = nullptr, size_t ext_bytes = 0, void* const* ext_peers = nullptr,
                   void* multicast_base = nullptr);
  static void Destroy();

  int rank() const { return rank_; }
  int world() const { return world_; }
  int dev() const { return dev_; }

  // gathers nbytes from every rank into recv[rank*nbytes ...]
  void AllGather(const void* send, void* recv, size_t nbytes);
  // gathers one int64 vector per rank (all the same length) -> [world][n] row-major
  std::vector<int64_t> AllGatherI64(const std::vector<int64_t>& mine);

  // ---- arena
  void* ArenaAlloc(size_t bytes);  // bump allocation, 512-byte aligned; nullptr when exhausted
  bool InArena(const void* p) const {
    return p >= arena_ && p < static_cast<const char*>(arena_) + arena_bytes_;
  }
  int64_t OffsetOf(const void* p) const {
    return static_cast<const char*>(p) - static_cast<const char*>(arena_);
  }
  void* PeerPtr(int peer, int64_t offset) const { return static_cast<char*>(peer_base_[peer]) + offset; }
  bool has_multicast() const { return mc_base_ != nullptr; }
  void* McPtr(int64_t offset) const { return static_cast<char*>(mc_base_) + offset; }

  // ---- in-kernel barrier state
  uint32_t* const* d_pads() const { return d_pads_; }  // device array [world] of pad pointers
  uint32_t* d_counter() const { return d_counter_; }   // "CTAs finished" counter of this rank
  uint32_t NextEpoch() { return ++epoch_; }
  // barrier part of a fused launch: pads, counter, rank / world, a fresh epoch, the time-out policy
  void FillLaunch(struct DenseLaunch* L);
  // false when the ranks' arenas could not be IPC-mapped (no GPU peer access between them, or
  // B200KV_GROUP_NO_IPC=1): stores then use the NCCL fallback collective instead of peer memory
  bool ipc_ok() const { return ipc_ok_; }
  // NCCL communicator over the group's ranks, created on first use (unique id from rank 0 through
  // the launcher's all-gather callback). Collective: every rank must call it at the same point.
  void* NcclCommunicator();
  // host-only test hook (no GPU): builds the shared-memory mailbox of a `world`-rank job through
  // the launcher's callback and runs `rounds` all-gathers of `n` int64 values over it; out receives
  // the last round's result, *used_mailbox whether shared memory (not the callback) carried them
  static void TestMailbox(int rank, int world, B200KVAllGatherFnC fn, void* ctx, const int64_t* mine,
                          int n, int rounds, int64_t* out, int* used_mailbox);

 private:
  PeerGroup() {}
  int rank_ = 0, world_ = 1, dev_ = 0;
  B200KVAllGatherFnC fn_ = nullptr;
  void* ctx_ = nullptr;
  void* arena_ = nullptr;
  size_t arena_bytes_ = 0, arena_used_ = 0;
  void* peer_base_[kMaxDevices] = {nullptr};
  void* mc_base_ = nullptr;
  bool external_arena_ = false;
  uint32_t* pads_[kMaxDevices] = {nullptr};
  uint32_t** d_pads_ = nullptr;
  uint32_t* d_counter_ = nullptr;
  uint32_t epoch_ = 0;
  bool ipc_ok_ = true;
  void* nccl_comm_ = nullptr;
  // host-side mailbox between the ranks of the node (POSIX shared memory, created at Init through
  // ONE exchange over the launcher's callback): later all-gathers of a few KB -- operand offsets
  // when a launch is planned, the (ids, rows, count) descriptor of every row_sparse push -- take a
  // microsecond of spinning on shared memory instead of a trip through the launcher's transport
  void InitMailbox();
  bool MailboxAllGather(const void* send, void* recv, size_t nbytes);
  char* mbox_ = nullptr;
  size_t mbox_bytes_ = 0;
  uint64_t mbox_seq_ = 0;
};

// Pure host logic, testable without a GPU: rank-major gather of equal-length int64 vectors through
// the launcher's all-gather callback.
std::vector<int64_t> GatherI64(B200KVAllGatherFnC fn, void* ctx, int world,
                               const std::vector<int64_t>& mine);

}  // namespace b200kv
