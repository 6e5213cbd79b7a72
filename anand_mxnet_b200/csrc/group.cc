// group.cc -- see group.h.
#include "group.h"

#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "kernels.h"

namespace b200kv {

static PeerGroup* g_group = nullptr;

PeerGroup* PeerGroup::Get() { return g_group; }

std::vector<int64_t> GatherI64(B200KVAllGatherFnC fn, void* ctx, int world,
                               const std::vector<int64_t>& mine) {
  KV_CHECK(fn != nullptr) << "peer group has no all-gather callback";
  // lengths must agree: gather them first so a mismatch is an error, not a memory fault
  int64_t n = static_cast<int64_t>(mine.size());
  std::vector<int64_t> lens(world, 0);
  KV_CHECK_EQ(fn(&n, lens.data(), sizeof(int64_t), ctx), 0) << "all-gather callback failed";
  for (int r = 0; r < world; ++r) {
    KV_CHECK_EQ(lens[r], n) << "ranks issued different KVStore calls (rank " << r << " exchanges "
                            << lens[r] << " values, this rank " << n << ")";
  }
  std::vector<int64_t> all(static_cast<size_t>(world) * mine.size());
  if (n > 0) {
    KV_CHECK_EQ(fn(mine.data(), all.data(), mine.size() * sizeof(int64_t), ctx), 0)
        << "all-gather callback failed";
  }
  return all;
}

void PeerGroup::AllGather(const void* send, void* recv, size_t nbytes) {
  KV_CHECK_EQ(fn_(send, recv, nbytes, ctx_), 0) << "all-gather callback failed";
}

std::vector<int64_t> PeerGroup::AllGatherI64(const std::vector<int64_t>& mine) {
  return GatherI64(fn_, ctx_, world_, mine);
}

void PeerGroup::Init(int rank, int world, int dev, B200KVAllGatherFnC fn, void* ctx, void* ext_arena,
                     size_t ext_bytes, void* const* ext_peers, void* multicast_base) {
  KV_CHECK(g_group == nullptr) << "peer group already initialised";
  KV_CHECK(world >= 1 && world <= kMaxDevices) << "world size must be 1.." << kMaxDevices;
  KV_CHECK(rank >= 0 && rank < world);
  KV_CHECK(fn != nullptr);
  Engine* eng = Engine::Get();
  KV_CHECK(dev >= 0 && dev < eng->NumDevices()) << "invalid gpu id " << dev;
  PeerGroup* g = new PeerGroup();
  g->rank_ = rank;
  g->world_ = world;
  g->dev_ = dev;
  g->fn_ = fn;
  g->ctx_ = ctx;
  DeviceGuard guard(dev);
  const bool external = ext_arena != nullptr;
  if (external) {
    // arena allocated and peer-mapped by the launcher (torch symmetric memory: same mechanism,
    // plus an NVSwitch multicast mapping of all ranks' arenas when the fabric supports it)
    g->arena_ = ext_arena;
    g->arena_bytes_ = ext_bytes;
    g->mc_base_ = multicast_base;
    g->external_arena_ = true;
  } else {
    const char* mb = std::getenv("B200KV_IPC_ARENA_MB");
    g->arena_bytes_ = static_cast<size_t>(mb ? std::max(64, std::atoi(mb)) : 6144) << 20;
    KV_CUDA(cudaMalloc(&g->arena_, g->arena_bytes_));
  }
  uint32_t* pad = nullptr;
  KV_CUDA(cudaMalloc(reinterpret_cast<void**>(&pad), kPadWords * sizeof(uint32_t)));
  KV_CUDA(cudaMemset(pad, 0, kPadWords * sizeof(uint32_t)));
  KV_CUDA(cudaMalloc(reinterpret_cast<void**>(&g->d_counter_), 256));
  KV_CUDA(cudaMemset(g->d_counter_, 0, 256));
  KV_CUDA(cudaDeviceSynchronize());
  // exchange (arena, pad) IPC handles
  struct Handles {
    cudaIpcMemHandle_t arena, pad;
    int dev;
    int pad_[3];
  };
  Handles mine;
  std::memset(&mine, 0, sizeof(mine));
  if (!external) KV_CUDA(cudaIpcGetMemHandle(&mine.arena, g->arena_));
  KV_CUDA(cudaIpcGetMemHandle(&mine.pad, pad));
  mine.dev = dev;
  std::vector<Handles> all(world);
  g->AllGather(&mine, all.data(), sizeof(Handles));
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      g->peer_base_[r] = g->arena_;
      g->pads_[r] = pad;
      continue;
    }
    KV_CHECK(all[r].dev != dev) << "ranks " << rank << " and " << r << " share gpu " << dev;
    void* p = nullptr;
    if (external) {
      g->peer_base_[r] = ext_peers[r];
    } else {
      KV_CUDA(cudaIpcOpenMemHandle(&p, all[r].arena, cudaIpcMemLazyEnablePeerAccess));
      g->peer_base_[r] = p;
    }
    KV_CUDA(cudaIpcOpenMemHandle(&p, all[r].pad, cudaIpcMemLazyEnablePeerAccess));
    g->pads_[r] = static_cast<uint32_t*>(p);
  }
  KV_CUDA(cudaMalloc(reinterpret_cast<void**>(&g->d_pads_), kMaxDevices * sizeof(uint32_t*)));
  KV_CUDA(cudaMemcpy(g->d_pads_, g->pads_, kMaxDevices * sizeof(uint32_t*), cudaMemcpyHostToDevice));
  g_group = g;
  // everyone has mapped everyone before any kernel may touch a peer
  int64_t one = 1;
  std::vector<int64_t> sync(world);
  g->AllGather(&one, sync.data(), sizeof(int64_t));
}

void PeerGroup::Destroy() {
  if (g_group == nullptr) return;
  PeerGroup* g = g_group;
  try {
    Engine::Get()->WaitAll();
    int64_t one = 1;
    std::vector<int64_t> sync(g->world_);
    g->AllGather(&one, sync.data(), sizeof(int64_t));  // nobody unmaps while a peer still runs
  } catch (...) {
  }
  DeviceGuard guard(g->dev_);
  for (int r = 0; r < g->world_; ++r) {
    if (r == g->rank_) continue;
    if (!g->external_arena_ && g->peer_base_[r]) cudaIpcCloseMemHandle(g->peer_base_[r]);
    if (g->pads_[r]) cudaIpcCloseMemHandle(g->pads_[r]);
  }
  // the arena itself is left to process teardown: pooled blocks carved from it may still be cached
  g_group = nullptr;
  delete g;
}

void PeerGroup::FillLaunch(DenseLaunch* L) {
  // B200KV_PEER_TIMEOUT_S: how long a rank waits inside a kernel for a peer that has not launched
  // the matching call before it reports an error (0 = wait for ever, like NCCL). Default 10 min.
  static const unsigned long long timeout_ns = []() {
    const char* z = std::getenv("B200KV_PEER_TIMEOUT_S");
    const double s = z ? std::atof(z) : 600.0;
    return s <= 0 ? 0ULL : static_cast<unsigned long long>(s * 1e9);
  }();
  L->signal_pads = d_pads_;
  L->counter = d_counter_;
  L->rank = rank_;
  L->world = world_;
  L->epoch = NextEpoch();
  L->timeout_ns = timeout_ns;
  DeviceGuard guard(dev_);
  L->err_word = Engine::Get()->DeviceErrorWord();
}

void* PeerGroup::ArenaAlloc(size_t bytes) {
  const size_t r = (bytes + 511) & ~static_cast<size_t>(511);
  if (arena_used_ + r > arena_bytes_) return nullptr;
  void* p = static_cast<char*>(arena_) + arena_used_;
  arena_used_ += r;
  return p;
}

}  // namespace b200kv
