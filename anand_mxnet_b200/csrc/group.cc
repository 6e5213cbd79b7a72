// group.cc -- see group.h.
#include "group.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "kernels.h"
#include "nccl_dyn.h"

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

// ---- shared-memory mailbox -------------------------------------------------------------------
namespace {
constexpr size_t kMboxPayload = 4096;
struct MboxSlot {                       // one per (rank, parity)
  std::atomic<uint64_t> seq;
  char pad[56];
  char payload[kMboxPayload];
};
}  // namespace

void PeerGroup::InitMailbox() {
  if (std::getenv("B200KV_NO_MAILBOX") != nullptr) return;
  char name[64];
  std::memset(name, 0, sizeof(name));
  const size_t bytes = sizeof(MboxSlot) * 2 * static_cast<size_t>(world_);
  int fd = -1;
  if (rank_ == 0) {
    std::snprintf(name, sizeof(name), "/b200kv_%d_%llx", static_cast<int>(getpid()),
                  static_cast<unsigned long long>(std::chrono::steady_clock::now().time_since_epoch().count()));
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd >= 0 && ftruncate(fd, static_cast<off_t>(bytes)) != 0) {
      close(fd);
      shm_unlink(name);
      fd = -1;
    }
    if (fd < 0) name[0] = '\0';   // no shared memory here: everybody keeps the callback
  }
  std::vector<char> names(static_cast<size_t>(world_) * sizeof(name));
  KV_CHECK_EQ(fn_(name, names.data(), sizeof(name), ctx_), 0) << "all-gather callback failed";
  const char* shared = names.data();   // rank 0's entry
  int64_t ok = 0;
  if (shared[0] != '\0') {
    if (rank_ != 0) fd = shm_open(shared, O_RDWR, 0600);
    if (fd >= 0) {
      void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      if (p != MAP_FAILED) {
        mbox_ = static_cast<char*>(p);
        mbox_bytes_ = bytes;
        ok = 1;
      }
    }
  }
  // (a freshly truncated segment is zero-filled: every slot starts at seq 0)
  std::vector<int64_t> oks(world_);
  KV_CHECK_EQ(fn_(&ok, oks.data(), sizeof(int64_t), ctx_), 0) << "all-gather callback failed";
  if (rank_ == 0 && shared[0] != '\0') shm_unlink(shared);   // mappings stay valid; nothing is left behind
  for (int64_t v : oks) {
    if (v == 0 && mbox_ != nullptr) {   // a rank could not map it (another host / container limits)
      munmap(mbox_, mbox_bytes_);
      mbox_ = nullptr;
    }
  }
}

bool PeerGroup::MailboxAllGather(const void* send, void* recv, size_t nbytes) {
  if (mbox_ == nullptr) return false;
  static const double timeout_s = []() {
    const char* z = std::getenv("B200KV_PEER_TIMEOUT_S");
    return z ? std::atof(z) : 600.0;
  }();
  MboxSlot* slots = reinterpret_cast<MboxSlot*>(mbox_);
  const char* src = static_cast<const char*>(send);
  char* dst = static_cast<char*>(recv);
  for (size_t off = 0; off < nbytes || off == 0; off += kMboxPayload) {
    const size_t n = std::min(kMboxPayload, nbytes - off);
    const uint64_t seq = ++mbox_seq_;
    // two slots per rank: a rank can only be ONE gather ahead of the slowest (gather s+1 completes
    // only after everybody posted s+1, i.e. finished reading s), so slot (s & 1) is free at s+2
    MboxSlot& mine = slots[(seq & 1) * world_ + rank_];
    std::memcpy(mine.payload, src + off, n);
    mine.seq.store(seq, std::memory_order_release);
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < world_; ++r) {
      MboxSlot& s = slots[(seq & 1) * world_ + r];
      uint64_t spins = 0;
      while (s.seq.load(std::memory_order_acquire) < seq) {
        if ((++spins & 0xfff) == 0) {
          if (spins > (1u << 16)) usleep(50);   // a peer busy elsewhere: stop burning the core
          if (timeout_s > 0 &&
              std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
            KV_FATAL << "peer group: rank " << r << " did not join a collective of rank " << rank_
                     << " within B200KV_PEER_TIMEOUT_S; the ranks issued different KVStore calls, "
                     << "or that rank died";
          }
        }
      }
      std::memcpy(dst + static_cast<size_t>(r) * nbytes + off, s.payload, n);
    }
    if (nbytes == 0) break;
  }
  return true;
}

void PeerGroup::TestMailbox(int rank, int world, B200KVAllGatherFnC fn, void* ctx, const int64_t* mine,
                            int n, int rounds, int64_t* out, int* used_mailbox) {
  PeerGroup g;
  g.rank_ = rank;
  g.world_ = world;
  g.fn_ = fn;
  g.ctx_ = ctx;
  g.InitMailbox();
  *used_mailbox = g.mbox_ != nullptr ? 1 : 0;
  std::vector<int64_t> v(mine, mine + n);
  std::vector<int64_t> all;
  for (int it = 0; it < rounds; ++it) {
    for (auto& x : v) x += 1;                 // every round carries different values
    all = g.AllGatherI64(v);
  }
  std::memcpy(out, all.data(), all.size() * sizeof(int64_t));
  if (g.mbox_ != nullptr) munmap(g.mbox_, g.mbox_bytes_);
}

void PeerGroup::AllGather(const void* send, void* recv, size_t nbytes) {
  if (MailboxAllGather(send, recv, nbytes)) return;
  KV_CHECK_EQ(fn_(send, recv, nbytes, ctx_), 0) << "all-gather callback failed";
}

std::vector<int64_t> PeerGroup::AllGatherI64(const std::vector<int64_t>& mine) {
  if (mbox_ == nullptr) return GatherI64(fn_, ctx_, world_, mine);
  // same protocol as GatherI64 (lengths first, so a mismatch is an error instead of a memory
  // fault), over the shared-memory mailbox
  int64_t n = static_cast<int64_t>(mine.size());
  std::vector<int64_t> lens(world_, 0);
  AllGather(&n, lens.data(), sizeof(int64_t));
  for (int r = 0; r < world_; ++r) {
    KV_CHECK_EQ(lens[r], n) << "ranks issued different KVStore calls (rank " << r << " exchanges "
                            << lens[r] << " values, this rank " << n << ")";
  }
  std::vector<int64_t> all(static_cast<size_t>(world_) * mine.size());
  if (n > 0) AllGather(mine.data(), all.data(), mine.size() * sizeof(int64_t));
  return all;
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
  g->InitMailbox();
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
  const char* no_ipc = std::getenv("B200KV_GROUP_NO_IPC");
  bool ipc_ok = !(no_ipc != nullptr && no_ipc[0] != '\0' && no_ipc[0] != '0');
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      g->peer_base_[r] = g->arena_;
      g->pads_[r] = pad;
      continue;
    }
    KV_CHECK(all[r].dev != dev) << "ranks " << rank << " and " << r << " share gpu " << dev;
    if (!ipc_ok) continue;
    void* p = nullptr;
    if (external) {
      g->peer_base_[r] = ext_peers[r];
    } else if (cudaIpcOpenMemHandle(&p, all[r].arena, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
      g->peer_base_[r] = p;
    } else {
      cudaGetLastError();   // no peer access to that GPU: the group falls back to NCCL
      ipc_ok = false;
      continue;
    }
    if (cudaIpcOpenMemHandle(&p, all[r].pad, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
      g->pads_[r] = static_cast<uint32_t*>(p);
    } else {
      cudaGetLastError();
      ipc_ok = false;
    }
  }
  {
    // the fallback is a property of the GROUP: one rank without peer memory puts everybody on NCCL
    int64_t ok = ipc_ok ? 1 : 0;
    std::vector<int64_t> oks(world);
    g->AllGather(&ok, oks.data(), sizeof(int64_t));
    for (int64_t v : oks) ipc_ok = ipc_ok && v != 0;
    g->ipc_ok_ = ipc_ok;
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
  if (g->nccl_comm_ != nullptr) {
    Nccl::Get()->CommDestroy(g->nccl_comm_);
    g->nccl_comm_ = nullptr;
  }
  for (int r = 0; r < g->world_; ++r) {
    if (r == g->rank_) continue;
    if (!g->external_arena_ && g->peer_base_[r]) cudaIpcCloseMemHandle(g->peer_base_[r]);
    if (g->pads_[r]) cudaIpcCloseMemHandle(g->pads_[r]);
  }
  if (g->mbox_ != nullptr) munmap(g->mbox_, g->mbox_bytes_);
  // the arena itself is left to process teardown: pooled blocks carved from it may still be cached
  g_group = nullptr;
  delete g;
}

void* PeerGroup::NcclCommunicator() {
  if (nccl_comm_ != nullptr) return nccl_comm_;
  Nccl* n = Nccl::Get();
  NcclUniqueId id;
  std::memset(&id, 0, sizeof(id));
  if (rank_ == 0) n->GetUniqueId(&id);
  std::vector<NcclUniqueId> ids(world_);
  AllGather(&id, ids.data(), sizeof(NcclUniqueId));
  DeviceGuard guard(dev_);
  nccl_comm_ = n->CommInitRank(world_, ids[0], rank_);
  return nccl_comm_;
}

void PeerGroup::FillLaunch(DenseLaunch* L) {
  KV_CHECK(ipc_ok_) << "this peer group has no NVLink peer memory (NCCL fallback mode): in-kernel "
                    << "barriers and peer loads are unavailable";
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
