// engine.cc -- see engine.h.
#include "engine.h"

#include <algorithm>

namespace b200kv {

Engine* Engine::Get() {
  static Engine* inst = new Engine();  // intentionally leaked: no static-destruction-order races
  return inst;
}

Engine::Engine() {}

void Engine::Init() {
  if (inited_) return;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    KV_FATAL << "libb200kv needs a CUDA device and found none (" << cudaGetErrorString(e)
             << "); there is no CPU fallback for the KVStore path";
  }
  n = std::min(n, kMaxDevices);
  devs_.resize(n);
  inited_ = true;
}

int Engine::NumDevices() {
  Init();
  return static_cast<int>(devs_.size());
}

cudaStream_t Engine::Stream(int dev) {
  Init();
  KV_CHECK(dev >= 0 && dev < static_cast<int>(devs_.size())) << "invalid gpu id " << dev;
  Dev& d = devs_[dev];
  if (d.own == nullptr) {
    DeviceGuard g(dev);
    KV_CUDA(cudaStreamCreateWithFlags(&d.own, cudaStreamNonBlocking));
    d.ring.resize(16);
    for (auto& ev : d.ring) KV_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (d.cur == nullptr) d.cur = d.own;
  }
  return d.cur;
}

void Engine::SetStream(int dev, cudaStream_t s) {
  cudaStream_t old = Stream(dev);
  Dev& d = devs_[dev];
  cudaStream_t next = s ? s : d.own;
  if (next == old) return;
  // work already issued on the old stream stays ordered before work on the new one
  DeviceGuard g(dev);
  cudaEvent_t ev = RecordLatest(dev);
  KV_CUDA(cudaStreamWaitEvent(next, ev, 0));
  d.cur = next;
}

cudaEvent_t Engine::RecordLatest(int dev) {
  Dev& d = devs_[dev];
  Stream(dev);
  if (d.latest == nullptr || d.recorded < d.issued) {
    DeviceGuard g(dev);
    d.ring_pos = (d.ring_pos + 1) % d.ring.size();
    d.latest = d.ring[d.ring_pos];
    KV_CUDA(cudaEventRecord(d.latest, d.cur));
    d.recorded = d.issued;
  }
  return d.latest;
}

uint64_t Engine::Issue(int dev) {
  Stream(dev);
  return ++devs_[dev].issued;
}

void Engine::StreamWait(int dev, Tag t) {
  if (t.dev < 0 || t.dev == dev) return;  // same device: stream order
  Stream(dev);
  Dev& d = devs_[dev];
  if (d.waited[t.dev] >= t.seq) return;
  if (devs_[t.dev].completed >= t.seq) return;
  cudaEvent_t ev = RecordLatest(t.dev);
  DeviceGuard g(dev);
  KV_CUDA(cudaStreamWaitEvent(d.cur, ev, 0));
  d.waited[t.dev] = devs_[t.dev].recorded;
}

void Engine::HostWait(Tag t) {
  if (t.dev < 0) return;
  Init();
  Dev& d = devs_[t.dev];
  if (d.completed >= t.seq) return;
  cudaEvent_t ev = RecordLatest(t.dev);
  uint64_t upto = d.recorded;
  KV_CUDA(cudaEventSynchronize(ev));
  d.completed = std::max(d.completed, upto);
}

void Engine::BeginRead(int dev, const Var& v) { StreamWait(dev, v.writer); }

void Engine::BeginWrite(int dev, const Var& v) {
  StreamWait(dev, v.writer);
  if (v.has_readers) {
    for (int e = 0; e < static_cast<int>(devs_.size()); ++e) {
      if (v.reader_seq[e]) StreamWait(dev, Tag{e, v.reader_seq[e]});
    }
  }
}

void Engine::MarkRead(int dev, uint64_t seq, Var* v) {
  v->reader_seq[dev] = seq;
  v->has_readers = true;
}

void Engine::MarkWrite(int dev, uint64_t seq, Var* v) {
  v->writer = Tag{dev, seq};
  if (v->has_readers) {
    for (int e = 0; e < kMaxDevices; ++e) v->reader_seq[e] = 0;
    v->has_readers = false;
  }
}

void Engine::WaitToRead(const Var& v) { HostWait(v.writer); }

void Engine::WaitToWrite(const Var& v) {
  HostWait(v.writer);
  if (v.has_readers) {
    for (int e = 0; e < static_cast<int>(devs_.size()); ++e) {
      if (v.reader_seq[e]) HostWait(Tag{e, v.reader_seq[e]});
    }
  }
}

void Engine::WaitAll() {
  if (!inited_) return;
  for (int e = 0; e < static_cast<int>(devs_.size()); ++e) {
    Dev& d = devs_[e];
    if (d.cur == nullptr) continue;
    DeviceGuard g(e);
    uint64_t upto = d.issued;
    KV_CUDA(cudaStreamSynchronize(d.cur));
    d.completed = std::max(d.completed, upto);
  }
}

void Engine::JoinStreams(const std::vector<int>& devs) {
  if (devs.size() < 2) return;
  // hub = devs[0]: it waits for every other stream, then every other stream waits for it.
  // Issue() on every stream first so that a fresh event is recorded even when the only new work on
  // a (caller-provided) stream was enqueued by the caller's framework and never counted here.
  const int hub = devs[0];
  for (size_t i = 1; i < devs.size(); ++i) {
    uint64_t s = Issue(devs[i]);
    StreamWait(hub, Tag{devs[i], s});
  }
  uint64_t seq = Issue(hub);
  for (size_t i = 1; i < devs.size(); ++i) StreamWait(devs[i], Tag{hub, seq});
}

// ---------------------------------------------------------------------------------------------
size_t Engine::RoundSize(size_t bytes) {
  if (bytes == 0) bytes = 1;
  if (bytes <= (1u << 20)) return (bytes + 511) & ~static_cast<size_t>(511);
  return (bytes + (2u << 20) - 1) & ~static_cast<size_t>((2u << 20) - 1);
}

void* Engine::Alloc(int dev, size_t bytes) {
  Stream(dev);
  Dev& d = devs_[dev];
  size_t r = RoundSize(bytes);
  auto it = d.pool.find(r);
  if (it != d.pool.end()) {
    void* p = it->second;
    d.pool.erase(it);
    return p;
  }
  DeviceGuard g(dev);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, r);
  if (e != cudaSuccess) {
    cudaGetLastError();
    // release the cache and retry once
    for (auto& kv : d.pool) cudaFree(kv.second);
    d.pool.clear();
    KV_CUDA(cudaMalloc(&p, r));
  }
  d.bytes += r;
  return p;
}

void Engine::Free(int dev, void* p, size_t bytes) {
  if (p == nullptr || !inited_) return;
  devs_[dev].pool.emplace(RoundSize(bytes), p);
}

void* Engine::AllocPinned(size_t bytes) {
  Init();
  size_t r = RoundSize(bytes);
  auto it = pinned_pool_.find(r);
  if (it != pinned_pool_.end()) {
    void* p = it->second;
    pinned_pool_.erase(it);
    return p;
  }
  void* p = nullptr;
  KV_CUDA(cudaHostAlloc(&p, r, cudaHostAllocPortable));
  return p;
}

void Engine::FreePinned(void* p, size_t bytes) {
  if (p == nullptr) return;
  pinned_pool_.emplace(RoundSize(bytes), p);
}

size_t Engine::BytesAllocated(int dev) {
  Init();
  return devs_[dev].bytes;
}

int Engine::EnablePeerAccess(const std::vector<int>& devs) {
  Init();
  int enabled = 0;
  for (int a : devs) {
    DeviceGuard g(a);
    for (int b : devs) {
      if (a == b) continue;
      if (peer_[a][b]) {
        ++enabled;
        continue;
      }
      int can = 0;
      KV_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
      if (!can) continue;
      cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        peer_[a][b] = true;
        ++enabled;
      } else {
        cudaGetLastError();
      }
    }
  }
  return enabled;
}

bool Engine::PeerEnabled(int a, int b) { return a == b || peer_[a][b]; }

void Engine::CountLaunch(const char* name, uint64_t algorithmic_bytes) {
  ++launch_count;
  last_kernel = name;
  last_kernel_bytes = algorithmic_bytes;
}

void Engine::Shutdown() {
  if (!inited_) return;
  WaitAll();
}

}  // namespace b200kv
