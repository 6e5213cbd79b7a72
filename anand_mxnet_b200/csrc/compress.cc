// compress.cc -- KVStore::CompressedReduce: the 2-bit gradient-compression variant of the device
// reduce (reference: CommDevice::ReduceCompressed, src/kvstore/comm.h:552-596; parameters
// src/kvstore/gradient_compression.h:41-51: type='2bit', threshold=0.5).
#include <algorithm>

#include "kvstore.h"

namespace b200kv {

NDArray KVStore::CompressedReduce(KeyEntry& e, const std::vector<NDArray>& srcs_in) {
  KV_CHECK_EQ(e.dtype, kFloat32) << "2-bit gradient compression needs float32 gradients";
  KV_CHECK(!dist_) << "gradient compression is not supported by the one-rank-per-GPU store yet";
  Engine* eng = Engine::Get();
  if (e.striped) EnsureWhole(e, devset_[0]);
  if (e.home < 0) {
    int pick = 0;
    for (auto& s : srcs_in) {
      if (s.on_gpu()) { pick = s.dev(); break; }
    }
    EnsureOnDevice(e, pick);
  }
  const int home = e.home;
  const size_t n = e.size;
  const int64_t words = static_cast<int64_t>((n + 15) / 16);
  const size_t nsrc = srcs_in.size();
  if (e.gc_residual.size() < nsrc) {
    e.gc_residual.resize(nsrc);
    e.gc_compressed.resize(nsrc);
  }
  std::vector<int> parts{home};
  std::vector<const uint32_t*> comp_ptrs(nsrc);
  for (size_t i = 0; i < nsrc; ++i) {
    NDArray src = srcs_in[i];
    if (!src.on_gpu()) src = StageSrc(e, i, src, home), CopyFromTo(srcs_in[i], src);
    const int d = src.dev();
    if (std::find(parts.begin(), parts.end(), d) == parts.end()) parts.push_back(d);
    NDArray& res = e.gc_residual[i];
    if (res.is_none() || res.dev() != d) {
      // buf.residual[i] = 0 on the source's device (comm.h:569-571)
      res = NDArray(e.shape, Context::GPU(d), kFloat32);
      DeviceGuard g(d);
      eng->BeginWrite(d, *res.var());   // first writer of a possibly recycled block
      KV_CUDA(cudaMemsetAsync(res.data(), 0, res.ByteSize(), eng->Stream(d)));
      eng->MarkWrite(d, eng->Issue(d), res.var());
      e.gc_compressed[i] = NDArray({words}, Context::GPU(d), kInt32);
    }
    NDArray& comp = e.gc_compressed[i];
    // quantise where the gradient lives: what crosses NVLink afterwards is 1/16 of it
    eng->BeginRead(d, *src.var());
    eng->BeginWrite(d, *res.var());
    eng->BeginWrite(d, *comp.var());
    {
      DeviceGuard g(d);
      LaunchQuantize2Bit(static_cast<const float*>(src.data()), static_cast<float*>(res.data()),
                         static_cast<uint32_t*>(comp.data()), n, gc_threshold_, eng->Stream(d));
    }
    eng->CountLaunch("quantize_2bit", n * 12 + words * 4);
    const uint64_t seq = eng->Issue(d);
    eng->MarkRead(d, seq, src.var());
    eng->MarkWrite(d, seq, res.var());
    eng->MarkWrite(d, seq, comp.var());
    comp_ptrs[i] = static_cast<const uint32_t*>(comp.data());
  }
  if (parts.size() > 1) {
    const int en = eng->EnablePeerAccess(parts);
    KV_CHECK_EQ(en, static_cast<int>(parts.size() * (parts.size() - 1)))
        << "GPU peer access is not available between all participating devices";
    eng->JoinStreams(parts);
  }
  if (e.gc_merged.is_none() || e.gc_merged.dev() != home) {
    e.gc_merged = NDArray(e.shape, Context::GPU(home), kFloat32);
  }
  for (size_t i = 0; i < nsrc; ++i) eng->BeginRead(home, *e.gc_compressed[i].var());
  eng->BeginWrite(home, *e.gc_merged.var());
  {
    DeviceGuard g(home);
    LaunchDequantizeSum2Bit(comp_ptrs.data(), static_cast<int>(nsrc),
                            static_cast<float*>(e.gc_merged.data()), n, gc_threshold_,
                            eng->Stream(home));
  }
  eng->CountLaunch("dequantize_sum_2bit", n * 4 + nsrc * words * 4);
  const uint64_t seq = eng->Issue(home);
  eng->MarkWrite(home, seq, e.gc_merged.var());
  for (size_t i = 0; i < nsrc; ++i) eng->MarkRead(home, seq, e.gc_compressed[i].var());
  if (parts.size() > 1) eng->JoinStreams(parts);
  return e.gc_merged;
}

}  // namespace b200kv
