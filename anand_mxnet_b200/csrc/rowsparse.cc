// rowsparse.cc -- host side of row_sparse push / row_sparse_pull (KVStore members).
//
// Reference flow: CommDevice::ReduceRowSparse (src/kvstore/comm.h:478-501) copies every source to
// the owner GPU and sums there; KVStoreLocal::PullRowSparseImpl / Unique (kvstore_local.h:263-283,
// 428-472) sort+unique the requested ids and CommDevice::BroadcastRowSparse (comm.h:618-674)
// retains those rows. Here the owner GPU reads the peers' rows directly over NVLink (no staging
// copies of values), and the union works on the ids present, not on a table-high flag array.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>

#include "kvstore.h"
#include "scalar_parse.h"

namespace b200kv {

namespace {

// blocking read of one int64 the device just produced (the reference blocks at the same point:
// ndarray_function.cu:158-166, kvstore_utils.cu:77-85)
int64_t ReadCount(int dev, const int64_t* d_count) {
  Engine* eng = Engine::Get();
  int64_t h = 0;
  DeviceGuard g(dev);
  KV_CUDA(cudaMemcpyAsync(&h, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, eng->Stream(dev)));
  KV_CUDA(cudaStreamSynchronize(eng->Stream(dev)));
  return h;
}

struct Scratch {
  int dev;
  void* p;
  size_t bytes;
  Scratch(int d, size_t b) : dev(d), p(Engine::Get()->Alloc(d, b)), bytes(b) {}
  ~Scratch() { Engine::Get()->Free(dev, p, bytes); }
};

}  // namespace

void KVStore::PushRowSparse(KeyEntry& e, const std::vector<NDArray>& srcs_in) {
  KV_CHECK_EQ(e.stype, kRowSparseStorage)
      << "key " << e.key << " was initialised dense but a row_sparse value was pushed";
  KV_CHECK_EQ(e.dtype, kFloat32) << "row_sparse keys are float32 on this path";
  KV_CHECK(srcs_in.size() <= static_cast<size_t>(kMaxSrc));
  Engine* eng = Engine::Get();
  if (e.home < 0) {
    int pick = 0;
    for (auto& s : srcs_in) {
      if (s.on_gpu()) { pick = s.dev(); break; }
    }
    e.home = pick;
  }
  const int home = e.home;
  if (!e.rsp.on_gpu()) e.rsp = e.rsp.Copy(Context::GPU(home));
  std::vector<NDArray> srcs;
  std::vector<int> parts{home};
  for (auto& s : srcs_in) {
    KV_CHECK(s.shape() == e.shape) << "push: shape mismatch for row_sparse key " << e.key;
    KV_CHECK_EQ(s.dtype(), e.dtype) << "push: dtype mismatch for key " << e.key;
    if (!s.on_gpu()) {
      srcs.push_back(s.Copy(Context::GPU(home)));
    } else {
      srcs.push_back(s);
      if (std::find(parts.begin(), parts.end(), s.dev()) == parts.end()) parts.push_back(s.dev());
    }
  }
  if (parts.size() > 1) {
    int en = eng->EnablePeerAccess(parts);
    KV_CHECK_EQ(en, static_cast<int>(parts.size() * (parts.size() - 1)))
        << "GPU peer access is not available between all participating devices";
  }
  // ---- union of ids
  int64_t total = 0;
  std::vector<const int64_t*> h_idx;
  std::vector<const float*> h_val;
  std::vector<int64_t> h_n;
  for (auto& s : srcs) {
    if (!s.storage_initialized()) continue;  // all-zero source
    h_idx.push_back(s.row_ids());
    h_val.push_back(static_cast<const float*>(s.data()));
    h_n.push_back(s.nnr());
    total += s.nnr();
  }
  NDArray merged = NDArray::RowSparse(e.shape, Context::GPU(home), e.dtype);
  const int64_t row_len = static_cast<int64_t>(e.rsp.RowLength());
  if (total > 0) {
    merged.CheckAndAllocRows(total);
    for (auto& s : srcs) eng->BeginRead(s.dev(), *s.var());
    if (parts.size() > 1) eng->JoinStreams(parts);
    DeviceGuard g(home);
    cudaStream_t st = eng->Stream(home);
    const int nsrc = static_cast<int>(h_idx.size());
    Scratch ws(home, RspUnionWorkspaceBytes(total));
    Scratch tbl(home, 1024 + nsrc * 3 * sizeof(void*));
    int64_t* d_count = static_cast<int64_t*>(tbl.p);
    LaunchRspUnion(h_idx.data(), h_n.data(), nsrc, total, merged.row_ids(), d_count, ws.p, ws.bytes, st);
    eng->CountLaunch("rsp_union(cub sort+unique)", total * 16);
    const int64_t nnr = ReadCount(home, d_count);
    merged.SetNnr(nnr);
    // ---- in-order accumulation
    char* t = static_cast<char*>(tbl.p) + 256;
    KV_CUDA(cudaMemcpyAsync(t, h_idx.data(), nsrc * sizeof(void*), cudaMemcpyHostToDevice, st));
    KV_CUDA(cudaMemcpyAsync(t + nsrc * sizeof(void*), h_val.data(), nsrc * sizeof(void*),
                            cudaMemcpyHostToDevice, st));
    KV_CUDA(cudaMemcpyAsync(t + 2 * nsrc * sizeof(void*), h_n.data(), nsrc * sizeof(int64_t),
                            cudaMemcpyHostToDevice, st));
    RspSumLaunch L;
    L.src_idx = reinterpret_cast<const int64_t* const*>(t);
    L.src_val = reinterpret_cast<const float* const*>(t + nsrc * sizeof(void*));
    L.src_nrows = reinterpret_cast<const int64_t*>(t + 2 * nsrc * sizeof(void*));
    L.nsrc = nsrc;
    L.out_idx = merged.row_ids();
    L.out_val = static_cast<float*>(merged.data());
    L.nnr = nnr;
    L.row_len = row_len;
    LaunchRspSum(L, st);
    eng->CountLaunch("rsp_sum", static_cast<uint64_t>(total + nnr) * row_len * 4);
    if (parts.size() > 1) eng->JoinStreams(parts);
    uint64_t seq = eng->Issue(home);
    eng->MarkWrite(home, seq, merged.var());
    for (auto& s : srcs) {
      uint64_t sq = s.dev() == home ? seq : eng->Issue(s.dev());
      eng->MarkRead(s.dev(), sq, s.var());
    }
  }
  // ---- consume the merged gradient
  if (opt_.enabled && (opt_.kind == kOptSGD || opt_.kind == kOptAdam)) {
    KV_CHECK(opt_.lazy_update) << "lazy_update=False for row_sparse gradients is a next-row item";
    KV_CHECK_EQ(e.rsp.nnr(), e.shape[0])
        << "key " << e.key << ": the stored row_sparse weight must hold every row for sparse "
        << "optimizer updates (initialise it from a dense weight, as gluon does)";
    // Optimizer._update_count, then SGD._update_impl's non-aggregated branch / Adam.update
    auto it = opt_.count.find(e.key);
    int c = (it == opt_.count.end() ? opt_.begin_num_update : it->second) + 1;
    opt_.count[e.key] = c;
    opt_.num_update = std::max(opt_.num_update, c);
    if (merged.nnr() == 0) return;
    auto lm = opt_.lr_mult.find(e.key);
    auto wm = opt_.wd_mult.find(e.key);
    double lrd = opt_.lr * (lm == opt_.lr_mult.end() ? 1.0 : lm->second);
    double wdd = opt_.wd * (wm == opt_.wd_mult.end() ? 1.0 : wm->second);
    RspUpdateLaunch U;
    DevState& s = e.dev[home];
    const std::vector<int64_t> dshape = e.shape;
    auto zero_state = [&](NDArray* a) {
      if (!a->is_none()) return;
      *a = NDArray(dshape, Context::GPU(home), kFloat32);
      DeviceGuard g(home);
      KV_CUDA(cudaMemsetAsync(a->data(), 0, a->ByteSize(), eng->Stream(home)));
      eng->MarkWrite(home, eng->Issue(home), a->var());
    };
    if (opt_.kind == kOptSGD) {
      U.opt = opt_.momentum != 0.0 ? kOptSGD : kOptSGDSingle;  // sgd_mom_update / sgd_update
      if (opt_.momentum != 0.0) zero_state(&s.s1);
      U.momentum = opt_.momentum > 0 ? ScalarParam(opt_.momentum) : 0.f;
    } else {
      U.opt = kOptAdam;
      zero_state(&s.s1);
      zero_state(&s.s2);
      const double coef1 = 1.0 - std::pow(opt_.beta1, c), coef2 = 1.0 - std::pow(opt_.beta2, c);
      lrd *= std::sqrt(coef2) / coef1;
      U.beta1 = ScalarParam(opt_.beta1);
      U.beta2 = ScalarParam(opt_.beta2);
      U.eps = ScalarParam(opt_.eps);
    }
    U.lr = ScalarParam(lrd);
    U.wd = ScalarParam(wdd);
    U.rescale = ScalarParam(opt_.rescale);
    U.clip = opt_.clip != 0.0 ? ScalarParam(opt_.clip) : -1.f;
    U.w = static_cast<float*>(e.rsp.data());
    U.s1 = s.s1.is_none() ? nullptr : static_cast<float*>(s.s1.data());
    U.s2 = s.s2.is_none() ? nullptr : static_cast<float*>(s.s2.data());
    U.gidx = merged.row_ids();
    U.gval = static_cast<const float*>(merged.data());
    U.nrows = merged.nnr();
    U.row_len = row_len;
    DeviceGuard g(home);
    eng->BeginWrite(home, *e.rsp.var());
    LaunchRspUpdate(U, eng->Stream(home));
    eng->CountLaunch("rsp_update", static_cast<uint64_t>(merged.nnr()) * row_len * 4 * 3);
    uint64_t seq = eng->Issue(home);
    eng->MarkWrite(home, seq, e.rsp.var());
    eng->MarkRead(home, seq, merged.var());
    return;
  }
  if (updater_ != nullptr && !opt_.enabled) {
    NDArray* recv_h = new NDArray(merged);
    NDArray* local_h = new NDArray(e.rsp);
    if (key_type_ == 0 && str_updater_ != nullptr) {
      str_updater_(reverse_str_key_dict_[e.key].c_str(), recv_h, local_h, updater_handle_);
    } else {
      updater_(e.key, recv_h, local_h, updater_handle_);
    }
    return;
  }
  e.rsp = merged;  // no updater: local = merged (kvstore_local.h:237-243)
}

// KVStoreLocal::Unique (kvstore_local.h:428-472): ids of any integer/float dtype and any shape ->
// ascending unique int64 in a row_sparse container whose aux_shape is the unique count.
NDArray KVStore::UniqueRowIds(const NDArray& row_ids, int dev, int64_t* count) {
  Engine* eng = Engine::Get();
  const int64_t n = static_cast<int64_t>(row_ids.Size());
  NDArray ids_dev = row_ids.on_gpu() && row_ids.dev() == dev ? row_ids : row_ids.Copy(Context::GPU(dev));
  NDArray out({std::max<int64_t>(n, 1)}, Context::GPU(dev), kInt64);
  *count = 0;
  if (n == 0) return out;
  DeviceGuard g(dev);
  cudaStream_t st = eng->Stream(dev);
  eng->BeginRead(dev, *ids_dev.var());
  NDArray ids64 = ids_dev;
  if (ids_dev.dtype() != kInt64) {
    ids64 = NDArray({n}, Context::GPU(dev), kInt64);
    LaunchCast(ids64.data(), kInt64, ids_dev.data(), ids_dev.dtype(), n, st);
    eng->CountLaunch("cast", 0);
  }
  Scratch ws(dev, UniqueWorkspaceBytes(n));
  Scratch cnt(dev, 256);
  LaunchUnique(static_cast<const int64_t*>(ids64.data()), n, static_cast<int64_t*>(out.data()),
               static_cast<int64_t*>(cnt.p), ws.p, ws.bytes, st);
  eng->CountLaunch("unique(cub sort+unique)", n * 16);
  uint64_t seq = eng->Issue(dev);
  eng->MarkRead(dev, seq, ids_dev.var());
  eng->MarkWrite(dev, seq, out.var());
  *count = ReadCount(dev, static_cast<int64_t*>(cnt.p));
  return out;  // the first *count entries are the ascending unique ids
}

void KVStore::PullRowSparse(const std::vector<int>& keys, const std::vector<NDArray>& outs,
                            const std::vector<NDArray>& row_ids, int) {
  KV_CHECK_EQ(keys.size(), outs.size());
  KV_CHECK_EQ(keys.size(), row_ids.size());
  // GroupKVPairsPullRsp (kvstore_local.h:352-371): stable by key, storage types validated
  std::vector<size_t> order(keys.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
  for (size_t i : order) {
    KV_CHECK_EQ(outs[i].stype(), kRowSparseStorage)
        << "Expected row_sparse storage type for row_sparse_pull values, but detected storage type "
        << outs[i].stype();
    KV_CHECK_EQ(row_ids[i].stype(), kDefaultStorage)
        << "Expected default storage type for row_sparse_pull rowids, but detected storage type "
        << row_ids[i].stype();
  }
  for (size_t i : order) {
    KeyEntry& e = Entry(keys[i]);
    KV_CHECK_EQ(e.stype, kRowSparseStorage) << "PullRowSparse expects row_sparse src NDArray";
    PullRowSparseOne(e, outs[i], row_ids[i]);
  }
}

void KVStore::PullRowSparseOne(KeyEntry& e, const NDArray& out, const NDArray& row_ids) {
  Engine* eng = Engine::Get();
  if (e.home < 0) e.home = out.on_gpu() ? out.dev() : (row_ids.on_gpu() ? row_ids.dev() : 0);
  const int home = e.home;
  if (!e.rsp.on_gpu()) e.rsp = e.rsp.Copy(Context::GPU(home));
  KV_CHECK(out.shape() == e.shape) << "row_sparse_pull: out shape mismatch for key " << e.key;
  KV_CHECK_EQ(out.dtype(), e.dtype) << "row_sparse_pull: out dtype mismatch for key " << e.key;
  if (out.SameStorage(e.rsp)) {
    std::cerr << "The output of row_sparse_pull() on key " << e.key << " refers to the same NDArray "
              << "as the one stored in KVStore. Consider a new NDArray buffer for the output.\n";
  }
  int64_t m = 0;
  NDArray uniq = UniqueRowIds(row_ids, home, &m);
  // retain on the owner, straight into `out` when it lives on a GPU (peer store over NVLink)
  const bool direct = out.on_gpu();
  NDArray target = direct ? out : NDArray::RowSparse(e.shape, Context::GPU(home), e.dtype);
  if (m == 0 || !e.rsp.storage_initialized()) {
    // FillZerosRspImpl (sparse_retain-inl.h:271-275)
    eng->WaitToWrite(*out.var());
    out.SetNnr(0);
    return;
  }
  if (direct && out.dev() != home) {
    int en = eng->EnablePeerAccess({home, out.dev()});
    KV_CHECK_EQ(en, 2) << "GPU peer access is not available between gpu " << home << " and gpu "
                       << out.dev();
  }
  target.CheckAndAllocRows(m);
  eng->BeginRead(home, *e.rsp.var());
  eng->BeginWrite(home, *target.var());
  if (direct && out.dev() != home) eng->JoinStreams({home, out.dev()});
  RetainLaunch L;
  L.src_idx = e.rsp.row_ids();
  L.src_val = static_cast<const float*>(e.rsp.data());
  L.src_nnr = e.rsp.nnr();
  L.src_dense_rows = e.rsp.nnr() == e.shape[0] ? 1 : 0;  // sparse_retain-inl.h:290
  L.ids = static_cast<const int64_t*>(uniq.data());
  L.nids = m;
  L.row_len = static_cast<int64_t>(e.rsp.RowLength());
  L.out_idx = target.row_ids();
  L.out_val = static_cast<float*>(target.data());
  {
    DeviceGuard g(home);
    LaunchRetain(L, eng->Stream(home));
  }
  eng->CountLaunch("sparse_retain", static_cast<uint64_t>(m) * (L.row_len * 8 + 16));
  if (direct && out.dev() != home) eng->JoinStreams({home, out.dev()});
  uint64_t seq = eng->Issue(home);
  eng->MarkRead(home, seq, e.rsp.var());
  eng->MarkRead(home, seq, uniq.var());
  if (direct) {
    uint64_t sq = out.dev() == home ? seq : eng->Issue(out.dev());
    eng->MarkWrite(out.dev(), sq, out.var());
  } else {
    eng->MarkWrite(home, seq, target.var());
    CopyFromTo(target, out);
  }
}

}  // namespace b200kv
