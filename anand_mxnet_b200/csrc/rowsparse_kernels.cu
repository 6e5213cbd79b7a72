// rowsparse_kernels.cu -- row_sparse push / row_sparse_pull kernels for sm_100a.
//
// Reference semantics (bit-exact targets, all integer/index results identical):
//   * reduce: out ids = ascending unique union of the sources' ids; out rows start at 0.0f and the
//     sources are accumulated in list order (src/ndarray/ndarray_function.cc:59-175 on CPU,
//     ndarray_function.cu:104-190 on GPU -- the GPU version marks a flag array as long as the whole
//     table (1M rows -> 8 MB scan); here the union is a radix sort + unique over the <= N*nnr ids
//     actually present, independent of table height);
//   * unique: sort + unique of the requested row ids (src/kvstore/kvstore_utils.cu:43-97);
//   * retain: every requested id is emitted, rows copied where present else zero
//     (src/operator/tensor/sparse_retain-inl.h:121-150,262-323);
//   * lazy optimizer updates over the gradient's rows (optimizer_op-inl.h:426-475,749-801,1350-1408).
//
// Row gathers/scatters are HBM-bound: one warp per row, 16-byte accesses along the row, ids looked
// up once per warp by binary search (sources are sorted), no atomics (the union makes every output
// row single-writer and the in-order source loop keeps the float association deterministic).
#include <cub/cub.cuh>

#include "common.h"
#include "kernels.h"
#include "opt_math.cuh"

namespace b200kv {
namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
template <int OPT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) rsp_update_kernel(RspUpdateLaunch p) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= p.nrows) return;
  const int lane = threadIdx.x & 31;
  Hyper h{p.lr, p.wd, p.momentum, p.rescale, p.clip, p.beta1, p.beta2, p.eps};
  const int64_t wrow = p.gidx[row];
  float* w = p.w + wrow * p.row_len;
  float* s1 = p.s1 ? p.s1 + wrow * p.row_len : nullptr;
  float* s2 = p.s2 ? p.s2 + wrow * p.row_len : nullptr;
  const float* g = p.gval + row * p.row_len;
  const bool vec = (p.row_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.w) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.gval) & 15) == 0) &&
                   (!p.s1 || (reinterpret_cast<uintptr_t>(p.s1) & 15) == 0) &&
                   (!p.s2 || (reinterpret_cast<uintptr_t>(p.s2) & 15) == 0);
  auto one = [&](float wv, float gv, float& a, float& b) -> float {
    if (OPT == kOptAdam) return step_adam_lazy(wv, gv, a, b, h);
    return step<OPT>(wv, gv, a, b, OPT == kOptSGD, h);
  };
  if (vec) {
    const int64_t nv = p.row_len / 4;
    for (int64_t v = lane; v < nv; v += 32) {
      float4 wv = reinterpret_cast<float4*>(w)[v];
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + v);
      float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
      if (OPT == kOptSGD || OPT == kOptAdam) a = reinterpret_cast<float4*>(s1)[v];
      if (OPT == kOptAdam) b = reinterpret_cast<float4*>(s2)[v];
      wv.x = one(wv.x, gv.x, a.x, b.x);
      wv.y = one(wv.y, gv.y, a.y, b.y);
      wv.z = one(wv.z, gv.z, a.z, b.z);
      wv.w = one(wv.w, gv.w, a.w, b.w);
      reinterpret_cast<float4*>(w)[v] = wv;
      if (OPT == kOptSGD || OPT == kOptAdam) reinterpret_cast<float4*>(s1)[v] = a;
      if (OPT == kOptAdam) reinterpret_cast<float4*>(s2)[v] = b;
    }
  } else {
    for (int64_t j = lane; j < p.row_len; j += 32) {
      float a = 0.f, b = 0.f;
      if (OPT == kOptSGD || OPT == kOptAdam) a = s1[j];
      if (OPT == kOptAdam) b = s2[j];
      w[j] = one(w[j], g[j], a, b);
      if (OPT == kOptSGD || OPT == kOptAdam) s1[j] = a;
      if (OPT == kOptAdam) s2[j] = b;
    }
  }
}

// out_val[r] = 0.0f; for s in list order: if s holds out_idx[r]: out_val[r] += src_val[s][row]
__global__ void __launch_bounds__(kWarpsPerBlock * 32) rsp_sum_kernel(RspSumLaunch p) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= p.nnr) return;
  const int lane = threadIdx.x & 31;
  const int64_t id = p.out_idx[r];
  float* out = p.out_val + r * p.row_len;
  const bool vec = (p.row_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out_val) & 15) == 0);
  // positions of this row in every source (-1: absent); nsrc <= kMaxSrc
  int64_t pos[kMaxSrc];
  for (int s = 0; s < p.nsrc; ++s) {
    const int64_t n = p.src_nrows[s];
    const int64_t* idx = p.src_idx[s];
    const int64_t lb = lower_bound_i64(idx, n, id);
    pos[s] = (lb < n && idx[lb] == id) ? lb : -1;
  }
  if (vec) {
    const int64_t nv = p.row_len / 4;
    for (int64_t v = lane; v < nv; v += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < p.nsrc; ++s) {
        if (pos[s] < 0) continue;
        const float* sv = p.src_val[s] + pos[s] * p.row_len;
        if ((reinterpret_cast<uintptr_t>(sv) & 15) == 0) {
          const float4 x = __ldcs(reinterpret_cast<const float4*>(sv) + v);
          acc.x = __fadd_rn(acc.x, x.x); acc.y = __fadd_rn(acc.y, x.y);
          acc.z = __fadd_rn(acc.z, x.z); acc.w = __fadd_rn(acc.w, x.w);
        } else {
          acc.x = __fadd_rn(acc.x, sv[4 * v]); acc.y = __fadd_rn(acc.y, sv[4 * v + 1]);
          acc.z = __fadd_rn(acc.z, sv[4 * v + 2]); acc.w = __fadd_rn(acc.w, sv[4 * v + 3]);
        }
      }
      reinterpret_cast<float4*>(out)[v] = acc;
    }
  } else {
    for (int64_t j = lane; j < p.row_len; j += 32) {
      float acc = 0.f;
      for (int s = 0; s < p.nsrc; ++s) {
        if (pos[s] >= 0) acc = __fadd_rn(acc, p.src_val[s][pos[s] * p.row_len + j]);
      }
      out[j] = acc;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) retain_kernel(RetainLaunch p) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= p.nids) return;
  const int lane = threadIdx.x & 31;
  const int64_t id = p.ids[i];
  if (lane == 0) p.out_idx[i] = id;
  int64_t j = -1;
  if (p.src_dense_rows) {
    j = id;
  } else if (p.src_nnr > 0) {
    const int64_t lb = lower_bound_i64(p.src_idx, p.src_nnr, id);
    if (lb < p.src_nnr && p.src_idx[lb] == id) j = lb;
  }
  float* out = p.out_val + i * p.row_len;
  const float* src = j >= 0 ? p.src_val + j * p.row_len : nullptr;
  const bool vec = (p.row_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out_val) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.src_val) & 15) == 0);
  if (vec) {
    const int64_t nv = p.row_len / 4;
    for (int64_t v = lane; v < nv; v += 32) {
      const float4 x = src ? __ldcs(reinterpret_cast<const float4*>(src) + v)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      __stcs(reinterpret_cast<float4*>(out) + v, x);
    }
  } else {
    for (int64_t c = lane; c < p.row_len; c += 32) out[c] = src ? src[c] : 0.f;
  }
}

struct UnionLayout {
  size_t off_concat, off_sorted, off_temp, temp_bytes, total;
};

UnionLayout Layout(int64_t n) {
  UnionLayout l;
  const size_t ids = (static_cast<size_t>(n) * sizeof(int64_t) + 255) & ~static_cast<size_t>(255);
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, static_cast<const int64_t*>(nullptr),
                                 static_cast<int64_t*>(nullptr), static_cast<int>(n));
  cub::DeviceSelect::Unique(nullptr, t2, static_cast<const int64_t*>(nullptr),
                            static_cast<int64_t*>(nullptr), static_cast<int64_t*>(nullptr),
                            static_cast<int>(n));
  l.off_concat = 0;
  l.off_sorted = ids;
  l.off_temp = 2 * ids;
  l.temp_bytes = (std::max(t1, t2) + 255) & ~static_cast<size_t>(255);
  l.total = l.off_temp + l.temp_bytes + 256;
  return l;
}

}  // namespace

void LaunchRspUpdate(const RspUpdateLaunch& p, cudaStream_t stream) {
  if (p.nrows <= 0 || p.row_len <= 0) return;
  const int blocks = static_cast<int>((p.nrows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  switch (p.opt) {
    case kOptSGDSingle: rsp_update_kernel<kOptSGDSingle><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p); break;
    case kOptSGD: rsp_update_kernel<kOptSGD><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p); break;
    case kOptAdam: rsp_update_kernel<kOptAdam><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p); break;
    default: KV_FATAL << "row_sparse update: unsupported optimizer kind " << p.opt;
  }
  KV_CUDA(cudaGetLastError());
}

size_t RspUnionWorkspaceBytes(int64_t total_ids) { return Layout(std::max<int64_t>(total_ids, 1)).total; }
size_t UniqueWorkspaceBytes(int64_t n) { return Layout(std::max<int64_t>(n, 1)).total; }

void LaunchUnique(const int64_t* ids, int64_t n, int64_t* out, int64_t* d_count, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0) {
    KV_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), stream));
    return;
  }
  KV_CHECK(n < (1LL << 31)) << "too many row ids";
  UnionLayout l = Layout(n);
  KV_CHECK(workspace_bytes >= l.total) << "unique: workspace too small";
  char* ws = static_cast<char*>(workspace);
  int64_t* sorted = reinterpret_cast<int64_t*>(ws + l.off_sorted);
  size_t tb = l.temp_bytes;
  KV_CUDA(cub::DeviceRadixSort::SortKeys(ws + l.off_temp, tb, ids, sorted, static_cast<int>(n), 0,
                                         64, stream));
  tb = l.temp_bytes;
  KV_CUDA(cub::DeviceSelect::Unique(ws + l.off_temp, tb, sorted, out, d_count, static_cast<int>(n),
                                    stream));
}

void LaunchRspUnion(const int64_t* const* src_idx, const int64_t* src_nrows, int nsrc,
                    int64_t total_ids, int64_t* out_idx, int64_t* d_nnr, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream) {
  // src_idx / src_nrows are HOST arrays of device pointers / counts here (gather by memcpy)
  UnionLayout l = Layout(std::max<int64_t>(total_ids, 1));
  KV_CHECK(workspace_bytes >= l.total) << "rsp union: workspace too small";
  char* ws = static_cast<char*>(workspace);
  int64_t* concat = reinterpret_cast<int64_t*>(ws + l.off_concat);
  int64_t off = 0;
  for (int s = 0; s < nsrc; ++s) {
    if (src_nrows[s] == 0) continue;
    KV_CUDA(cudaMemcpyAsync(concat + off, src_idx[s], src_nrows[s] * sizeof(int64_t),
                            cudaMemcpyDefault, stream));
    off += src_nrows[s];
  }
  LaunchUnique(concat, total_ids, out_idx, d_nnr, workspace, workspace_bytes, stream);
}

void LaunchRspSum(const RspSumLaunch& p, cudaStream_t stream) {
  if (p.nnr <= 0 || p.row_len <= 0) return;
  KV_CHECK(p.nsrc <= kMaxSrc);
  const int blocks = static_cast<int>((p.nnr + kWarpsPerBlock - 1) / kWarpsPerBlock);
  rsp_sum_kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p);
  KV_CUDA(cudaGetLastError());
}

void LaunchRetain(const RetainLaunch& p, cudaStream_t stream) {
  if (p.nids <= 0) return;
  const int blocks = static_cast<int>((p.nids + kWarpsPerBlock - 1) / kWarpsPerBlock);
  retain_kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(p);
  KV_CUDA(cudaGetLastError());
}

}  // namespace b200kv
