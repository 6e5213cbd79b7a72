// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product library.
//
// The row_sparse half of oracle/_ref/libmxref.so: the reference's OWN code for
//   * the row_sparse reduce (src/ndarray/ndarray_function.cc:59-153: GetUniqueRspRowIdx and
//     ElementwiseSumRspImpl -- sorted unique union of the sources' row ids, rows zero-initialised and
//     accumulated source by source in list order), and
//   * sparse_retain (src/operator/tensor/sparse_retain-inl.h:121-150, 205-262: the kernels behind
//     row_sparse_pull),
//   * the grouping of a call's (key, value) pairs (KVStoreLocal::GroupKVPairs, kvstore_local.h:377-407),
// so that oracle/kvoracle.c's restatement of them is pinned to live reference code, not only to the
// reference's unit-test identities.
//
// The retain kernels are header templates over plain pointers and are called directly. The reduce
// functions are templates in a .cc file that take mxnet::NDArray objects, which cannot exist
// without libmxnet (Storage / Engine). oracle/Makefile therefore EXTRACTS the two function
// templates from the reference file, where it lies, into oracle/_ref/ for the duration of the
// compile (deleted afterwards; never committed or shipped) and this file compiles them unchanged against a stand-in `NDArray` that offers exactly
// the accessors they use (data / aux_data / storage_shape / aux_shape / storage_initialized /
// storage_type), backed by the reference's real TBlob / TShape / mshadow types.
// common::ParallelSort (src/common/utils.h) is std::sort here: same result for integers.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>
#include <dmlc/logging.h>
#include <mshadow/tensor.h>
#include <mxnet/base.h>
#include <mxnet/tensor_blob.h>
#include <mxnet/op_attr_types.h>
#include "operator/tensor/sparse_retain-inl.h"

namespace mxref_sparse {

using mshadow::cpu;
using mshadow::index_t;
enum { kRowSparseStorage = 1 };
namespace rowsparse { enum { kIdx = 0 }; }
namespace common {
template <typename It>
void ParallelSort(It first, It last, int) { std::sort(first, last); }
}  // namespace common

// stand-in for mxnet::NDArray (row_sparse, float32 values, int64 row ids)
struct NDArray {
  float* val = nullptr;
  int64_t* idx = nullptr;
  size_t nnr = 0, row_len = 0;
  int storage_type() const { return kRowSparseStorage; }
  bool storage_initialized() const { return nnr > 0; }
  mxnet::TBlob data() const {
    return mxnet::TBlob(val, mxnet::TShape(mshadow::Shape2(nnr, row_len)), cpu::kDevMask, mshadow::kFloat32);
  }
  mxnet::TBlob aux_data(int) const {
    return mxnet::TBlob(idx, mxnet::TShape(mshadow::Shape1(nnr)), cpu::kDevMask, mshadow::kInt64);
  }
  mxnet::TShape storage_shape() const { return mxnet::TShape(mshadow::Shape2(nnr, row_len)); }
  mxnet::TShape aux_shape(int) const { return mxnet::TShape(mshadow::Shape1(nnr)); }
};

// ---- the reference's two function templates, extracted at build time (see oracle/Makefile)
#include "_ref/excerpt_rsp_unique.inc"
#include "_ref/excerpt_rsp_sum.inc"

}  // namespace mxref_sparse

// KVStoreLocal::GroupKVPairs (src/kvstore/kvstore_local.h:377-407), a member template of a class that
// cannot be instantiated without libmxnet: the function template alone is cut out of the header at
// build time (oracle/Makefile) and compiled unchanged, as a free function, with V = position in the call.
namespace mxref_group {
#include "_ref/excerpt_group_kv_pairs.inc"
}  // namespace mxref_group

extern "C" {

// Groups the n (key, position) pairs of one call the way the reference does: uniq[] = distinct keys
// ascending, counts[] = values per key, positions[] = the call positions in the order the reference
// hands them to the reduce (group after group). Returns the number of distinct keys.
int64_t mxref_group_kv_pairs(int64_t n, const int* keys, int* uniq, int* counts, int* positions) {
  if (n == 0) return 0;      // the reference dereferences idx[0] unconditionally
  std::vector<int> k(keys, keys + n), v(n), uk;
  for (int64_t i = 0; i < n; ++i) v[i] = static_cast<int>(i);
  std::vector<std::vector<int>> gv;
  mxref_group::GroupKVPairs(k, v, &uk, &gv, [](int, int, bool) { return true; }, false);
  int64_t p = 0;
  for (size_t g = 0; g < uk.size(); ++g) {
    uniq[g] = uk[g];
    counts[g] = static_cast<int>(gv[g].size());
    for (int pos : gv[g]) positions[p++] = pos;
  }
  return static_cast<int64_t>(uk.size());
}

// ElementwiseSumRsp (ndarray_function.cc:155-176) over `nsrc` row_sparse sources: returns the
// number of union rows; out_idx / out_val must hold sum(nnr) rows.
int64_t mxref_rsp_reduce(int nsrc, const int64_t* const* idx, const int64_t* nnr, const float* const* val,
                         int64_t row_len, int64_t* out_idx, float* out_val, int nthreads) {
  using namespace mxref_sparse;
  std::vector<NDArray> nds(nsrc);
  for (int i = 0; i < nsrc; ++i) {
    nds[i].val = const_cast<float*>(val[i]);
    nds[i].idx = const_cast<int64_t*>(idx[i]);
    nds[i].nnr = static_cast<size_t>(nnr[i]);
    nds[i].row_len = static_cast<size_t>(row_len);
  }
  std::vector<int64_t> uniq;
  GetUniqueRspRowIdx(nds, &uniq);
  NDArray out;
  out.val = out_val;
  out.idx = out_idx;
  out.nnr = uniq.size();                 // out->CheckAndAlloc({Shape1(uniq_row_idx.size())})
  out.row_len = static_cast<size_t>(row_len);
  std::memset(out_val, 0, uniq.size() * row_len * sizeof(float));   // out->data() = 0
  if (!uniq.empty()) ElementwiseSumRspImpl<float, int64_t>(nullptr, nds, uniq, &out, nthreads < 1 ? 1 : nthreads);
  return static_cast<int64_t>(uniq.size());
}

// SparseRetainOpForwardRspImpl (sparse_retain-inl.h:262-323) with the reference's kernels: output
// zero-filled, ids copied, rows copied where the source holds them. `src_dense_rows`: the source
// holds every row of the table (the :290 branch); `row_block`: use SparseRetainRspRowBlockKernel
// (needs ascending ids) instead of the per-id thread kernel.
void mxref_sparse_retain(const int64_t* src_idx, int64_t src_nnr, const float* src_val, int64_t row_len,
                         const int64_t* ids, int64_t nids, int src_dense_rows, int row_block,
                         int64_t* out_idx, float* out_val, int nthreads) {
  using namespace mxnet::op;
  if (nthreads < 1) nthreads = 1;   // Kernel<OP, cpu>::Launch: omp parallel for over the work items
  std::memset(out_val, 0, static_cast<size_t>(nids) * row_len * sizeof(float));   // Kernel<set_zero>
  if (src_dense_rows) {
#pragma omp parallel for num_threads(nthreads)
    for (int i = 0; i < nids; ++i) {
      SparseRetainCopyIndices::Map(i, out_idx, const_cast<int64_t*>(ids));
      SparseRetainCopyRetainedRowsFromDnsPerRow::Map(i, out_val, src_val, ids, static_cast<size_t>(row_len));
    }
  } else if (row_block) {
    const size_t seg_len = 7;   // any segmentation must give the same result
    const int nseg = static_cast<int>((nids + seg_len - 1) / seg_len);
#pragma omp parallel for num_threads(nthreads)
    for (int i = 0; i < nseg; ++i) {
      SparseRetainRspRowBlockKernel::Map(i, out_val, out_idx, src_val, src_idx, ids, static_cast<size_t>(nids),
                                         static_cast<size_t>(src_nnr), static_cast<size_t>(row_len), seg_len);
    }
  } else {
#pragma omp parallel for num_threads(nthreads)
    for (int i = 0; i < nids; ++i) {
      SparseRetainRspThreadKernel::Map(i, out_val, out_idx, src_val, src_idx, ids,
                                       static_cast<size_t>(src_nnr), static_cast<size_t>(row_len));
    }
  }
}

// UniqueImpl<cpu> (src/kvstore/kvstore_utils.cc:31-44): sort + std::unique in place; returns the count
int64_t mxref_unique(int64_t* ids, int64_t n) {
  mxref_sparse::common::ParallelSort(ids, ids + n, 1);
  return std::unique(ids, ids + n) - ids;
}

}  // extern "C"
