// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product library.
//
// oracle/_ref/libmxref.so: a thin extern "C" harness around the *reference's own* CPU arithmetic
// for the KVStore hot path, compiled from the reference headers where they lie under
// /root/reference (see oracle/Makefile; no reference source is copied into this repository).
//
// It exists to (1) pin oracle/kvoracle.c (our plain-C restatement) bit-for-bit against the real
// reference code, (2) generate tests/golden/*.npz (oracle/gen_golden.py), and (3) optionally serve
// as bench.py's cpu_baseline of kind "reference" (the .so travels to the GPU box; the sources do
// not need to).
//
// Each entry point names the reference function it calls:
//   CommCPU::ReduceSumCPU<float>         src/kvstore/comm.h:357-392   (groups-of-4 association)
//   CommCPU::ReduceSumCPUImpl semantics  src/kvstore/comm.h:394-410   (4096-element OMP tasks)
//   SGDKernel / SGDMomKernel             src/operator/optimizer_op-inl.h:382-397, 601-618
//   MP_SGDKernel / MP_SGDMomKernel       src/operator/optimizer_op-inl.h:655-674, 700-725
//   MultiSGDKernel<float,MOM,MP>         src/operator/optimizer_op-inl.h:207-258
//   AdamUpdateKernel                     src/operator/optimizer_op-inl.h:1292-1314
//   SGDDnsRspKernel<req,cpu>             src/operator/optimizer_op-inl.h:453-475
//   SGDMomDnsRspDnsKernel<req,cpu>       src/operator/optimizer_op-inl.h:749-773
//   AdamDnsRspDnsKernel<req,cpu>         src/operator/optimizer_op-inl.h:1350-1380
//   quantize_2bit / dequantize_2bit      src/kvstore/gradient_compression-inl.h:40-132
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>
#include <omp.h>
#include <dmlc/logging.h>
#include <mshadow/tensor.h>

// comm.h keeps ReduceSumCPU private; the harness needs to call the reference code itself.
#define private public
#define protected public
#include "kvstore/comm.h"
#undef private
#undef protected
#include "operator/optimizer_op-inl.h"
#include "kvstore/gradient_compression-inl.h"

using mxnet::kWriteTo;
namespace op = mxnet::op;

// Kernel<OP,cpu>::Launch semantics (src/operator/mxnet_op.h:928-958): serial below 2 threads,
// otherwise a plain `omp parallel for` over the element index.
template <typename F>
static inline void launch_n(size_t n, int nthreads, F f) {
  if (nthreads < 2) {
    for (int64_t i = 0; i < static_cast<int64_t>(n); ++i) f(i);
  } else {
#pragma omp parallel for num_threads(nthreads)
    for (int64_t i = 0; i < static_cast<int64_t>(n); ++i) f(i);
  }
}
#define LAUNCH(N, NT, BODY) launch_n((N), (NT), [&](int64_t i) { BODY; })

extern "C" {

int mxref_abi_version() { return 1; }

// dmlc::stof (3rdparty/dmlc-core/include/dmlc/strtonum.h:467): how FieldEntry<float>::Set parses
// every scalar op parameter (lr, wd, rescale_grad, ...). Pins csrc/scalar_parse.cc.
float mxref_dmlc_stof(const char* s) { return dmlc::stof(std::string(s)); }

// CommCPU::ReduceSumCPU<float>: sums n buffers of `size` floats into ptrs[0], reference association.
void mxref_reduce_sum_cpu(float** ptrs, int n, size_t size) {
  std::vector<float*> d(ptrs, ptrs + n);
  mxnet::kvstore::CommCPU::ReduceSumCPU<float>(d, 0, static_cast<mxnet::index_t>(size));
}

// CommCPU::ReduceSumCPUImpl with explicit thread count / bigarray bound (comm.h:394-410). The
// member function needs a CommCPU object (NDArray/engine); the partitioning is restated here and
// each task calls the reference ReduceSumCPU<float> above.
void mxref_reduce_sum_cpu_impl(float** ptrs, int n, size_t total, int nthreads, size_t bigarray_bound) {
  std::vector<float*> d(ptrs, ptrs + n);
  const size_t step = std::min(bigarray_bound, static_cast<size_t>(4 << 10));
  long ntask = (total + step - 1) / step;  // NOLINT
  if (total < bigarray_bound || nthreads <= 1) {
    mxnet::kvstore::CommCPU::ReduceSumCPU<float>(d, 0, static_cast<mxnet::index_t>(total));
  } else {
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long j = 0; j < ntask; ++j) {  // NOLINT
      size_t k = static_cast<size_t>(j);
      size_t begin = std::min(k * step, total);
      size_t end = std::min((k + 1) * step, total);
      mxnet::kvstore::CommCPU::ReduceSumCPU<float>(d, begin, static_cast<mxnet::index_t>(end - begin));
    }
  }
}

void mxref_sgd_update(size_t n, float* out, const float* w, const float* g, float clip, float lr,
                      float wd, float rescale, int nthreads) {
  LAUNCH(n, nthreads, op::SGDKernel::Map(static_cast<int>(i), out, w, g, clip, lr, wd, rescale, kWriteTo));
}

void mxref_sgd_mom_update(size_t n, float* out, float* mom, const float* w, const float* g,
                          float clip, float momentum, float lr, float wd, float rescale,
                          int nthreads) {
  LAUNCH(n, nthreads, op::SGDMomKernel::Map(static_cast<int>(i), out, mom, w, g, clip, momentum, lr, wd,
                                            rescale, kWriteTo));
}

// fp16 weights/grads with fp32 master weights (mp_sgd_update / mp_sgd_mom_update).
void mxref_mp_sgd_update_f16(size_t n, uint16_t* out, const uint16_t* w, const uint16_t* g,
                             float* w32, float clip, float lr, float wd, float rescale,
                             int nthreads) {
  using mshadow::half::half_t;
  LAUNCH(n, nthreads,
         op::MP_SGDKernel::Map(static_cast<int>(i), reinterpret_cast<half_t*>(out),
                               reinterpret_cast<const half_t*>(w),
                               reinterpret_cast<const half_t*>(g), w32, clip, lr, wd, rescale,
                               kWriteTo));
}

void mxref_mp_sgd_mom_update_f16(size_t n, uint16_t* out, float* mom, const uint16_t* w,
                                 const uint16_t* g, float* w32, float clip, float momentum,
                                 float lr, float wd, float rescale, int nthreads) {
  using mshadow::half::half_t;
  LAUNCH(n, nthreads,
         op::MP_SGDMomKernel::Map(static_cast<int>(i), reinterpret_cast<half_t*>(out), mom,
                                  reinterpret_cast<const half_t*>(w),
                                  reinterpret_cast<const half_t*>(g), w32, clip, momentum, lr, wd,
                                  rescale, kWriteTo));
}

// multi_sgd_update / multi_sgd_mom_update over `count` fp32 tensors (count <= 60).
void mxref_multi_sgd_update(int count, const size_t* sizes, float** weights, float** grads,
                            float** moms /* may be null */, float** outs, const float* lrs,
                            const float* wds, float clip, float rescale, float momentum,
                            int nthreads) {
  op::MultiSGDKernelParam<float, float> p;
  p.count = count;
  p.max_size = 0;
  for (int k = 0; k < count; ++k) {
    p.sizes[k] = sizes[k];
    p.max_size = std::max(p.max_size, sizes[k]);
    p.weights[k] = weights[k];
    p.grads[k] = grads[k];
    p.mom[k] = moms ? moms[k] : nullptr;
    p.weights32[k] = nullptr;
    p.out_data[k] = outs[k];
    p.lrs[k] = lrs[k];
    p.wds[k] = wds[k];
  }
  p.clip_gradient = clip;
  p.rescale_grad = rescale;
  p.momentum = momentum;
  if (moms) {
    LAUNCH(p.max_size, nthreads, (op::MultiSGDKernel<float, true, false>::Map(static_cast<int>(i), p, kWriteTo)));
  } else {
    LAUNCH(p.max_size, nthreads, (op::MultiSGDKernel<float, false, false>::Map(static_cast<int>(i), p, kWriteTo)));
  }
}

// multi_mp_sgd_update / multi_mp_sgd_mom_update, fp16 weights+grads, fp32 master + momentum.
void mxref_multi_mp_sgd_update_f16(int count, const size_t* sizes, uint16_t** weights,
                                   uint16_t** grads, float** moms /* may be null */,
                                   float** weights32, uint16_t** outs, const float* lrs,
                                   const float* wds, float clip, float rescale, float momentum,
                                   int nthreads) {
  using mshadow::half::half_t;
  op::MultiSGDKernelParam<half_t, float> p;
  p.count = count;
  p.max_size = 0;
  for (int k = 0; k < count; ++k) {
    p.sizes[k] = sizes[k];
    p.max_size = std::max(p.max_size, sizes[k]);
    p.weights[k] = reinterpret_cast<half_t*>(weights[k]);
    p.grads[k] = reinterpret_cast<half_t*>(grads[k]);
    p.mom[k] = moms ? moms[k] : nullptr;
    p.weights32[k] = weights32[k];
    p.out_data[k] = reinterpret_cast<half_t*>(outs[k]);
    p.lrs[k] = lrs[k];
    p.wds[k] = wds[k];
  }
  p.clip_gradient = clip;
  p.rescale_grad = rescale;
  p.momentum = momentum;
  if (moms) {
    LAUNCH(p.max_size, nthreads, (op::MultiSGDKernel<float, true, true>::Map(static_cast<int>(i), p, kWriteTo)));
  } else {
    LAUNCH(p.max_size, nthreads, (op::MultiSGDKernel<float, false, true>::Map(static_cast<int>(i), p, kWriteTo)));
  }
}

void mxref_adam_update(size_t n, float* out, float* mean, float* var, const float* w,
                       const float* g, float clip, float rescale, float beta1, float beta2,
                       float lr, float wd, float eps, int nthreads) {
  LAUNCH(n, nthreads,
         op::AdamUpdateKernel::Map(static_cast<int>(i), out, mean, var, w, g, clip, rescale, beta1, beta2, lr,
                                   wd, eps, kWriteTo));
}

// lazy row_sparse updates: one Map call per gradient row (the <req,cpu> specialisations).
void mxref_sgd_rsp_update(size_t nrows, int64_t row_len, float* out, const float* w,
                          const int64_t* gidx, const float* gval, float clip, float lr, float wd,
                          float rescale, int nthreads) {
  LAUNCH(nrows, nthreads,
         (op::SGDDnsRspKernel<kWriteTo, mshadow::cpu>::Map(static_cast<int>(i), row_len, out, w, gidx, gval,
                                                           clip, lr, wd, rescale)));
}

void mxref_sgd_mom_rsp_update(size_t nrows, int64_t row_len, float* out, float* mom,
                              const float* w, const int64_t* gidx, const float* gval, float clip,
                              float momentum, float lr, float wd, float rescale, int nthreads) {
  LAUNCH(nrows, nthreads,
         (op::SGDMomDnsRspDnsKernel<kWriteTo, mshadow::cpu>::Map(static_cast<int>(i), row_len, out, mom, w,
                                                                 gidx, gval, clip, momentum, lr,
                                                                 wd, rescale)));
}

void mxref_adam_rsp_update(size_t nrows, int64_t row_len, float* out, float* mean, float* var,
                           const float* w, const int64_t* gidx, const float* gval, float clip,
                           float beta1, float beta2, float lr, float wd, float eps, float rescale,
                           int nthreads) {
  LAUNCH(nrows, nthreads,
         (op::AdamDnsRspDnsKernel<kWriteTo, mshadow::cpu>::Map(static_cast<int>(i), row_len, out, mean, var, w,
                                                               gidx, gval, clip, beta1, beta2, lr,
                                                               wd, eps, rescale)));
}

// 2-bit gradient compression with residual (adjacent path).
void mxref_quantize_2bit(size_t n, float* compressed, const float* grad, float* residual,
                         float neg_threshold, float pos_threshold) {
  size_t nblocks = (n + 15) / 16;
  for (size_t b = 0; b < nblocks; ++b) {
    mxnet::kvstore::quantize_2bit::Map(static_cast<int>(b), static_cast<int>(n), compressed, const_cast<float*>(grad),
                                       residual, neg_threshold, pos_threshold);
  }
}

void mxref_dequantize_2bit(size_t n, float* out, const float* compressed, float neg_threshold,
                           float pos_threshold) {
  for (size_t i = 0; i < n; ++i) {
    mxnet::kvstore::dequantize_2bit::Map(static_cast<int>(i), out, const_cast<float*>(compressed), neg_threshold,
                                         pos_threshold);
  }
}

}  // extern "C"
