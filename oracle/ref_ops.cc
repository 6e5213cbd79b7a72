// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product library.
//
// Second translation unit of oracle/_ref/libmxref.so: the reference's own CPU FCompute functions
// for the multi-tensor optimizer operators of SURVEY 8(f)-f1, called through ONE generic entry
// point (op name + TBlobs + the string attribute dict, i.e. the same plumbing MXImperativeInvokeEx
// feeds them: dmlc::Parameter parsing included).
//
// The CPU kernels of these operators live in the reference's .cc files next to their
// NNVM_REGISTER_OP blocks. The files are compiled where they lie under /root/reference (nothing is
// copied); NNVM_REGISTER_OP is redefined to a do-nothing builder first, so no nnvm registry /
// libmxnet symbol is needed. What runs is the reference's code:
//   multi_sum_sq                        src/operator/contrib/multi_sum_sq.cc:64-90 (CalcSumSq)
//   multi_lars                          src/operator/contrib/multi_lars-inl.h:61-98
//   preloaded_multi_[mp_]sgd[_mom]_update  src/operator/contrib/preloaded_multi_sgd-inl.h:154-330
//   _adamw_update / _mp_adamw_update    src/operator/contrib/adamw-inl.h:108-215, adamw.cc:117-141
//   _multi_adamw_update / _multi_mp_..  src/operator/contrib/adamw-inl.h:322-497
//   lamb_update_phase1/2, mp_lamb_...   src/operator/optimizer_op-inl.h:1566-1930
//   standard (non-lazy) sparse updates   src/operator/optimizer_op.cc:108-139, 195-229 (the
//       <req,cpu> kernels; their Impl functions need NDArray / Resource objects, so the row-flag
//       prefix sum of optimizer_op.cc:173-182 is rebuilt here and the kernels are called per row)
//   _multi_lamb_update / _multi_mp_..   src/operator/contrib/multi_lamb.cc:33-170 (Step1/Step2
//       kernels; the temp-space orchestration of multi_lamb-inl.h:268-338 needs the engine's
//       Resource manager and is restated below with a plain workspace)
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
#include <dmlc/memory_io.h>
#include <mshadow/tensor.h>
#include <nnvm/op.h>
#include "operator/optimizer_op-inl.h"

namespace mxref {
struct FakeOp {
  template <typename... A> FakeOp& describe(A&&...) { return *this; }
  template <typename... A> FakeOp& set_num_inputs(A&&...) { return *this; }
  template <typename... A> FakeOp& set_num_outputs(A&&...) { return *this; }
  template <typename... A> FakeOp& set_attr_parser(A&&...) { return *this; }
  template <typename T, typename... A> FakeOp& set_attr(A&&...) { return *this; }
  template <typename... A> FakeOp& add_argument(A&&...) { return *this; }
  template <typename... A> FakeOp& add_arguments(A&&...) { return *this; }
  template <typename... A> FakeOp& add_alias(A&&...) { return *this; }
};
}  // namespace mxref
#undef NNVM_REGISTER_OP
#define NNVM_REGISTER_OP(OpName) static ::mxref::FakeOp __fake_op_##OpName = ::mxref::FakeOp()
#include "operator/optimizer_op.cc"   // SGDMomStdDnsRspDnsKernel<req,cpu>, AdamStdDnsRspDnsKernel<req,cpu>
#include "operator/contrib/multi_sum_sq.cc"
#include "operator/contrib/multi_lars.cc"
#include "operator/contrib/preloaded_multi_sgd.cc"
#include "operator/contrib/adamw.cc"
#include "operator/contrib/multi_lamb.cc"

using namespace mxnet;

// Glue the reference keeps in translation units this harness does not compile:
//  * the OMP thread-count policy object consulted by Kernel<OP,cpu>::Launch (src/engine/openmp.cc):
//    here simply a harness-controlled number (mxref_set_omp_threads), default 1;
static int g_omp_threads = 1;
namespace mxnet {
namespace engine {
OpenMP::OpenMP() : omp_num_threads_set_in_environment_(false) {}
OpenMP* OpenMP::Get() {
  static OpenMP inst;
  return &inst;
}
int OpenMP::GetRecommendedOMPThreadCount(bool) const { return g_omp_threads; }
}  // namespace engine
// optimizer_op.cc's FComputeEx functions (never called here) reference a few libmxnet symbols;
// the library must still load, so they get aborting definitions.
namespace op {
namespace mxnet_op {
template <>
bool tuned_op<set_to_int<0>, int64_t>::UseOMP(size_t, size_t) { return false; }
}  // namespace mxnet_op
}  // namespace op
Storage* Storage::Get() { LOG(FATAL) << "mxref: Storage is not part of the harness"; return nullptr; }
void NDArray::SetTBlob() const { LOG(FATAL) << "mxref: NDArray is not part of the harness"; }
void* Resource::get_space_internal(size_t) const {
  LOG(FATAL) << "mxref: Resource is not part of the harness";
  return nullptr;
}
}  // namespace mxnet

namespace {

typedef void (*Parser)(nnvm::NodeAttrs*);
typedef void (*Compute)(const nnvm::NodeAttrs&, const OpContext&, const std::vector<TBlob>&,
                        const std::vector<OpReqType>&, const std::vector<TBlob>&);

// MultiLAMB<cpu, ...> (multi_lamb-inl.h:268-338) with the workspace taken from a std::vector
// instead of ctx.requested[kTempSpace]; every arithmetic step calls the reference functions.
template <template <typename> class MPTypeChooser, int input_stride>
void MultiLambCPU(const nnvm::NodeAttrs& attrs, const OpContext& ctx, const std::vector<TBlob>& inputs,
                  const std::vector<OpReqType>& req, const std::vector<TBlob>& outputs) {
  using namespace mxnet::op;
  auto param = nnvm::get<MultiLAMBParam>(attrs.parsed);
  mshadow::Stream<cpu>* s = ctx.get_stream<cpu>();
  MSHADOW_REAL_TYPE_SWITCH(inputs[0].type_flag_, DType, {
    using MPDType = typename MPTypeChooser<DType>::type;
    MultiLAMBKernelParam<DType, MPDType> kernel_params;
    FillMultiLAMBKernelParam<cpu, DType, MPDType, MultiLAMBParam, input_stride>(
        attrs, ctx, inputs, outputs, &kernel_params);
    std::vector<TBlob> weights;
    for (size_t index = 0; index < kernel_params.ntensors; ++index) {
      weights.emplace_back(inputs[index * input_stride]);
    }
    std::vector<float> temp_g(kernel_params.total_size), r1(kernel_params.ntensors),
        r2(kernel_params.ntensors);
    std::vector<int> block_to_tensor(kernel_params.nchunks + 1), block_to_chunk(kernel_params.nchunks + 1);
    std::vector<TBlob> temp_g_tblobs;
    size_t pos = 0;
    for (size_t index = 0; index < kernel_params.ntensors; ++index) {
      mshadow::Tensor<cpu, 1, float> aux(temp_g.data() + pos, mshadow::Shape1(kernel_params.sizes[index]), s);
      temp_g_tblobs.emplace_back(TBlob(aux));
      pos += kernel_params.sizes[index];
    }
    MultiSumSqRun<cpu>(weights, kernel_params.ntensors, r1.data(), ctx);
    CallKernel1<MPDType, DType>(s, kernel_params, param, temp_g.data(), block_to_tensor.data(),
                                block_to_chunk.data());
    MultiSumSqRun<cpu>(temp_g_tblobs, kernel_params.ntensors, r2.data(), ctx);
    CallKernel2<MPDType, DType>(s, kernel_params, param, r1.data(), r2.data(), temp_g.data(),
                                block_to_tensor.data(), block_to_chunk.data(), req[0]);
  });
}

struct Entry { Parser parse; Compute fn; };

const std::map<std::string, Entry>& Table() {
  using namespace mxnet::op;
  static std::map<std::string, Entry> t = {
      // the dense optimizer operators of the KVStore path, as optimizer_op.cc:322-703 registers them
      {"sgd_update", {ParamParser<SGDParam>, SGDUpdate<cpu>}},
      {"sgd_mom_update", {ParamParser<SGDMomParam>, SGDMomUpdate<cpu>}},
      {"mp_sgd_update", {ParamParser<SGDParam>, MP_SGDUpdate<cpu>}},
      {"mp_sgd_mom_update", {ParamParser<SGDMomParam>, MP_SGDMomUpdate<cpu>}},
      {"multi_sgd_update", {ParamParser<MultiSGDParam>, MultiSGDUpdate<cpu, type_identity, 2>}},
      {"multi_sgd_mom_update", {ParamParser<MultiSGDMomParam>, MultiSGDMomUpdate<cpu, type_identity, 3>}},
      {"multi_mp_sgd_update", {ParamParser<MultiSGDParam>, MultiSGDUpdate<cpu, single_precision, 3>}},
      {"multi_mp_sgd_mom_update", {ParamParser<MultiSGDMomParam>, MultiSGDMomUpdate<cpu, single_precision, 4>}},
      {"adam_update", {ParamParser<AdamParam>, AdamUpdate<cpu>}},
      {"multi_sum_sq", {ParamParser<MultiSumSqParam>, MultiSumSq<cpu>}},
      {"multi_lars", {ParamParser<LARSParam>, MultiLARS<cpu>}},
      {"preloaded_multi_sgd_update",
       {ParamParser<PreloadedMultiSGDParam>, PreloadedMultiSGDUpdate<cpu, preloaded_type_identity, 2>}},
      {"preloaded_multi_sgd_mom_update",
       {ParamParser<PreloadedMultiSGDMomParam>, PreloadedMultiSGDMomUpdate<cpu, preloaded_type_identity, 3>}},
      {"preloaded_multi_mp_sgd_update",
       {ParamParser<PreloadedMultiSGDParam>, PreloadedMultiSGDUpdate<cpu, preloaded_single_precision, 3>}},
      {"preloaded_multi_mp_sgd_mom_update",
       {ParamParser<PreloadedMultiSGDMomParam>, PreloadedMultiSGDMomUpdate<cpu, preloaded_single_precision, 4>}},
      {"_adamw_update", {ParamParser<AdamWParam>, MPUpdate<cpu, AdamWUpdate<cpu>>}},
      {"_mp_adamw_update", {ParamParser<AdamWParam>, MPUpdate<cpu, MPAdamWUpdate<cpu>>}},
      {"_multi_adamw_update", {ParamParser<MultiAdamWParam>, multiMPUpdate<cpu, false>}},
      {"_multi_mp_adamw_update", {ParamParser<MultiAdamWParam>, multiMPUpdate<cpu, true>}},
      {"lamb_update_phase1", {ParamParser<LambUpdatePhaseOneParam>, LambUpdatePhaseOne<cpu>}},
      {"lamb_update_phase2", {ParamParser<LambUpdatePhaseTwoParam>, LambUpdatePhaseTwo<cpu>}},
      {"mp_lamb_update_phase1", {ParamParser<LambUpdatePhaseOneParam>, MPLambUpdatePhaseOne<cpu>}},
      {"mp_lamb_update_phase2", {ParamParser<LambUpdatePhaseTwoParam>, MPLambUpdatePhaseTwo<cpu>}},
      {"_multi_lamb_update", {ParamParser<MultiLAMBParam>, MultiLambCPU<LAMBTypeIdentity, 4>}},
      {"_multi_mp_lamb_update", {ParamParser<MultiLAMBParam>, MultiLambCPU<LAMBSinglePrecision, 5>}},
  };
  return t;
}

}  // namespace

extern "C" {

// Runs the reference FCompute<cpu> of `op`. Arrays are flat: dtype flags are mshadow's
// (0 f32, 1 f64, 2 f16, 3 u8, 4 i32, 5 i8, 6 i64); sizes in elements. Returns 0, or -1 with the
// dmlc::Error text in `err`.
int mxref_op_invoke(const char* op, int nin, void** in_ptr, const int* in_dtype, const int64_t* in_size,
                    int nout, void** out_ptr, const int* out_dtype, const int64_t* out_size,
                    int nparam, const char** keys, const char** vals, char* err, int errlen) {
  try {
    auto it = Table().find(op);
    if (it == Table().end()) throw dmlc::Error(std::string("unknown operator ") + op);
    nnvm::NodeAttrs attrs;
    attrs.name = op;
    for (int i = 0; i < nparam; ++i) attrs.dict[keys[i]] = vals[i];
    it->second.parse(&attrs);
    OpContext ctx;
    ctx.is_train = false;
    ctx.need_grad = false;
    ctx.run_ctx.stream = nullptr;
    std::vector<TBlob> in, out;
    for (int i = 0; i < nin; ++i) {
      in.emplace_back(in_ptr[i], mxnet::TShape({static_cast<dim_t>(in_size[i])}), 1 /*cpu::kDevMask*/,
                      in_dtype[i]);
    }
    for (int i = 0; i < nout; ++i) {
      out.emplace_back(out_ptr[i], mxnet::TShape({static_cast<dim_t>(out_size[i])}), 1, out_dtype[i]);
    }
    std::vector<OpReqType> req(nout, kWriteTo);
    it->second.fn(attrs, ctx, in, req, out);
    return 0;
  } catch (const std::exception& e) {
    if (err && errlen > 0) {
      std::strncpy(err, e.what(), errlen - 1);
      err[errlen - 1] = '\0';
    }
    return -1;
  }
}

// standard (non-lazy) updates with a row_sparse gradient over a dense weight of `num_rows` rows:
// prefix_sum as optimizer_op.cc:173-182 (row flags, inclusive scan), then one Map per weight row.
static std::vector<nnvm::dim_t> RowPrefixSum(int64_t num_rows, const int64_t* gidx, int64_t nnr) {
  std::vector<nnvm::dim_t> ps(num_rows, 0);
  for (int64_t i = 0; i < nnr; ++i) ps[gidx[i]] = 1;
  for (int64_t i = 1; i < num_rows; ++i) ps[i] += ps[i - 1];
  return ps;
}

void mxref_sgd_mom_std_rsp_update(int64_t num_rows, int64_t row_len, float* w, float* mom,
                                  const int64_t* gidx, const float* gval, int64_t nnr, float clip,
                                  float momentum, float lr, float wd, float rescale) {
  std::vector<nnvm::dim_t> ps = RowPrefixSum(num_rows, gidx, nnr);
  for (int64_t i = 0; i < num_rows; ++i) {
    op::SGDMomStdDnsRspDnsKernel<kWriteInplace, cpu>::Map(static_cast<int>(i), row_len, w, mom, w, gidx,
                                                        gval, ps.data(), clip, momentum, lr, wd, rescale);
  }
}

void mxref_adam_std_rsp_update(int64_t num_rows, int64_t row_len, float* w, float* mean, float* var,
                               const int64_t* gidx, const float* gval, int64_t nnr, float clip,
                               float beta1, float beta2, float lr, float wd, float eps, float rescale) {
  std::vector<nnvm::dim_t> ps = RowPrefixSum(num_rows, gidx, nnr);
  for (int64_t i = 0; i < num_rows; ++i) {
    op::AdamStdDnsRspDnsKernel<kWriteInplace, cpu>::Map(static_cast<int>(i), row_len, w, mean, var, w, gidx,
                                                      gval, ps.data(), clip, beta1, beta2, lr, wd, eps,
                                                      rescale);
  }
}

// non-lazy sgd_update (optimizer_op-inl.h:505-528): whole-weight scale by DType(1 - lr*wd) through
// the reference's op_with_req<mul> kernel, then SGDDnsRspKernel<req,cpu> with wd = 0 on the rows
void mxref_sgd_std_rsp_update(int64_t num_rows, int64_t row_len, float* w, const int64_t* gidx,
                              const float* gval, int64_t nnr, float clip, float lr, float wd,
                              float rescale) {
  const float scale = static_cast<float>(1 - lr * wd);
  for (int64_t i = 0; i < num_rows * row_len; ++i) {
    op::mxnet_op::op_with_req<op::mshadow_op::mul, kWriteInplace>::Map(static_cast<int>(i), w, w, scale);
  }
  for (int64_t r = 0; r < nnr; ++r) {
    op::SGDDnsRspKernel<kWriteInplace, cpu>::Map(static_cast<int>(r), row_len, w, w, gidx, gval, clip, lr,
                                                 0.f, rescale);
  }
}

// NDArray::Save (src/ndarray/ndarray.cc:1596-1670) for a CPU-resident value, field by field in the
// reference's order; the sub-structures are written by the reference's own TShape::Save
// (include/mxnet/tuple.h:704-713) and Context::Save (include/mxnet/base.h:157-160). NDArray itself
// needs libmxnet (storage, engine), hence the restated composition. nnr < 0: dense.
// Returns the number of bytes written (or needed, when `cap` is too small).
size_t mxref_ndarray_save(int ndim, const int64_t* shape, int dev_type, int dev_id, int type_flag,
                          const void* data, size_t data_bytes, int64_t nnr, const int64_t* idx,
                          char* out, size_t cap) {
  std::string buf;
  dmlc::MemoryStringStream strm(&buf);
  const uint32_t magic = 0xF993fac9;  // NDARRAY_V2_MAGIC (ndarray.cc:1590)
  strm.Write(&magic, sizeof(magic));
  int32_t stype = nnr < 0 ? kDefaultStorage : kRowSparseStorage;
  strm.Write(&stype, sizeof(stype));
  mxnet::TShape tshape(shape, shape + ndim);
  if (nnr >= 0) {
    mxnet::TShape sshape = tshape;
    sshape[0] = nnr;
    sshape.Save(&strm);
  }
  tshape.Save(&strm);
  Context ctx = Context::Create(static_cast<Context::DeviceType>(dev_type), dev_id);
  ctx.Save(&strm);
  int32_t tf = type_flag;
  strm.Write(&tf, sizeof(tf));
  if (nnr >= 0) {
    int32_t aux = mshadow::kInt64;
    strm.Write(&aux, sizeof(aux));
    mxnet::TShape ashape(1, nnr);
    ashape.Save(&strm);
  }
  strm.Write(data, data_bytes);
  if (nnr > 0) strm.Write(idx, static_cast<size_t>(nnr) * sizeof(int64_t));
  if (buf.size() <= cap) std::memcpy(out, buf.data(), buf.size());
  return buf.size();
}

void mxref_set_omp_threads(int n) { g_omp_threads = n < 1 ? 1 : n; }

int mxref_op_known(const char* op) { return Table().count(op) ? 1 : 0; }

}  // extern "C"
