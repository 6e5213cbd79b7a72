/*
 * TEST INFRASTRUCTURE ONLY.
 *
 * kvoracle: a plain-C CPU restatement of the reference's (Apache MXNet 1.6, anandj91/anand-mxnet)
 * KVStore hot-path arithmetic -- the `kvstore('local')` reducer, the SGD / Adam optimizer kernels
 * and the row_sparse reduce / unique / retain steps. It is the *checker* for the CUDA product under
 * anand_mxnet_b200/csrc and bench.py's `cpu_baseline` ("port") leg. Nothing in the product path may
 * include, link, load or call it; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function here bit-for-bit against
 * (a) oracle/_ref/libmxref.so -- the reference's own headers compiled in place (oracle/Makefile) --
 * when that library is present, and (b) the committed fixtures tests/golden/ (npz files) that were
 * generated from it by oracle/gen_golden.py, plus the known-answer values SURVEY.md 8(c) records.
 *
 * All arithmetic is IEEE binary32, round-to-nearest-even, NO fused multiply-add (the reference CPU
 * build is `-O3 -msse3`, Makefile:109 / mshadow.mk:11,29; compile this file with -ffp-contract=off).
 * Each function cites the reference file:line whose expression tree it follows.
 */
#ifndef KVORACLE_H_
#define KVORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int kvo_abi_version(void);

/* ---- dense reduce -------------------------------------------------------------------------- */
/* 'local' association, src/kvstore/comm.h:357-392 (CommCPU::ReduceSumCPU):
 *   acc = g0; for i = 1,5,9,...: acc = acc + (((g[i]+g[i+1])+g[i+2])+g[i+3])   (tail: 1..3 terms) */
void kvo_reduce_local(const float* const* src, int n, size_t size, float* out, int nthreads);
/* 'device' association, src/ndarray/ndarray_function-inl.h:387-434 (ElementwiseSum<DEVICE>):
 *   N<=4: ((g0+g1)+g2)+g3 ; N>4: acc=g0; acc+=g1; ... -- both are the plain left fold. */
void kvo_reduce_device(const float* const* src, int n, size_t size, float* out, int nthreads);

/* ---- dense optimizers (fp32) --------------------------------------------------------------- */
/* clip < 0 means "no clipping" (mshadow_op.h:912-922 applied only if clip_gradient >= 0). */
/* sgd_update, optimizer_op-inl.h:388-397:  w' = (1-lr*wd)*w - (lr*rescale)*g | - lr*clip(rescale*g) */
void kvo_sgd_update(size_t n, float* out, const float* w, const float* g, float clip, float lr,
                    float wd, float rescale, int nthreads);
/* sgd_mom_update, optimizer_op-inl.h:607-617 */
void kvo_sgd_mom_update(size_t n, float* out, float* mom, const float* w, const float* g,
                        float clip, float momentum, float lr, float wd, float rescale, int nthreads);
/* multi_sgd_update / multi_sgd_mom_update for ONE tensor of the list, optimizer_op-inl.h:225-258.
 * mom == NULL: has_momentum=false (mom term is momentum*0). */
void kvo_multi_sgd_update(size_t n, float* out, float* mom, const float* w, const float* g,
                          float clip, float momentum, float lr, float wd, float rescale,
                          int nthreads);
/* adam_update, optimizer_op-inl.h:1302-1312; lr already carries the Python-side bias correction */
void kvo_adam_update(size_t n, float* out, float* mean, float* var, const float* w, const float* g,
                     float clip, float rescale, float beta1, float beta2, float lr, float wd,
                     float eps, int nthreads);
/* mx.optimizer.Test, python/mxnet/optimizer/optimizer.py:2031-2046: w += g*rescale */
void kvo_test_update(size_t n, float* out, const float* w, const float* g, float rescale,
                     int nthreads);

/* ---- mixed precision: 16-bit weights/grads + fp32 master (mp_sgd*, multi_mp_sgd*) ----------- */
/* kind: 0 = IEEE fp16 (reference dtype 2), 1 = bfloat16 (B200 extension; mshadow 1.6 has none) */
float kvo_half_to_float(uint16_t h, int kind);
uint16_t kvo_float_to_half(float f, int kind);
/* mp_sgd_update / mp_sgd_mom_update (single-tensor association), optimizer_op-inl.h:661-674,706-724 */
void kvo_mp_sgd_update(size_t n, uint16_t* out, float* w32, const uint16_t* g, int kind, float clip,
                       float lr, float wd, float rescale, int nthreads);
void kvo_mp_sgd_mom_update(size_t n, uint16_t* out, float* mom, float* w32, const uint16_t* g,
                           int kind, float clip, float momentum, float lr, float wd, float rescale,
                           int nthreads);
/* multi_mp_sgd_update / multi_mp_sgd_mom_update for one tensor, optimizer_op-inl.h:225-258 with
 * has_mixed_precision=true. mom may be NULL. */
void kvo_multi_mp_sgd_update(size_t n, uint16_t* out, float* mom, float* w32, const uint16_t* g,
                             int kind, float clip, float momentum, float lr, float wd,
                             float rescale, int nthreads);

/* ---- row_sparse ---------------------------------------------------------------------------- */
/* sort + unique of int64 ids in place; returns the unique count (kvstore_utils.cc:32-44) */
size_t kvo_unique_i64(int64_t* ids, size_t n);
/* row_sparse sum, src/ndarray/ndarray_function.cc:59-175: out_idx = ascending unique union of the
 * sources' row ids; out rows start at 0.0f and sources are accumulated in list order.
 * out_idx must hold sum(nrows) ids and out_val sum(nrows)*row_len floats; returns nnr. */
size_t kvo_rsp_reduce(int nsrc, const int64_t* const* idx, const size_t* nrows,
                      const float* const* val, size_t row_len, int64_t* out_idx, float* out_val);
/* sparse_retain on a row_sparse source, src/operator/tensor/sparse_retain-inl.h:121-150,262-323:
 * out_idx = ids verbatim; out rows = source row if present else 0. src_dense_rows != 0 means the
 * source holds every row of the table (idx used as row position, :290-313). */
void kvo_sparse_retain(const int64_t* src_idx, size_t src_nnr, const float* src_val,
                       size_t row_len, const int64_t* ids, size_t nids, int src_dense_rows,
                       int64_t* out_idx, float* out_val);
/* lazy updates over the rows listed in the gradient (cpu specialisations):
 * optimizer_op-inl.h:453-475 (sgd), :749-773 (sgd mom), :1350-1380 (adam) */
void kvo_sgd_rsp_update(size_t nrows, size_t row_len, float* w, const int64_t* gidx,
                        const float* gval, float clip, float lr, float wd, float rescale);
void kvo_sgd_mom_rsp_update(size_t nrows, size_t row_len, float* w, float* mom, const int64_t* gidx,
                            const float* gval, float clip, float momentum, float lr, float wd,
                            float rescale);
void kvo_adam_rsp_update(size_t nrows, size_t row_len, float* w, float* mean, float* var,
                         const int64_t* gidx, const float* gval, float clip, float beta1,
                         float beta2, float lr, float wd, float eps, float rescale);

/* ---- 2-bit gradient compression (adjacent row f2), gradient_compression-inl.h:40-132 -------- */
void kvo_quantize_2bit(size_t n, uint32_t* compressed, const float* grad, float* residual,
                       float threshold);
void kvo_dequantize_2bit(size_t n, float* out, const uint32_t* compressed, float threshold);

/* ---- multi-tensor optimizer operators ---------------------------------------- */
/* see kvoracle.c for the reference file:line of each; kind: -1 fp32, 0 fp16, 1 bf16 */
float kvo_sum_sq(const void* x, size_t n, int kind);
void kvo_multi_lars(size_t n, float* out, const float* lrs, const float* w_sum_sq,
                    const float* g_sum_sq, const float* wds, float eta, float eps, float rescale);
void kvo_adamw_update(size_t n, float* out, float* mean, float* var, const float* w, float* g,
                      float rescale, float clip, float beta1, float beta2, float eta, float lr,
                      float wd, float eps);
void kvo_mp_adamw_update(size_t n, uint16_t* out, float* mean, float* var, float* w32,
                         const uint16_t* g, int kind, float rescale, float clip, float beta1,
                         float beta2, float eta, float lr, float wd, float eps);
void kvo_multi_adamw_update(size_t n, float* w, const float* g, uint16_t* w16, const uint16_t* g16,
                            int kind, float* mean, float* var, float rescale, float clip,
                            float beta1, float beta2, float eta, float lr, float wd, float eps);
void kvo_lamb_phase1(size_t n, float* out, float* mean, float* var, const float* w, const float* g,
                     const uint16_t* g16, int kind, float clip, float rescale, float beta1,
                     float beta1_t, float beta2, float beta2_t, float wd, float eps,
                     int bias_correction);
void kvo_lamb_phase2(size_t n, float* out, uint16_t* out16, int kind, const float* w, const float* g,
                     float r1, float r2, float lr, float lower_bound, float upper_bound);
void kvo_multi_lamb_step1(size_t n, float* temp_g, float* mean, float* var, const float* w,
                          const float* g, const uint16_t* g16, int kind, float clip, float rescale,
                          float beta1, float beta2, float eps, float wd, int step_count,
                          int bias_correction);
void kvo_multi_lamb_step2(size_t n, float* w, uint16_t* w16, int kind, const float* temp_g,
                          float sum_sq_w, float sum_sq_g, float lr, float lower_bound,
                          float upper_bound);

/* ---- standard (non-lazy) updates with a row_sparse gradient over a dense weight (8f-f3) ------ */
void kvo_sgd_std_rsp_update(size_t num_rows, size_t row_len, float* w, const int64_t* gidx,
                            const float* gval, size_t nnr, float clip, float lr, float wd,
                            float rescale);
void kvo_sgd_mom_std_rsp_update(size_t num_rows, size_t row_len, float* w, float* mom,
                                const int64_t* gidx, const float* gval, size_t nnr, float clip,
                                float momentum, float lr, float wd, float rescale);
void kvo_adam_std_rsp_update(size_t num_rows, size_t row_len, float* w, float* mean, float* var,
                             const int64_t* gidx, const float* gval, size_t nnr, float clip,
                             float beta1, float beta2, float lr, float wd, float eps,
                             float rescale);

#ifdef __cplusplus
}
#endif
#endif /* KVORACLE_H_ */
