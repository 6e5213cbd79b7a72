// multi_tensor_kernels.cu -- the multi-tensor optimizer operators next to the KVStore path
// (SURVEY 8f-f1): multi_sum_sq, multi_lars, _[mp_]adamw_update, _multi_[mp_]adamw_update,
// [mp_]lamb_update_phase1/2, _multi_[mp_]lamb_update.
//
// One launch covers every tensor of the call: a device table of MTTensor descriptors + a list of
// <= 8192-element chunks (one 256-thread CTA each, 16-byte accesses, streaming loads/stores), the
// per-tensor scalars (lr / wd / eta / bias-correction terms) travelling by value in the kernel
// parameters. All of it is HBM-bound elementwise work; the only reductions are sums of squares,
// done deterministically: per-thread sequential partial -> fixed shuffle tree -> one partial per
// chunk -> a second kernel folds a tensor's partials in a fixed tree (same result every run, which
// the reference's test asserts: tests/python/gpu/test_operator_gpu.py:279).
//
// Arithmetic is IEEE fp32 with explicit round-to-nearest intrinsics (no FMA contraction) in the
// reference's CPU expression trees -- each functor cites the file:line it follows -- so every
// elementwise result is bit-identical to the reference / oracle/kvoracle.c. Sums of squares differ
// from the reference CPU's sequential loop in association only (tolerance 1e-5 relative, the
// reference's own bound).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "kernels.h"

namespace b200kv {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float clipf(float x, float b) { return x > b ? b : (x < -b ? -b : x); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqr(float a) { return __fsqrt_rn(a); }

template <typename T> struct Cv;
template <> struct Cv<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Cv<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Cv<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 4 consecutive elements
template <typename T> struct L4 {
  static __device__ __forceinline__ void ld(const T* p, float (&x)[4]) {
    const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(p));
    const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = Cv<T>::to_f(h[i]);
  }
  static __device__ __forceinline__ void st(T* p, const float (&x)[4]) {
    uint2 raw;
    T* h = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = Cv<T>::from_f(x[i]);
    __stcs(reinterpret_cast<uint2*>(p), raw);
  }
};
template <> struct L4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&x)[4]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&x)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3]));
  }
};
template <typename T> struct L1 {
  static __device__ __forceinline__ void ld(const T* p, float (&x)[1]) { x[0] = Cv<T>::to_f(*p); }
  static __device__ __forceinline__ void st(T* p, const float (&x)[1]) { *p = Cv<T>::from_f(x[0]); }
};
template <typename T, int V> struct LS;
template <typename T> struct LS<T, 4> : L4<T> {};
template <typename T> struct LS<T, 1> : L1<T> {};

// deterministic block sum: fixed shuffle tree, then the warp leaders in index order
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float s_part[kThreads / 32];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
  __syncthreads();  // s_part may still be read by the previous call
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) r = __fadd_rn(r, s_part[i]);
  }
  return r;  // valid in thread 0
}

// ---------------------------------------------------------------------------------------------
// functors: V elements of tensor t starting at element `base`; s = launch scalars; k = tensor
// index inside the launch (per-tensor scalars s.t[k])

// multi_sum_sq (src/operator/contrib/multi_sum_sq.cc:64-78; GPU form multi_sum_sq.cu:92-118:
// widen to float, square, accumulate in float)
template <typename T> struct SumSqOp {
  static constexpr bool kReduce = true;
  template <int V>
  static __device__ __forceinline__ void run(const MTTensor& t, const MTScalars&, int, uint32_t base,
                                             float& acc0, float&) {
    float x[V];
    LS<T, V>::ld(static_cast<const T*>(t.p[0]) + base, x);
#pragma unroll
    for (int i = 0; i < V; ++i) acc0 = add(acc0, mul(x[i], x[i]));
  }
};

// _adamw_update (adamw-inl.h:176-208, mshadow expression form; the reference writes the rescaled /
// clipped gradient back into `grad`), _mp_adamw_update (adamw-inl.h:108-131) and the multi-tensor
// forms (adamw-inl.h:340-372: mean = beta1*(mean - g) + g). s.t[k] = (lr, wd, eta, -).
// s.f: 0 clip, 1 beta1, 2 beta2, 3 eps. s.d0 = rescale_grad (device scalar): a zero / inf / nan
// value skips the update (adamw-inl.h:449-461).
template <typename T, bool MP, bool MULTI> struct AdamWOp {
  static constexpr bool kReduce = false;
  template <int V>
  static __device__ __forceinline__ void run(const MTTensor& t, const MTScalars& s, int k,
                                             uint32_t base, float&, float&) {
    const float rescale = *s.d0;
    if (!isfinite(rescale) || rescale == 0.f) return;
    const float lr = s.t[k].x, wd = s.t[k].y, eta = s.t[k].z;
    const float clip = s.f[0], beta1 = s.f[1], beta2 = s.f[2], eps = s.f[3];
    float w[V], g[V], m[V], v[V];
    if (MP) LS<float, V>::ld(static_cast<const float*>(t.p[4]) + base, w);
    else LS<T, V>::ld(static_cast<const T*>(t.p[0]) + base, w);
    LS<T, V>::ld(static_cast<const T*>(t.p[1]) + base, g);
    LS<float, V>::ld(static_cast<const float*>(t.p[2]) + base, m);
    LS<float, V>::ld(static_cast<const float*>(t.p[3]) + base, v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float sg = mul(rescale, g[i]);
      if (clip >= 0.f) sg = clipf(sg, clip);
      g[i] = sg;
      if (MULTI) {
        m[i] = add(mul(beta1, sub(m[i], sg)), sg);
        const float adj = mul(sg, sg);
        v[i] = add(mul(beta2, sub(v[i], adj)), adj);
      } else {
        m[i] = add(mul(beta1, m[i]), mul(sub(1.f, beta1), sg));
        v[i] = add(mul(beta2, v[i]), mul(sub(1.f, beta2), mul(sg, sg)));
      }
      w[i] = sub(w[i], mul(eta, add(dvd(mul(lr, m[i]), add(sqr(v[i]), eps)), mul(wd, w[i]))));
    }
    LS<float, V>::st(static_cast<float*>(t.p[2]) + base, m);
    LS<float, V>::st(static_cast<float*>(t.p[3]) + base, v);
    if (MP) LS<float, V>::st(static_cast<float*>(t.p[4]) + base, w);
    LS<T, V>::st(static_cast<T*>(t.p[5]) + base, w);
    if (!MP && !MULTI) LS<T, V>::st(static_cast<T*>(t.p[1]) + base, g);  // the reference's side effect
  }
};

// lamb_update_phase1 / mp_lamb_update_phase1 (optimizer_op-inl.h:1621-1648, 1772-1801).
// s.f: 0 clip, 1 rescale, 2 beta1, 3 beta1_t, 4 beta2, 5 beta2_t, 6 wd, 7 eps, 8 bias_correction.
// With a 16-bit gradient the reference's `grad * rescale` is a half_t product (rounded to half).
template <typename T, bool MP> struct LambPhase1Op {
  static constexpr bool kReduce = false;
  template <int V>
  static __device__ __forceinline__ void run(const MTTensor& t, const MTScalars& s, int,
                                             uint32_t base, float&, float&) {
    const float clip = s.f[0], rescale = s.f[1], beta1 = s.f[2], beta1_t = s.f[3], beta2 = s.f[4],
                beta2_t = s.f[5], wd = s.f[6], eps = s.f[7];
    const bool bias = s.f[8] != 0.f;
    float w[V], g[V], m[V], v[V], out[V];
    if (MP) LS<float, V>::ld(static_cast<const float*>(t.p[4]) + base, w);
    else LS<T, V>::ld(static_cast<const T*>(t.p[0]) + base, w);
    LS<T, V>::ld(static_cast<const T*>(t.p[1]) + base, g);
    LS<float, V>::ld(static_cast<const float*>(t.p[2]) + base, m);
    LS<float, V>::ld(static_cast<const float*>(t.p[3]) + base, v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float gr = mul(g[i], rescale);
      if (sizeof(T) == 2) gr = Cv<T>::to_f(Cv<T>::from_f(gr));
      if (clip >= 0.f) gr = clipf(gr, clip);
      m[i] = add(mul(beta1, m[i]), mul(sub(1.f, beta1), gr));
      v[i] = add(mul(beta2, v[i]), mul(mul(sub(1.f, beta2), gr), gr));
      float r = add(dvd(m[i], add(sqr(v[i]), eps)), mul(wd, w[i]));
      if (bias) {
        // mean_hat divides in DOUBLE (`1. - beta1_t`), var_hat in float (`1 - beta2_t`)
        const float mean_hat = static_cast<float>(
            __ddiv_rn(static_cast<double>(m[i]), __dsub_rn(1.0, static_cast<double>(beta1_t))));
        const float var_hat = dvd(v[i], sub(1.f, beta2_t));
        r = add(dvd(mean_hat, add(sqr(var_hat), eps)), mul(wd, w[i]));
      }
      out[i] = r;
    }
    LS<float, V>::st(static_cast<float*>(t.p[2]) + base, m);
    LS<float, V>::st(static_cast<float*>(t.p[3]) + base, v);
    LS<float, V>::st(static_cast<float*>(t.p[5]) + base, out);
  }
};

// lamb_update_phase2 / mp_lamb_update_phase2 (optimizer_op-inl.h:1705-1729, 1862-1886).
// s.f: 0 lr, 1 lower_bound, 2 upper_bound; s.d0 / s.d1 = r1 / r2 (device scalars).
// The mp form writes ONLY the 16-bit output from weight32 - lr*g; weight32 stays as it was.
template <typename T, bool MP> struct LambPhase2Op {
  static constexpr bool kReduce = false;
  template <int V>
  static __device__ __forceinline__ void run(const MTTensor& t, const MTScalars& s, int,
                                             uint32_t base, float&, float&) {
    float lr = s.f[0];
    const float lb = s.f[1], ub = s.f[2];
    float r1 = *s.d0;
    const float r2 = *s.d1;
    if (lb >= 0) r1 = fmaxf(r1, lb);
    if (ub >= 0) r1 = fminf(r1, ub);
    if (!(r1 == 0.0f || r2 == 0.0f)) lr = dvd(mul(lr, r1), r2);
    float w[V], g[V];
    if (MP) LS<float, V>::ld(static_cast<const float*>(t.p[4]) + base, w);
    else LS<T, V>::ld(static_cast<const T*>(t.p[0]) + base, w);
    LS<float, V>::ld(static_cast<const float*>(t.p[1]) + base, g);
#pragma unroll
    for (int i = 0; i < V; ++i) w[i] = sub(w[i], mul(lr, g[i]));
    LS<T, V>::st(static_cast<T*>(t.p[5]) + base, w);
  }
};

// _multi_[mp_]lamb_update step 1 (multi_lamb.cc:33-77) fused with BOTH sums of squares the
// reference computes in separate passes (multi_lamb-inl.h:325-331): sum w^2 of the weight INPUT
// (the 16-bit weights in the mp form) and sum temp_g^2. s.t[k] = (lr, wd, 1-beta1^t, 1-beta2^t)
// (the powf terms are evaluated on the host with the C library, as the reference CPU does).
// s.f: 0 clip, 1 rescale, 2 beta1, 3 beta2, 4 eps, 5 bias_correction.
template <typename T, bool MP> struct MultiLambStep1Op {
  static constexpr bool kReduce = true;
  template <int V>
  static __device__ __forceinline__ void run(const MTTensor& t, const MTScalars& s, int k,
                                             uint32_t base, float& acc_w, float& acc_g) {
    const float wd = s.t[k].y, c1 = s.t[k].z, c2 = s.t[k].w;
    const float clip = s.f[0], rescale = s.f[1], beta1 = s.f[2], beta2 = s.f[3], eps = s.f[4];
    const bool bias = s.f[5] != 0.f;
    float wi[V], w[V], g[V], m[V], v[V], out[V];
    LS<T, V>::ld(static_cast<const T*>(t.p[0]) + base, wi);
    if (MP) {
      LS<float, V>::ld(static_cast<const float*>(t.p[4]) + base, w);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) w[i] = wi[i];
    }
    LS<T, V>::ld(static_cast<const T*>(t.p[1]) + base, g);
    LS<float, V>::ld(static_cast<const float*>(t.p[2]) + base, m);
    LS<float, V>::ld(static_cast<const float*>(t.p[3]) + base, v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float sg = mul(g[i], rescale);
      if (clip >= 0.f) sg = clipf(sg, clip);
      m[i] = add(mul(beta1, m[i]), mul(sub(1.f, beta1), sg));
      v[i] = add(mul(beta2, v[i]), mul(mul(sub(1.f, beta2), sg), sg));
      float r;
      if (bias) {
        r = add(dvd(dvd(m[i], c1), add(sqr(dvd(v[i], c2)), eps)), mul(wd, w[i]));
      } else {
        r = add(dvd(m[i], add(sqr(v[i]), eps)), mul(wd, w[i]));
      }
      out[i] = r;
      acc_w = add(acc_w, mul(wi[i], wi[i]));
      acc_g = add(acc_g, mul(r, r));
    }
    LS<float, V>::st(static_cast<float*>(t.p[2]) + base, m);
    LS<float, V>::st(static_cast<float*>(t.p[3]) + base, v);
    LS<float, V>::st(static_cast<float*>(t.p[5]) + base, out);
  }
};

// step 2 (multi_lamb.cc:79-118). s.t[k].x = lr; s.f: 0 lower_bound, 1 upper_bound;
// s.d0 / s.d1 = per-tensor sums of squares (weights / temp_g), indexed by t.aux.
template <typename T, bool MP> struct MultiLambStep2Op {
  static constexpr bool kReduce = false;
  template <int V>
  static __device__ __forceinline__ void run(const MTTensor& t, const MTScalars& s, int k,
                                             uint32_t base, float&, float&) {
    const float lb = s.f[0], ub = s.f[1];
    float r1 = sqr(s.d0[t.aux]);
    const float r2 = sqr(s.d1[t.aux]);
    if (lb >= 0) r1 = fmaxf(r1, lb);
    if (ub >= 0) r1 = fminf(r1, ub);
    const float r = (r1 == 0.0f || r2 == 0.0f) ? 1.0f : dvd(r1, r2);
    const float lr_adjusted = mul(s.t[k].x, r);
    float w[V], g[V];
    if (MP) LS<float, V>::ld(static_cast<const float*>(t.p[4]) + base, w);
    else LS<T, V>::ld(static_cast<const T*>(t.p[0]) + base, w);
    LS<float, V>::ld(static_cast<const float*>(t.p[5]) + base, g);
#pragma unroll
    for (int i = 0; i < V; ++i) w[i] = sub(w[i], mul(lr_adjusted, g[i]));
    if (MP) LS<float, V>::st(static_cast<float*>(t.p[4]) + base, w);
    LS<T, V>::st(static_cast<T*>(t.p[0]) + base, w);
  }
};

template <class Op>
__global__ void __launch_bounds__(kThreads) mt_kernel(const MTTensor* tensors, const MTChunk* chunks,
                                                      const MTScalars s, float* part0, float* part1) {
  const MTChunk c = chunks[blockIdx.x];
  const MTTensor t = tensors[c.tensor];
  float a0 = 0.f, a1 = 0.f;
  const uint32_t nvec = t.vec_ok ? c.len / 4 : 0;
  for (uint32_t v = threadIdx.x; v < nvec; v += kThreads) {
    Op::template run<4>(t, s, static_cast<int>(c.tensor), c.off + v * 4, a0, a1);
  }
  for (uint32_t e = nvec * 4 + threadIdx.x; e < c.len; e += kThreads) {
    Op::template run<1>(t, s, static_cast<int>(c.tensor), c.off + e, a0, a1);
  }
  if (Op::kReduce) {
    const float r0 = block_sum(a0);
    if (threadIdx.x == 0) part0[blockIdx.x] = r0;
    if (part1 != nullptr) {
      const float r1 = block_sum(a1);
      if (threadIdx.x == 0) part1[blockIdx.x] = r1;
    }
  }
}

// out[t.aux] = sum of the tensor's chunk partials, fixed order: one CTA per tensor
__global__ void __launch_bounds__(kThreads) mt_finalize_kernel(const MTTensor* tensors,
                                                               const float* part0, const float* part1,
                                                               float* out0, float* out1) {
  const MTTensor t = tensors[blockIdx.x];
  const uint32_t n = (t.size + kMTChunk - 1) / kMTChunk;
  float a0 = 0.f, a1 = 0.f;
  for (uint32_t i = threadIdx.x; i < n; i += kThreads) {
    a0 = __fadd_rn(a0, part0[t.first_chunk + i]);
    if (part1) a1 = __fadd_rn(a1, part1[t.first_chunk + i]);
  }
  const float r0 = block_sum(a0);
  if (threadIdx.x == 0) out0[t.aux] = r0;
  if (part1) {
    const float r1 = block_sum(a1);
    if (threadIdx.x == 0) out1[t.aux] = r1;
  }
}

// multi_lars (multi_lars-inl.h:61-74)
__global__ void lars_kernel(int n, float* out, const float* lrs, const float* wsq, const float* gsq,
                            const float* wds, float eta, float eps, float rescale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float w_norm = sqr(wsq[i]);
  const bool valid = w_norm > 0.f && gsq[i] > 0.f;
  out[i] = valid ? dvd(mul(mul(lrs[i], eta), w_norm),
                       add(add(mul(sqr(gsq[i]), rescale), mul(wds[i], w_norm)), eps))
                 : lrs[i];
}

template <class Op>
void launch(const MTLaunch& L, cudaStream_t st) {
  mt_kernel<Op><<<L.n_chunks, kThreads, 0, st>>>(L.tensors, L.chunks, L.s, L.part0, L.part1);
}

template <template <typename, bool> class Op>
void launch_tm(const MTLaunch& L, cudaStream_t st) {
  if (L.dtype == kFloat32) { KV_CHECK(!L.mp); return launch<Op<float, false>>(L, st); }
  KV_CHECK(L.mp) << "16-bit weights need the multi-precision (fp32 master copy) form of this operator";
  if (L.dtype == kFloat16) return launch<Op<__half, true>>(L, st);
  if (L.dtype == kBfloat16) return launch<Op<__nv_bfloat16, true>>(L, st);
  KV_FATAL << "unsupported dtype " << DTypeName(L.dtype);
}

template <bool MULTI>
void launch_adamw(const MTLaunch& L, cudaStream_t st) {
  if (L.dtype == kFloat32) { KV_CHECK(!L.mp); return launch<AdamWOp<float, false, MULTI>>(L, st); }
  KV_CHECK(L.mp) << "16-bit weights need the multi-precision (fp32 master copy) form of this operator";
  if (L.dtype == kFloat16) return launch<AdamWOp<__half, true, MULTI>>(L, st);
  if (L.dtype == kBfloat16) return launch<AdamWOp<__nv_bfloat16, true, MULTI>>(L, st);
  KV_FATAL << "unsupported dtype " << DTypeName(L.dtype);
}

}  // namespace

void LaunchMultiTensor(const MTLaunch& L, cudaStream_t st) {
  if (L.n_chunks <= 0) return;
  switch (L.op) {
    case kMTSumSq:
      if (L.dtype == kFloat32) launch<SumSqOp<float>>(L, st);
      else if (L.dtype == kFloat16) launch<SumSqOp<__half>>(L, st);
      else if (L.dtype == kBfloat16) launch<SumSqOp<__nv_bfloat16>>(L, st);
      else KV_FATAL << "multi_sum_sq: unsupported dtype " << DTypeName(L.dtype);
      break;
    case kMTAdamW: launch_adamw<false>(L, st); break;
    case kMTMultiAdamW: launch_adamw<true>(L, st); break;
    case kMTLambPhase1: launch_tm<LambPhase1Op>(L, st); break;
    case kMTLambPhase2: launch_tm<LambPhase2Op>(L, st); break;
    case kMTMultiLambStep1: launch_tm<MultiLambStep1Op>(L, st); break;
    case kMTMultiLambStep2: launch_tm<MultiLambStep2Op>(L, st); break;
    default: KV_FATAL << "unknown multi-tensor op " << L.op;
  }
  KV_CUDA(cudaGetLastError());
}

void LaunchMultiTensorFinalize(const MTTensor* tensors, int n_tensors, const float* part0,
                               const float* part1, float* out0, float* out1, cudaStream_t st) {
  if (n_tensors <= 0) return;
  mt_finalize_kernel<<<n_tensors, kThreads, 0, st>>>(tensors, part0, part1, out0, out1);
  KV_CUDA(cudaGetLastError());
}

void LaunchMultiLars(int n, float* out, const float* lrs, const float* wsq, const float* gsq,
                     const float* wds, float eta, float eps, float rescale, cudaStream_t st) {
  if (n <= 0) return;
  lars_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, out, lrs, wsq, gsq, wds, eta, eps, rescale);
  KV_CUDA(cudaGetLastError());
}

}  // namespace b200kv
