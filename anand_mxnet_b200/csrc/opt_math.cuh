// opt_math.cuh -- per-element optimizer arithmetic shared by the dense and row_sparse kernels.
//
// IEEE fp32, round-to-nearest, NO fused multiply-add: every operation is an explicit __f*_rn
// intrinsic (which the compiler never contracts), so results are bit-identical to the reference's
// CPU build and to oracle/kvoracle.c. Each branch cites the reference expression tree it follows.
#pragma once
#include "kernels.h"

namespace b200kv {

struct Hyper {
  float lr, wd, momentum, rescale, clip, beta1, beta2, eps;
};

// mshadow_op::clip (src/operator/mshadow_op.h:912-922)
__device__ __forceinline__ float clipf(float x, float b) { return x > b ? b : (x < -b ? -b : x); }

template <int OPT>
__device__ __forceinline__ float step(float w, float g, float& s1, float& s2, bool has_mom,
                                      const Hyper& h) {
  if (OPT == kOptAssign) {
    return g;
  } else if (OPT == kOptSGD) {
    // MultiSGDKernel / SGDMomKernel / MP_SGDMomKernel / SGDMomDnsRspDnsKernel
    // (optimizer_op-inl.h:232-254, 607-617, 706-724, 749-773)
    //   mom = momentum*mom - lr*wd*w - lr*rescale*g        | ... - lr*clip(rescale*g, c)
    //   w   = w + mom
    const float m0 = has_mom ? s1 : 0.f;
    const float a = __fsub_rn(__fmul_rn(h.momentum, m0), __fmul_rn(__fmul_rn(h.lr, h.wd), w));
    float m1;
    if (h.clip >= 0.f) {
      m1 = __fsub_rn(a, __fmul_rn(h.lr, clipf(__fmul_rn(h.rescale, g), h.clip)));
    } else {
      m1 = __fsub_rn(a, __fmul_rn(__fmul_rn(h.lr, h.rescale), g));
    }
    if (has_mom) s1 = m1;
    return __fadd_rn(w, m1);
  } else if (OPT == kOptSGDSingle) {
    // SGDKernel / MP_SGDKernel / SGDDnsRspKernel (optimizer_op-inl.h:388-397, 661-674, 453-475)
    //   w = (1-lr*wd)*w - (lr*rescale)*g                   | ... - lr*clip(rescale*g, c)
    const float a = __fmul_rn(__fsub_rn(1.f, __fmul_rn(h.lr, h.wd)), w);
    if (h.clip >= 0.f) {
      return __fsub_rn(a, __fmul_rn(h.lr, clipf(__fmul_rn(h.rescale, g), h.clip)));
    }
    return __fsub_rn(a, __fmul_rn(__fmul_rn(h.lr, h.rescale), g));
  } else if (OPT == kOptAdam) {
    // AdamUpdateKernel (optimizer_op-inl.h:1302-1312)
    float gr = __fadd_rn(__fmul_rn(g, h.rescale), __fmul_rn(w, h.wd));
    if (h.clip >= 0.f) gr = clipf(gr, h.clip);
    s1 = __fadd_rn(__fmul_rn(h.beta1, s1), __fmul_rn(__fsub_rn(1.f, h.beta1), gr));
    s2 = __fadd_rn(__fmul_rn(h.beta2, s2), __fmul_rn(__fmul_rn(__fsub_rn(1.f, h.beta2), gr), gr));
    return __fsub_rn(w, __fdiv_rn(__fmul_rn(h.lr, s1), __fadd_rn(__fsqrt_rn(s2), h.eps)));
  } else if (OPT == kOptTest) {
    // optimizer.py:2044-2046: weight[:] += grad * rescale_grad
    return __fadd_rn(w, __fmul_rn(g, h.rescale));
  }
  return w;
}

// AdamDnsRspDnsKernel<req,cpu> (optimizer_op-inl.h:1350-1380): the lazy CPU kernel squares the
// CLIPPED value first -- (1-beta2)*(c*c) -- unlike the dense kernel's ((1-beta2)*g)*g; without
// clipping both agree. The oracle is the CPU path, so the row_sparse kernel follows it.
__device__ __forceinline__ float step_adam_lazy(float w, float g, float& s1, float& s2,
                                                const Hyper& h) {
  const float gr = __fadd_rn(__fmul_rn(g, h.rescale), __fmul_rn(w, h.wd));
  if (h.clip >= 0.f) {
    const float c = clipf(gr, h.clip);
    s1 = __fadd_rn(__fmul_rn(h.beta1, s1), __fmul_rn(__fsub_rn(1.f, h.beta1), c));
    s2 = __fadd_rn(__fmul_rn(h.beta2, s2), __fmul_rn(__fsub_rn(1.f, h.beta2), __fmul_rn(c, c)));
  } else {
    s1 = __fadd_rn(__fmul_rn(h.beta1, s1), __fmul_rn(__fsub_rn(1.f, h.beta1), gr));
    s2 = __fadd_rn(__fmul_rn(h.beta2, s2), __fmul_rn(__fmul_rn(__fsub_rn(1.f, h.beta2), gr), gr));
  }
  return __fsub_rn(w, __fdiv_rn(__fmul_rn(h.lr, s1), __fadd_rn(__fsqrt_rn(s2), h.eps)));
}

}  // namespace b200kv
