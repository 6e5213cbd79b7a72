/*
 * TEST INFRASTRUCTURE ONLY -- see kvoracle.h for scope, parity status and citation conventions.
 * Plain C11, IEEE binary32, no FMA contraction (build with -ffp-contract=off, no -mfma).
 */
#include "kvoracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

int kvo_abi_version(void) { return 1; }

#define PFOR _Pragma("omp parallel for schedule(static) num_threads(nt) if (nt > 1)")

static inline int nt_(int nthreads) { return nthreads < 1 ? 1 : nthreads; }

/* mshadow_op::clip, src/operator/mshadow_op.h:912-922 */
static inline float clipf(float x, float bound) {
  if (x > bound) return bound;
  if (x < -bound) return -bound;
  return x;
}

/* ------------------------------------------------------------------------------------------ */
/* dense reduce                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* src/kvstore/comm.h:357-392. `in_0 += in_1 + in_2 + in_3 + in_4` is an mshadow expression that
 * evaluates per element as in_0 = in_0 + (((in_1 + in_2) + in_3) + in_4). */
void kvo_reduce_local(const float* const* src, int n, size_t size, float* out, int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t e = 0; e < (int64_t)size; ++e) {
    float acc = src[0][e];
    for (int i = 1; i < n; i += 4) {
      float t = src[i][e];
      const int lim = (n - i) < 4 ? (n - i) : 4;
      for (int k = 1; k < lim; ++k) t = t + src[i + k][e];
      acc = acc + t;
    }
    out[e] = acc;
  }
}

/* src/ndarray/ndarray_function-inl.h:402-431: 2/3/4 inputs are one left-associated expression,
 * more inputs are `out = in0` then `out += in_i` -- the same left fold. */
void kvo_reduce_device(const float* const* src, int n, size_t size, float* out, int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t e = 0; e < (int64_t)size; ++e) {
    float acc = src[0][e];
    for (int i = 1; i < n; ++i) acc = acc + src[i][e];
    out[e] = acc;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* dense optimizers                                                                           */
/* ------------------------------------------------------------------------------------------ */

/* optimizer_op-inl.h:388-397 (SGDKernel) */
void kvo_sgd_update(size_t n, float* out, const float* w, const float* g, float clip, float lr,
                    float wd, float rescale, int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    if (clip >= 0.0f) {
      out[i] = (1.f - lr * wd) * w[i] - (lr)*clipf(rescale * g[i], clip);
    } else {
      out[i] = (1.f - lr * wd) * w[i] - (lr * rescale) * g[i];
    }
  }
}

/* optimizer_op-inl.h:607-617 (SGDMomKernel): a*b*c parses as (a*b)*c, a-b-c as (a-b)-c */
void kvo_sgd_mom_update(size_t n, float* out, float* mom, const float* w, const float* g,
                        float clip, float momentum, float lr, float wd, float rescale,
                        int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    if (clip >= 0.0f) {
      mom[i] = momentum * mom[i] - lr * wd * w[i] - lr * clipf(rescale * g[i], clip);
    } else {
      mom[i] = momentum * mom[i] - lr * wd * w[i] - lr * rescale * g[i];
    }
    out[i] = w[i] + mom[i];
  }
}

/* optimizer_op-inl.h:225-258 (MultiSGDKernel<float,has_momentum,false>), one tensor */
void kvo_multi_sgd_update(size_t n, float* out, float* mom, const float* w, const float* g,
                          float clip, float momentum, float lr, float wd, float rescale,
                          int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    float wv = w[i];
    float m = mom ? mom[i] : 0.0f;
    if (clip >= 0.0f) {
      m = momentum * m - lr * wd * wv - lr * clipf(rescale * g[i], clip);
    } else {
      m = momentum * m - lr * wd * wv - lr * rescale * g[i];
    }
    if (mom) mom[i] = m;
    wv = wv + m;
    out[i] = wv;
  }
}

/* optimizer_op-inl.h:1302-1312 (AdamUpdateKernel) */
void kvo_adam_update(size_t n, float* out, float* mean, float* var, const float* w, const float* g,
                     float clip, float rescale, float beta1, float beta2, float lr, float wd,
                     float eps, int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    float gr = g[i] * rescale + w[i] * wd;
    if (clip >= 0.f) gr = clipf(gr, clip);
    mean[i] = beta1 * mean[i] + (1.f - beta1) * gr;
    var[i] = beta2 * var[i] + (1.f - beta2) * gr * gr;
    out[i] = w[i] - lr * mean[i] / (sqrtf(var[i]) + eps);
  }
}

/* python/mxnet/optimizer/optimizer.py:2044-2046: weight[:] += grad * self.rescale_grad */
void kvo_test_update(size_t n, float* out, const float* w, const float* g, float rescale,
                     int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const float t = g[i] * rescale;
    out[i] = w[i] + t;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* 16-bit storage helpers (round-to-nearest-even, like F16C / cuda __float2half_rn / mshadow    */
/* half.h with MSHADOW_HALF_ROUND_TO_NEAREST=1, 3rdparty/mshadow/mshadow/half.h:16-20,197-260)  */
/* ------------------------------------------------------------------------------------------ */

static inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static float fp16_to_float(uint16_t h) {
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  if (exp == 0) {
    if (man == 0) return u2f(sign);
    /* subnormal: normalise */
    int e = -1;
    do {
      ++e;
      man <<= 1;
    } while ((man & 0x400u) == 0);
    man &= 0x3ffu;
    return u2f(sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13));
  }
  if (exp == 31) return u2f(sign | 0x7f800000u | (man << 13));
  return u2f(sign | ((exp + (127 - 15)) << 23) | (man << 13));
}

static uint16_t float_to_fp16(float f) {
  const uint32_t x = f2u(f);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) { /* inf / nan */
    return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? (0x200u | ((ax >> 13) & 0x3ffu)) : 0));
  }
  if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* rounds to >= 65520 -> inf */
  if (ax < 0x33000001u) return (uint16_t)sign;               /* <= 2^-25 -> 0 (ties to even) */
  int32_t e = (int32_t)(ax >> 23) - 127;
  uint32_t man = (ax & 0x7fffffu) | 0x800000u;
  uint32_t shift;
  uint32_t base;
  if (e < -14) { /* subnormal half */
    shift = (uint32_t)(13 + (-14 - e));
    base = 0;
  } else {
    shift = 13;
    base = (uint32_t)(e + 15) << 10;
    man &= 0x7fffffu;
  }
  uint32_t q = man >> shift;
  const uint32_t rem = man & ((1u << shift) - 1u);
  const uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) ++q;
  return (uint16_t)(sign | (base + q)); /* carry into the exponent is the correct rounding */
}

static float bf16_to_float(uint16_t h) { return u2f((uint32_t)h << 16); }

static uint16_t float_to_bf16(float f) {
  uint32_t x = f2u(f);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u); /* quiet nan */
  x += 0x7fffu + ((x >> 16) & 1u);
  return (uint16_t)(x >> 16);
}

float kvo_half_to_float(uint16_t h, int kind) { return kind ? bf16_to_float(h) : fp16_to_float(h); }
uint16_t kvo_float_to_half(float f, int kind) { return kind ? float_to_bf16(f) : float_to_fp16(f); }

/* optimizer_op-inl.h:661-674 (MP_SGDKernel) */
void kvo_mp_sgd_update(size_t n, uint16_t* out, float* w32, const uint16_t* g, int kind, float clip,
                       float lr, float wd, float rescale, int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    float w = w32[i];
    const float gf = kvo_half_to_float(g[i], kind);
    if (clip >= 0.0f) {
      w = (1.f - lr * wd) * w - (lr)*clipf(rescale * gf, clip);
    } else {
      w = (1.f - lr * wd) * w - (lr * rescale) * gf;
    }
    w32[i] = w;
    out[i] = kvo_float_to_half(w, kind);
  }
}

/* optimizer_op-inl.h:706-724 (MP_SGDMomKernel) */
void kvo_mp_sgd_mom_update(size_t n, uint16_t* out, float* mom, float* w32, const uint16_t* g,
                           int kind, float clip, float momentum, float lr, float wd, float rescale,
                           int nthreads) {
  kvo_multi_mp_sgd_update(n, out, mom, w32, g, kind, clip, momentum, lr, wd, rescale, nthreads);
}

/* optimizer_op-inl.h:225-258 with has_mixed_precision=true (identical tree to MP_SGDMomKernel) */
void kvo_multi_mp_sgd_update(size_t n, uint16_t* out, float* mom, float* w32, const uint16_t* g,
                             int kind, float clip, float momentum, float lr, float wd,
                             float rescale, int nthreads) {
  const int nt = nt_(nthreads);
  PFOR
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    float w = w32[i];
    float m = mom ? mom[i] : 0.0f;
    const float gf = kvo_half_to_float(g[i], kind);
    if (clip >= 0.0f) {
      m = momentum * m - lr * wd * w - lr * clipf(rescale * gf, clip);
    } else {
      m = momentum * m - lr * wd * w - lr * rescale * gf;
    }
    if (mom) mom[i] = m;
    w = w + m;
    w32[i] = w;
    out[i] = kvo_float_to_half(w, kind);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* row_sparse                                                                                 */
/* ------------------------------------------------------------------------------------------ */

static int cmp_i64(const void* a, const void* b) {
  const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return (x > y) - (x < y);
}

/* src/kvstore/kvstore_utils.cc:32-44: ParallelSort + std::unique */
size_t kvo_unique_i64(int64_t* ids, size_t n) {
  if (n == 0) return 0;
  qsort(ids, n, sizeof(int64_t), cmp_i64);
  size_t m = 1;
  for (size_t i = 1; i < n; ++i) {
    if (ids[i] != ids[m - 1]) ids[m++] = ids[i];
  }
  return m;
}

/* src/ndarray/ndarray_function.cc:59-175 (GetUniqueRspRowIdx + ElementwiseSumRspImpl). Source row
 * ids are ascending & unique per source (row_sparse invariant), so the reference's merge-walk
 * visits, for each output row, the sources in list order: out_row (0.0f) += src_row. */
size_t kvo_rsp_reduce(int nsrc, const int64_t* const* idx, const size_t* nrows,
                      const float* const* val, size_t row_len, int64_t* out_idx, float* out_val) {
  size_t total = 0;
  for (int s = 0; s < nsrc; ++s) {
    memcpy(out_idx + total, idx[s], nrows[s] * sizeof(int64_t));
    total += nrows[s];
  }
  const size_t nnr = kvo_unique_i64(out_idx, total);
  memset(out_val, 0, nnr * row_len * sizeof(float));
  for (int s = 0; s < nsrc; ++s) {
    size_t o = 0;
    for (size_t r = 0; r < nrows[s]; ++r) {
      const int64_t id = idx[s][r];
      while (o < nnr && out_idx[o] < id) ++o;
      if (o == nnr) break;
      if (out_idx[o] == id) {
        float* dst = out_val + o * row_len;
        const float* srcp = val[s] + r * row_len;
        for (size_t j = 0; j < row_len; ++j) dst[j] += srcp[j];
      }
    }
  }
  return nnr;
}

/* src/operator/tensor/sparse_retain-inl.h:121-150 (binary search per requested id) and :290-313
 * (source holds every row: idx is the row position). Output is zero-filled first (:283). */
void kvo_sparse_retain(const int64_t* src_idx, size_t src_nnr, const float* src_val,
                       size_t row_len, const int64_t* ids, size_t nids, int src_dense_rows,
                       int64_t* out_idx, float* out_val) {
  memset(out_val, 0, nids * row_len * sizeof(float));
  for (size_t i = 0; i < nids; ++i) {
    const int64_t irow = ids[i];
    out_idx[i] = irow;
    if (src_dense_rows) {
      memcpy(out_val + i * row_len, src_val + (size_t)irow * row_len, row_len * sizeof(float));
      continue;
    }
    int64_t j = -1, left = 0, right = (int64_t)src_nnr - 1;
    while (left <= right) {
      const int64_t m = left + (right - left) / 2;
      if (src_idx[m] == irow) {
        j = m;
        break;
      } else if (src_idx[m] < irow) {
        left = m + 1;
      } else {
        right = m - 1;
      }
    }
    if (j >= 0) memcpy(out_val + i * row_len, src_val + (size_t)j * row_len, row_len * sizeof(float));
  }
}

/* optimizer_op-inl.h:453-475 (SGDDnsRspKernel<req,cpu>) */
void kvo_sgd_rsp_update(size_t nrows, size_t row_len, float* w, const int64_t* gidx,
                        const float* gval, float clip, float lr, float wd, float rescale) {
  for (size_t i = 0; i < nrows; ++i) {
    for (size_t j = 0; j < row_len; ++j) {
      const size_t di = (size_t)gidx[i] * row_len + j, gi = i * row_len + j;
      if (clip >= 0.0f) {
        w[di] = (1.f - lr * wd) * w[di] - (lr)*clipf(rescale * gval[gi], clip);
      } else {
        w[di] = (1.f - lr * wd) * w[di] - (lr * rescale) * gval[gi];
      }
    }
  }
}

/* optimizer_op-inl.h:749-773 (SGDMomDnsRspDnsKernel<req,cpu>): rate = lr*wd hoisted */
void kvo_sgd_mom_rsp_update(size_t nrows, size_t row_len, float* w, float* mom, const int64_t* gidx,
                            const float* gval, float clip, float momentum, float lr, float wd,
                            float rescale) {
  const float rate = lr * wd;
  for (size_t i = 0; i < nrows; ++i) {
    for (size_t j = 0; j < row_len; ++j) {
      const size_t di = (size_t)gidx[i] * row_len + j, gi = i * row_len + j;
      if (clip >= 0.0f) {
        mom[di] = momentum * mom[di] - rate * w[di] - lr * clipf(rescale * gval[gi], clip);
      } else {
        mom[di] = momentum * mom[di] - rate * w[di] - lr * rescale * gval[gi];
      }
      w[di] = w[di] + mom[di];
    }
  }
}

/* optimizer_op-inl.h:1350-1380 (AdamDnsRspDnsKernel<req,cpu>). NOTE the clipped branch squares the
 * clipped value first -- (1-beta2) * (c*c) -- unlike the dense kernel's ((1-beta2)*g)*g. */
void kvo_adam_rsp_update(size_t nrows, size_t row_len, float* w, float* mean, float* var,
                         const int64_t* gidx, const float* gval, float clip, float beta1,
                         float beta2, float lr, float wd, float eps, float rescale) {
  for (size_t i = 0; i < nrows; ++i) {
    for (size_t j = 0; j < row_len; ++j) {
      const size_t di = (size_t)gidx[i] * row_len + j, gi = i * row_len + j;
      const float gr = gval[gi] * rescale + w[di] * wd;
      if (clip >= 0.0f) {
        const float c = clipf(gr, clip);
        mean[di] = beta1 * mean[di] + (1.f - beta1) * c;
        var[di] = beta2 * var[di] + (1.f - beta2) * (c * c);
      } else {
        mean[di] = beta1 * mean[di] + (1.f - beta1) * gr;
        var[di] = beta2 * var[di] + (1.f - beta2) * gr * gr;
      }
      w[di] = w[di] - lr * mean[di] / (sqrtf(var[di]) + eps);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* 2-bit compression, src/kvstore/gradient_compression-inl.h:40-132                             */
/* ------------------------------------------------------------------------------------------ */

void kvo_quantize_2bit(size_t n, uint32_t* compressed, const float* grad, float* residual,
                       float threshold) {
  static const uint8_t posbits[4] = {0xc0, 0x30, 0x0c, 0x03};
  static const uint8_t negbits[4] = {0x80, 0x20, 0x08, 0x02};
  const float pos = threshold, neg = -1 * threshold;
  const size_t nblocks = (n + 15) / 16;
  for (size_t b = 0; b < nblocks; ++b) {
    uint8_t bytes[4] = {0, 0, 0, 0};
    const size_t start = b << 4;
    const size_t end = (start + 16 <= n) ? start + 16 : n;
    for (size_t i = start; i < end; ++i) {
      uint8_t* cur = bytes + ((i - start) >> 2);
      residual[i] += grad[i];
      if (residual[i] >= pos) {
        *cur |= posbits[i & 3];
        residual[i] -= pos;
      } else if (residual[i] <= neg) {
        *cur |= negbits[i & 3];
        residual[i] -= neg;
      }
    }
    memcpy(&compressed[b], bytes, 4);
  }
}

void kvo_dequantize_2bit(size_t n, float* out, const uint32_t* compressed, float threshold) {
  static const uint8_t posbits[4] = {0xc0, 0x30, 0x0c, 0x03};
  static const uint8_t negbits[4] = {0x80, 0x20, 0x08, 0x02};
  const float pos = threshold, neg = -1 * threshold;
  for (size_t i = 0; i < n; ++i) {
    const uint8_t* ch = (const uint8_t*)(compressed + (i >> 4)) + ((i & 15) >> 2);
    const int col = (int)(i & 3);
    const uint8_t masked = (uint8_t)(*ch & posbits[col]);
    if (masked == posbits[col]) {
      out[i] = pos;
    } else if (masked == negbits[col]) {
      out[i] = neg;
    } else {
      out[i] = 0;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* multi-tensor optimizer operators: LARS, AdamW, LAMB                          */
/* ------------------------------------------------------------------------------------------ */

/* multi_sum_sq CPU, src/operator/contrib/multi_sum_sq.cc:64-78 (CalcSumSq): one float
 * accumulator per array, elements in index order. kind: -1 fp32, 0 fp16, 1 bf16. For fp16 the
 * reference multiplies in half_t (mshadow MSHADOW_HALF_OPERATOR: the product is rounded to half)
 * and adds the rounded product in float. */
float kvo_sum_sq(const void* x, size_t n, int kind) {
  float sum = 0.f;
  if (kind < 0) {
    const float* a = (const float*)x;
    for (size_t j = 0; j < n; ++j) sum += a[j] * a[j];
  } else {
    const uint16_t* a = (const uint16_t*)x;
    for (size_t j = 0; j < n; ++j) {
      const float v = kvo_half_to_float(a[j], kind);
      sum += kvo_half_to_float(kvo_float_to_half(v * v, kind), kind);
    }
  }
  return sum;
}

/* multi_lars, src/operator/contrib/multi_lars-inl.h:61-74 (MultiLARSKernel) */
void kvo_multi_lars(size_t n, float* out, const float* lrs, const float* w_sum_sq,
                    const float* g_sum_sq, const float* wds, float eta, float eps, float rescale) {
  for (size_t i = 0; i < n; ++i) {
    const float w_norm = sqrtf(w_sum_sq[i]);
    const int valid = w_norm > 0. && g_sum_sq[i] > 0.;
    out[i] = valid ? lrs[i] * eta * w_norm / (sqrtf(g_sum_sq[i]) * rescale + wds[i] * w_norm + eps)
                   : lrs[i];
  }
}

/* preloaded_multi_[mp_]sgd[_mom]_update: PreloadedMultiSGDKernel
 * (src/operator/contrib/preloaded_multi_sgd-inl.h:170-203) is MultiSGDKernel with lr / wd read
 * from arrays -- the same expression tree, so kvo_multi_sgd_update / kvo_multi_mp_sgd_update with
 * lr = lrs[k], wd = wds[k] restate it (tests pin that against the reference FCompute). */

/* _adamw_update fp32, src/operator/contrib/adamw-inl.h:176-208 (AdamWUpdate, mshadow expression
 * templates). NOTE the reference writes the rescaled (and clipped) gradient back into `grad`. */
void kvo_adamw_update(size_t n, float* out, float* mean, float* var, const float* w, float* g,
                      float rescale, float clip, float beta1, float beta2, float eta, float lr,
                      float wd, float eps) {
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (size_t i = 0; i < n; ++i) {
    float gr = rescale * g[i];
    if (clip >= 0.0f) gr = clipf(gr, clip);
    g[i] = gr;
    mean[i] = beta1 * mean[i] + omb1 * gr;
    var[i] = beta2 * var[i] + omb2 * (gr * gr);
    out[i] = w[i] - eta * (lr * mean[i] / (sqrtf(var[i]) + eps) + wd * w[i]);
  }
}

/* _mp_adamw_update, adamw-inl.h:108-131 (MPAdamWKernel): 16-bit weight/grad, fp32 master */
void kvo_mp_adamw_update(size_t n, uint16_t* out, float* mean, float* var, float* w32,
                         const uint16_t* g, int kind, float rescale, float clip, float beta1,
                         float beta2, float eta, float lr, float wd, float eps) {
  for (size_t i = 0; i < n; ++i) {
    float w = w32[i];
    float sg = rescale * kvo_half_to_float(g[i], kind);
    if (clip >= 0.0f) sg = clipf(sg, clip);
    const float m = mean[i] = beta1 * mean[i] + (1.0f - beta1) * sg;
    const float v = var[i] = beta2 * var[i] + (1.0f - beta2) * (sg * sg);
    w = w - eta * (lr * m / (sqrtf(v) + eps) + wd * w);
    w32[i] = w;
    out[i] = kvo_float_to_half(w, kind);
  }
}

/* _multi_adamw_update / _multi_mp_adamw_update for ONE tensor, adamw-inl.h:340-372
 * (MultiMPAdamWKernel): note the different association of the moment updates,
 * mean = beta1*(mean - g) + g. w16/g16 != NULL selects the mixed-precision form. */
void kvo_multi_adamw_update(size_t n, float* w, const float* g, uint16_t* w16, const uint16_t* g16,
                            int kind, float* mean, float* var, float rescale, float clip,
                            float beta1, float beta2, float eta, float lr, float wd, float eps) {
  for (size_t i = 0; i < n; ++i) {
    float wv = w[i]; /* fp32 weight, or the fp32 master copy */
    float sg = rescale * (g16 ? kvo_half_to_float(g16[i], kind) : g[i]);
    if (clip >= 0.0f) sg = clipf(sg, clip);
    const float m = beta1 * (mean[i] - sg) + sg;
    const float adj = sg * sg;
    const float v = beta2 * (var[i] - adj) + adj;
    mean[i] = m;
    var[i] = v;
    wv = wv - eta * (lr * m / (sqrtf(v) + eps) + wd * wv);
    w[i] = wv;
    if (w16) w16[i] = kvo_float_to_half(wv, kind);
  }
}

/* lamb_update_phase1 / mp_lamb_update_phase1, optimizer_op-inl.h:1621-1648, 1772-1801.
 * beta1_t / beta2_t = float(pow(double(beta), double(t))) (:1660-1661). The bias-corrected mean
 * divides in DOUBLE (`1. - beta1_t`), the variance in float (`1 - beta2_t`).
 * g16 != NULL: gradient is 16-bit and w is the fp32 master copy. */
void kvo_lamb_phase1(size_t n, float* out, float* mean, float* var, const float* w, const float* g,
                     const uint16_t* g16, int kind, float clip, float rescale, float beta1,
                     float beta1_t, float beta2, float beta2_t, float wd, float eps,
                     int bias_correction) {
  for (size_t i = 0; i < n; ++i) {
    float gr;
    if (g16) {
      /* `grad_data[i] * rescale_grad` with a half_t gradient is mshadow's half_t operator*: the
       * product is rounded to half before it is widened to float (3rdparty/mshadow half.h) */
      gr = kvo_half_to_float(kvo_float_to_half(kvo_half_to_float(g16[i], kind) * rescale, kind), kind);
    } else {
      gr = g[i] * rescale;
    }
    if (clip >= 0.f) gr = clipf(gr, clip);
    mean[i] = beta1 * mean[i] + (1.f - beta1) * gr;
    var[i] = beta2 * var[i] + (1.f - beta2) * gr * gr;
    float r = mean[i] / (sqrtf(var[i]) + eps) + wd * w[i];
    if (bias_correction) {
      const float mean_hat = (float)((double)mean[i] / (1. - (double)beta1_t));
      const float var_hat = var[i] / (1 - beta2_t);
      r = mean_hat / (sqrtf(var_hat) + eps) + wd * w[i];
    }
    out[i] = r;
  }
}

/* lamb_update_phase2 / mp_lamb_update_phase2, optimizer_op-inl.h:1705-1729, 1862-1886.
 * out16 != NULL: mixed precision -- the reference writes ONLY the 16-bit output from
 * weight32 - lr*g and leaves weight32 itself untouched. */
void kvo_lamb_phase2(size_t n, float* out, uint16_t* out16, int kind, const float* w, const float* g,
                     float r1, float r2, float lr, float lower_bound, float upper_bound) {
  float new_r1 = r1;
  if (lower_bound >= 0) new_r1 = new_r1 > lower_bound ? new_r1 : lower_bound;
  if (upper_bound >= 0) new_r1 = new_r1 < upper_bound ? new_r1 : upper_bound;
  if (new_r1 == 0.0f || r2 == 0.0f) {
    lr = lr * 1.0f;
  } else {
    lr = lr * new_r1 / r2;
  }
  for (size_t i = 0; i < n; ++i) {
    const float v = w[i] - lr * g[i];
    if (out16) out16[i] = kvo_float_to_half(v, kind); else out[i] = v;
  }
}

/* _multi_lamb_update step 1 for ONE tensor, src/operator/contrib/multi_lamb.cc:33-77
 * (MultiLAMBKernelStep1); power::Map on floats is powf. w is fp32 (or the master copy). */
void kvo_multi_lamb_step1(size_t n, float* temp_g, float* mean, float* var, const float* w,
                          const float* g, const uint16_t* g16, int kind, float clip, float rescale,
                          float beta1, float beta2, float eps, float wd, int step_count,
                          int bias_correction) {
  for (size_t i = 0; i < n; ++i) {
    float sg = (g16 ? kvo_half_to_float(g16[i], kind) : g[i]) * rescale;
    if (clip >= 0.0f) sg = clipf(sg, clip);
    const float m = beta1 * mean[i] + (1.0f - beta1) * sg;
    const float v = beta2 * var[i] + (1.0f - beta2) * sg * sg;
    mean[i] = m;
    var[i] = v;
    float r;
    if (bias_correction) {
      const float mean_hat = m / (1.0f - powf(beta1, (float)step_count));
      const float var_hat = v / (1.0f - powf(beta2, (float)step_count));
      r = mean_hat / (sqrtf(var_hat) + eps) + wd * w[i];
    } else {
      r = m / (sqrtf(v) + eps) + wd * w[i];
    }
    temp_g[i] = r;
  }
}

/* _multi_lamb_update step 2 for ONE tensor, multi_lamb.cc:79-118 (MultiLAMBKernelStep2) */
void kvo_multi_lamb_step2(size_t n, float* w, uint16_t* w16, int kind, const float* temp_g,
                          float sum_sq_w, float sum_sq_g, float lr, float lower_bound,
                          float upper_bound) {
  float r1 = sqrtf(sum_sq_w);
  const float r2 = sqrtf(sum_sq_g);
  if (lower_bound >= 0) r1 = r1 > lower_bound ? r1 : lower_bound;
  if (upper_bound >= 0) r1 = r1 < upper_bound ? r1 : upper_bound;
  float r;
  if (r1 == 0.0f || r2 == 0.0f) r = 1.0f; else r = r1 / r2;
  const float lr_adjusted = lr * r;
  for (size_t i = 0; i < n; ++i) {
    float wv = w[i];
    wv -= lr_adjusted * temp_g[i];
    w[i] = wv;
    if (w16) w16[i] = kvo_float_to_half(wv, kind);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* standard (non-lazy) updates with a row_sparse gradient                        */
/* ------------------------------------------------------------------------------------------ */

/* sgd_update, lazy_update=False (optimizer_op-inl.h:505-528): the whole weight is scaled by
 * DType(1 - lr*wd), then SGDDnsRspKernel runs over the gradient's rows with wd = 0. */
void kvo_sgd_std_rsp_update(size_t num_rows, size_t row_len, float* w, const int64_t* gidx,
                            const float* gval, size_t nnr, float clip, float lr, float wd,
                            float rescale) {
  const float scale = (float)(1 - lr * wd);
  for (size_t i = 0; i < num_rows * row_len; ++i) w[i] = w[i] * scale;
  kvo_sgd_rsp_update(nnr, row_len, w, gidx, gval, clip, lr, 0.f, rescale);
}

/* sgd_mom_update, lazy_update=False (optimizer_op.cc:108-139 SGDMomStdDnsRspDnsKernel<req,cpu>):
 * every row is updated; rows absent from the gradient use grad = 0. */
void kvo_sgd_mom_std_rsp_update(size_t num_rows, size_t row_len, float* w, float* mom,
                                const int64_t* gidx, const float* gval, size_t nnr, float clip,
                                float momentum, float lr, float wd, float rescale) {
  const float rate = lr * wd;
  size_t k = 0; /* gradient rows are ascending: walk them alongside the weight rows */
  for (size_t r = 0; r < num_rows; ++r) {
    const int present = k < nnr && (size_t)gidx[k] == r;
    const float* g = present ? gval + k * row_len : NULL;
    for (size_t j = 0; j < row_len; ++j) {
      const size_t i = r * row_len + j;
      const float grad = present ? g[j] : 0.f;
      if (clip >= 0.0f) {
        mom[i] = momentum * mom[i] - rate * w[i] - lr * clipf(rescale * grad, clip);
      } else {
        mom[i] = momentum * mom[i] - rate * w[i] - lr * rescale * grad;
      }
      w[i] = w[i] + mom[i];
    }
    if (present) ++k;
  }
}

/* adam_update, lazy_update=False (optimizer_op.cc:195-229 AdamStdDnsRspDnsKernel<req,cpu>):
 * absent rows see grad_rescaled = w*wd; the variance squares first: (1-beta2)*(g'*g'). */
void kvo_adam_std_rsp_update(size_t num_rows, size_t row_len, float* w, float* mean, float* var,
                             const int64_t* gidx, const float* gval, size_t nnr, float clip,
                             float beta1, float beta2, float lr, float wd, float eps,
                             float rescale) {
  size_t k = 0;
  for (size_t r = 0; r < num_rows; ++r) {
    const int present = k < nnr && (size_t)gidx[k] == r;
    const float* g = present ? gval + k * row_len : NULL;
    for (size_t j = 0; j < row_len; ++j) {
      const size_t i = r * row_len + j;
      float gr = present ? (g[j] * rescale + w[i] * wd) : (w[i] * wd);
      if (clip >= 0.0f) gr = clipf(gr, clip);
      mean[i] = beta1 * mean[i] + (1.f - beta1) * gr;
      var[i] = beta2 * var[i] + (1.f - beta2) * (gr * gr);
      w[i] = w[i] - lr * mean[i] / (sqrtf(var[i]) + eps);
    }
    if (present) ++k;
  }
}
