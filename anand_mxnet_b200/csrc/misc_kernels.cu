// misc_kernels.cu -- small elementwise kernels behind the imperative-op surface (custom updaters
// written in the host language: `local += recv`, `weight[:] += grad * rescale`, casts, fills).
// Not on the measured hot path; kept simple: 16-byte vector body + scalar tail, grid-stride.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "kernels.h"

namespace b200kv {
namespace {

template <typename T> __device__ __forceinline__ float ToF(T x);
template <> __device__ __forceinline__ float ToF<float>(float x) { return x; }
template <> __device__ __forceinline__ float ToF<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float ToF<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> __device__ __forceinline__ float ToF<double>(double x) { return static_cast<float>(x); }
template <> __device__ __forceinline__ float ToF<int32_t>(int32_t x) { return static_cast<float>(x); }
template <> __device__ __forceinline__ float ToF<int64_t>(int64_t x) { return static_cast<float>(x); }
template <> __device__ __forceinline__ float ToF<uint8_t>(uint8_t x) { return static_cast<float>(x); }
template <> __device__ __forceinline__ float ToF<int8_t>(int8_t x) { return static_cast<float>(x); }

template <typename T> __device__ __forceinline__ T FromF(float x);
template <> __device__ __forceinline__ float FromF<float>(float x) { return x; }
template <> __device__ __forceinline__ __half FromF<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 FromF<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ double FromF<double>(float x) { return x; }
template <> __device__ __forceinline__ int32_t FromF<int32_t>(float x) { return static_cast<int32_t>(x); }
template <> __device__ __forceinline__ int64_t FromF<int64_t>(float x) { return static_cast<int64_t>(x); }
template <> __device__ __forceinline__ uint8_t FromF<uint8_t>(float x) { return static_cast<uint8_t>(x); }
template <> __device__ __forceinline__ int8_t FromF<int8_t>(float x) { return static_cast<int8_t>(x); }

// mshadow evaluates 16-bit binary ops in float and rounds once (half.h:45-66); fp32 ops are plain
// IEEE adds/muls (no contraction possible: one op per element).
template <typename T>
__global__ void ew_kernel(int op, T* out, const T* a, const T* b, float scalar, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float x = (op == kEwFill) ? 0.f : ToF<T>(a[i]);
    float r;
    switch (op) {
      case kEwCopy: r = x; break;
      case kEwAdd: r = __fadd_rn(x, ToF<T>(b[i])); break;
      case kEwSub: r = __fsub_rn(x, ToF<T>(b[i])); break;
      case kEwMul: r = __fmul_rn(x, ToF<T>(b[i])); break;
      case kEwAddScalar: r = __fadd_rn(x, scalar); break;
      case kEwMulScalar: r = __fmul_rn(x, scalar); break;
      case kEwSqrt: r = __fsqrt_rn(x); break;
      default: r = scalar; break;
    }
    out[i] = FromF<T>(r);
  }
}

template <typename TO, typename TI>
__global__ void cast_kernel(TO* out, const TI* in, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    out[i] = FromF<TO>(ToF<TI>(in[i]));
  }
}

inline int GridFor(size_t n) {
  size_t blocks = (n + 255) / 256;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
}

template <typename T>
void ew_launch(int op, void* out, const void* a, const void* b, float scalar, size_t n,
               cudaStream_t s) {
  ew_kernel<T><<<GridFor(n), 256, 0, s>>>(op, static_cast<T*>(out), static_cast<const T*>(a),
                                          static_cast<const T*>(b), scalar, n);
}

template <typename TO>
void cast_from(void* out, const void* in, int in_dtype, size_t n, cudaStream_t s) {
  const int g = GridFor(n);
  switch (in_dtype) {
    case kFloat32: cast_kernel<TO, float><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const float*>(in), n); break;
    case kFloat16: cast_kernel<TO, __half><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const __half*>(in), n); break;
    case kBfloat16: cast_kernel<TO, __nv_bfloat16><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const __nv_bfloat16*>(in), n); break;
    case kFloat64: cast_kernel<TO, double><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const double*>(in), n); break;
    case kInt32: cast_kernel<TO, int32_t><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const int32_t*>(in), n); break;
    case kInt64: cast_kernel<TO, int64_t><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const int64_t*>(in), n); break;
    case kUint8: cast_kernel<TO, uint8_t><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const uint8_t*>(in), n); break;
    case kInt8: cast_kernel<TO, int8_t><<<g, 256, 0, s>>>(static_cast<TO*>(out), static_cast<const int8_t*>(in), n); break;
    default: KV_FATAL << "cast: unsupported source dtype " << DTypeName(in_dtype);
  }
}

}  // namespace

void LaunchElementwise(int op, int dtype, void* out, const void* a, const void* b, float scalar,
                       size_t n, cudaStream_t stream) {
  if (n == 0) return;
  switch (dtype) {
    case kFloat32: ew_launch<float>(op, out, a, b, scalar, n, stream); break;
    case kFloat16: ew_launch<__half>(op, out, a, b, scalar, n, stream); break;
    case kBfloat16: ew_launch<__nv_bfloat16>(op, out, a, b, scalar, n, stream); break;
    case kFloat64: ew_launch<double>(op, out, a, b, scalar, n, stream); break;
    case kInt32: ew_launch<int32_t>(op, out, a, b, scalar, n, stream); break;
    case kInt64: ew_launch<int64_t>(op, out, a, b, scalar, n, stream); break;
    default: KV_FATAL << "elementwise op: unsupported dtype " << DTypeName(dtype);
  }
  KV_CUDA(cudaGetLastError());
}

void LaunchCast(void* out, int out_dtype, const void* in, int in_dtype, size_t n,
                cudaStream_t stream) {
  if (n == 0) return;
  switch (out_dtype) {
    case kFloat32: cast_from<float>(out, in, in_dtype, n, stream); break;
    case kFloat16: cast_from<__half>(out, in, in_dtype, n, stream); break;
    case kBfloat16: cast_from<__nv_bfloat16>(out, in, in_dtype, n, stream); break;
    case kFloat64: cast_from<double>(out, in, in_dtype, n, stream); break;
    case kInt32: cast_from<int32_t>(out, in, in_dtype, n, stream); break;
    case kInt64: cast_from<int64_t>(out, in, in_dtype, n, stream); break;
    default: KV_FATAL << "cast: unsupported target dtype " << DTypeName(out_dtype);
  }
  KV_CUDA(cudaGetLastError());
}

}  // namespace b200kv
