// compress_kernels.cu -- 2-bit gradient compression with error-feedback residual on the device
// path (reference: src/kvstore/gradient_compression-inl.h:40-132 quantize_2bit / dequantize_2bit,
// src/kvstore/comm.h:552-596 CommDevice::ReduceCompressed).
//
// B200 shape of the same algorithm: every source GPU quantises its own gradient locally (16 fp32
// -> one 32-bit word: 64 B read + 64 B residual read/write + 4 B written per block), so what
// crosses NVLink afterwards is 1/16 of the gradient; the owner GPU then decodes the peers' words
// straight from peer memory and sums the decoded values in ElementwiseSum order -- no compressed
// recv buffers, no per-source fp32 copy buffers. Bit layout and thresholds are the reference's, so
// the compressed words and residuals are bit-identical to its CPU kernels (oracle/kvoracle.c).
#include "common.h"
#include "kernels.h"

namespace b200kv {
namespace {

// one thread per block of 16 values
__global__ void __launch_bounds__(256) quantize_2bit_kernel(const float* __restrict__ grad,
                                                            float* __restrict__ residual,
                                                            uint32_t* __restrict__ compressed,
                                                            size_t n, float pos) {
  const float neg = -1 * pos;
  const size_t nblocks = (n + 15) / 16;
  for (size_t b = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; b < nblocks;
       b += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t start = b << 4;
    const size_t end = (start + 16 <= n) ? start + 16 : n;
    uint32_t word = 0;
    for (size_t i = start; i < end; ++i) {
      float r = __fadd_rn(residual[i], grad[i]);
      const int byte = static_cast<int>((i - start) >> 2);
      const int shift = byte * 8 + (6 - 2 * static_cast<int>(i & 3));  // posbits {c0,30,0c,03}
      if (r >= pos) {
        word |= (3u << shift);
        r = __fsub_rn(r, pos);
      } else if (r <= neg) {
        word |= (2u << shift);
        r = __fsub_rn(r, neg);
      }
      residual[i] = r;
    }
    compressed[b] = word;
  }
}

constexpr int kMaxCompSrc = 16;
struct CompSrcs {
  const uint32_t* p[kMaxCompSrc];
};

__global__ void __launch_bounds__(256) dequantize_sum_2bit_kernel(CompSrcs srcs, int nsrc,
                                                                  float* __restrict__ merged,
                                                                  size_t n, float pos) {
  const float neg = -1 * pos;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int shift = static_cast<int>((i & 15) >> 2) * 8 + (6 - 2 * static_cast<int>(i & 3));
    float acc = 0.f;
    for (int s = 0; s < nsrc; ++s) {
      const uint32_t code = (srcs.p[s][i >> 4] >> shift) & 3u;
      const float v = code == 3u ? pos : (code == 2u ? neg : 0.f);
      acc = (s == 0) ? v : __fadd_rn(acc, v);
    }
    merged[i] = acc;
  }
}

inline int GridFor(size_t work) {
  size_t blocks = (work + 255) / 256;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks));
}

}  // namespace

void LaunchQuantize2Bit(const float* grad, float* residual, uint32_t* compressed, size_t n,
                        float threshold, cudaStream_t stream) {
  if (n == 0) return;
  quantize_2bit_kernel<<<GridFor((n + 15) / 16), 256, 0, stream>>>(grad, residual, compressed, n,
                                                                   threshold);
  KV_CUDA(cudaGetLastError());
}

void LaunchDequantizeSum2Bit(const uint32_t* const* compressed, int nsrc, float* merged, size_t n,
                             float threshold, cudaStream_t stream) {
  if (n == 0) return;
  KV_CHECK(nsrc >= 1 && nsrc <= kMaxCompSrc);
  CompSrcs s;
  for (int i = 0; i < nsrc; ++i) s.p[i] = compressed[i];  // host array of device pointers
  dequantize_sum_2bit_kernel<<<GridFor(n), 256, 0, stream>>>(s, nsrc, merged, n, threshold);
  KV_CUDA(cudaGetLastError());
}

}  // namespace b200kv
