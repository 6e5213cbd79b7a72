// peer_probe.cu -- measurement harness (not product code): which access pattern moves the fused
// reduce-scatter + update + all-gather traffic of the dense KVStore kernel fastest over NVLink?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/peer_probe tools/peer_probe.cu
//   ./peer_probe <ngpu> [melems=25.5]
//
// One process, N GPUs with peer access (same NVLink path as the IPC mappings of the rank-per-GPU
// store). Every GPU holds g (gradient), w, m (momentum), out; stripes of 32768 elements rotate over
// the GPUs; a GPU sums its stripes over all GPUs' g, applies w += (m = mu*m - lr*sum), stores w, m
// locally and out to EVERY GPU. Variants differ only in how the bytes move. Reported: ms per step
// (max over GPUs, CUDA events, 10 back-to-back steps after 3 warm-ups) and the all-reduce bus
// bandwidth per GPU, S*2(N-1)/N/t (tools/bandwidth/measure.py:137-138).
#include <cuda_runtime.h>
#include <cstdint>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kMaxN = 8;
constexpr int kChunk = 4096;
constexpr int kStripe = 32768;
constexpr int kThreads = 256;

struct Args {
  const float* g[kMaxN];
  float* out[kMaxN];
  float* w;
  float* m;
  int n, rank;
  long long nelem;
  int n_chunks_owned;   // chunks this rank owns
  int n_chunks_all;
};

__device__ __forceinline__ long long owned_chunk_base(const Args& a, int j) {
  // j-th owned chunk -> element offset. stripes rotate: stripe s owned by s % n; 8 chunks per stripe
  const int per = kStripe / kChunk;
  const long long stripe = (long long)(j / per) * a.n + a.rank;
  return stripe * kStripe + (long long)(j % per) * kChunk;
}

__device__ __forceinline__ float4 ldcs(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void stwb(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ void upd(float4& w, float4& m, float4 g) {
  m.x = 0.9f * m.x - 0.1f * g.x; m.y = 0.9f * m.y - 0.1f * g.y; m.z = 0.9f * m.z - 0.1f * g.z; m.w = 0.9f * m.w - 0.1f * g.w;
  w = add4(w, m);
}

// V0: the product's current shape: one CTA per owned chunk, one vector at a time per thread
template <int N>
__global__ void __launch_bounds__(kThreads) v0_chunk_cta(Args a) {
  const long long base = owned_chunk_base(a, blockIdx.x);
  if (base >= a.nelem) return;
  for (int v = threadIdx.x; v < kChunk / 4; v += kThreads) {
    const long long e = base + v * 4;
    if (e + 4 > a.nelem) break;
    float4 g[N];
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = ldcs(a.g[i] + e);
    float4 w = ldcs(a.w + e), m = ldcs(a.m + e);
    float4 s = g[0];
#pragma unroll
    for (int i = 1; i < N; ++i) s = add4(s, g[i]);
    upd(w, m, s);
    stcs(a.m + e, m);
    stcs(a.w + e, w);
#pragma unroll
    for (int i = 0; i < N; ++i) stcs(a.out[i] + e, w);
  }
}


// ---- V0B: V0 plus the product's cross-GPU barriers (signal pads, st.release.sys / ld.acquire.sys)
// mode 1 = start + end barrier (product), 2 = start barrier only; timeline of GPU 0 in `tl`
struct Bar {
  unsigned* pads[kMaxN];   // per GPU: [2][kMaxN][32] words (start flags, end flags), peer mapped
  unsigned* counter;       // this GPU's finished-CTA counter
  unsigned epoch;
  unsigned long long* tl;  // [4] accumulators: start-wait, data, end-wait, (kernel begin stamp)
};
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void bar_signal_wait(const Args& a, const Bar& b, int phase, bool signal) {
  const int t = threadIdx.x;
  if (t < a.n && t != a.rank) {
    const int base = phase * kMaxN * 32;
    if (signal) st_release_sys(b.pads[t] + base + a.rank * 32, b.epoch);
    const unsigned* mine = b.pads[a.rank] + base + t * 32;
    while ((int)(ld_acquire_sys(mine) - b.epoch) < 0) __nanosleep(32);
  }
}
template <int N, int MODE>
__global__ void __launch_bounds__(kThreads) v0b_barriers(Args a, Bar b) {
  __shared__ unsigned s_last;
  unsigned long long t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) t0 = gtime();
  bar_signal_wait(a, b, 0, blockIdx.x == 0);
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0 && b.tl) { const unsigned long long t1 = gtime(); atomicAdd(&b.tl[0], t1 - t0); b.tl[3] = t1; }
  const long long base = owned_chunk_base(a, blockIdx.x);
  for (int v = threadIdx.x; v < kChunk / 4; v += kThreads) {
    const long long e = base + v * 4;
    float4 g[N];
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = ldcs(a.g[i] + e);
    float4 w = ldcs(a.w + e), m = ldcs(a.m + e);
    float4 s = g[0];
#pragma unroll
    for (int i = 1; i < N; ++i) s = add4(s, g[i]);
    upd(w, m, s);
    stcs(a.m + e, m);
    stcs(a.w + e, w);
#pragma unroll
    for (int i = 0; i < N; ++i) stcs(a.out[i] + e, w);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned done = atomicAdd(b.counter, 1u);
    s_last = (done == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    unsigned long long t2 = 0;
    if (threadIdx.x == 0) { t2 = gtime(); if (b.tl) atomicAdd(&b.tl[1], t2 - b.tl[3]); }
    __threadfence_system();
    if (MODE == 1) { bar_signal_wait(a, b, 1, true); __syncthreads(); }
    if (threadIdx.x == 0) { *b.counter = 0; if (b.tl) atomicAdd(&b.tl[2], gtime() - t2); }
  }
}

// V1: U vectors per thread, every load issued before the first store; grid = chunks or persistent
template <int N, int U, bool PERSIST, bool WB>
__global__ void __launch_bounds__(kThreads) v1_unrolled(Args a) {
  const int step = PERSIST ? gridDim.x : 1 << 30;
  for (int j = blockIdx.x; j < a.n_chunks_owned; j += step) {
    const long long base = owned_chunk_base(a, j);
    if (base >= a.nelem) return;
#pragma unroll 1
    for (int v0 = 0; v0 < kChunk / 4; v0 += kThreads * U) {
      float4 g[U][N], w[U], m[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + (long long)(v0 + u * kThreads + threadIdx.x) * 4;
#pragma unroll
        for (int i = 0; i < N; ++i) g[u][i] = ldcs(a.g[i] + e);
        w[u] = ldcs(a.w + e);
        m[u] = ldcs(a.m + e);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + (long long)(v0 + u * kThreads + threadIdx.x) * 4;
        float4 s = g[u][0];
#pragma unroll
        for (int i = 1; i < N; ++i) s = add4(s, g[u][i]);
        upd(w[u], m[u], s);
        stcs(a.m + e, m[u]);
        stcs(a.w + e, w[u]);
#pragma unroll
        for (int i = 0; i < N; ++i) { if (WB) stwb(a.out[i] + e, w[u]); else stcs(a.out[i] + e, w[u]); }
      }
    }
    if (!PERSIST) return;
  }
}

// V2: all-gather + redundant update: every GPU reads ALL of every peer's gradient, no peer stores
template <int N, int U>
__global__ void __launch_bounds__(kThreads) v2_allgather(Args a) {
  for (int j = blockIdx.x; j < a.n_chunks_all; j += gridDim.x) {
    const long long base = (long long)j * kChunk;
#pragma unroll 1
    for (int v0 = 0; v0 < kChunk / 4; v0 += kThreads * U) {
      float4 g[U][N], w[U], m[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + (long long)(v0 + u * kThreads + threadIdx.x) * 4;
        if (e + 4 > a.nelem) continue;
#pragma unroll
        for (int i = 0; i < N; ++i) g[u][i] = ldcs(a.g[i] + e);
        w[u] = ldcs(a.w + e);
        m[u] = ldcs(a.m + e);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + (long long)(v0 + u * kThreads + threadIdx.x) * 4;
        if (e + 4 > a.nelem) continue;
        float4 s = g[u][0];
#pragma unroll
        for (int i = 1; i < N; ++i) s = add4(s, g[u][i]);
        upd(w[u], m[u], s);
        stcs(a.m + e, m[u]);
        stcs(a.w + e, w[u]);
        stcs(a.out[a.rank] + e, w[u]);
      }
    }
  }
}

// V5 / V6: pure peer read / pure peer write of the same bytes the fused kernel moves per direction
template <int N>
__global__ void __launch_bounds__(kThreads) v5_pure_read(Args a, float* sink) {
  float4 acc = make_float4(0, 0, 0, 0);
  for (int j = blockIdx.x; j < a.n_chunks_owned; j += gridDim.x) {
    const long long base = owned_chunk_base(a, j);
    if (base >= a.nelem) break;
    for (int v = threadIdx.x; v < kChunk / 4; v += kThreads) {
      const long long e = base + v * 4;
#pragma unroll
      for (int i = 0; i < N; ++i) if (i != a.rank) acc = add4(acc, ldcs(a.g[i] + e));
    }
  }
  if (acc.x == 12345.f) sink[0] = acc.x + acc.y + acc.z + acc.w;
}
template <int N>
__global__ void __launch_bounds__(kThreads) v6_pure_write(Args a) {
  for (int j = blockIdx.x; j < a.n_chunks_owned; j += gridDim.x) {
    const long long base = owned_chunk_base(a, j);
    if (base >= a.nelem) break;
    for (int v = threadIdx.x; v < kChunk / 4; v += kThreads) {
      const long long e = base + v * 4;
      const float4 w = make_float4(1.f, 2.f, 3.f, (float)v);
#pragma unroll
      for (int i = 0; i < N; ++i) if (i != a.rank) stcs(a.out[i] + e, w);
    }
  }
}

// ---- V3: TMA bulk tiles. One elected thread streams the chunk's N gradient tiles + w + m into a
// shared-memory ring (cp.async.bulk + mbarrier), all threads combine from smem into an out tile,
// the elected thread bulk-stores it to w, m (local) and every GPU's out.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(K) : "memory"); }

// TILE floats per tile (TILE*4 bytes); stage = N+2 input tiles; 2 output tiles (w, m) double-buffered
template <int N, int TILE, int STAGES>
__global__ void __launch_bounds__(kThreads) v3_tma(Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  float* in = reinterpret_cast<float*>(smem);                       // [STAGES][N+2][TILE]
  float* outb = in + (size_t)STAGES * (N + 2) * TILE;                // [2][2][TILE]  (w, m)
  constexpr int kTilesPerChunk = kChunk / TILE;
  const int n_tiles = a.n_chunks_owned * kTilesPerChunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto tile_elem = [&](int t) -> long long {
    return owned_chunk_base(a, t / kTilesPerChunk) + (long long)(t % kTilesPerChunk) * TILE;
  };
  auto issue = [&](int k) {   // k-th tile of this CTA
    const int t = blockIdx.x + k * gridDim.x;
    const int s = k % STAGES;
    const long long e = tile_elem(t);
    float* dst = in + (size_t)s * (N + 2) * TILE;
    mbar_expect_tx(&full[s], (N + 2) * TILE * 4);
#pragma unroll
    for (int i = 0; i < N; ++i) tma_load(dst + i * TILE, a.g[i] + e, TILE * 4, &full[s]);
    tma_load(dst + N * TILE, a.w + e, TILE * 4, &full[s]);
    tma_load(dst + (N + 1) * TILE, a.m + e, TILE * 4, &full[s]);
  };
  int my_tiles = 0;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (tile_elem(t) + TILE > a.nelem) break;
    ++my_tiles;
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < STAGES - 1 && k < my_tiles; ++k) issue(k);
  }
  for (int k = 0; k < my_tiles; ++k) {
    const int s = k % STAGES;
    const int ob = k & 1;
    if (threadIdx.x == 0) {
      if (k + STAGES - 1 < my_tiles) issue(k + STAGES - 1);   // stage (k-1)%STAGES was consumed at k-1
      // the bulk stores that read out buffer `ob` two tiles ago must have finished reading it
      tma_wait_read<1>();
    }
    mbar_wait(&full[s], (k / STAGES) & 1);
    __syncthreads();   // thread 0's wait_read is visible: out buffer `ob` is free
    const float* src = in + (size_t)s * (N + 2) * TILE;
    float* wo = outb + (size_t)ob * 2 * TILE;
    float* mo = wo + TILE;
    for (int v = threadIdx.x; v < TILE / 4; v += kThreads) {
      float4 sum = reinterpret_cast<const float4*>(src)[v];
#pragma unroll
      for (int i = 1; i < N; ++i) sum = add4(sum, reinterpret_cast<const float4*>(src + i * TILE)[v]);
      float4 w = reinterpret_cast<const float4*>(src + N * TILE)[v];
      float4 m = reinterpret_cast<const float4*>(src + (N + 1) * TILE)[v];
      upd(w, m, sum);
      reinterpret_cast<float4*>(wo)[v] = w;
      reinterpret_cast<float4*>(mo)[v] = m;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long e = tile_elem(blockIdx.x + k * gridDim.x);
      tma_store(a.w + e, wo, TILE * 4);
      tma_store(a.m + e, mo, TILE * 4);
#pragma unroll
      for (int i = 0; i < N; ++i) tma_store(a.out[i] + e, wo, TILE * 4);
      tma_commit();
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Gpu {
  float *g, *w, *m, *out, *sink;
  cudaStream_t st;
  cudaEvent_t e0, e1;
};

template <typename F>
static double run_variant(const char* name, int n, long long nelem, std::vector<Gpu>& G, F launch, double bytes_dir_factor) {
  const int warm = 3, iters = 10;
  for (int it = 0; it < warm + iters; ++it) {
    if (it == warm) {
      for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(G[d].st)); }
      for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(G[d].e0, G[d].st)); }
    }
    for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); launch(d); }
  }
  double worst = 0;
  for (int d = 0; d < n; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaEventRecord(G[d].e1, G[d].st));
    CK(cudaStreamSynchronize(G[d].st));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, G[d].e0, G[d].e1));
    worst = std::max(worst, (double)ms / iters);
  }
  const double S = (double)nelem * 4;
  const double bus = S * bytes_dir_factor / (worst * 1e-3) / 1e9;
  printf("%-44s %8.4f ms   %7.1f GB/s bus/GPU\n", name, worst, bus);
  fflush(stdout);
  return worst;
}

template <int N>
static void run_all(long long nelem) {
  const int n = N;
  std::vector<Gpu> G(n);
  for (int d = 0; d < n; ++d) {
    CK(cudaSetDevice(d));
    for (int p = 0; p < n; ++p) if (p != d) { cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e); cudaGetLastError(); }
    const size_t bytes = (size_t)nelem * 4 + (1 << 20);
    CK(cudaMalloc(&G[d].g, bytes)); CK(cudaMalloc(&G[d].w, bytes)); CK(cudaMalloc(&G[d].m, bytes)); CK(cudaMalloc(&G[d].out, bytes));
    CK(cudaMalloc(&G[d].sink, 256));
    CK(cudaMemset(G[d].g, 0, bytes)); CK(cudaMemset(G[d].w, 0, bytes)); CK(cudaMemset(G[d].m, 0, bytes));
    CK(cudaStreamCreateWithFlags(&G[d].st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&G[d].e0)); CK(cudaEventCreate(&G[d].e1));
  }
  const int chunks_all = (int)(nelem / kChunk);
  std::vector<Args> A(n);
  for (int d = 0; d < n; ++d) {
    Args& a = A[d];
    for (int i = 0; i < n; ++i) { a.g[i] = G[i].g; a.out[i] = G[i].out; }
    a.w = G[d].w; a.m = G[d].m; a.n = n; a.rank = d; a.nelem = (nelem / (kStripe * n)) * (kStripe * n);
    a.n_chunks_all = (int)(a.nelem / kChunk);
    a.n_chunks_owned = a.n_chunks_all / n;
  }
  (void)chunks_all;
  const double f = 2.0 * (n - 1) / n;
  printf("---- N=%d GPUs, %.1f M elements (%.1f MB) per GPU, %d owned chunks per GPU\n", n, nelem / 1e6, nelem * 4 / 1e6, A[0].n_chunks_owned);
  run_variant("V5 pure peer read  (S(N-1)/N in)", n, nelem, G, [&](int d) { v5_pure_read<N><<<148 * 8, kThreads, 0, G[d].st>>>(A[d], G[d].sink); }, (n - 1.0) / n);
  run_variant("V6 pure peer write (S(N-1)/N out)", n, nelem, G, [&](int d) { v6_pure_write<N><<<148 * 8, kThreads, 0, G[d].st>>>(A[d]); }, (n - 1.0) / n);
  run_variant("V0 chunk-per-CTA, 1 vector/thread (product r1)", n, nelem, G, [&](int d) { v0_chunk_cta<N><<<A[d].n_chunks_owned, kThreads, 0, G[d].st>>>(A[d]); }, f);
  {
    // barrier emulation (same process: plain peer-mapped pads)
    std::vector<unsigned*> pads(n), counters(n);
    std::vector<unsigned long long*> tls(n);
    for (int d = 0; d < n; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaMalloc(&pads[d], 2 * kMaxN * 32 * 4)); CK(cudaMemset(pads[d], 0, 2 * kMaxN * 32 * 4));
      CK(cudaMalloc(&counters[d], 256)); CK(cudaMemset(counters[d], 0, 256));
      CK(cudaMalloc(&tls[d], 64)); CK(cudaMemset(tls[d], 0, 64));
      CK(cudaDeviceSynchronize());
    }
    unsigned epoch = 0;
    auto mk = [&](int d, unsigned ep) { Bar b; for (int i = 0; i < n; ++i) b.pads[i] = pads[i]; b.counter = counters[d]; b.epoch = ep; b.tl = tls[d]; return b; };
    for (int mode = 1; mode <= 2; ++mode) {
      for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(tls[d], 0, 64)); CK(cudaDeviceSynchronize()); }
      int calls = 0;
      char nm[96];
      snprintf(nm, sizeof nm, "V0B V0 + %s", mode == 1 ? "start+end barriers (product)" : "start barrier only");
      run_variant(nm, n, nelem, G, [&](int d) {
        if (d == 0) { ++epoch; ++calls; }
        if (mode == 1) v0b_barriers<N, 1><<<A[d].n_chunks_owned, kThreads, 0, G[d].st>>>(A[d], mk(d, epoch));
        else v0b_barriers<N, 2><<<A[d].n_chunks_owned, kThreads, 0, G[d].st>>>(A[d], mk(d, epoch));
      }, f);
      unsigned long long h[4];
      CK(cudaSetDevice(0)); CK(cudaMemcpy(h, tls[0], 32, cudaMemcpyDeviceToHost));
      printf("      GPU0 per step: start-barrier wait %.2f us, data phase %.2f us, end-barrier wait %.2f us (%d steps)\n",
             h[0] / 1e3 / calls, h[1] / 1e3 / calls, h[2] / 1e3 / calls, calls);
    }
  }
  run_variant("V1 chunk-per-CTA, U=2", n, nelem, G, [&](int d) { v1_unrolled<N, 2, false, false><<<A[d].n_chunks_owned, kThreads, 0, G[d].st>>>(A[d]); }, f);
  if (N <= 4) run_variant("V1 chunk-per-CTA, U=4", n, nelem, G, [&](int d) { v1_unrolled<N, (N <= 4 ? 4 : 2), false, false><<<A[d].n_chunks_owned, kThreads, 0, G[d].st>>>(A[d]); }, f);
  for (int k : {1, 2, 4, 8}) {
    char nm[96];
    snprintf(nm, sizeof nm, "V1 persistent 148x%d CTAs, U=2", k);
    run_variant(nm, n, nelem, G, [&](int d) { v1_unrolled<N, 2, true, false><<<148 * k, kThreads, 0, G[d].st>>>(A[d]); }, f);
    if (N <= 4) {
      snprintf(nm, sizeof nm, "V1 persistent 148x%d CTAs, U=4", k);
      run_variant(nm, n, nelem, G, [&](int d) { v1_unrolled<N, (N <= 4 ? 4 : 2), true, false><<<148 * k, kThreads, 0, G[d].st>>>(A[d]); }, f);
    }
  }
  run_variant("V1 persistent 148x4, U=2, default-policy stores", n, nelem, G, [&](int d) { v1_unrolled<N, 2, true, true><<<148 * 4, kThreads, 0, G[d].st>>>(A[d]); }, f);
  for (int k : {2, 4, 8}) {
    char nm[96];
    snprintf(nm, sizeof nm, "V2 all-gather + redundant update 148x%d, U=2", k);
    run_variant(nm, n, nelem, G, [&](int d) { v2_allgather<N, 2><<<148 * k, kThreads, 0, G[d].st>>>(A[d]); }, f);
  }
  {
    // TMA: 4 KB tiles (1024 floats): stage = (N+2)*4 KB
    constexpr int TILE = 1024;
    constexpr int ST = (N <= 2) ? 6 : (N <= 4 ? 4 : 3);
    const int smem = ST * (N + 2) * TILE * 4 + 2 * 2 * TILE * 4;
    for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaFuncSetAttribute(v3_tma<N, TILE, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); }
    for (int k : {1, 2, 3}) {
      if ((size_t)k * (smem + 1024) > 227 * 1024) continue;
      char nm[96];
      snprintf(nm, sizeof nm, "V3 TMA bulk tiles 4KB x%d stages, 148x%d CTAs", ST, k);
      run_variant(nm, n, nelem, G, [&](int d) { v3_tma<N, TILE, ST><<<148 * k, kThreads, smem, G[d].st>>>(A[d]); }, f);
    }
  }
  {
    constexpr int TILE = 4096;   // 16 KB tiles = one chunk
    constexpr int ST = (N <= 2) ? 3 : 2;
    const int smem = ST * (N + 2) * TILE * 4 + 2 * 2 * TILE * 4;
    if (smem <= 227 * 1024) {
      for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaFuncSetAttribute(v3_tma<N, TILE, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); }
      char nm[96];
      snprintf(nm, sizeof nm, "V3 TMA bulk tiles 16KB x%d stages, 148x1 CTAs", ST);
      run_variant(nm, n, nelem, G, [&](int d) { v3_tma<N, TILE, ST><<<148, kThreads, smem, G[d].st>>>(A[d]); }, f);
    }
  }
  for (int d = 0; d < n; ++d) {
    CK(cudaSetDevice(d));
    cudaFree(G[d].g); cudaFree(G[d].w); cudaFree(G[d].m); cudaFree(G[d].out); cudaFree(G[d].sink);
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2;
  const double me = argc > 2 ? atof(argv[2]) : 25.55;
  const long long nelem = (long long)(me * 1e6);
  int have = 0;
  CK(cudaGetDeviceCount(&have));
  if (have < n) { printf("need %d GPUs, have %d\n", n, have); return 2; }
  switch (n) {
    case 2: run_all<2>(nelem); break;
    case 4: run_all<4>(nelem); break;
    case 8: run_all<8>(nelem); break;
    default: printf("ngpu must be 2, 4 or 8\n"); return 2;
  }
  return 0;
}
