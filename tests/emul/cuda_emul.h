// Host emulation of the small CUDA subset the SIMT kernels in avlen_b200/csrc use.
// TEST INFRASTRUCTURE: lets the CPU test-suite execute the *same kernel source*
// (one OS thread per CUDA thread, std::barrier for __syncthreads/__syncwarp,
// a per-warp exchange buffer for shuffles) so index algebra and barrier
// placement are checked against the oracle before any GPU time is spent.
// Not used by the product; tcgen05/TMA kernels are excluded (#ifndef AVL_HOST_EMUL).
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

extern thread_local dim3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

namespace emul {
struct BlockCtx {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<std::array<uint32_t, 32>> warp_xchg;
};
extern thread_local BlockCtx* ctx;
extern thread_local int linear_tid;
}  // namespace emul

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __align__(x)
#define __launch_bounds__(...)
#define __constant__ static

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDeviceToHost = 2,
       cudaMemcpyHostToDevice = 1 };
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return 0; }

static inline void __syncthreads() { emul::ctx->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::ctx->warp_bar[emul::linear_tid >> 5]->arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
static inline T __shfl_xchg(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  int w = emul::linear_tid >> 5, lane = emul::linear_tid & 31;
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  emul::ctx->warp_xchg[w][lane] = bits;
  emul::ctx->warp_bar[w]->arrive_and_wait();
  uint32_t r = emul::ctx->warp_xchg[w][src_lane & 31];
  emul::ctx->warp_bar[w]->arrive_and_wait();
  T out;
  std::memcpy(&out, &r, 4);
  return out;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) { return __shfl_xchg(v, (emul::linear_tid & 31) ^ mask); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int lane) { return __shfl_xchg(v, lane); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
  int lane = emul::linear_tid & 31;
  return __shfl_xchg(v, lane + d < 32 ? lane + d : lane);
}

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }

static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
#define __expf(a) expf(a)
#define __logf(a) logf(a)
using std::min;
using std::max;
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
static inline double cospi(double x) { return cos(M_PI * x); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }

static inline int atomicExch(int* p, int v) { return reinterpret_cast<std::atomic<int>*>(p)->exchange(v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return reinterpret_cast<std::atomic<unsigned>*>(p)->fetch_add(v); }
static inline int atomicAdd(int* p, int v) { return reinterpret_cast<std::atomic<int>*>(p)->fetch_add(v); }
static inline float atomicAdd(float* p, float v) {
  auto* a = reinterpret_cast<std::atomic<uint32_t>*>(p);
  uint32_t old = a->load();
  for (;;) {
    float f;
    std::memcpy(&f, &old, 4);
    f += v;
    uint32_t nw;
    std::memcpy(&nw, &f, 4);
    if (a->compare_exchange_weak(old, nw)) { float r; std::memcpy(&r, &old, 4); return r; }
  }
}

static inline double atomicAdd(double* p, double v) {
  auto* a = reinterpret_cast<std::atomic<uint64_t>*>(p);
  uint64_t old = a->load();
  for (;;) {
    double f;
    std::memcpy(&f, &old, 8);
    f += v;
    uint64_t nw;
    std::memcpy(&nw, &f, 8);
    if (a->compare_exchange_weak(old, nw)) { double r; std::memcpy(&r, &old, 8); return r; }
  }
}

extern unsigned char smem_raw[];

namespace emul {
// Runs `body` once per CUDA thread: blocks sequentially, threads of a block concurrently.
inline void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
  const int nthreads = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        BlockCtx c;
        c.block_bar = std::make_unique<std::barrier<>>(nthreads);
        int nw = (nthreads + 31) / 32;
        for (int w = 0; w < nw; ++w) {
          int cnt = std::min(32, nthreads - 32 * w);
          c.warp_bar.push_back(std::make_unique<std::barrier<>>(cnt));
        }
        c.warp_xchg.resize(nw);
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
          ts.emplace_back([&, t] {
            ctx = &c;
            linear_tid = t;
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            blockIdx = dim3(bx, by, bz);
            blockDim = block;
            gridDim = grid;
            body();
          });
        }
        for (auto& th : ts) th.join();
      }
}
}  // namespace emul

#define AVL_EMUL_DEFINE_GLOBALS                                   \
  thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;       \
  namespace emul { thread_local BlockCtx* ctx = nullptr; thread_local int linear_tid = 0; } \
  alignas(1024) unsigned char smem_raw[256 * 1024];
