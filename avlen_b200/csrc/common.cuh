// Shared helpers for the avlen_b200 CUDA library (sm_100a only).
#pragma once
#ifdef AVL_HOST_EMUL
// CPU emulation build used only by tests/ (tests/emul/cuda_emul.h)
#include "cuda_emul.h"
#define AVL_DYN_SMEM(name) unsigned char* name = ::smem_raw
#define AVL_LAUNCH(kern, grid, block, smem, stream, ...) \
  emul::launch(dim3(grid), dim3(block), [&] { kern(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define AVL_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define AVL_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#include <stdint.h>
#include <math.h>

#define AVL_OK 0
#define AVL_ERR_ARG (-1)         // null pointer / negative size / inconsistent shapes
#define AVL_ERR_UNSUPPORTED (-2) // size outside what the kernels are built for
#define AVL_ERR_CUDA (-3)        // a CUDA runtime call failed (see avl_last_cuda_error)

extern "C" int avl_set_cuda_error(int e);

#define AVL_CUDA_CHECK(expr)                                  \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      avl_set_cuda_error((int)_e);                            \
      return AVL_ERR_CUDA;                                    \
    }                                                         \
  } while (0)

extern "C" void avl_count_launch();
extern "C" void avl_add_launches(long long n);
extern "C" void avl_bump_config_epoch();
extern "C" long long avl_config_epoch();
#define AVL_LAUNCH_CHECK()                \
  do {                                    \
    avl_count_launch();                   \
    AVL_CUDA_CHECK(cudaGetLastError());   \
  } while (0)

#define AVL_API extern "C" __attribute__((visibility("default")))

static inline int avl_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (all threads must call); result valid in all threads
__device__ __forceinline__ float block_sum(float v, float* red /* >=33 floats smem */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) red[32] = r;
  }
  __syncthreads();
  return red[32];
}

int avl_num_sms();

// ---- programmatic dependent launch (PDL): a kernel that calls avl_pdl_wait() before its first access to global memory
// written by earlier kernels may be LAUNCHED while its predecessor in the stream still runs (launch attribute added by
// avl_pdl_attr): its prologue (barrier / TMEM set-up, weight staging, the launch latency itself) overlaps the predecessor's
// tail.  Every such kernel triggers its own dependents only AFTER its wait, so completion is transitive along the chain.
#ifndef AVL_HOST_EMUL
__device__ __forceinline__ void avl_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void avl_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
extern "C" int avl_pdl_enabled();
extern "C" int avl_wide_stores();
// rows of `ld` elements of `esize` bytes starting at p are 32-byte aligned (and wide stores are on)
static inline int avl_rows_32b(const void* p, long long ld, int esize) {
  return (avl_wide_stores() && ((ld * esize) & 31) == 0 && ((uintptr_t)p & 31) == 0) ? 1 : 0;
}
static inline void avl_pdl_attr(cudaLaunchAttribute* at, unsigned* n) {
  if (avl_pdl_enabled()) {
    at[*n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[*n].val.programmaticStreamSerializationAllowed = 1;
    ++*n;
  }
}
// <<<>>>-style launch with the programmatic-serialization attribute (for kernels that call avl_pdl_wait())
template <typename... KArgs, typename... Args>
static inline void avl_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  unsigned n = 0;
  avl_pdl_attr(at, &n);
  cfg.attrs = at;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through cudaGetLastError at the caller
}
#define AVL_LAUNCH_PDL(kern, grid, block, smem, stream, ...) avl_launch_pdl(kern, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)
#else
static inline void avl_pdl_wait() {}
static inline void avl_pdl_trigger() {}
#define AVL_LAUNCH_PDL(kern, grid, block, smem, stream, ...) AVL_LAUNCH(kern, grid, block, smem, stream, __VA_ARGS__)
#endif
