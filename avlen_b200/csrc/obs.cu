// Observation ingest for the compact rollout storage: RolloutStorage keeps rgb as uint8 and depth
// as fp16 (the reference stores fp32, ss_baselines/savi/models/rollout_storage.py:58-63, after inflating uint8 frames on
// the host, common/utils.py:149-154), and the encoders' first op — x / 255 then the exact 2x2 area mean
// (smt_cnn.py:83-95, common/utils.py:515-517) — reads those types directly.  A minibatch of the PPO update is addressed
// by a per-row sample index into the time-major storage, so the (T*N_mb, 128, 128, C) copies the reference's generator
// stacks (rollout_storage.py:716-760) never exist.
#include "common.cuh"

#ifndef AVL_HOST_EMUL
#include <cuda_fp16.h>
namespace {

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float ld_as_float<unsigned char>(const unsigned char* p) { return (float)__ldg(p); }

// one thread per output pixel: reads the 2 x 2 x C source block (two contiguous runs of 2 C elements), writes Cp floats
template <typename T>
__global__ void resize_half_typed_kernel(const T* __restrict__ x, const long long* __restrict__ sample_index, float* y, int N,
                                         int H, int W, int C, int Cp, float scale) {
  const int OH = H >> 1, OW = W >> 1;
  const long long total = (long long)N * OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % OW);
    long long t = i / OW;
    const int oh = (int)(t % OH);
    const long long n = t / OH;
    const long long src = sample_index ? sample_index[n] : n;
    const T* p = x + ((src * H + 2 * oh) * W + 2 * ow) * C;
    const T* q = p + (long long)W * C;
    float* o = y + i * Cp;
    for (int c = 0; c < C; ++c) {
      // the reference divides by 255 first, then averages (smt_cnn.py:83-86): same operation order as resize_half_kernel
      const float a = ld_as_float(p + c) * scale, b = ld_as_float(p + C + c) * scale;
      const float d = ld_as_float(q + c) * scale, e = ld_as_float(q + C + c) * scale;
      o[c] = (a + b + d + e) * 0.25f;
    }
    for (int c = C; c < Cp; ++c) o[c] = 0.f;
  }
}

}  // namespace

// dtype: 0 fp32, 1 fp16, 2 uint8.  sample_index (optional, device int64 [N]): row n reads source sample sample_index[n].
AVL_API int avl_resize_half_typed(const void* x, int dtype, const long long* sample_index, float* y, int N, int H, int W,
                                  int C, int C_out, float scale, void* stream) {
  if (N < 0 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 1 || C_out < C || dtype < 0 || dtype > 2) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !y) return AVL_ERR_ARG;
  const long long total = (long long)N * (H / 2) * (W / 2);
  const int blocks = avl_div_up(total, 256) > 148 * 32 ? 148 * 32 : avl_div_up(total, 256);
  cudaStream_t cs = (cudaStream_t)stream;
  if (dtype == 0)
    resize_half_typed_kernel<float><<<blocks, 256, 0, cs>>>((const float*)x, sample_index, y, N, H, W, C, C_out, scale);
  else if (dtype == 1)
    resize_half_typed_kernel<__half><<<blocks, 256, 0, cs>>>((const __half*)x, sample_index, y, N, H, W, C, C_out, scale);
  else
    resize_half_typed_kernel<unsigned char><<<blocks, 256, 0, cs>>>((const unsigned char*)x, sample_index, y, N, H, W, C,
                                                                    C_out, scale);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL

// ------------------------------------------------------------------------------------------ fused storage insert
// RolloutStorage.insert (ss_baselines/savi/models/rollout_storage.py:214-295) writes ~20-35 small tensors into slot
// ``step`` of the time-major stores: as torch ``copy_`` calls that is one launch (and ~8 us of host time) EACH, every
// rollout step.  Here all of them are one launch: a table of (dst, src, bytes, kind) segments travels in the kernel
// parameters, one CTA column per segment.  kind 0: byte copy; 1: fp32 -> uint8 (rgb into the compact store);
// 2: fp32 -> fp16 (depth); 3: int64 -> fp32; 4: fp32 -> int64.
#ifndef AVL_HOST_EMUL
namespace {

constexpr int MC_MAX = 48;
struct MultiCopy {
  void* dst[MC_MAX];
  const void* src[MC_MAX];
  long long n[MC_MAX];      // bytes (kind 0) or elements
  unsigned char kind[MC_MAX];
  int count;
};

__global__ void multi_copy_kernel(MultiCopy m) {
  const int s = blockIdx.y;
  if (s >= m.count) return;
  const long long n = m.n[s];
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const int kind = m.kind[s];
  if (kind == 0) {
    const uintptr_t a = (uintptr_t)m.dst[s] | (uintptr_t)m.src[s];
    if (!(a & 15) && !(n & 15)) {
      const uint4* src = (const uint4*)m.src[s];
      uint4* dst = (uint4*)m.dst[s];
      for (long long i = i0; i < (n >> 4); i += stride) dst[i] = src[i];
    } else if (!(a & 3) && !(n & 3)) {
      const uint32_t* src = (const uint32_t*)m.src[s];
      uint32_t* dst = (uint32_t*)m.dst[s];
      for (long long i = i0; i < (n >> 2); i += stride) dst[i] = src[i];
    } else {
      const unsigned char* src = (const unsigned char*)m.src[s];
      unsigned char* dst = (unsigned char*)m.dst[s];
      for (long long i = i0; i < n; i += stride) dst[i] = src[i];
    }
  } else if (kind == 1) {
    const float* src = (const float*)m.src[s];
    unsigned char* dst = (unsigned char*)m.dst[s];
    for (long long i = i0; i < n; i += stride) dst[i] = (unsigned char)src[i];  // torch copy_: truncation toward zero
  } else if (kind == 2) {
    const float* src = (const float*)m.src[s];
    __half* dst = (__half*)m.dst[s];
    for (long long i = i0; i < n; i += stride) dst[i] = __float2half_rn(src[i]);
  } else if (kind == 3) {
    const long long* src = (const long long*)m.src[s];
    float* dst = (float*)m.dst[s];
    for (long long i = i0; i < n; i += stride) dst[i] = (float)src[i];
  } else {
    const float* src = (const float*)m.src[s];
    long long* dst = (long long*)m.dst[s];
    for (long long i = i0; i < n; i += stride) dst[i] = (long long)src[i];
  }
}

}  // namespace

// dst / src: host arrays of ``count`` device pointers; n: bytes (kind 0) or elements; kind as above.  count <= 48.
AVL_API int avl_multi_copy(int count, void* const* dst, const void* const* src, const long long* n, const unsigned char* kind,
                           void* stream) {
  if (count < 0 || count > MC_MAX) return AVL_ERR_ARG;
  if (count == 0) return AVL_OK;
  if (!dst || !src || !n || !kind) return AVL_ERR_ARG;
  MultiCopy m;
  m.count = count;
  long long biggest = 0;
  for (int i = 0; i < count; ++i) {
    if (!dst[i] || !src[i] || n[i] < 0 || kind[i] > 4) return AVL_ERR_ARG;
    m.dst[i] = dst[i]; m.src[i] = src[i]; m.n[i] = n[i]; m.kind[i] = kind[i];
    const long long units = kind[i] == 0 ? (n[i] + 15) / 16 : n[i];
    if (units > biggest) biggest = units;
  }
  int bx = avl_div_up(biggest, 256 * 4);
  if (bx < 1) bx = 1;
  if (bx > 512) bx = 512;
  multi_copy_kernel<<<dim3(bx, count), 256, 0, (cudaStream_t)stream>>>(m);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL

// ------------------------------------------------------------------ RIR bank lookup + spectrogram cache
// The reference reads `{azimuth}/{receiver}_{source}.wav` from disk on every audio miss (soundspaces/simulator.py:650-659)
// and keeps per-simulator dicts `_audiogoal_cache` / `_spectrogram_cache` keyed by (source, receiver, azimuth)
// (:711-734), cleared when the scene or the sound changes (:393-395); the clip position `_audio_index` advances only
// when the audiogoal is actually computed (:668).  Here a scene's RIRs are ONE packed tensor in HBM with a dense
// (azimuth, receiver, source) -> (offset, length) table, and every env owns a dense spectrogram cache over the same
// key space:  lookup  -> RIR descriptors of the step + hit flags (a hit is rendered as "silent": the render kernel skips it)
//             commit  -> hits copy their cached spectrogram out, misses store theirs and advance the clip position.
#ifndef AVL_HOST_EMUL
namespace {

__global__ void spec_cache_lookup_kernel(int n, int V, const int* __restrict__ src, const int* __restrict__ recv,
                                         const int* __restrict__ az, const long long* __restrict__ tab_off,
                                         const int* __restrict__ tab_len, const unsigned char* __restrict__ clear,
                                         unsigned char* valid, const int* __restrict__ silent_in, long long* rir_off,
                                         int* rir_len, int* silent_out, unsigned char* hit) {
  const int i = blockIdx.x;
  const int S = V * V * 4;
  unsigned char* vrow = valid + (size_t)i * S;
  const bool clr = clear && clear[i];
  if (clr)
    for (int k = threadIdx.x; k < S; k += blockDim.x) vrow[k] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int key = (src[i] * V + recv[i]) * 4 + az[i];
    const int t = (az[i] * V + recv[i]) * V + src[i];
    rir_off[i] = tab_off[t];
    rir_len[i] = tab_len[t];
    const unsigned char h = clr ? 0 : vrow[key];
    hit[i] = h;
    silent_out[i] = h ? 1 : silent_in[i];
  }
}

__global__ void spec_cache_commit_kernel(int n, int V, int E, const int* __restrict__ src, const int* __restrict__ recv,
                                         const int* __restrict__ az, const unsigned char* __restrict__ hit,
                                         const int* __restrict__ silent_in, float* cache, unsigned char* valid, float* spec,
                                         int* index, const int* __restrict__ clip_secs) {
  const int i = blockIdx.x;
  const int S = V * V * 4;
  const int key = (src[i] * V + recv[i]) * 4 + az[i];
  float* c = cache + ((size_t)i * S + key) * E;
  float* s = spec + (size_t)i * E;
  if (hit[i]) {
    for (int k = threadIdx.x; k < E; k += blockDim.x) s[k] = c[k];
  } else {
    for (int k = threadIdx.x; k < E; k += blockDim.x) c[k] = s[k];
    if (threadIdx.x == 0) {
      valid[(size_t)i * S + key] = 1;
      if (!silent_in[i]) index[i] = (index[i] + 1) % clip_secs[i];  // simulator.py:668 (inside the non-silent branch)
    }
  }
}

}  // namespace

AVL_API int avl_spec_cache_lookup(int n, int V, const int* src, const int* recv, const int* az, const long long* tab_off,
                                  const int* tab_len, const unsigned char* clear, unsigned char* valid, const int* silent_in,
                                  long long* rir_off, int* rir_len, int* silent_out, unsigned char* hit, void* stream) {
  if (n < 0 || V < 1) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!src || !recv || !az || !tab_off || !tab_len || !valid || !silent_in || !rir_off || !rir_len || !silent_out || !hit)
    return AVL_ERR_ARG;
  spec_cache_lookup_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(n, V, src, recv, az, tab_off, tab_len, clear, valid,
                                                               silent_in, rir_off, rir_len, silent_out, hit);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_spec_cache_commit(int n, int V, int E, const int* src, const int* recv, const int* az,
                                  const unsigned char* hit, const int* silent_in, float* cache, unsigned char* valid,
                                  float* spec, int* index, const int* clip_secs, void* stream) {
  if (n < 0 || V < 1 || E < 1) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!src || !recv || !az || !hit || !silent_in || !cache || !valid || !spec || !index || !clip_secs) return AVL_ERR_ARG;
  spec_cache_commit_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(n, V, E, src, recv, az, hit, silent_in, cache, valid, spec,
                                                               index, clip_secs);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
