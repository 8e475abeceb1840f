// GroupNorm(16) (+ residual) (+ ReLU) on NHWC in ONE pass over HBM (smt_resnet.py:22-33, nn.GroupNorm eps 1e-5).
//
// A thread-block CLUSTER owns one sample: each of its CL CTAs stages 1/CL of the sample's pixels in shared memory
// while accumulating per-channel sums, the per-group partial sums are exchanged through distributed shared memory
// (every CTA reads its peers' partials after a cluster barrier), and the slice is normalised straight from shared
// memory.  x is read once and y written once — the two-pass kernels (statistics pass + apply pass, conv.cu) read x
// twice and need three launches.  With 64 environments in a rollout step this also turns 64 CTAs into 64 * CL.
//
// IN16 / OUT16: x (and the residual) / y are stored as fp16 (the encoders' widest activations, see conv_halo_tc.cu);
// statistics and the normalisation run in fp32 on the staged fp32 copy either way.
#include "common.cuh"

#ifndef AVL_HOST_EMUL
#include <cooperative_groups.h>
#include <cuda_fp16.h>
namespace cg = cooperative_groups;

namespace {

// fp16 storage guard: the convolution that produced an fp16 tensor SATURATES to +-65504 instead of overflowing to inf
// (conv_halo_tc.cu); this kernel reads every element of such a tensor anyway and raises a sticky flag when it meets a
// saturated value, so that the host can fall back to fp32 storage (avl_f16_overflow, nn.check_f16_overflow).
__device__ int g_f16_overflow = 0;

constexpr int GNC_THREADS = 256;
constexpr int GNC_MAX_SLICE = 48 * 1024;  // bytes of one CTA's slice (dynamic shared memory)

__device__ __forceinline__ float4 gn_cvt4(uint2 u) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <bool F16>
__device__ __forceinline__ float4 gn_ld4(const void* base, size_t i) {  // 4 consecutive channels, element index 4 * i
  if (F16) return gn_cvt4(__ldg(reinterpret_cast<const uint2*>(base) + i));
  return __ldg(reinterpret_cast<const float4*>(base) + i);
}
template <bool F16>
__device__ __forceinline__ void gn_st4(void* base, size_t i, float4 o) {
  if (F16) {
    const __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(base)[i] = u;
  } else {
    reinterpret_cast<float4*>(base)[i] = o;
  }
}

template <bool IN16, bool OUT16>
__global__ void __launch_bounds__(GNC_THREADS) gn_cluster_kernel(const void* __restrict__ x,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 const void* __restrict__ residual, void* y, int HW,
                                                                 int C, int groups, float eps, int relu, int cl,
                                                                 int pix_per_cta) {
  extern __shared__ __align__(16) unsigned char gsm[];
  __shared__ float acc_s[512], acc_q[512];  // per channel (C <= 512)
  __shared__ float red_s[1024], red_q[1024];  // per (slot, channel) partials, summed in a fixed order (deterministic)
  __shared__ double part[64][2];            // this CTA's per-group (sum, sumsq)
  __shared__ float g_mean[64], g_rstd[64];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  const int n = blockIdx.x / cl;
  const int nq = C >> 2;  // float4 per pixel; divides GNC_THREADS
  const int p0 = rank * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  const int n4 = max(0, p1 - p0) * nq;
  const size_t base4 = ((size_t)n * HW + p0) * nq;  // float4-sized element index of this CTA's slice
  // the slice is staged in its storage type: an fp16 sample needs half the shared memory, i.e. half the cluster
  // size (fewer barriers per sample) and twice the bytes in flight per CTA
  float4* tile = reinterpret_cast<float4*>(gsm);
  uint2* tile16 = reinterpret_cast<uint2*>(gsm);
  avl_pdl_wait();     // (programmatic dependent launch: this grid may have been launched before its producer finished)
  avl_pdl_trigger();  // the next convolution's prologue may start now

  // ---- pass over HBM: stage + per-thread sums (a thread always sees the same 4 channels: nq divides the stride)
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  unsigned sat = 0u;
  auto absorb = [&](const float4& v) {
    s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1); q2 = fmaf(v.z, v.z, q2); q3 = fmaf(v.w, v.w, q3);
  };
  int i = tid;
  if (IN16) {
    // an fp16 load is 8 bytes per thread: eight of them are issued before the first is consumed, so that a CTA keeps as
    // many bytes in flight as the fp32 variant does (the fp16 variant ran at 3.1 TB/s against 5.3, r01_halo_f16_bench.txt)
    constexpr int B = 8;
    for (; i + (B - 1) * GNC_THREADS < n4; i += B * GNC_THREADS) {
      uint2 u[B];
#pragma unroll
      for (int k = 0; k < B; ++k) u[k] = __ldg(reinterpret_cast<const uint2*>(x) + base4 + i + k * GNC_THREADS);
#pragma unroll
      for (int k = 0; k < B; ++k) {
        tile16[i + k * GNC_THREADS] = u[k];
        // |h| >= 65504 (0x7BFF): a saturated (or non-finite) fp16 value
        sat |= __vcmpgeu2(u[k].x & 0x7fff7fffu, 0x7bff7bffu) | __vcmpgeu2(u[k].y & 0x7fff7fffu, 0x7bff7bffu);
        absorb(gn_cvt4(u[k]));
      }
    }
  }
  for (; i < n4; i += GNC_THREADS) {
    float4 v;
    if (IN16) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(x) + base4 + i);
      tile16[i] = u;
      sat |= __vcmpgeu2(u.x & 0x7fff7fffu, 0x7bff7bffu) | __vcmpgeu2(u.y & 0x7fff7fffu, 0x7bff7bffu);
      v = gn_cvt4(u);
    } else {
      v = __ldg(reinterpret_cast<const float4*>(x) + base4 + i);
      tile[i] = v;
    }
    absorb(v);
  }
  if (IN16 && sat) atomicOr(&g_f16_overflow, 1);
  // lanes that share a channel quad inside the warp (nq < 32) combine first
  for (int o = 16; o >= nq && o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    q2 += __shfl_xor_sync(0xffffffffu, q2, o); q3 += __shfl_xor_sync(0xffffffffu, q3, o);
  }
  const int lane = tid & 31;
  const int c0 = (tid % nq) * 4;
  // slot = warp (nq < 32: lanes < nq hold the warp's totals) or tid / nq (nq >= 32: every thread its own quad)
  const int n_slots = nq < 32 ? GNC_THREADS / 32 : GNC_THREADS / nq;
  if (nq >= 32 || lane < nq) {
    const int slot = nq < 32 ? (tid >> 5) : tid / nq;
    float* ds = red_s + (slot * nq + (tid % nq)) * 4;
    float* dq = red_q + (slot * nq + (tid % nq)) * 4;
    ds[0] = s0; ds[1] = s1; ds[2] = s2; ds[3] = s3;
    dq[0] = q0; dq[1] = q1; dq[2] = q2; dq[3] = q3;
  }
  __syncthreads();
  for (int c = tid; c < C; c += GNC_THREADS) {
    float S = 0.f, Q = 0.f;
    for (int k = 0; k < n_slots; ++k) {
      S += red_s[(k * nq + (c >> 2)) * 4 + (c & 3)];
      Q += red_q[(k * nq + (c >> 2)) * 4 + (c & 3)];
    }
    acc_s[c] = S;
    acc_q[c] = Q;
  }
  __syncthreads();
  const int cpg = C / groups;
  if (tid < groups) {
    double S = 0.0, Q = 0.0;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) { S += (double)acc_s[c]; Q += (double)acc_q[c]; }
    part[tid][0] = S;
    part[tid][1] = Q;
  }
  cluster.sync();
  if (tid < groups) {
    double S = 0.0, Q = 0.0;
    for (int r = 0; r < cl; ++r) {
      const double* rp = cluster.map_shared_rank(&part[0][0], r);
      S += rp[tid * 2];
      Q += rp[tid * 2 + 1];
    }
    const double cnt = (double)HW * cpg;
    const double m = S / cnt;
    double var = Q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    g_mean[tid] = (float)m;
    g_rstd[tid] = (float)(1.0 / sqrt(var + (double)eps));
  }
  cluster.sync();  // also: no CTA may retire while a peer still reads its partials
  // ---- normalise from shared memory
  float a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j, g = c / cpg;
    a[j] = g_rstd[g] * __ldg(gamma + c);
    b[j] = __ldg(beta + c) - g_mean[g] * a[j];
  }
  auto emit = [&](int i, const float4& r) {
    const float4 v = IN16 ? gn_cvt4(tile16[i]) : tile[i];
    float4 o = make_float4(fmaf(v.x, a[0], b[0]), fmaf(v.y, a[1], b[1]), fmaf(v.z, a[2], b[2]), fmaf(v.w, a[3], b[3]));
    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    gn_st4<OUT16>(y, base4 + i, o);
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  i = tid;
  if (IN16 && residual) {  // residual loads batched like the loads of the first pass
    constexpr int B = 8;
    for (; i + (B - 1) * GNC_THREADS < n4; i += B * GNC_THREADS) {
      uint2 u[B];
#pragma unroll
      for (int k = 0; k < B; ++k) u[k] = __ldg(reinterpret_cast<const uint2*>(residual) + base4 + i + k * GNC_THREADS);
#pragma unroll
      for (int k = 0; k < B; ++k) emit(i + k * GNC_THREADS, gn_cvt4(u[k]));
    }
  }
  for (; i < n4; i += GNC_THREADS) emit(i, residual ? gn_ld4<IN16>(residual, base4 + i) : zero4);
}


// ------------------------------------------------------------------------------------------- backward, one pass
// Forward was y = act(gn(x) * gamma + beta (+ residual)).  With g = dy * (y > 0) (ReLU) — also the gradient of the
// residual branch — dx = rstd * (g*gamma - mean_grp(g*gamma) - xhat * mean_grp(g*gamma*xhat)).
// Same decomposition as the forward kernel: a cluster per sample, every CTA stages its slice of x AND of g in shared
// memory while accumulating the per-channel sums (x, x^2, g, g*x), the per-group sums cross CTAs through distributed
// shared memory in rank order, and dx is produced from shared memory: x, dy, y are read ONCE and dx (+ dres) written
// once (the two-pass kernel of nn_bwd.cu reads the three tensors twice).  dgamma / dbeta: per-sample partial rows
// (plain stores), summed over samples in order by gn_param_reduce_kernel — deterministic, no atomics.
__global__ void __launch_bounds__(GNC_THREADS) gn_cluster_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                     const float* __restrict__ dy,
                                                                     const float* __restrict__ gamma, float* dx, float* dres,
                                                                     float* pgamma, float* pbeta, int HW, int C, int groups,
                                                                     float eps, int relu, int cl, int pix_per_cta) {
  extern __shared__ __align__(16) unsigned char gsm[];
  __shared__ double part[64][4];    // per group: S, Q, sum gamma*G, sum gamma*GX   (this CTA)
  __shared__ float g_mean[64], g_rstd[64], g_m1[64], g_m2[64];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  const int n = blockIdx.x / cl;
  const int nq = C >> 2;
  const int p0 = rank * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  const int n4 = max(0, p1 - p0) * nq;
  const size_t base4 = ((size_t)n * HW + p0) * nq;
  float4* tx = reinterpret_cast<float4*>(gsm);
  float4* tg = tx + (size_t)pix_per_cta * nq;
  // reduction scratch behind the two slices, sized by C (as static arrays for C <= 512 they cost 24 KB and the third
  // CTA of an SM): red[k][slot * C + c] per (slot, channel) partials, summed in a fixed order; ch[k][c] per-channel
  // totals of this CTA (k: sum x, sum x^2, sum g, sum g*x)
  const int n_slots = nq < 32 ? GNC_THREADS / 32 : GNC_THREADS / nq;
  const int RS = n_slots * C;
  float* red = reinterpret_cast<float*>(tg + (size_t)pix_per_cta * nq);
  float* ch = red + 4 * RS;
  float a[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[k][j] = 0.f;
  // two iterations per trip, all six 16-byte loads issued before the first use: the kernel is latency-bound at 16
  // warps per SM (profiles/r02_gn_bwd_l1_ncu_full.txt: long-scoreboard stalls 9.2 per issue with one iteration in flight)
  auto absorb = [&](int i, const float4& xv, float4 gv, const float4& yv) {
    if (relu) {
      if (!(yv.x > 0.f)) gv.x = 0.f;
      if (!(yv.y > 0.f)) gv.y = 0.f;
      if (!(yv.z > 0.f)) gv.z = 0.f;
      if (!(yv.w > 0.f)) gv.w = 0.f;
    }
    tx[i] = xv;
    tg[i] = gv;
    if (dres) reinterpret_cast<float4*>(dres)[base4 + i] = gv;
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[0][j] += xs[j];
      a[1][j] = fmaf(xs[j], xs[j], a[1][j]);
      a[2][j] += gs[j];
      a[3][j] = fmaf(gs[j], xs[j], a[3][j]);
    }
  };
  const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
  int i = tid;
  for (; i + GNC_THREADS < n4; i += 2 * GNC_THREADS) {
    const int i2 = i + GNC_THREADS;
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(x) + base4 + i);
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(x) + base4 + i2);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy) + base4 + i);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(dy) + base4 + i2);
    const float4 y0 = relu ? __ldg(reinterpret_cast<const float4*>(y) + base4 + i) : one4;
    const float4 y1 = relu ? __ldg(reinterpret_cast<const float4*>(y) + base4 + i2) : one4;
    absorb(i, x0, g0, y0);
    absorb(i2, x1, g1, y1);
  }
  for (; i < n4; i += GNC_THREADS) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(x) + base4 + i);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy) + base4 + i);
    const float4 y0 = relu ? __ldg(reinterpret_cast<const float4*>(y) + base4 + i) : one4;
    absorb(i, x0, g0, y0);
  }
  for (int o = 16; o >= nq && o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[k][j] += __shfl_xor_sync(0xffffffffu, a[k][j], o);
  const int lane = tid & 31;
  const int c0 = (tid % nq) * 4;
  if (nq >= 32 || lane < nq) {
    const int slot = nq < 32 ? (tid >> 5) : tid / nq;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[k * RS + (slot * nq + (tid % nq)) * 4 + j] = a[k][j];
  }
  __syncthreads();
  for (int c = tid; c < C; c += GNC_THREADS) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float S = 0.f;
      for (int sl = 0; sl < n_slots; ++sl) S += red[k * RS + (sl * nq + (c >> 2)) * 4 + (c & 3)];
      ch[k * C + c] = S;
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  if (tid < groups) {
    double S = 0.0, Q = 0.0, P1 = 0.0, P2 = 0.0;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) {
      const double ga = (double)__ldg(gamma + c);
      S += (double)ch[c];
      Q += (double)ch[C + c];
      P1 += ga * (double)ch[2 * C + c];
      P2 += ga * (double)ch[3 * C + c];
    }
    part[tid][0] = S; part[tid][1] = Q; part[tid][2] = P1; part[tid][3] = P2;
  }
  cluster.sync();
  if (tid < groups) {
    double S = 0.0, Q = 0.0, P1 = 0.0, P2 = 0.0;
    for (int r = 0; r < cl; ++r) {
      const double* rp = cluster.map_shared_rank(&part[0][0], r);
      S += rp[tid * 4]; Q += rp[tid * 4 + 1]; P1 += rp[tid * 4 + 2]; P2 += rp[tid * 4 + 3];
    }
    const double cnt = (double)HW * cpg;
    const double m = S / cnt;
    double var = Q / cnt - m * m;
    if (var < 0.0) var = 0.0;
    const double rs = 1.0 / sqrt(var + (double)eps);
    g_mean[tid] = (float)m;
    g_rstd[tid] = (float)rs;
    g_m1[tid] = (float)(P1 / cnt);
    g_m2[tid] = (float)(rs * (P2 - m * P1) / cnt);
  }
  __syncthreads();
  if (rank == 0 && (pgamma || pbeta)) {  // per-sample parameter-gradient rows: channel sums over the whole cluster
    for (int c = tid; c < C; c += GNC_THREADS) {
      float G = 0.f, GX = 0.f;
      for (int r = 0; r < cl; ++r) {
        const float* rc = cluster.map_shared_rank(ch, r);
        G += rc[2 * C + c];
        GX += rc[3 * C + c];
      }
      const int g = c / cpg;
      if (pgamma) pgamma[(size_t)n * C + c] = g_rstd[g] * (GX - g_mean[g] * G);
      if (pbeta) pbeta[(size_t)n * C + c] = G;
    }
  }
  cluster.sync();  // no CTA may retire (or overwrite ch / part) while a peer still reads them
  if (!dx) return;
  float ga[4], mu[4], rs[4], m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j, g = c / cpg;
    ga[j] = __ldg(gamma + c);
    mu[j] = g_mean[g]; rs[j] = g_rstd[g]; m1[j] = g_m1[g]; m2[j] = g_m2[g];
  }
  for (int i = tid; i < n4; i += GNC_THREADS) {
    const float4 xv = tx[i], gv = tg[i];
    float4 o;
    o.x = rs[0] * (gv.x * ga[0] - m1[0] - (xv.x - mu[0]) * rs[0] * m2[0]);
    o.y = rs[1] * (gv.y * ga[1] - m1[1] - (xv.y - mu[1]) * rs[1] * m2[1]);
    o.z = rs[2] * (gv.z * ga[2] - m1[2] - (xv.z - mu[2]) * rs[2] * m2[2]);
    o.w = rs[3] * (gv.w * ga[3] - m1[3] - (xv.w - mu[3]) * rs[3] * m2[3]);
    reinterpret_cast<float4*>(dx)[base4 + i] = o;
  }
}

// out[c] (+)= sum over the N per-sample rows.  Block = 32 channels x 32 row lanes: lane j adds rows j, j + 32, ... and
// the 32 partial sums are added in lane order — the order depends on the launch shape only (deterministic).
__global__ void __launch_bounds__(1024) gn_param_reduce_kernel(const float* __restrict__ rows, int N, int C, float* out,
                                                                int accumulate) {
  __shared__ float sm[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (int n = threadIdx.y; n < N; n += 32) s += rows[(size_t)n * C + c];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

}  // namespace

// Returns AVL_ERR_UNSUPPORTED (nothing launched) for shapes outside this kernel: the caller falls back to the
// two-pass kernels.  x, y, residual must be 16-byte aligned.  in16: x and residual are fp16; out16: y is fp16.
int avl_groupnorm_cluster_typed(const void* x, int in16, const float* gamma, const float* beta, const void* residual,
                                void* y, int out16, int N, int HW, int C, int groups, float eps, int relu,
                                void* stream) {
  if (N < 0 || HW < 1 || C < 1 || groups < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !gamma || !beta || !y) return AVL_ERR_ARG;
  if ((C & 3) || C > 512 || groups > 64 || C % groups || GNC_THREADS % (C >> 2)) return AVL_ERR_UNSUPPORTED;
  if (((uintptr_t)x & 15) || ((uintptr_t)y & 15) || ((uintptr_t)residual & 15)) return AVL_ERR_UNSUPPORTED;
  const int esz = in16 ? 2 : 4;
  const long long sample_bytes = (long long)HW * C * esz;  // staged in shared memory in its storage type
  int cl = 1;
  while (cl < 8 && sample_bytes / cl > 32 * 1024) cl <<= 1;
  if (cl > HW) return AVL_ERR_UNSUPPORTED;
  const int pix_per_cta = avl_div_up(HW, cl);
  const size_t smem = (size_t)pix_per_cta * C * esz;
  if (smem > (size_t)GNC_MAX_SLICE || (long long)N * cl > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(gn_cluster_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GNC_MAX_SLICE));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(gn_cluster_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GNC_MAX_SLICE));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(gn_cluster_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GNC_MAX_SLICE));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(gn_cluster_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GNC_MAX_SLICE));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(N * cl));
  cfg.blockDim = dim3(GNC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  unsigned nat = 1;
  avl_pdl_attr(at, &nat);
  cfg.attrs = at;
  cfg.numAttrs = nat;
  auto kern = in16 ? (out16 ? gn_cluster_kernel<true, true> : gn_cluster_kernel<true, false>)
                   : (out16 ? gn_cluster_kernel<false, true> : gn_cluster_kernel<false, false>);
  AVL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, x, gamma, beta, residual, y, HW, C, groups, eps, relu, cl, pix_per_cta));
  avl_count_launch();
  return AVL_OK;
}

// 1 if an fp16-stored activation tensor saturated since the last reset (synchronises the device).
AVL_API int avl_f16_overflow(int reset) {
  int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_f16_overflow, sizeof(int)) != cudaSuccess) return AVL_ERR_CUDA;
  if (reset && v) {
    const int z = 0;
    if (cudaMemcpyToSymbol(g_f16_overflow, &z, sizeof(int)) != cudaSuccess) return AVL_ERR_CUDA;
  }
  return v;
}

AVL_API int avl_groupnorm_fwd_cluster(const float* x, const float* gamma, const float* beta, const float* residual,
                                      float* y, int N, int HW, int C, int groups, float eps, int relu, void* stream) {
  return avl_groupnorm_cluster_typed(x, 0, gamma, beta, residual, y, 0, N, HW, C, groups, eps, relu, stream);
}

// Test entry of the fp16-storage variants (x / residual fp16; y fp16 when out16).
AVL_API int avl_groupnorm_fwd_cluster_f16(const void* x, const float* gamma, const float* beta, const void* residual,
                                          void* y, int out16, int N, int HW, int C, int groups, float eps, int relu,
                                          void* stream) {
  return avl_groupnorm_cluster_typed(x, 1, gamma, beta, residual, y, out16, N, HW, C, groups, eps, relu, stream);
}

// One-pass cluster GroupNorm backward (see gn_cluster_bwd_kernel).  dgamma / dbeta are ACCUMULATED INTO (+=), like
// avl_groupnorm_bwd; scratch: 2 * N * C floats.  -2 (nothing launched): shape not covered -> avl_groupnorm_bwd.
AVL_API int avl_groupnorm_bwd_cluster(const float* x, const float* y, const float* dy, const float* gamma, float* dx,
                                      float* dres, float* dgamma, float* dbeta, int N, int HW, int C, int groups, float eps,
                                      int relu, float* scratch, void* stream) {
  if (N < 0 || HW < 1 || C < 1 || groups < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !dy || !gamma || (relu && !y) || ((dgamma || dbeta) && !scratch)) return AVL_ERR_ARG;
  if ((C & 3) || C > 512 || groups > 64 || C % groups || GNC_THREADS % (C >> 2)) return AVL_ERR_UNSUPPORTED;
  if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15) || ((uintptr_t)y & 15) || ((uintptr_t)dx & 15) || ((uintptr_t)dres & 15))
    return AVL_ERR_UNSUPPORTED;
  const long long sample_bytes = (long long)HW * C * 4;
  const int nq = C >> 2;
  const int n_slots = nq < 32 ? GNC_THREADS / 32 : GNC_THREADS / nq;
  const size_t scratch_bytes = (size_t)4 * (n_slots + 1) * C * sizeof(float);  // reduction scratch (see the kernel)
  // cluster size: the smallest one whose CTAs (x slice + g slice + scratch) fit three to an SM — the kernel alternates
  // between a load phase and reduction / cluster barriers, and two CTAs per SM left the memory system idle too often
  int cl = 1;
  while (cl < 8 && 2 * sample_bytes / cl + (long long)scratch_bytes > 72 * 1024) cl <<= 1;
  if (cl > HW) return AVL_ERR_UNSUPPORTED;
  const int pix_per_cta = avl_div_up(HW, cl);
  const size_t slices = (size_t)pix_per_cta * C * 4 * 2;  // x slice + g slice
  if (slices > 2 * (size_t)GNC_MAX_SLICE || (long long)N * cl > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  const size_t smem = slices + scratch_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(gn_cluster_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2 * GNC_MAX_SLICE + 24 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(N * cl));
  cfg.blockDim = dim3(GNC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  float* pg = dgamma ? scratch : nullptr;
  float* pb = dbeta ? scratch + (size_t)N * C : nullptr;
  AVL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gn_cluster_bwd_kernel, x, y, dy, gamma, dx, dres, pg, pb, HW, C, groups, eps, relu,
                                    cl, pix_per_cta));
  avl_count_launch();
  if (dgamma) {
    gn_param_reduce_kernel<<<avl_div_up(C, 32), dim3(32, 32), 0, (cudaStream_t)stream>>>(pg, N, C, dgamma, 1);
    AVL_LAUNCH_CHECK();
  }
  if (dbeta) {
    gn_param_reduce_kernel<<<avl_div_up(C, 32), dim3(32, 32), 0, (cudaStream_t)stream>>>(pb, N, C, dbeta, 1);
    AVL_LAUNCH_CHECK();
  }
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
