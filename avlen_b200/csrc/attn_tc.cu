// Variable-length multi-head self-attention of the scene-memory transformer on the tensor cores
// (reference: ss_baselines/savi/models/smt_state_encoder.py:160-166 -> nn.TransformerEncoderLayer.self_attn,
// torch MultiheadAttention with key_padding_mask; here the masked slots are already compacted away).
//
// Q K^T and P V (forward) and the five products of the backward are warp-level TF32 MMAs (mma.sync m16n8k8) with
// every operand split as x = hi + lo (3xTF32: lo*hi + hi*lo + hi*hi, fp32 accumulate), which reproduces fp32
// matmul to ~1e-6 — the transformer is held to the reference's fp32 results (1e-3), so plain TF32 is not enough.
// Softmax stays in registers: the accumulator fragment of an m16n8 tile gives each lane two rows (g, g + 8) and two
// adjacent columns; row statistics are finished by two shuffles inside the quad.
//
// The accumulator fragment of S feeds the next product as its A operand WITHOUT a shuffle: an m16n8k8 A fragment holds
// columns (t, t + 4) of rows (g, g + 8), an accumulator holds columns (2t, 2t + 1).  The reduction index of an MMA can
// be permuted freely if both operands agree, so the B fragment of the second product is loaded from rows
// (2t, 2t + 1) instead of (t, t + 4).
//
// One CTA per (sample, head).  K / V (forward) or Q / K / V / dO (backward) of the head live in shared memory with a
// row pitch of 36 floats: both B-fragment access patterns ([row g][col t] and [row 2t][col g]) are conflict-free.
// The backward recomputes the probabilities from the saved log-sum-exp (no V^2 storage) and is deterministic: phase A
// owns 16 keys per warp (dK, dV), phase B owns 16 queries per warp (dQ); every output element is written once.
#include "common.cuh"

#ifndef AVL_HOST_EMUL
namespace {

constexpr int AT_HD = 32;
constexpr int AT_LD = 36;        // shared-memory row pitch in floats
constexpr int AT_WARPS = 8;
constexpr int AT_MAXV = 320;

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// x = hi + lo exactly; hi carries the 10 mantissa bits the tensor core reads, lo the rest (the core truncates lo itself)
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

struct Frag {  // an A fragment (16 x 8) in hi / lo form
  uint32_t h[4], l[4];
};
__device__ __forceinline__ void frag_from(Frag& f, float a0, float a1, float a2, float a3) {
  split(a0, f.h[0], f.l[0]);
  split(a1, f.h[1], f.l[1]);
  split(a2, f.h[2], f.l[2]);
  split(a3, f.h[3], f.l[3]);
}
// c += A * B at fp32 accuracy (small terms first)
__device__ __forceinline__ void mma3(float (&c)[4], const Frag& a, float b0, float b1) {
  uint32_t bh0, bl0, bh1, bl1;
  split(b0, bh0, bl0);
  split(b1, bh1, bl1);
  mma_tf32(c, a.l[0], a.l[1], a.l[2], a.l[3], bh0, bh1);
  mma_tf32(c, a.h[0], a.h[1], a.h[2], a.h[3], bl0, bl1);
  mma_tf32(c, a.h[0], a.h[1], a.h[2], a.h[3], bh0, bh1);
}

// A fragments of a 16 x 32 row block held in shared memory (pitch AT_LD): rows r0 + g, r0 + g + 8
__device__ __forceinline__ void load_block_frags(Frag (&f)[4], const float* base, int r0, int g, int t) {
  const float* p0 = base + (r0 + g) * AT_LD + t;
  const float* p1 = p0 + 8 * AT_LD;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) frag_from(f[ks], p0[8 * ks], p1[8 * ks], p0[8 * ks + 4], p1[8 * ks + 4]);
}

// c(16 x 8) += A(16 x 32) * Bs[rows n0 .. n0 + 7][32]^T   (B row = output column)
__device__ __forceinline__ void mma_rows_t(float (&c)[4], const Frag (&a)[4], const float* Bs, int n0, int g, int t) {
  const float* p = Bs + (n0 + g) * AT_LD + t;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) mma3(c, a[ks], p[8 * ks], p[8 * ks + 4]);
}

// acc(16 x 32) += P(16 x 8, accumulator fragment) * Bs[rows k0 .. k0 + 7][32]   (reduction index permuted, see header)
__device__ __forceinline__ void mma_acc_rows(float (&acc)[4][4], const float (&p)[4], const float* Bs, int k0, int g,
                                             int t) {
  Frag a;
  frag_from(a, p[0], p[2], p[1], p[3]);
  const float* r0 = Bs + (k0 + 2 * t) * AT_LD + g;
  const float* r1 = r0 + AT_LD;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) mma3(acc[nb], a, r0[8 * nb], r1[8 * nb]);
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// qkv: [R, 3 D] packed rows (q | k | v), heads of 32; off[b] .. off[b + 1] are sample b's rows.
__global__ void __launch_bounds__(AT_WARPS * 32, 3)
attn_self_fwd_tc_kernel(const float* __restrict__ qkv, const int* __restrict__ off, float* out, float* lse, int D,
                        float scale, int vcap) {
  AVL_DYN_SMEM(smem_raw);
  avl_pdl_wait();
  avl_pdl_trigger();
  const int b = blockIdx.x, h = blockIdx.y;
  const int r0 = off[b], V = min(off[b + 1] - r0, vcap);
  if (V <= 0) return;
  const int Vp = (V + 31) & ~31;  // key chunks of 32; the padding rows are zero (0 * garbage must not be NaN)
  float* Ks = reinterpret_cast<float*>(smem_raw);
  float* Vs = Ks + ((vcap + 31) & ~31) * AT_LD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ld = 3 * D;
  for (int u = threadIdx.x; u < Vp * 8; u += blockDim.x) {
    const int j = u >> 3, c = u & 7;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (j < V) {
      const float* row = qkv + (size_t)(r0 + j) * ld + h * AT_HD + 4 * c;
      kv = *reinterpret_cast<const float4*>(row + D);
      vv = *reinterpret_cast<const float4*>(row + 2 * D);
    }
    *reinterpret_cast<float4*>(Ks + j * AT_LD + 4 * c) = kv;
    *reinterpret_cast<float4*>(Vs + j * AT_LD + 4 * c) = vv;
  }
  __syncthreads();
  // gridDim.z CTAs share a (sample, head): CTA z takes the query blocks z, z + gridDim.z, ... (each CTA stages all keys)
  for (int qb = blockIdx.z + gridDim.z * warp; qb * 16 < V; qb += gridDim.z * AT_WARPS) {
    const int i0 = qb * 16 + g, i1 = i0 + 8;
    Frag q[4];
    {
      const float* p0 = qkv + (size_t)(r0 + min(i0, V - 1)) * ld + h * AT_HD + t;
      const float* p1 = qkv + (size_t)(r0 + min(i1, V - 1)) * ld + h * AT_HD + t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        frag_from(q[ks], __ldg(p0 + 8 * ks) * scale, __ldg(p1 + 8 * ks) * scale, __ldg(p0 + 8 * ks + 4) * scale,
                  __ldg(p1 + 8 * ks + 4) * scale);
    }
    float o[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kc = 0; kc < Vp; kc += 32) {
      float s[4][4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
        mma_rows_t(s[nb], q, Ks, kc + 8 * nb, g, t);
      }
      float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const int key = kc + 8 * nb + 2 * t;
        if (key >= V) s[nb][0] = s[nb][2] = -INFINITY;
        if (key + 1 >= V) s[nb][1] = s[nb][3] = -INFINITY;
        x0 = fmaxf(x0, fmaxf(s[nb][0], s[nb][1]));
        x1 = fmaxf(x1, fmaxf(s[nb][2], s[nb][3]));
      }
      const float n0 = fmaxf(m0, quad_max(x0)), n1 = fmaxf(m1, quad_max(x1));  // finite: every chunk holds a valid key
      const float al0 = __expf(m0 - n0), al1 = __expf(m1 - n1);                // 0 on the first chunk
      m0 = n0;
      m1 = n1;
      l0 *= al0;
      l1 *= al1;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        o[nb][0] *= al0; o[nb][1] *= al0; o[nb][2] *= al1; o[nb][3] *= al1;
        s[nb][0] = __expf(s[nb][0] - n0); s[nb][1] = __expf(s[nb][1] - n0);
        s[nb][2] = __expf(s[nb][2] - n1); s[nb][3] = __expf(s[nb][3] - n1);
        l0 += s[nb][0] + s[nb][1];
        l1 += s[nb][2] + s[nb][3];
      }
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) mma_acc_rows(o, s[kb], Vs, kc + 8 * kb, g, t);
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int H = D / AT_HD;
    if (i0 < V) {
      float* orow = out + (size_t)(r0 + i0) * D + h * AT_HD + 2 * t;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<float2*>(orow + 8 * nb) = make_float2(o[nb][0] * inv0, o[nb][1] * inv0);
      if (lse && t == 0) lse[(size_t)(r0 + i0) * H + h] = m0 + __logf(l0);
    }
    if (i1 < V) {
      float* orow = out + (size_t)(r0 + i1) * D + h * AT_HD + 2 * t;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<float2*>(orow + 8 * nb) = make_float2(o[nb][2] * inv1, o[nb][3] * inv1);
      if (lse && t == 0) lse[(size_t)(r0 + i1) * H + h] = m1 + __logf(l1);
    }
  }
}

// Backward with recomputation.  dqkv receives (dq | dk | dv) rows.
__global__ void __launch_bounds__(AT_WARPS * 32, 2)
attn_self_bwd_tc_kernel(const float* __restrict__ qkv, const int* __restrict__ off, const float* __restrict__ out,
                        const float* __restrict__ lse, const float* __restrict__ dout, float* dqkv, int D, float scale,
                        int vcap) {
  AVL_DYN_SMEM(smem_raw);
  const int b = blockIdx.x, h = blockIdx.y;
  const int r0 = off[b], V = min(off[b + 1] - r0, vcap);
  if (V <= 0) return;
  const int Vp = (V + 15) & ~15;
  const int cap = (vcap + 15) & ~15;
  float* Qs = reinterpret_cast<float*>(smem_raw);  // pre-scaled
  float* Ks = Qs + cap * AT_LD;
  float* Vs = Ks + cap * AT_LD;
  float* Gs = Vs + cap * AT_LD;                    // dO
  float* Ls = Gs + cap * AT_LD;                    // lse   (+inf on padding rows: their probabilities are 0)
  float* Ds = Ls + cap;                            // D_i = dO_i . O_i
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ld = 3 * D, H = D / AT_HD;
  for (int u = threadIdx.x; u < Vp * 8; u += blockDim.x) {  // Vp * 8 is a multiple of 32: warp-uniform trip count
    const int j = u >> 3, c = u & 7;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv, gv = qv;
    float dsum = 0.f;
    if (j < V) {
      const float* row = qkv + (size_t)(r0 + j) * ld + h * AT_HD + 4 * c;
      qv = *reinterpret_cast<const float4*>(row);
      qv.x *= scale; qv.y *= scale; qv.z *= scale; qv.w *= scale;
      kv = *reinterpret_cast<const float4*>(row + D);
      vv = *reinterpret_cast<const float4*>(row + 2 * D);
      gv = *reinterpret_cast<const float4*>(dout + (size_t)(r0 + j) * D + h * AT_HD + 4 * c);
      const float4 ov = *reinterpret_cast<const float4*>(out + (size_t)(r0 + j) * D + h * AT_HD + 4 * c);
      dsum = gv.x * ov.x + gv.y * ov.y + gv.z * ov.z + gv.w * ov.w;
    }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 4);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    *reinterpret_cast<float4*>(Qs + j * AT_LD + 4 * c) = qv;
    *reinterpret_cast<float4*>(Ks + j * AT_LD + 4 * c) = kv;
    *reinterpret_cast<float4*>(Vs + j * AT_LD + 4 * c) = vv;
    *reinterpret_cast<float4*>(Gs + j * AT_LD + 4 * c) = gv;
    if (c == 0) {
      Ds[j] = dsum;
      Ls[j] = j < V ? lse[(size_t)(r0 + j) * H + h] : INFINITY;
    }
  }
  __syncthreads();
  const int nb16 = Vp >> 4;
  for (int u = warp; u < 2 * nb16; u += AT_WARPS) {
    if (u < nb16) {
      // ---- owner = 16 keys: dK_j = sum_i dS_ij Qs_i (Qs carries the scale), dV_j = sum_i P_ij dO_i
      const int j0 = u * 16;
      Frag kf[4], vf[4];
      load_block_frags(kf, Ks, j0, g, t);
      load_block_frags(vf, Vs, j0, g, t);
      float dk[4][4], dv[4][4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        dk[nb][0] = dk[nb][1] = dk[nb][2] = dk[nb][3] = 0.f;
        dv[nb][0] = dv[nb][1] = dv[nb][2] = dv[nb][3] = 0.f;
      }
      for (int i0 = 0; i0 < Vp; i0 += 8) {
        float st[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_rows_t(st, kf, Qs, i0, g, t);   // S^T: rows = keys, columns = queries i0 + 2t, i0 + 2t + 1
        mma_rows_t(dp, vf, Gs, i0, g, t);   // dP^T = V dO^T
        const float2 li = *reinterpret_cast<const float2*>(Ls + i0 + 2 * t);
        const float2 di = *reinterpret_cast<const float2*>(Ds + i0 + 2 * t);
        float p[4], ds[4];
        p[0] = __expf(st[0] - li.x); p[1] = __expf(st[1] - li.y);
        p[2] = __expf(st[2] - li.x); p[3] = __expf(st[3] - li.y);
        ds[0] = p[0] * (dp[0] - di.x); ds[1] = p[1] * (dp[1] - di.y);
        ds[2] = p[2] * (dp[2] - di.x); ds[3] = p[3] * (dp[3] - di.y);
        mma_acc_rows(dv, p, Gs, i0, g, t);
        mma_acc_rows(dk, ds, Qs, i0, g, t);
      }
      const int ja = j0 + g, jb = ja + 8;
      if (ja < V) {
        float* krow = dqkv + (size_t)(r0 + ja) * ld + D + h * AT_HD + 2 * t;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          *reinterpret_cast<float2*>(krow + 8 * nb) = make_float2(dk[nb][0], dk[nb][1]);
          *reinterpret_cast<float2*>(krow + D + 8 * nb) = make_float2(dv[nb][0], dv[nb][1]);
        }
      }
      if (jb < V) {
        float* krow = dqkv + (size_t)(r0 + jb) * ld + D + h * AT_HD + 2 * t;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          *reinterpret_cast<float2*>(krow + 8 * nb) = make_float2(dk[nb][2], dk[nb][3]);
          *reinterpret_cast<float2*>(krow + D + 8 * nb) = make_float2(dv[nb][2], dv[nb][3]);
        }
      }
    } else {
      // ---- owner = 16 queries: dQ_i = scale * sum_j dS_ij K_j
      const int i0 = (u - nb16) * 16;
      Frag qf[4], gf[4];
      load_block_frags(qf, Qs, i0, g, t);
      load_block_frags(gf, Gs, i0, g, t);
      const float la = Ls[i0 + g], lb = Ls[i0 + g + 8], da = Ds[i0 + g], db = Ds[i0 + g + 8];
      float dq[4][4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) dq[nb][0] = dq[nb][1] = dq[nb][2] = dq[nb][3] = 0.f;
      for (int j0 = 0; j0 < Vp; j0 += 8) {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_rows_t(s, qf, Ks, j0, g, t);
        mma_rows_t(dp, gf, Vs, j0, g, t);
        const bool k0 = j0 + 2 * t < V, k1 = j0 + 2 * t + 1 < V;  // zero rows of K score 0, not -inf: mask the padding
        float ds[4];
        ds[0] = k0 ? __expf(s[0] - la) * (dp[0] - da) : 0.f;
        ds[1] = k1 ? __expf(s[1] - la) * (dp[1] - da) : 0.f;
        ds[2] = k0 ? __expf(s[2] - lb) * (dp[2] - db) : 0.f;
        ds[3] = k1 ? __expf(s[3] - lb) * (dp[3] - db) : 0.f;
        mma_acc_rows(dq, ds, Ks, j0, g, t);
      }
      const int ia = i0 + g, ib = ia + 8;
      if (ia < V) {
        float* qrow = dqkv + (size_t)(r0 + ia) * ld + h * AT_HD + 2 * t;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<float2*>(qrow + 8 * nb) = make_float2(dq[nb][0] * scale, dq[nb][1] * scale);
      }
      if (ib < V) {
        float* qrow = dqkv + (size_t)(r0 + ib) * ld + h * AT_HD + 2 * t;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<float2*>(qrow + 8 * nb) = make_float2(dq[nb][2] * scale, dq[nb][3] * scale);
      }
    }
  }
}

size_t fwd_smem(int vcap) { return (size_t)2 * ((vcap + 31) & ~31) * AT_LD * sizeof(float); }
size_t bwd_smem(int vcap) { return ((size_t)4 * ((vcap + 15) & ~15) * AT_LD + 2 * ((vcap + 15) & ~15)) * sizeof(float); }

static int g_attn_tc = 1;
static int g_attn_qsplit = 0;
static bool g_attr_set = false;

int ensure_attrs() {
  if (g_attr_set) return AVL_OK;
  AVL_CUDA_CHECK(cudaFuncSetAttribute(attn_self_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem(AT_MAXV)));
  AVL_CUDA_CHECK(cudaFuncSetAttribute(attn_self_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem(AT_MAXV)));
  g_attr_set = true;
  return AVL_OK;
}

}  // namespace

AVL_API int avl_get_tensor_cores(void);

// CTAs per (sample, head) of the forward kernel: 0 (default) automatic, 1 .. 4 forced (diagnostic).  Returns old.
AVL_API int avl_set_attn_qsplit(int n) {
  int old = g_attn_qsplit;
  g_attn_qsplit = n < 0 ? 0 : (n > 4 ? 4 : n);
  avl_bump_config_epoch();
  return old;
}

// 1 (default): self-attention runs on the tensor cores (3xTF32 warp MMAs) whenever the tensor-core level is >= 1;
// 0: the register-tiled fp32 kernels.  Returns the old value.
AVL_API int avl_set_attn_tc(int on) {
  int old = g_attn_tc;
  g_attn_tc = on ? 1 : 0;
  return old;
}

// Internal (smt.cu; C linkage, hidden visibility): AVL_ERR_UNSUPPORTED = use the fp32 SIMT kernel.
extern "C" int avl_attn_self_fwd_tc_try(const float* qkv, const int* off, int B, int D, float* out, float* lse, float scale, int vcap,
                             cudaStream_t stream) {
  if (!g_attn_tc || avl_get_tensor_cores() < 1 || vcap > AT_MAXV || (D & 31) || (((uintptr_t)qkv | (uintptr_t)out) & 15))
    return AVL_ERR_UNSUPPORTED;
  int rc = ensure_attrs();
  if (rc) return rc;
  // rollout batches: B * heads CTAs are little more than one wave at three CTAs per SM, and 10 query blocks over 8 warps
  // leave a CTA as long as its two-block warps — with the query blocks dealt out over two CTAs every warp has one block
  int nz = g_attn_qsplit;
  if (nz <= 0) nz = ((long long)B * (D / AT_HD) <= 1024 && vcap > 16 * AT_WARPS) ? 2 : 1;
  AVL_LAUNCH_PDL(attn_self_fwd_tc_kernel, dim3(B, D / AT_HD, nz), AT_WARPS * 32, fwd_smem(vcap), stream, qkv, off, out, lse, D,
                 scale, vcap);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

extern "C" int avl_attn_self_bwd_tc_try(const float* qkv, const int* off, int B, int D, const float* out, const float* lse,
                             const float* dout, float* dqkv, float scale, int vcap, cudaStream_t stream) {
  if (!g_attn_tc || avl_get_tensor_cores() < 1 || vcap > AT_MAXV || (D & 31) ||
      (((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dqkv) & 15))
    return AVL_ERR_UNSUPPORTED;
  int rc = ensure_attrs();
  if (rc) return rc;
  attn_self_bwd_tc_kernel<<<dim3(B, D / AT_HD), AT_WARPS * 32, bwd_smem(vcap), stream>>>(qkv, off, out, lse, dout, dqkv, D,
                                                                                         scale, vcap);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
