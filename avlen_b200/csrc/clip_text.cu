// Row L: the CLIP ViT-B/32 TEXT tower as AVLEN calls it (ss_baselines/savi/ppo/policy.py:761-762, :844-851:
// `clip.load("ViT-B/32")`, `self.clip.encode_text(all_dialog).float()` under no_grad; frozen,
// ddppo_trainer.py:401-403).  openai/CLIP (unpinned git dependency, README.md:61) model.py `encode_text`:
//   x = token_embedding[tokens] + positional_embedding                                  (B, 77, 512)
//   layers (12) x { x += out_proj(causal MHA_8x64(ln_1(x))) ; x += c_proj(QuickGELU(c_fc(ln_2(x)))) }
//   x = ln_final(x)[b, argmax(tokens[b])] @ text_projection                             (B, 512)
//
// B200-first restructuring (identical results): the dialog tensor is all-zero for every env that has no active
// query (ppo_trainer.py:625-637 feeds all N rows every step), and every all-zero row encodes to the same vector.
// Rows are therefore compacted on the device: the distinct work is (number of non-zero rows + 1 shared zero row);
// all kernels are launched for the worst case and exit early on a device-side count — no host synchronisation.
// Dense layers run on the tcgen05 TF32 GEMM (gemm_tc.cu) when tensor cores are enabled (the reference runs this
// tower in fp16 on CUDA), fp32 SIMT otherwise.
#include "nn_kernels.cuh"
#ifndef AVL_HOST_EMUL
#include <cuda_fp16.h>
#endif

namespace {

constexpr int CL_W = 512, CL_HEADS = 8, CL_HD = 64, CL_FF = 2048, CL_MAXL = 128, CL_MAX_LAYERS = 48;
enum { CP_TOK = 0, CP_POS, CP_LAYER0 };
enum { CL_LN1_W = 0, CL_LN1_B, CL_IN_W, CL_IN_B, CL_OUT_W, CL_OUT_B, CL_LN2_W, CL_LN2_B, CL_FC_W, CL_FC_B, CL_PROJ_W,
       CL_PROJ_B, CL_PER_LAYER };
// after the per-layer entries: ln_final.weight, ln_final.bias, text_projection
static inline int cp_lnf_w(int layers) { return CP_LAYER0 + layers * CL_PER_LAYER; }

// slot[b] = packed index of sample b (non-zero rows first, in order); all-zero rows share slot n_active.
// counts[0] = n_active + (any zero row ? 1 : 0) distinct sequences, counts[1] = that * L rows, eot[slot] = argmax.
__global__ void clip_compact_kernel(const long long* __restrict__ tokens, int B, int L, int dedupe, int* slot,
                                    int* src_of_slot, int* eot, int* counts) {
  __shared__ int buf[1024];
  __shared__ int carry, any_zero, first_zero;
  if (threadIdx.x == 0) { carry = 0; any_zero = 0; first_zero = -1; }
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int b = base + threadIdx.x;
    int nz = 0, am = 0;
    if (b < B) {
      long long best = tokens[(size_t)b * L];
      for (int t = 0; t < L; ++t) {
        long long v = tokens[(size_t)b * L + t];
        if (v != 0) nz = 1;
        if (v > best) { best = v; am = t; }
      }
      if (!dedupe) nz = 1;
    }
    buf[threadIdx.x] = nz;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      int t = ((int)threadIdx.x >= d) ? buf[threadIdx.x - d] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (b < B) {
      if (nz) {
        int s = carry + buf[threadIdx.x] - 1;
        slot[b] = s;
        src_of_slot[s] = b;
        eot[s] = am;
      } else {
        slot[b] = -1;  // patched below
        atomicExch(&any_zero, 1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  const int n_active = carry;
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    if (slot[b] < 0) {
      slot[b] = n_active;
      src_of_slot[n_active] = b;  // any all-zero row (they are identical); racing writers store equivalent rows
      eot[n_active] = 0;
    }
  if (threadIdx.x == 0) {
    int n = n_active + (any_zero ? 1 : 0);
    counts[0] = n;
    counts[1] = n * L;
  }
}

__global__ void clip_embed_kernel(const long long* __restrict__ tokens, const int* __restrict__ src_of_slot,
                                  const int* __restrict__ counts, const float* __restrict__ tok_emb,
                                  const float* __restrict__ pos_emb, int L, int vocab, float* x) {
  const int s = blockIdx.x / L, t = blockIdx.x % L;
  if (s >= counts[0]) return;
  long long id = tokens[(size_t)src_of_slot[s] * L + t];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const float* e = tok_emb + (size_t)id * CL_W;
  const float* p = pos_emb + (size_t)t * CL_W;
  float* o = x + (size_t)blockIdx.x * CL_W;
  for (int c = threadIdx.x; c < CL_W; c += blockDim.x) o[c] = e[c] + p[c];
}

// causal multi-head attention, head dim 64; qkv rows [q | k | v] of width 3*512; gridDim.z CTAs per (sequence, head)
template <bool OUT16>
__global__ void __launch_bounds__(256) clip_attn_kernel(const float* __restrict__ qkv, const int* __restrict__ counts,
                                                        int L, void* out_) {
  AVL_DYN_SMEM(smem_raw);
  avl_pdl_wait();
  avl_pdl_trigger();
  const int s = blockIdx.x, h = blockIdx.y;
  if (s >= counts[0]) return;
  float* Ks = reinterpret_cast<float*>(smem_raw);  // [L][65]
  float* Vs = Ks + L * 65;                         // [L][65]
  float* Ps = Vs + L * 65;                         // [8][CL_MAXL]
  float* Qs = Ps + 8 * CL_MAXL;                    // [8][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = 3 * CL_W;
  const float* base = qkv + (size_t)s * L * ld + h * CL_HD;
  for (int i = threadIdx.x; i < L * CL_HD; i += blockDim.x) {
    int j = i >> 6, d = i & 63;
    Ks[j * 65 + d] = base[(size_t)j * ld + CL_W + d];
    Vs[j * 65 + d] = base[(size_t)j * ld + 2 * CL_W + d];
  }
  __syncthreads();
  float* ps = Ps + warp * CL_MAXL;
  float* qs = Qs + warp * CL_HD;
  // gridDim.z CTAs share one (sequence, head): queries are dealt out round-robin over (CTA, warp) — at rollout batch a
  // handful of sequences is live and the ~10 queries a warp walked one after the other were the kernel's latency
  for (int i = warp + 8 * blockIdx.z; i < L; i += 8 * gridDim.z) {
    qs[lane] = base[(size_t)i * ld + lane] * 0.125f;
    qs[lane + 32] = base[(size_t)i * ld + lane + 32] * 0.125f;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) {
      float a = 0.f;
#pragma unroll 16
      for (int d = 0; d < CL_HD; ++d) a = fmaf(qs[d], Ks[j * 65 + d], a);
      ps[j] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j <= i; j += 32) {
      float p = __expf(ps[j] - mx);
      ps[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= i; ++j) {
      float p = ps[j];
      o0 = fmaf(p, Vs[j * 65 + lane], o0);
      o1 = fmaf(p, Vs[j * 65 + lane + 32], o1);
    }
    const float inv = 1.f / sum;
#ifndef AVL_HOST_EMUL
    if (OUT16) {  // (fp16 operand of the out-projection GEMM)
      __half* o = reinterpret_cast<__half*>(out_) + ((size_t)s * L + i) * CL_W + h * CL_HD;
      o[lane] = __float2half_rn(o0 * inv);
      o[lane + 32] = __float2half_rn(o1 * inv);
    } else
#endif
    {
      float* o = reinterpret_cast<float*>(out_) + ((size_t)s * L + i) * CL_W + h * CL_HD;
      o[lane] = o0 * inv;
      o[lane + 32] = o1 * inv;
    }
    __syncwarp();
  }
}

// x * sigmoid(1.702 x) in place (openai/CLIP QuickGELU)
__global__ void quickgelu_kernel(float* x, const int* __restrict__ rows_dev, long long rows_max, int cols) {
  avl_pdl_wait();
  avl_pdl_trigger();
  const long long n = min((long long)*rows_dev, rows_max) * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    x[i] = v / (1.f + __expf(-1.702f * v));
  }
}

#ifndef AVL_HOST_EMUL
// LayerNorm over 512 columns, warp per row, fp16 output (the A operand of the next fp16 GEMM); fp32 statistics
__global__ void __launch_bounds__(256) clip_ln16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __half* y,
                                                        const int* __restrict__ rows_dev, int rows_max) {
  avl_pdl_wait();
  avl_pdl_trigger();
  const int rows = min(*rows_dev, rows_max);
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * CL_W);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / CL_W);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / CL_W) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    const __half2 lo = __floats2half2_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
    const __half2 hi = __floats2half2_rn((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&lo);
    u.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y + (size_t)row * CL_W + 4 * c4) = u;
  }
}
#endif

// xe[s, :] = x[s*L + eot[s], :]
__global__ void clip_take_eot_kernel(const float* __restrict__ x, const int* __restrict__ eot,
                                     const int* __restrict__ counts, int L, float* xe) {
  const int s = blockIdx.x;
  if (s >= counts[0]) return;
  const float* src = x + ((size_t)s * L + eot[s]) * CL_W;
  for (int c = threadIdx.x; c < CL_W; c += blockDim.x) xe[(size_t)s * CL_W + c] = src[c];
}

// out[b, :] = emb[slot[b], :]
__global__ void clip_scatter_kernel(const float* __restrict__ emb, const int* __restrict__ slot, int B, int cols,
                                    float* out) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const float* src = emb + (size_t)slot[b] * cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) out[(size_t)b * cols + c] = src[c];
}

#ifndef AVL_HOST_EMUL
AVL_API int avl_tc_gemm(const float* A, long long lda, const float* B, float* C, long long ldc, int M, int N, int K,
                        const float* scale, const float* bias, const float* residual, long long ldr, int relu,
                        const int* m_dev, void* stream);
AVL_API int avl_get_tensor_cores(void);
static bool clip_tc_ok(int rows) { return avl_get_tensor_cores() >= 1 && rows >= 512; }
#else
static bool clip_tc_ok(int) { return false; }
static int clip_tc_stub(const float*, long long, const float*, float*, long long, int, int, int, const float*,
                        const float*, const float*, long long, int, const int*, void*) { return 0; }
#define avl_tc_gemm clip_tc_stub
#endif

struct ClipCtx {
  cudaStream_t s;
  int err = 0;
  void check() {
    avl_count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && !err) { avl_set_cuda_error((int)e); err = AVL_ERR_CUDA; }
  }
};

// Y[rows, N] = X[rows, K] W[N, K]^T + b (+ residual)
static void clip_lin(ClipCtx& c, const float* X, const float* W, const float* b, const float* residual, float* Y,
                     int rows, int N, int K, const int* rows_dev) {
  if (clip_tc_ok(rows)) {
    int rc = avl_tc_gemm(X, K, W, Y, N, rows, N, K, nullptr, b, residual, N, 0, rows_dev, c.s);
    if (rc && !c.err) c.err = rc;
    return;
  }
  GemmEpilogue ep;
  ep.bias = b; ep.scale = nullptr; ep.residual = residual; ep.ldr = N; ep.relu = 0; ep.accumulate = 0;
  ep.m_dev = rows_dev; ep.k_dev = nullptr;
  ConvGeom g = {};
  dim3 grid(avl_div_up(rows, GBM), avl_div_up(N, GBN), 1);
  auto kern = gemm_kernel<false, true, true>;
  GemmOperand A = {X, (long long)K, 1}, Bo = {W, (long long)K, 1};
  AVL_LAUNCH(kern, grid, GTHREADS, 0, c.s, A, Bo, Y, (long long)N, rows, N, K, g, ep, K);
  c.check();
}

static void clip_ln(ClipCtx& c, const float* x, const float* g, const float* b, float* y, const int* rows_dev, int rows) {
  AVL_LAUNCH_PDL(layernorm_fwd_kernel, avl_div_up(rows, 8), 256, 0, c.s, x, (const float*)nullptr, g, b, y, (float*)nullptr,
                 (float*)nullptr, rows_dev, rows, CL_W, 1e-5f);
  c.check();
}

struct ClipBufs {
  int *slot, *src, *eot, *counts;
  float *X, *XN, *QKV, *ATT, *H, *XE, *EMB;
  void *XN16, *ATT16, *H16;  // fp16 operands of the kind::f16 GEMMs (half the size of their fp32 twins)
};
static size_t clip_layout(char* base, ClipBufs& b, size_t B, size_t L) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    char* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  const size_t S = B + 1, R = S * L;
  b.slot = (int*)take(4 * S); b.src = (int*)take(4 * S); b.eot = (int*)take(4 * S); b.counts = (int*)take(16);
  b.X = (float*)take(4 * R * CL_W); b.XN = (float*)take(4 * R * CL_W); b.QKV = (float*)take(4 * R * 3 * CL_W);
  b.ATT = (float*)take(4 * R * CL_W); b.H = (float*)take(4 * R * CL_FF);
  b.XE = (float*)take(4 * S * CL_W); b.EMB = (float*)take(4 * S * CL_W);
  b.XN16 = take(2 * R * CL_W); b.ATT16 = take(2 * R * CL_W); b.H16 = take(2 * R * CL_FF);
  return off + 256;
}

}  // namespace

AVL_API int avl_clip_text_param_count(int layers) { return cp_lnf_w(layers) + 3; }

AVL_API long long avl_clip_text_workspace_bytes(int B, int L) {
  ClipBufs b;
  return (long long)clip_layout(nullptr, b, (size_t)B, (size_t)L);
}

// tokens (B, L) int64 (clip.tokenize layout: SOT, ids, EOT = largest id, zero padding; L <= 77 positions);
// params: avl_clip_text_param_count(layers) device pointers in the order of avlen_b200/savi/models/clip_text.py::CLIP_PARAM_KEYS;
// out (B, 512) fp32.  dedupe != 0: all-zero rows are encoded once (see header).  counts_out (optional, device,
// 2 ints): distinct sequences / rows actually processed.
#ifndef AVL_HOST_EMUL
extern "C" int avl_tc_gemm_tma_f16(const void* A, long long lda, const void* W, void* C, long long ldc, int M, int N, int K,
                                   const float* bias, const float* residual, long long ldr, int act, int out16,
                                   const int* m_dev, cudaStream_t stream);  // gemm_tma.cu
#endif
static int clip_forward(int B, int L, int vocab, int layers, const long long* tokens, const float* const* params,
                        const void* const* params16, float* out, void* workspace, int dedupe, void* stream);

AVL_API int avl_clip_text_forward(int B, int L, int vocab, int layers, const long long* tokens,
                                  const float* const* params, float* out, void* workspace, int dedupe, void* stream) {
  return clip_forward(B, L, vocab, layers, tokens, params, nullptr, out, workspace, dedupe, stream);
}

// The same tower with its four linears per layer on fp16 operands (tcgen05 kind::f16, fp32 accumulation, fp32 residual
// stream, LayerNorm / softmax statistics in fp32) — the dtype the reference itself runs CLIP in on CUDA (policy.py:761,
// `clip.load` converts the weights to fp16).  params16: 4 * layers device pointers to fp16 copies of
// (in_proj_weight, out_proj.weight, c_fc.weight, c_proj.weight) per layer; NULL = avl_clip_text_forward.
AVL_API int avl_clip_text_forward_f16(int B, int L, int vocab, int layers, const long long* tokens,
                                      const float* const* params, const void* const* params16, float* out, void* workspace,
                                      int dedupe, void* stream) {
  return clip_forward(B, L, vocab, layers, tokens, params, params16, out, workspace, dedupe, stream);
}

static int clip_forward(int B, int L, int vocab, int layers, const long long* tokens, const float* const* params,
                        const void* const* params16, float* out, void* workspace, int dedupe, void* stream) {
  if (B < 0 || L < 1 || L > CL_MAXL || vocab < 1 || layers < 1 || layers > CL_MAX_LAYERS) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!tokens || !params || !out || !workspace) return AVL_ERR_ARG;
  ClipBufs b;
  clip_layout(static_cast<char*>(workspace), b, (size_t)B, (size_t)L);
  ClipCtx c{(cudaStream_t)stream};
  const int S = B + 1, R = S * L;  // worst case: every row distinct (+ the unused shared slot)
  const size_t attn_smem = (size_t)(2 * L * 65 + 8 * CL_MAXL + 8 * CL_HD) * sizeof(float);
#ifndef AVL_HOST_EMUL
  static bool attr = false;
  if (!attr) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(clip_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(clip_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
#endif
  AVL_LAUNCH(clip_compact_kernel, 1, 1024, 0, c.s, tokens, B, L, dedupe, b.slot, b.src, b.eot, b.counts);
  c.check();
  const int* n_seq = b.counts;
  const int* n_rows = b.counts + 1;
  AVL_LAUNCH(clip_embed_kernel, S * L, 128, 0, c.s, tokens, b.src, n_seq, params[CP_TOK], params[CP_POS], L, vocab, b.X);
  c.check();
  int ew = avl_div_up((long long)R * CL_FF, 1024);
  int cap = avl_num_sms() * 16;
  if (ew > cap) ew = cap;
#ifndef AVL_HOST_EMUL
  const bool f16 = params16 != nullptr && clip_tc_ok(R);
  for (int l = 0; f16 && l < layers; ++l) {
    const float* const* P = params + CP_LAYER0 + l * CL_PER_LAYER;
    const void* const* W16 = params16 + 4 * l;
    cudaStream_t cs = (cudaStream_t)stream;
    int rc;
    AVL_LAUNCH_PDL(clip_ln16_kernel, avl_div_up(R, 8), 256, 0, cs, b.X, P[CL_LN1_W], P[CL_LN1_B], (__half*)b.XN16, n_rows, R);
    c.check();
    rc = avl_tc_gemm_tma_f16(b.XN16, CL_W, W16[0], b.QKV, 3 * CL_W, R, 3 * CL_W, CL_W, P[CL_IN_B], nullptr, 0, 0, 0, n_rows, cs);
    if (rc && !c.err) c.err = rc;
    AVL_LAUNCH_PDL(clip_attn_kernel<true>, dim3(S, CL_HEADS, S <= 64 ? 4 : 1), 256, attn_smem, cs, b.QKV, n_seq, L, b.ATT16);
    c.check();
    rc = avl_tc_gemm_tma_f16(b.ATT16, CL_W, W16[1], b.X, CL_W, R, CL_W, CL_W, P[CL_OUT_B], b.X, CL_W, 0, 0, n_rows, cs);  // x += out_proj
    if (rc && !c.err) c.err = rc;
    AVL_LAUNCH_PDL(clip_ln16_kernel, avl_div_up(R, 8), 256, 0, cs, b.X, P[CL_LN2_W], P[CL_LN2_B], (__half*)b.XN16, n_rows, R);
    c.check();
    // c_fc with QuickGELU in the epilogue, fp16 result = operand of c_proj (no separate activation pass)
    rc = avl_tc_gemm_tma_f16(b.XN16, CL_W, W16[2], b.H16, CL_FF, R, CL_FF, CL_W, P[CL_FC_B], nullptr, 0, 2, 1, n_rows, cs);
    if (rc && !c.err) c.err = rc;
    rc = avl_tc_gemm_tma_f16(b.H16, CL_FF, W16[3], b.X, CL_W, R, CL_W, CL_FF, P[CL_PROJ_B], b.X, CL_W, 0, 0, n_rows, cs);  // x += c_proj
    if (rc && !c.err) c.err = rc;
  }
#else
  const bool f16 = false;
#endif
  for (int l = 0; !f16 && l < layers; ++l) {
    const float* const* P = params + CP_LAYER0 + l * CL_PER_LAYER;
    clip_ln(c, b.X, P[CL_LN1_W], P[CL_LN1_B], b.XN, n_rows, R);
    clip_lin(c, b.XN, P[CL_IN_W], P[CL_IN_B], nullptr, b.QKV, R, 3 * CL_W, CL_W, n_rows);
    AVL_LAUNCH_PDL(clip_attn_kernel<false>, dim3(S, CL_HEADS, S <= 64 ? 4 : 1), 256, attn_smem, c.s, b.QKV, n_seq, L, (void*)b.ATT);
    c.check();
    clip_lin(c, b.ATT, P[CL_OUT_W], P[CL_OUT_B], b.X, b.X, R, CL_W, CL_W, n_rows);  // x += out_proj(att)
    clip_ln(c, b.X, P[CL_LN2_W], P[CL_LN2_B], b.XN, n_rows, R);
    clip_lin(c, b.XN, P[CL_FC_W], P[CL_FC_B], nullptr, b.H, R, CL_FF, CL_W, n_rows);
    AVL_LAUNCH_PDL(quickgelu_kernel, ew, 256, 0, c.s, b.H, n_rows, (long long)R, CL_FF);
    c.check();
    clip_lin(c, b.H, P[CL_PROJ_W], P[CL_PROJ_B], b.X, b.X, R, CL_W, CL_FF, n_rows);  // x += c_proj(h)
  }
  AVL_LAUNCH(clip_take_eot_kernel, S, 128, 0, c.s, b.X, b.eot, n_seq, L, b.XE);
  c.check();
  clip_ln(c, b.XE, params[cp_lnf_w(layers)], params[cp_lnf_w(layers) + 1], b.XN, n_seq, S);
  {  // EMB[s, n] = sum_k XN[s, k] * text_projection[k, n]
    GemmEpilogue ep;
    ep.bias = nullptr; ep.scale = nullptr; ep.residual = nullptr; ep.ldr = 0; ep.relu = 0; ep.accumulate = 0;
    ep.m_dev = n_seq; ep.k_dev = nullptr;
    ConvGeom g = {};
    dim3 grid(avl_div_up(S, GBM), avl_div_up(CL_W, GBN), 1);
    auto kern = gemm_kernel<false, true, false>;
    GemmOperand A = {b.XN, (long long)CL_W, 1}, Bo = {params[cp_lnf_w(layers) + 2], 1, (long long)CL_W};
    AVL_LAUNCH(kern, grid, GTHREADS, 0, c.s, A, Bo, b.EMB, (long long)CL_W, S, CL_W, CL_W, g, ep, CL_W);
    c.check();
  }
  AVL_LAUNCH(clip_scatter_kernel, B, 128, 0, c.s, b.EMB, b.slot, B, CL_W, out);
  c.check();
  return c.err;
}

// (synchronising) distinct sequences / rows processed by the last forward on this workspace
AVL_API int avl_clip_text_status(int B, int L, void* workspace, int* n_sequences /* host */, int* n_rows /* host */) {
  ClipBufs b;
  clip_layout(static_cast<char*>(workspace), b, (size_t)B, (size_t)L);
  int h[2] = {0, 0};
  AVL_CUDA_CHECK(cudaMemcpy(h, b.counts, sizeof(h), cudaMemcpyDeviceToHost));
  if (n_sequences) *n_sequences = h[0];
  if (n_rows) *n_rows = h[1];
  return AVL_OK;
}
