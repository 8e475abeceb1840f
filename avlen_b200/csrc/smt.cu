// Scene-memory-transformer state encoder forward + backward, and the generic
// dense ops it is built from (exported for the parity tests and reused by the other heads).
//
// Reference: ss_baselines/savi/models/smt_state_encoder.py:109-276 (single_forward, _encode_pose,
// _compute_relative_pose, _format_pose) around torch.nn.Transformer (1 encoder + 1 decoder layer,
// post-norm, ReLU FFN, final encoder/decoder LayerNorm, key-padding mask = (1 - mask) > 0).
//
// B200-first restructuring (results identical up to fp32 summation order):
//  * masked memory slots can never influence the output (they are ignored as keys in the encoder and
//    in the decoder), so every sample's valid slots are compacted into a packed token list and all
//    per-token work (pose encoding, fusion MLP, encoder layer, K/V projection) runs on
//    R = sum_b V_b rows instead of B*(M+1) rows;
//  * the memory is read straight from the single-copy ring buffer (total, N, dim) through a per-row
//    environment index, so the (M, T*N, dim) minibatch copies of recurrent_generator never exist;
//  * the row count R lives on the device (no host synchronisation): kernels are launched for the
//    worst case and exit early.
#include "nn_kernels.cuh"

namespace {

// ------------------------------------------------------------------------------ token compaction
__device__ __forceinline__ bool slot_valid(float m) { return !((1.f - m) > 0.f); }  // smt_state_encoder.py:107

__global__ void smt_count_kernel(const float* __restrict__ masks, int B, int M, int pretraining, int* cnt) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int c = 0;
  if (!pretraining)
    for (int s = lane; s < M; s += 32) c += slot_valid(masks[(size_t)b * M + s]) ? 1 : 0;
  c = (int)(warp_sum((float)c) + 0.5f);
  if (lane == 0) cnt[b] = c + 1;  // + current observation (always valid, :131)
}

// exclusive scan of cnt[0..B) -> off[0..B], total; single block of 1024 threads
__global__ void smt_scan_kernel(const int* __restrict__ cnt, int B, int* off, int* total, int rows_cap, int* err) {
  __shared__ int buf[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < B) ? cnt[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      int t = ((int)threadIdx.x >= d) ? buf[threadIdx.x - d] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < B) off[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    off[B] = carry;
    *total = carry;
    if (carry > rows_cap) *err = 1;
  }
}

// tok_slot[r] = memory slot of packed row r (M for the current observation), tok_sample[r] = b.
// Slots are emitted in increasing slot order followed by the current observation (the reference's
// concatenation order, :146).
__global__ void smt_fill_kernel(const float* __restrict__ masks, const int* __restrict__ off, int B, int M,
                                int pretraining, int rows_cap, int* tok_slot, int* tok_sample) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int pos = off[b];
  if (!pretraining) {
    for (int s0 = 0; s0 < M; s0 += 32) {
      int s = s0 + lane;
      bool v = (s < M) && slot_valid(masks[(size_t)b * M + s]);
      int flag = v ? 1 : 0, incl = flag;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_sync(0xffffffffu, incl, (lane >= d) ? lane - d : lane);
        if (lane >= d) incl += t;
      }
      int tot = __shfl_sync(0xffffffffu, incl, 31);
      int p = pos + incl - flag;
      if (v && p < rows_cap) {
        tok_slot[p] = s;
        tok_sample[p] = b;
      }
      pos += tot;
    }
  }
  if (lane == 0 && pos < rows_cap) {
    tok_slot[pos] = M;
    tok_sample[pos] = b;
  }
}

// ------------------------------------------------------------- token gather + relative pose encoding
// smt_state_encoder.py:210-276.  One warp per packed row.  xin row = [f[:pi], pose_enc(16), f[pi+4:]]
// where f is the memory slot's feature vector (or the current observation's), pose5 row =
// [x, y, cos h, sin h, exp(-t)] of the token relative to the sample's agent pose (saved for backward).
__global__ void smt_gather_kernel(const float* __restrict__ x, const float* __restrict__ memory,
                                  const int* __restrict__ env_index, const int* __restrict__ tok_slot,
                                  const int* __restrict__ tok_sample, const int* __restrict__ total, int rows_cap,
                                  int M, int n_mem_envs, int F, int pi, const float* __restrict__ pose_w,
                                  const float* __restrict__ pose_b, float* xin, float* pose5) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int R = min(*total, rows_cap);
  if (r >= R) return;
  const int b = tok_sample[r], slot = tok_slot[r];
  const float* xa = x + (size_t)b * F;
  const float* f = xa;
  if (slot < M) {
    const int env = env_index ? env_index[b] : b;
    f = memory + ((size_t)slot * n_mem_envs + env) * F;
  }
  const int Fin = F + 12;
  float* o = xin + (size_t)r * Fin;
  for (int c = lane; c < pi; c += 32) o[c] = f[c];
  for (int c = pi + 4 + lane; c < F; c += 32) o[c + 12] = f[c];
  // relative pose (agent a = current obs of sample b, token t)
  const float ax = xa[pi], ay = xa[pi + 1], ah = xa[pi + 2];
  const float tx = f[pi], ty = f[pi + 1], th = f[pi + 2], tt = f[pi + 3];
  const float heading_a = -ah, heading_b = -th;
  const float dx = ax - tx, dy = ay - ty;
  const float rab = sqrtf(dx * dx + dy * dy);
  float phi = atan2f(ty - ay, tx - ax);
  phi = phi - heading_a;
  const float xab = rab * cosf(phi), yab = rab * sinf(phi);
  float hab = heading_b - heading_a;
  hab = atan2f(sinf(hab), cosf(hab));
  hab = -hab;
  float p5[5] = {xab, yab, cosf(hab), sinf(hab), expf(-tt)};
  if (lane < 5) pose5[(size_t)r * 8 + lane] = p5[lane];
  if (lane < 16) {
    float acc = pose_b[lane];
#pragma unroll
    for (int i = 0; i < 5; ++i) acc = fmaf(pose_w[lane * 5 + i], p5[i], acc);
    o[pi + lane] = acc;
  }
}

// dx[b, c] (+)= gxin[row of the current token of b, mapped column]; pose columns get no gradient.
__global__ void smt_scatter_dx_kernel(const float* __restrict__ gxin, const int* __restrict__ off, int B, int F,
                                      int pi, float* dx) {
  const int b = blockIdx.x;
  const int r = off[b + 1] - 1;  // the current observation is the last token of the sample
  const int Fin = F + 12;
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float g = 0.f;
    if (c < pi) g = gxin[(size_t)r * Fin + c];
    else if (c >= pi + 4) g = gxin[(size_t)r * Fin + c + 12];
    dx[(size_t)b * F + c] += g;
  }
}

__global__ void add_inplace_kernel(float* a, const float* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    a[i] += b[i];
}

// =========================================================================== host-side launch helpers
struct Launcher {
  cudaStream_t s;
  int err = 0;
  void check() {
    avl_count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && !err) { avl_set_cuda_error((int)e); err = AVL_ERR_CUDA; }
  }
};

static GemmEpilogue make_ep(const float* bias, int relu, const int* m_dev) {
  GemmEpilogue ep;
  ep.bias = bias; ep.scale = nullptr; ep.residual = nullptr; ep.ldr = 0; ep.relu = relu; ep.accumulate = 0;
  ep.m_dev = m_dev; ep.k_dev = nullptr;
  return ep;
}

static void launch_gemm(Launcher& L, GemmOperand A, bool a_kc, GemmOperand B, bool b_kc, float* C, long long ldc,
                        int M, int N, int K, const GemmEpilogue& ep, int splits) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  if (a_kc && b_kc && splits <= 1 && M <= 4 * SK_BM && !ep.scale && !ep.residual && !ep.k_dev) {
    // few rows (rollout batch): many small CTAs instead of 1-2 tiles of the 128 x 64 kernel
    const int vec = (((A.s_row & 3) == 0 && (K & 3) == 0 && ((uintptr_t)A.p & 15) == 0) ? 1 : 0) |
                    (((B.s_row & 3) == 0 && (K & 3) == 0 && ((uintptr_t)B.p & 15) == 0) ? 2 : 0);
    AVL_LAUNCH_PDL(skinny_gemm_kernel, dim3(avl_div_up(M, SK_BM), avl_div_up(N, SK_BN)), SK_THREADS, 0, L.s, A.p, A.s_row, B.p,
               B.s_row, C, ldc, M, N, K, ep.bias, ep.relu, ep.accumulate, ep.m_dev, vec);
    L.check();
    return;
  }
  ConvGeom g = {};
  if (splits < 1) splits = 1;
  int kps = ((K + splits - 1) / splits + GBK - 1) / GBK * GBK;
  splits = (K + kps - 1) / kps;
  dim3 grid(avl_div_up(M, GBM), avl_div_up(N, GBN), splits);
  auto k11 = gemm_kernel<false, true, true>;
  auto k10 = gemm_kernel<false, true, false>;
  auto k01 = gemm_kernel<false, false, true>;
  auto k00 = gemm_kernel<false, false, false>;
  if (a_kc && b_kc) AVL_LAUNCH(k11, grid, GTHREADS, 0, L.s, A, B, C, ldc, M, N, K, g, ep, kps);
  else if (a_kc && !b_kc) AVL_LAUNCH(k10, grid, GTHREADS, 0, L.s, A, B, C, ldc, M, N, K, g, ep, kps);
  else if (!a_kc && b_kc) AVL_LAUNCH(k01, grid, GTHREADS, 0, L.s, A, B, C, ldc, M, N, K, g, ep, kps);
  else AVL_LAUNCH(k00, grid, GTHREADS, 0, L.s, A, B, C, ldc, M, N, K, g, ep, kps);
  L.check();
}

#ifndef AVL_HOST_EMUL
AVL_API int avl_tc_gemm(const float* A, long long lda, const float* B, float* C, long long ldc, int M, int N, int K,
                           const float* scale, const float* bias, const float* residual, long long ldr, int relu,
                           const int* m_dev, void* stream);
AVL_API int avl_get_tensor_cores(void);
static bool tc_ok(const float* X, long long ldx, const float* W, int rows, int K) {
  return avl_get_tensor_cores() >= 2 && rows >= 512 && (K & 3) == 0 && (ldx & 3) == 0 && (((uintptr_t)X | (uintptr_t)W) & 15) == 0;
}
// fp32-accurate tensor-core path (3xTF32, gemm_3xtf32.cu): level >= 1, enough rows to fill 128-row tiles
AVL_API int avl_tc_gemm_3x(const float* A, long long lda, const float* B, long long ldb, int b_transposed, float* C,
                           long long ldc, int M, int N, int K, const float* bias, const float* residual, long long ldr,
                           int relu, const int* m_dev, void* stream);
AVL_API int avl_tc_wgrad_3x(const float* dY, long long ldy, const float* X, long long ldx, float* dW, long long lddw,
                            int rows, int N, int K, const int* rows_dev, void* stream);
static bool x3_ok(const float* X, long long ldx, int rows, int K) {
  return avl_get_tensor_cores() >= 1 && rows >= 512 && (K & 3) == 0 && (ldx & 3) == 0 && ((uintptr_t)X & 15) == 0;
}
static int avl_get_tensor_cores_or0() { return avl_get_tensor_cores(); }
#else
static int avl_get_tensor_cores_or0() { return 0; }
static bool tc_ok(const float*, long long, const float*, int, int) { return false; }
static bool x3_ok(const float*, long long, int, int) { return false; }
static int avl_tc_gemm_3x(const float*, long long, const float*, long long, int, float*, long long, int, int, int,
                          const float*, const float*, long long, int, const int*, void*) { return AVL_ERR_UNSUPPORTED; }
static int avl_tc_wgrad_3x(const float*, long long, const float*, long long, float*, long long, int, int, int, const int*,
                           void*) { return AVL_ERR_UNSUPPORTED; }
static int avl_tc_gemm(const float*, long long, const float*, float*, long long, int, int, int, const float*,
                       const float*, const float*, long long, int, const int*, void*) { return 0; }
#endif

__global__ void transpose_kernel(const float* __restrict__ w, long long ldw, float* wt, int rows, int cols) {
  // wt[c][r] = w[r][c]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int c = i % cols, r = i / cols;
  wt[(size_t)c * rows + r] = w[(size_t)r * ldw + c];
}

// Y[rows, N] = X[rows, K] W[N, K]^T + b  (ReLU)
static void lin_fwd(Launcher& L, const float* X, long long ldx, const float* W, const float* b, float* Y,
                    long long ldy, int rows, int N, int K, int relu, const int* rows_dev) {
  if (tc_ok(X, ldx, W, rows, K)) {
    int rc = avl_tc_gemm(X, ldx, W, Y, ldy, rows, N, K, nullptr, b, nullptr, 0, relu, rows_dev, L.s);
    if (rc && !L.err) L.err = rc;
    return;
  }
  if (x3_ok(X, ldx, rows, K)) {
    int rc = avl_tc_gemm_3x(X, ldx, W, K, 0, Y, ldy, rows, N, K, b, nullptr, 0, relu, rows_dev, L.s);
    if (rc == AVL_OK) return;
    if (rc != AVL_ERR_UNSUPPORTED && !L.err) { L.err = rc; return; }
  }
  launch_gemm(L, {X, ldx, 1}, true, {W, (long long)K, 1}, true, Y, ldy, rows, N, K, make_ep(b, relu, rows_dev), 1);
}
// dX[rows, K] (+)= dY[rows, N] W[N, K]
static void lin_bwd_x(Launcher& L, const float* dY, long long ldy, const float* W, long long ldw, float* dX,
                      long long ldx, int rows, int N, int K, int accumulate, const int* rows_dev,
                      float* wt_scratch = nullptr) {
  if (wt_scratch && tc_ok(dY, ldy, wt_scratch, rows, N) && (ldx & 3) == 0) {
    // B operand must be [K][N] N-contiguous: transpose the (small) weight once, then dX = dY . (W^T)^T
    AVL_LAUNCH(transpose_kernel, avl_div_up((long long)N * K, 256), 256, 0, L.s, W, ldw, wt_scratch, N, K);
    L.check();
    int rc = avl_tc_gemm(dY, ldy, wt_scratch, dX, ldx, rows, K, N, nullptr, nullptr, accumulate ? dX : nullptr, ldx, 0,
                         rows_dev, L.s);
    if (rc && !L.err) L.err = rc;
    return;
  }
  if (x3_ok(dY, ldy, rows, N)) {  // dX = dY . W with B = W^T taken from the [N][K] weight by the split kernel
    int rc = avl_tc_gemm_3x(dY, ldy, W, ldw, 1, dX, ldx, rows, K, N, nullptr, accumulate ? dX : nullptr, ldx, 0, rows_dev,
                            L.s);
    if (rc == AVL_OK) return;
    if (rc != AVL_ERR_UNSUPPORTED && !L.err) { L.err = rc; return; }
  }
  GemmEpilogue ep = make_ep(nullptr, 0, rows_dev);
  ep.accumulate = accumulate;
  launch_gemm(L, {dY, ldy, 1}, true, {W, 1, ldw}, false, dX, ldx, rows, K, N, ep, 1);
}
// dW[N, K] += dY[rows, N]^T X[rows, K] ; db[N] += colsum(dY)
static void lin_bwd_w(Launcher& L, const float* dY, long long ldy, const float* X, long long ldx, float* dW,
                      long long lddw, float* db, int rows, int N, int K, const int* rows_dev) {
  bool done_w = false;
  if (dW && avl_get_tensor_cores_or0() >= 1 && rows >= 2048 && N >= 32) {  // tcgen05 3xTF32, MN-major operands straight from TMA
    int rc = avl_tc_wgrad_3x(dY, ldy, X, ldx, dW, lddw, rows, N, K, rows_dev, L.s);
    if (rc == AVL_OK) done_w = true;
    else if (rc != AVL_ERR_UNSUPPORTED && !L.err) { L.err = rc; return; }
  }
  if (dW && !done_w) {
    GemmEpilogue ep = make_ep(nullptr, 0, nullptr);
    ep.accumulate = 1;
    ep.k_dev = rows_dev;
    int tiles = avl_div_up(N, GBM) * avl_div_up(K, GBN);
    int splits = avl_div_up(4 * avl_num_sms(), tiles);
    int max_splits = avl_div_up(rows, 256);
    if (splits > max_splits) splits = max_splits;
    launch_gemm(L, {dY, 1, ldy}, false, {X, 1, ldx}, false, dW, lddw, N, K, rows, ep, splits);
  }
  if (db) {
    int gy = rows / 64;
    if (gy < 1) gy = 1;
    if (gy > 256) gy = 256;
    if ((N & 3) == 0 && (ldy & 3) == 0 && ((uintptr_t)dY & 15) == 0 && rows >= 512) {
      AVL_LAUNCH(colsum4_kernel, dim3(avl_div_up(N, 128), gy), 256, 0, L.s, dY, ldy, rows_dev, rows, N, db);
    } else {
      AVL_LAUNCH(colsum_kernel, dim3(avl_div_up(N, 128), gy), 128, 0, L.s, dY, ldy, rows_dev, rows, N, db);
    }
    L.check();
  }
}
static void ln_fwd(Launcher& L, const float* x, const float* res, const float* g, const float* b, float* y,
                   float* stats, const int* rows_dev, int rows, int cols) {
  if (rows <= 0) return;
  AVL_LAUNCH_PDL(layernorm_fwd_kernel, avl_div_up(rows, 8), 256, 0, L.s, x, res, g, b, y, stats, stats ? stats + rows : nullptr,
                                                            rows_dev, rows, cols, 1e-5f);
  L.check();
}
static void ln_bwd(Launcher& L, const float* x, const float* res, const float* g, const float* stats,
                   const float* dy, float* dx, float* dg, float* db, const int* rows_dev, int rows, int cols) {
  if (rows <= 0) return;
  int grid = avl_div_up(rows, 8 * 4);
  int cap = avl_num_sms() * 4;
  if (grid > cap) grid = cap;
  AVL_LAUNCH(layernorm_bwd_kernel, grid, 256, 2 * cols * sizeof(float), L.s, x, res, g, stats, stats + rows, dy, dx, dg, db,
                                                                    rows_dev, rows, cols);
  L.check();
}
static void relu_bwd(Launcher& L, float* dy, const float* y, const int* rows_dev, long long rows, int cols) {
  long long n = rows * cols;
  if (n <= 0) return;
  int grid = avl_div_up(n, 256 * 4);
  int cap = avl_num_sms() * 8;
  if (grid > cap) grid = cap;
  AVL_LAUNCH(relu_bwd_kernel, grid, 256, 0, L.s, dy, y, rows_dev, rows, cols);
  L.check();
}

static size_t attn_fwd_smem(int vcap) { return (size_t)(2 * vcap * ATT_HD) * sizeof(float); }
static size_t attn_bwd_smem(int vcap) { return (size_t)(4 * vcap * ATT_HD + 2 * vcap) * sizeof(float); }
static const size_t kAttnFwdSmem = attn_fwd_smem(ATT_MAXV);
static const size_t kAttnBwdSmem = attn_bwd_smem(ATT_MAXV);
static int attn_vcap(int rows_cap, int B) {  // per-sample token bound implied by the caller's row capacity
  int v = (rows_cap + B - 1) / B;
  if (v > ATT_MAXV) v = ATT_MAXV;
  if (v < 1) v = 1;
  return v;
}
// Tensor-core self-attention (attn_tc.cu, 3xTF32 warp MMAs) when the tensor-core level allows it; false = not taken.
#ifndef AVL_HOST_EMUL
extern "C" int avl_attn_self_fwd_tc_try(const float* qkv, const int* off, int B, int D, float* out, float* lse, float scale,
                                        int vcap, cudaStream_t stream);
extern "C" int avl_attn_self_bwd_tc_try(const float* qkv, const int* off, int B, int D, const float* out, const float* lse,
                             const float* dout, float* dqkv, float scale, int vcap, cudaStream_t stream);
static bool attn_fwd_tc(Launcher& L, const float* qkv, const int* off, int B, int D, float* out, float* lse, float scale,
                        int vcap) {
  int rc = avl_attn_self_fwd_tc_try(qkv, off, B, D, out, lse, scale, vcap, L.s);
  if (rc == AVL_ERR_UNSUPPORTED) return false;
  if (rc && !L.err) L.err = rc;
  return true;
}
static bool attn_bwd_tc(Launcher& L, const float* qkv, const int* off, int B, int D, const float* out, const float* lse,
                        const float* dout, float* dqkv, float scale, int vcap) {
  int rc = avl_attn_self_bwd_tc_try(qkv, off, B, D, out, lse, dout, dqkv, scale, vcap, L.s);
  if (rc == AVL_ERR_UNSUPPORTED) return false;
  if (rc && !L.err) L.err = rc;
  return true;
}
#else
static bool attn_fwd_tc(Launcher&, const float*, const int*, int, int, float*, float*, float, int) { return false; }
static bool attn_bwd_tc(Launcher&, const float*, const int*, int, int, const float*, const float*, const float*, float*,
                        float, int) { return false; }
#endif
static bool g_attn_attr_set = false;
static int ensure_attn_attrs() {
  if (g_attn_attr_set) return AVL_OK;
  AVL_CUDA_CHECK(cudaFuncSetAttribute(attn_self_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnFwdSmem));
  AVL_CUDA_CHECK(cudaFuncSetAttribute(attn_self_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnBwdSmem));
  g_attn_attr_set = true;
  return AVL_OK;
}

// ------------------------------------------------------------------------------ transformer block
// parameter table indices (shared with avlen_b200/savi/models/smt_state_encoder.py)
enum {
  TP_ENC_IN_W, TP_ENC_IN_B, TP_ENC_OUT_W, TP_ENC_OUT_B, TP_ENC_L1_W, TP_ENC_L1_B, TP_ENC_L2_W, TP_ENC_L2_B,
  TP_ENC_N1_W, TP_ENC_N1_B, TP_ENC_N2_W, TP_ENC_N2_B, TP_ENC_NORM_W, TP_ENC_NORM_B,
  TP_DEC_SA_IN_W, TP_DEC_SA_IN_B, TP_DEC_SA_OUT_W, TP_DEC_SA_OUT_B, TP_DEC_CA_IN_W, TP_DEC_CA_IN_B,
  TP_DEC_CA_OUT_W, TP_DEC_CA_OUT_B, TP_DEC_L1_W, TP_DEC_L1_B, TP_DEC_L2_W, TP_DEC_L2_B,
  TP_DEC_N1_W, TP_DEC_N1_B, TP_DEC_N2_W, TP_DEC_N2_B, TP_DEC_N3_W, TP_DEC_N3_B, TP_DEC_NORM_W, TP_DEC_NORM_B,
  TP_COUNT
};
enum { SP_POSE_W = TP_COUNT, SP_POSE_B, SP_FUS0_W, SP_FUS0_B, SP_FUS2_W, SP_FUS2_B, SP_COUNT };

struct Arena {
  char* base;
  size_t off = 0;
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct TfBufs {  // saved forward activations of the transformer on R packed rows / B target rows
  float *QKV, *ATT, *LSE, *AO, *X1, *ST1, *FF1, *FF2, *X2, *ST2, *MEM, *ST3, *KV, *PROBS;
  float *TV, *SA, *T1, *STT1, *Q, *C, *CO, *T2, *STT2, *DF1, *DF2, *T3, *STT3, *STT4;
  // backward scratch
  float *GA, *GB, *GC, *GQKV, *GKV, *gB1, *gB2, *gB3, *gB4, *WT;
};

static void tf_alloc(Arena& a, TfBufs& t, size_t R, size_t B, int D, int H, bool bwd) {
  t.QKV = a.take<float>(R * 3 * D); t.ATT = a.take<float>(R * D); t.LSE = a.take<float>(R * H);
  t.AO = a.take<float>(R * D); t.X1 = a.take<float>(R * D); t.ST1 = a.take<float>(2 * R);
  t.FF1 = a.take<float>(R * D); t.FF2 = a.take<float>(R * D); t.X2 = a.take<float>(R * D);
  t.ST2 = a.take<float>(2 * R); t.MEM = a.take<float>(R * D); t.ST3 = a.take<float>(2 * R);
  t.KV = a.take<float>(R * 2 * D); t.PROBS = a.take<float>(R * H);
  t.TV = a.take<float>(B * D); t.SA = a.take<float>(B * D); t.T1 = a.take<float>(B * D);
  t.STT1 = a.take<float>(2 * B); t.Q = a.take<float>(B * D); t.C = a.take<float>(B * D);
  t.CO = a.take<float>(B * D); t.T2 = a.take<float>(B * D); t.STT2 = a.take<float>(2 * B);
  t.DF1 = a.take<float>(B * D); t.DF2 = a.take<float>(B * D); t.T3 = a.take<float>(B * D);
  t.STT3 = a.take<float>(2 * B); t.STT4 = a.take<float>(2 * B);
  if (bwd) {
    t.GA = a.take<float>(R * D); t.GB = a.take<float>(R * D); t.GC = a.take<float>(R * D);
    t.GQKV = a.take<float>(R * 3 * D); t.GKV = a.take<float>(R * 2 * D);
    t.gB1 = a.take<float>(B * D); t.gB2 = a.take<float>(B * D); t.gB3 = a.take<float>(B * D);
    t.gB4 = a.take<float>(B * D);
    t.WT = a.take<float>((size_t)3 * D * 512);  // transposed-weight scratch (largest: in_proj 3D x D; fusion D x Fin)
  } else {
    t.GA = t.GB = t.GC = t.GQKV = t.GKV = t.gB1 = t.gB2 = t.gB3 = t.gB4 = t.WT = nullptr;
  }
}

// X0: [R, D] fused tokens, tgt: [B, D] decoder query, out: [B, D]
static void tf_forward(Launcher& L, const float* const* P, const TfBufs& t, const float* X0, const int* off,
                       const int* total, int Rcap, int B, int D, const float* tgt, float* out) {
  const int H = D / ATT_HD;
  const float scale = 1.0f / sqrtf((float)ATT_HD);
  const int vcap = attn_vcap(Rcap, B);
  lin_fwd(L, X0, D, P[TP_ENC_IN_W], P[TP_ENC_IN_B], t.QKV, 3 * D, Rcap, 3 * D, D, 0, total);
  if (!attn_fwd_tc(L, t.QKV, off, B, D, t.ATT, t.LSE, scale, vcap)) {
    AVL_LAUNCH(attn_self_fwd_kernel, dim3(B, H), ATT_WARPS * 32, attn_fwd_smem(vcap), L.s, t.QKV, off, t.ATT, t.LSE, D, scale, vcap);
    L.check();
  }
  lin_fwd(L, t.ATT, D, P[TP_ENC_OUT_W], P[TP_ENC_OUT_B], t.AO, D, Rcap, D, D, 0, total);
  ln_fwd(L, X0, t.AO, P[TP_ENC_N1_W], P[TP_ENC_N1_B], t.X1, t.ST1, total, Rcap, D);
  lin_fwd(L, t.X1, D, P[TP_ENC_L1_W], P[TP_ENC_L1_B], t.FF1, D, Rcap, D, D, 1, total);
  lin_fwd(L, t.FF1, D, P[TP_ENC_L2_W], P[TP_ENC_L2_B], t.FF2, D, Rcap, D, D, 0, total);
  ln_fwd(L, t.X1, t.FF2, P[TP_ENC_N2_W], P[TP_ENC_N2_B], t.X2, t.ST2, total, Rcap, D);
  ln_fwd(L, t.X2, nullptr, P[TP_ENC_NORM_W], P[TP_ENC_NORM_B], t.MEM, t.ST3, total, Rcap, D);
  // decoder: self-attention over a length-1 target is the identity softmax -> out_proj(v_proj(tgt))
  lin_fwd(L, tgt, D, P[TP_DEC_SA_IN_W] + (size_t)2 * D * D, P[TP_DEC_SA_IN_B] + 2 * D, t.TV, D, B, D, D, 0, nullptr);
  lin_fwd(L, t.TV, D, P[TP_DEC_SA_OUT_W], P[TP_DEC_SA_OUT_B], t.SA, D, B, D, D, 0, nullptr);
  ln_fwd(L, tgt, t.SA, P[TP_DEC_N1_W], P[TP_DEC_N1_B], t.T1, t.STT1, nullptr, B, D);
  lin_fwd(L, t.T1, D, P[TP_DEC_CA_IN_W], P[TP_DEC_CA_IN_B], t.Q, D, B, D, D, 0, nullptr);
  lin_fwd(L, t.MEM, D, P[TP_DEC_CA_IN_W] + (size_t)D * D, P[TP_DEC_CA_IN_B] + D, t.KV, 2 * D, Rcap, 2 * D, D, 0, total);
  if (D == 256)
    AVL_LAUNCH_PDL(attn_cross_fwd256_kernel, B, 256, H * ATT_MAXV * sizeof(float), L.s, t.Q, t.KV, off, t.C, t.PROBS, scale);
  else
    AVL_LAUNCH(attn_cross_fwd_kernel, B, H * 32, H * ATT_MAXV * sizeof(float), L.s, t.Q, t.KV, off, t.C, t.PROBS, D, scale);
  L.check();
  lin_fwd(L, t.C, D, P[TP_DEC_CA_OUT_W], P[TP_DEC_CA_OUT_B], t.CO, D, B, D, D, 0, nullptr);
  ln_fwd(L, t.T1, t.CO, P[TP_DEC_N2_W], P[TP_DEC_N2_B], t.T2, t.STT2, nullptr, B, D);
  lin_fwd(L, t.T2, D, P[TP_DEC_L1_W], P[TP_DEC_L1_B], t.DF1, D, B, D, D, 1, nullptr);
  lin_fwd(L, t.DF1, D, P[TP_DEC_L2_W], P[TP_DEC_L2_B], t.DF2, D, B, D, D, 0, nullptr);
  ln_fwd(L, t.T2, t.DF2, P[TP_DEC_N3_W], P[TP_DEC_N3_B], t.T3, t.STT3, nullptr, B, D);
  ln_fwd(L, t.T3, nullptr, P[TP_DEC_NORM_W], P[TP_DEC_NORM_B], out, t.STT4, nullptr, B, D);
}

static float* gp(float* const* G, int i) { return G ? G[i] : nullptr; }
static float* gp_off(float* const* G, int i, size_t o) { return (G && G[i]) ? G[i] + o : nullptr; }

// Backward of tf_forward.  gout: [B, D].  Writes gX0 [R, D] (gradient wrt the fused tokens) into
// t.GC and gtgt [B, D] into t.gB4; parameter gradients are ACCUMULATED into G[*] (null = skip).
static void tf_backward(Launcher& L, const float* const* P, float* const* G, const TfBufs& t, const float* X0,
                        const int* off, const int* total, int Rcap, int B, int D, const float* tgt,
                        const float* gout) {
  const int H = D / ATT_HD;
  const float scale = 1.0f / sqrtf((float)ATT_HD);
  const int vcap = attn_vcap(Rcap, B);
  const size_t DD = (size_t)D * D;
  // ---- decoder
  ln_bwd(L, t.T3, nullptr, P[TP_DEC_NORM_W], t.STT4, gout, t.gB1, gp(G, TP_DEC_NORM_W), gp(G, TP_DEC_NORM_B), nullptr, B, D);
  ln_bwd(L, t.T2, t.DF2, P[TP_DEC_N3_W], t.STT3, t.gB1, t.gB2, gp(G, TP_DEC_N3_W), gp(G, TP_DEC_N3_B), nullptr, B, D);
  // gB2 = grad wrt (T2 + DF2)
  lin_bwd_w(L, t.gB2, D, t.DF1, D, gp(G, TP_DEC_L2_W), D, gp(G, TP_DEC_L2_B), B, D, D, nullptr);
  lin_bwd_x(L, t.gB2, D, P[TP_DEC_L2_W], D, t.gB3, D, B, D, D, 0, nullptr);
  relu_bwd(L, t.gB3, t.DF1, nullptr, B, D);
  lin_bwd_w(L, t.gB3, D, t.T2, D, gp(G, TP_DEC_L1_W), D, gp(G, TP_DEC_L1_B), B, D, D, nullptr);
  lin_bwd_x(L, t.gB3, D, P[TP_DEC_L1_W], D, t.gB2, D, B, D, D, 1, nullptr);  // gB2 = gT2
  ln_bwd(L, t.T1, t.CO, P[TP_DEC_N2_W], t.STT2, t.gB2, t.gB1, gp(G, TP_DEC_N2_W), gp(G, TP_DEC_N2_B), nullptr, B, D);
  // gB1 = grad wrt (T1 + CO)
  lin_bwd_w(L, t.gB1, D, t.C, D, gp(G, TP_DEC_CA_OUT_W), D, gp(G, TP_DEC_CA_OUT_B), B, D, D, nullptr);
  lin_bwd_x(L, t.gB1, D, P[TP_DEC_CA_OUT_W], D, t.gB3, D, B, D, D, 0, nullptr);  // gB3 = gC
  AVL_LAUNCH(attn_cross_bwd_kernel, B, H * 32, H * ATT_MAXV * sizeof(float), L.s, t.Q, t.KV, off, t.PROBS, t.gB3, t.gB2,
                                                                         t.GKV, D, scale);  // gB2 = gQ
  L.check();
  lin_bwd_w(L, t.gB2, D, t.T1, D, gp(G, TP_DEC_CA_IN_W), D, gp(G, TP_DEC_CA_IN_B), B, D, D, nullptr);
  lin_bwd_x(L, t.gB2, D, P[TP_DEC_CA_IN_W], D, t.gB1, D, B, D, D, 1, nullptr);  // gB1 = gT1 (residual + q path)
  lin_bwd_w(L, t.GKV, 2 * D, t.MEM, D, gp_off(G, TP_DEC_CA_IN_W, DD), D, gp_off(G, TP_DEC_CA_IN_B, D), Rcap, 2 * D, D, total);
  lin_bwd_x(L, t.GKV, 2 * D, P[TP_DEC_CA_IN_W] + DD, D, t.GA, D, Rcap, 2 * D, D, 0, total, t.WT);  // GA = gMEM
  ln_bwd(L, tgt, t.SA, P[TP_DEC_N1_W], t.STT1, t.gB1, t.gB4, gp(G, TP_DEC_N1_W), gp(G, TP_DEC_N1_B), nullptr, B, D);
  // gB4 = grad wrt (tgt + SA)
  lin_bwd_w(L, t.gB4, D, t.TV, D, gp(G, TP_DEC_SA_OUT_W), D, gp(G, TP_DEC_SA_OUT_B), B, D, D, nullptr);
  lin_bwd_x(L, t.gB4, D, P[TP_DEC_SA_OUT_W], D, t.gB3, D, B, D, D, 0, nullptr);  // gB3 = gTV
  lin_bwd_w(L, t.gB3, D, tgt, D, gp_off(G, TP_DEC_SA_IN_W, 2 * DD), D, gp_off(G, TP_DEC_SA_IN_B, 2 * D), B, D, D, nullptr);
  lin_bwd_x(L, t.gB3, D, P[TP_DEC_SA_IN_W] + 2 * DD, D, t.gB4, D, B, D, D, 1, nullptr);  // gB4 = gtgt
  // ---- encoder
  ln_bwd(L, t.X2, nullptr, P[TP_ENC_NORM_W], t.ST3, t.GA, t.GB, gp(G, TP_ENC_NORM_W), gp(G, TP_ENC_NORM_B), total, Rcap, D);
  ln_bwd(L, t.X1, t.FF2, P[TP_ENC_N2_W], t.ST2, t.GB, t.GA, gp(G, TP_ENC_N2_W), gp(G, TP_ENC_N2_B), total, Rcap, D);
  // GA = grad wrt (X1 + FF2)
  lin_bwd_w(L, t.GA, D, t.FF1, D, gp(G, TP_ENC_L2_W), D, gp(G, TP_ENC_L2_B), Rcap, D, D, total);
  lin_bwd_x(L, t.GA, D, P[TP_ENC_L2_W], D, t.GB, D, Rcap, D, D, 0, total, t.WT);
  relu_bwd(L, t.GB, t.FF1, total, Rcap, D);
  lin_bwd_w(L, t.GB, D, t.X1, D, gp(G, TP_ENC_L1_W), D, gp(G, TP_ENC_L1_B), Rcap, D, D, total);
  lin_bwd_x(L, t.GB, D, P[TP_ENC_L1_W], D, t.GA, D, Rcap, D, D, 1, total, t.WT);  // GA = gX1
  ln_bwd(L, X0, t.AO, P[TP_ENC_N1_W], t.ST1, t.GA, t.GC, gp(G, TP_ENC_N1_W), gp(G, TP_ENC_N1_B), total, Rcap, D);
  // GC = grad wrt (X0 + AO)
  lin_bwd_w(L, t.GC, D, t.ATT, D, gp(G, TP_ENC_OUT_W), D, gp(G, TP_ENC_OUT_B), Rcap, D, D, total);
  lin_bwd_x(L, t.GC, D, P[TP_ENC_OUT_W], D, t.GB, D, Rcap, D, D, 0, total, t.WT);  // GB = gATT
  if (!attn_bwd_tc(L, t.QKV, off, B, D, t.ATT, t.LSE, t.GB, t.GQKV, scale, vcap)) {
    AVL_LAUNCH(attn_self_bwd_kernel, dim3(B, H), ATT_BWD_WARPS * 32, attn_bwd_smem(vcap), L.s, t.QKV, off, t.ATT, t.LSE, t.GB, t.GQKV, D, scale, vcap);
    L.check();
  }
  lin_bwd_w(L, t.GQKV, 3 * D, X0, D, gp(G, TP_ENC_IN_W), D, gp(G, TP_ENC_IN_B), Rcap, 3 * D, D, total);
  lin_bwd_x(L, t.GQKV, 3 * D, P[TP_ENC_IN_W], D, t.GC, D, Rcap, 3 * D, D, 1, total, t.WT);  // GC = gX0
}

// ------------------------------------------------------------------------------------- SMT context
struct SmtBufs {
  int *cnt, *off, *total, *err, *tok_slot, *tok_sample;
  float *XIN, *POSE5, *H1, *X0, *GH1, *GPOSE, *GXIN;
  TfBufs tf;
};

static size_t smt_layout(char* base, SmtBufs& s, size_t B, size_t R, int F, int D, bool bwd, bool need_dx) {
  Arena a{base};
  const int Fin = F + 12, H = D / ATT_HD;
  s.cnt = a.take<int>(B); s.off = a.take<int>(B + 1); s.total = a.take<int>(1); s.err = a.take<int>(1);
  s.tok_slot = a.take<int>(R); s.tok_sample = a.take<int>(R);
  s.XIN = a.take<float>(R * Fin); s.POSE5 = a.take<float>(R * 8); s.H1 = a.take<float>(R * D);
  s.X0 = a.take<float>(R * D);
  tf_alloc(a, s.tf, R, B, D, H, bwd);
  s.GH1 = bwd ? a.take<float>(R * D) : nullptr;
  s.GPOSE = bwd ? a.take<float>(R * 16) : nullptr;
  s.GXIN = (bwd && need_dx) ? a.take<float>(R * Fin) : nullptr;
  return a.off + 256;
}
}  // namespace

// ================================================================================= exported entry points

// Generic dense GEMM: C[M,N] (+)= A(m,k) B(n,k) with explicit element strides (sa_m, sa_k, sb_n, sb_k),
// optional bias[N], ReLU, C accumulate and split-K (atomic accumulation into C).
AVL_API int avl_gemm(const float* A, long long sa_m, long long sa_k, const float* B, long long sb_n, long long sb_k,
                     float* C, long long ldc, int M, int N, int K, const float* bias, int relu, int accumulate,
                     int splits, void* stream) {
  if (M < 0 || N < 0 || K < 0) return AVL_ERR_ARG;
  if (M == 0 || N == 0) return AVL_OK;
  if (!A || !B || !C) return AVL_ERR_ARG;
  if (splits > 1 && (bias || relu)) return AVL_ERR_ARG;
  Launcher L{(cudaStream_t)stream};
  GemmEpilogue ep = make_ep(bias, relu, nullptr);
  ep.accumulate = accumulate;
  launch_gemm(L, {A, sa_m, sa_k}, sa_k == 1, {B, sb_n, sb_k}, sb_k == 1, C, ldc, M, N, K, ep, splits);
  return L.err;
}

// y = LayerNorm(x (+ res)) ; stats (2*rows floats: mean then rstd) may be null.
AVL_API int avl_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                              float* stats, int rows, int cols, void* stream) {
  if (rows < 0 || cols < 32 || (cols & 31) || cols > 32 * LN_MAX_PER_LANE) return AVL_ERR_UNSUPPORTED;
  if (rows == 0) return AVL_OK;
  if (!x || !gamma || !beta || !y) return AVL_ERR_ARG;
  Launcher L{(cudaStream_t)stream};
  ln_fwd(L, x, res, gamma, beta, y, stats, nullptr, rows, cols);
  return L.err;
}

// dx (grad wrt x + res); dgamma / dbeta accumulated.
AVL_API int avl_layernorm_bwd(const float* x, const float* res, const float* gamma, const float* stats,
                              const float* dy, float* dx, float* dgamma, float* dbeta, int rows, int cols,
                              void* stream) {
  if (rows < 0 || cols < 32 || (cols & 31) || cols > 32 * LN_MAX_PER_LANE) return AVL_ERR_UNSUPPORTED;
  if (rows == 0) return AVL_OK;
  if (!x || !gamma || !stats || !dy || !dx) return AVL_ERR_ARG;
  Launcher L{(cudaStream_t)stream};
  ln_bwd(L, x, res, gamma, stats, dy, dx, dgamma, dbeta, nullptr, rows, cols);
  return L.err;
}

// Variable-length multi-head self-attention on packed rows (head dim 32).  off: [B+1] row offsets.
AVL_API int avl_attn_self_fwd(const float* qkv, const int* off, int B, int D, float* out, float* lse, void* stream) {
  if (B < 0 || D < 32 || D % 32) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!qkv || !off || !out) return AVL_ERR_ARG;
  int rc = ensure_attn_attrs();
  if (rc) return rc;
  {
    Launcher L{(cudaStream_t)stream};
    if (attn_fwd_tc(L, qkv, off, B, D, out, lse, 1.0f / sqrtf(32.f), ATT_MAXV)) return L.err;
  }
  AVL_LAUNCH(attn_self_fwd_kernel, dim3(B, D / 32), ATT_WARPS * 32, kAttnFwdSmem, (cudaStream_t)stream, 
      qkv, off, out, lse, D, 1.0f / sqrtf(32.f), ATT_MAXV);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_attn_self_bwd(const float* qkv, const int* off, int B, int D, const float* out, const float* lse,
                              const float* dout, float* dqkv, void* stream) {
  if (B < 0 || D < 32 || D % 32) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!qkv || !off || !out || !lse || !dout || !dqkv) return AVL_ERR_ARG;
  int rc = ensure_attn_attrs();
  if (rc) return rc;
  {
    Launcher L{(cudaStream_t)stream};
    if (attn_bwd_tc(L, qkv, off, B, D, out, lse, dout, dqkv, 1.0f / sqrtf(32.f), ATT_MAXV)) return L.err;
  }
  AVL_LAUNCH(attn_self_bwd_kernel, dim3(B, D / 32), ATT_BWD_WARPS * 32, kAttnBwdSmem, (cudaStream_t)stream, 
      qkv, off, out, lse, dout, dqkv, D, 1.0f / sqrtf(32.f), ATT_MAXV);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_attn_cross_fwd(const float* q, const float* kv, const int* off, int B, int D, float* out,
                               float* probs, void* stream) {
  if (B < 0 || D < 32 || D % 32 || D > 1024) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!q || !kv || !off || !out) return AVL_ERR_ARG;
  int H = D / 32;
  if (D == 256 && (((uintptr_t)q | (uintptr_t)kv) & 15) == 0)
    AVL_LAUNCH(attn_cross_fwd256_kernel, B, 256, H * ATT_MAXV * sizeof(float), (cudaStream_t)stream, q, kv, off, out, probs,
               1.0f / sqrtf(32.f));
  else
    AVL_LAUNCH(attn_cross_fwd_kernel, B, H * 32, H * ATT_MAXV * sizeof(float), (cudaStream_t)stream, q, kv, off, out, probs, D,
               1.0f / sqrtf(32.f));
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_attn_cross_bwd(const float* q, const float* kv, const int* off, const float* probs,
                               const float* dout, int B, int D, float* dq, float* dkv, void* stream) {
  if (B < 0 || D < 32 || D % 32 || D > 1024) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!q || !kv || !off || !probs || !dout || !dq || !dkv) return AVL_ERR_ARG;
  int H = D / 32;
  AVL_LAUNCH(attn_cross_bwd_kernel, B, H * 32, H * ATT_MAXV * sizeof(float), (cudaStream_t)stream, q, kv, off, probs, dout,
                                                                                          dq, dkv, D,
                                                                                          1.0f / sqrtf(32.f));
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

AVL_API int avl_smt_param_count(void) { return SP_COUNT; }

// Workspace bytes for avl_smt_forward (+ backward when with_backward != 0).
AVL_API long long avl_smt_workspace_bytes(int B, int rows_cap, int F, int D, int with_backward, int need_dx) {
  SmtBufs s;
  return (long long)smt_layout(nullptr, s, (size_t)B, (size_t)rows_cap, F, D, with_backward != 0, need_dx != 0);
}

// SMTStateEncoder.single_forward (smt_state_encoder.py:109-188).
//   x [B, F] current features (pose at columns [pi, pi+4)); memory [M, n_mem_envs, F] single-copy ring
//   buffer; env_index [B] (null: row b reads memory[:, b]); masks [B, M] float {0,1}; goal [B, D];
//   params: table of SP_COUNT device pointers (order = enum above); out [B, D].
//   rows_cap: upper bound on the number of valid tokens (<= B*(M+1)); exceeding it sets a sticky
//   error readable with avl_smt_status.  No host synchronisation happens here.
AVL_API int avl_smt_forward(int B, int M, int F, int D, int pi, int pretraining, int rows_cap, const float* x,
                            const float* memory, int n_mem_envs, const int* env_index, const float* masks,
                            const float* goal, const float* const* params, float* out, void* workspace,
                            int with_backward, int need_dx, void* stream) {
  if (B < 0 || M < 0 || F < 5 || D != 256 || pi < 0 || pi + 4 > F || rows_cap < B) return AVL_ERR_ARG;
  if (M + 1 > ATT_MAXV) return AVL_ERR_UNSUPPORTED;
  if (B == 0) return AVL_OK;
  if (!x || !goal || !params || !out || !workspace) return AVL_ERR_ARG;
  if (!pretraining && M > 0 && (!memory || !masks)) return AVL_ERR_ARG;
  int rc = ensure_attn_attrs();
  if (rc) return rc;
  SmtBufs s;
  smt_layout(static_cast<char*>(workspace), s, (size_t)B, (size_t)rows_cap, F, D, with_backward != 0, need_dx != 0);
  Launcher L{(cudaStream_t)stream};
  const int Fin = F + 12;
  cudaMemsetAsync(s.err, 0, sizeof(int), L.s);
  AVL_LAUNCH(smt_count_kernel, avl_div_up(B, 8), 256, 0, L.s, masks, B, M, pretraining, s.cnt);
  L.check();
  AVL_LAUNCH(smt_scan_kernel, 1, 1024, 0, L.s, s.cnt, B, s.off, s.total, rows_cap, s.err);
  L.check();
  AVL_LAUNCH(smt_fill_kernel, avl_div_up(B, 8), 256, 0, L.s, masks, s.off, B, M, pretraining, rows_cap, s.tok_slot, s.tok_sample);
  L.check();
  AVL_LAUNCH(smt_gather_kernel, avl_div_up(rows_cap, 8), 256, 0, L.s, x, memory, env_index, s.tok_slot, s.tok_sample, s.total,
                                                             rows_cap, M, n_mem_envs, F, pi, params[SP_POSE_W],
                                                             params[SP_POSE_B], s.XIN, s.POSE5);
  L.check();
  lin_fwd(L, s.XIN, Fin, params[SP_FUS0_W], params[SP_FUS0_B], s.H1, D, rows_cap, D, Fin, 1, s.total);
  lin_fwd(L, s.H1, D, params[SP_FUS2_W], params[SP_FUS2_B], s.X0, D, rows_cap, D, D, 0, s.total);
  tf_forward(L, params, s.tf, s.X0, s.off, s.total, rows_cap, B, D, goal, out);
  return L.err;
}

// Backward of avl_smt_forward (same workspace, which must have been laid out with with_backward=1).
//   gout [B, D]; grads: table of SP_COUNT device pointers, gradients are ACCUMULATED (null = skip);
//   dx [B, F] (accumulated) and dgoal [B, D] (overwritten) may be null.
AVL_API int avl_smt_backward(int B, int M, int F, int D, int pi, int rows_cap, const float* goal,
                             const float* const* params, float* const* grads, const float* gout, float* dx,
                             float* dgoal, void* workspace, void* stream) {
  if (B < 0 || D != 256 || rows_cap < B) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!goal || !params || !gout || !workspace) return AVL_ERR_ARG;
  SmtBufs s;
  const bool need_dx = dx != nullptr;
  smt_layout(static_cast<char*>(workspace), s, (size_t)B, (size_t)rows_cap, F, D, true, need_dx);
  Launcher L{(cudaStream_t)stream};
  const int Fin = F + 12;
  tf_backward(L, params, grads, s.tf, s.X0, s.off, s.total, rows_cap, B, D, goal, gout);
  float* gX0 = s.tf.GC;
  lin_bwd_w(L, gX0, D, s.H1, D, gp(grads, SP_FUS2_W), D, gp(grads, SP_FUS2_B), rows_cap, D, D, s.total);
  lin_bwd_x(L, gX0, D, params[SP_FUS2_W], D, s.GH1, D, rows_cap, D, D, 0, s.total, s.tf.WT);
  relu_bwd(L, s.GH1, s.H1, s.total, rows_cap, D);
  lin_bwd_w(L, s.GH1, D, s.XIN, Fin, gp(grads, SP_FUS0_W), Fin, gp(grads, SP_FUS0_B), rows_cap, D, Fin, s.total);
  if (gp(grads, SP_POSE_W) || gp(grads, SP_POSE_B)) {
    // gradient wrt the 16 encoded pose columns only: gPOSE[R,16] = GH1[R,D] . W0[:, pi:pi+16]
    lin_bwd_x(L, s.GH1, D, params[SP_FUS0_W] + pi, Fin, s.GPOSE, 16, rows_cap, D, 16, 0, s.total);
    lin_bwd_w(L, s.GPOSE, 16, s.POSE5, 8, gp(grads, SP_POSE_W), 5, gp(grads, SP_POSE_B), rows_cap, 16, 5, s.total);
  }
  if (need_dx) {
    lin_bwd_x(L, s.GH1, D, params[SP_FUS0_W], Fin, s.GXIN, Fin, rows_cap, D, Fin, 0, s.total, s.tf.WT);
    AVL_LAUNCH(smt_scatter_dx_kernel, B, 128, 0, L.s, s.GXIN, s.off, B, F, pi, dx);
    L.check();
  }
  if (dgoal) {
    cudaMemcpyAsync(dgoal, s.tf.gB4, (size_t)B * D * sizeof(float), cudaMemcpyDeviceToDevice, L.s);
  }
  return L.err;
}

// Reads back (synchronising) the number of packed rows and the overflow flag of the last forward.
AVL_API int avl_smt_status(int B, int rows_cap, int F, int D, void* workspace, int* total_rows, int* overflow) {
  SmtBufs s;
  smt_layout(static_cast<char*>(workspace), s, (size_t)B, (size_t)rows_cap, F, D, false, false);
  if (total_rows) AVL_CUDA_CHECK(cudaMemcpy(total_rows, s.total, sizeof(int), cudaMemcpyDeviceToHost));
  if (overflow) AVL_CUDA_CHECK(cudaMemcpy(overflow, s.err, sizeof(int), cudaMemcpyDeviceToHost));
  return AVL_OK;
}

// =====================================================================================================
// Row K: DialogStateEncoder.single_forward (ss_baselines/savi/models/dialog_state_encoder.py:114-155).
//   tokens of sample b = [state_memory[s, env(b)] for the valid slots s < K] + [x_att[b]]      (256-d each)
//   with a dialog: token <- fusion_encoder(cat[token, d_emb[b]])  (512 -> 256 ReLU -> 256)       (:138-140)
//   token += pe[agent_step[b]]   (sinusoidal table, max_len 100, dropout 0)                       (:142, :39)
//   out = Transformer(src = tokens, tgt = goal, key padding = invalid slots)[-1]                 (:147-152)
// Same packed-token design as the scene memory: only valid slots become rows.
// params: TP_COUNT transformer pointers followed by fusion_encoder.{0,2}.{weight,bias}.
namespace {
enum { DP_FUS0_W = TP_COUNT, DP_FUS0_B, DP_FUS2_W, DP_FUS2_B, DP_COUNT };

struct DlgBufs {
  int *cnt, *off, *total, *err, *tok_slot, *tok_sample;
  float *XIN, *PE, *H1, *X0, *GH1, *GXIN;
  TfBufs tf;
};

static size_t dlg_layout(char* base, DlgBufs& s, size_t B, size_t R, int D, bool bwd) {
  Arena a{base};
  s.cnt = a.take<int>(B); s.off = a.take<int>(B + 1); s.total = a.take<int>(1); s.err = a.take<int>(1);
  s.tok_slot = a.take<int>(R); s.tok_sample = a.take<int>(R);
  s.XIN = a.take<float>(R * 2 * D); s.PE = a.take<float>(R * D); s.H1 = a.take<float>(R * D);
  s.X0 = a.take<float>(R * D);
  tf_alloc(a, s.tf, R, B, D, D / ATT_HD, bwd);
  s.GH1 = bwd ? a.take<float>(R * D) : nullptr;
  s.GXIN = bwd ? a.take<float>(R * 2 * D) : nullptr;
  return a.off + 256;
}

// one warp per packed row: writes the token (+ d_emb) into XIN, the positional row into PE, and when there is no
// dialog X0 = token + pe directly.
__global__ void dlg_gather_kernel(const float* __restrict__ x_att, const float* __restrict__ mem,
                                  const int* __restrict__ env_index, const float* __restrict__ d_emb,
                                  const int* __restrict__ agent_step, const float* __restrict__ pe_table, int pe_len,
                                  const int* __restrict__ tok_slot, const int* __restrict__ tok_sample,
                                  const int* __restrict__ total, int rows_cap, int K, int n_mem_envs, int D,
                                  float* XIN, float* PE, float* X0) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int R = min(*total, rows_cap);
  if (r >= R) return;
  const int b = tok_sample[r], slot = tok_slot[r];
  const float* f = x_att + (size_t)b * D;
  if (slot < K) {
    const int env = env_index ? env_index[b] : b;
    f = mem + ((size_t)slot * n_mem_envs + env) * D;
  }
  int st = agent_step[b];
  st = st < 0 ? 0 : (st >= pe_len ? pe_len - 1 : st);
  const float* pe = pe_table + (size_t)st * D;
  if (d_emb) {
    float* o = XIN + (size_t)r * 2 * D;
    const float* de = d_emb + (size_t)b * D;
    for (int c = lane; c < D; c += 32) {
      o[c] = f[c];
      o[D + c] = de[c];
      PE[(size_t)r * D + c] = pe[c];
    }
  } else {
    for (int c = lane; c < D; c += 32) X0[(size_t)r * D + c] = f[c] + pe[c];
  }
}

// dx_att[b] = g[last row of b][:D] ; dd_emb[b] = sum over the rows of b of g[row][D:2D]   (ldg = row stride of g)
__global__ void dlg_scatter_kernel(const float* __restrict__ g, int ldg, const int* __restrict__ off, int D,
                                   float* dx_att, float* dd_emb) {
  const int b = blockIdx.x;
  const int r0 = off[b], r1 = off[b + 1];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    if (dx_att) dx_att[(size_t)b * D + c] = g[(size_t)(r1 - 1) * ldg + c];
    if (dd_emb) {
      float s = 0.f;
      for (int r = r0; r < r1; ++r) s += g[(size_t)r * ldg + D + c];
      dd_emb[(size_t)b * D + c] = s;
    }
  }
}
}  // namespace

AVL_API int avl_dialog_param_count(void) { return DP_COUNT; }

AVL_API long long avl_dialog_workspace_bytes(int B, int K, int D, int with_backward) {
  DlgBufs s;
  return (long long)dlg_layout(nullptr, s, (size_t)B, (size_t)B * (K + 1), D, with_backward != 0);
}

// x_att [B, D]; memory_state [K, n_mem_envs, D]; env_index [B] or NULL; masks [B, K] float; d_emb [B, D] or NULL
// (no dialog: the fusion MLP is skipped, :138); agent_step [B] int32; pe_table [pe_len, D]; goal [B, D]; out [B, D].
AVL_API int avl_dialog_forward(int B, int K, int D, const float* x_att, const float* memory_state, int n_mem_envs,
                               const int* env_index, const float* masks, const float* d_emb, const int* agent_step,
                               const float* pe_table, int pe_len, const float* goal, const float* const* params,
                               float* out, void* workspace, int with_backward, void* stream) {
  if (B < 0 || K < 0 || D != 256 || pe_len < 1) return AVL_ERR_ARG;
  if (K + 1 > ATT_MAXV) return AVL_ERR_UNSUPPORTED;
  if (B == 0) return AVL_OK;
  if (!x_att || !agent_step || !pe_table || !goal || !params || !out || !workspace) return AVL_ERR_ARG;
  if (K > 0 && (!memory_state || !masks)) return AVL_ERR_ARG;
  int rc = ensure_attn_attrs();
  if (rc) return rc;
  const int rows_cap = B * (K + 1);
  DlgBufs s;
  dlg_layout(static_cast<char*>(workspace), s, (size_t)B, (size_t)rows_cap, D, with_backward != 0);
  Launcher L{(cudaStream_t)stream};
  cudaMemsetAsync(s.err, 0, sizeof(int), L.s);
  AVL_LAUNCH(smt_count_kernel, avl_div_up(B, 8), 256, 0, L.s, masks, B, K, 0, s.cnt);
  L.check();
  AVL_LAUNCH(smt_scan_kernel, 1, 1024, 0, L.s, s.cnt, B, s.off, s.total, rows_cap, s.err);
  L.check();
  AVL_LAUNCH(smt_fill_kernel, avl_div_up(B, 8), 256, 0, L.s, masks, s.off, B, K, 0, rows_cap, s.tok_slot, s.tok_sample);
  L.check();
  AVL_LAUNCH(dlg_gather_kernel, avl_div_up(rows_cap, 8), 256, 0, L.s, x_att, memory_state, env_index, d_emb, agent_step,
             pe_table, pe_len, s.tok_slot, s.tok_sample, s.total, rows_cap, K, n_mem_envs, D, s.XIN, s.PE, s.X0);
  L.check();
  if (d_emb) {
    lin_fwd(L, s.XIN, 2 * D, params[DP_FUS0_W], params[DP_FUS0_B], s.H1, D, rows_cap, D, 2 * D, 1, s.total);
    // X0 = H1 W2^T + b2 + PE  (positional rows as the epilogue residual)
    GemmEpilogue ep = make_ep(params[DP_FUS2_B], 0, s.total);
    ep.residual = s.PE;
    ep.ldr = D;
    launch_gemm(L, {s.H1, (long long)D, 1}, true, {params[DP_FUS2_W], (long long)D, 1}, true, s.X0, D, rows_cap, D, D, ep, 1);
  }
  tf_forward(L, params, s.tf, s.X0, s.off, s.total, rows_cap, B, D, goal, out);
  return L.err;
}

// grads: DP_COUNT table (accumulated, NULL = skip); dx_att [B, D], dd_emb [B, D], dgoal [B, D] overwritten (NULL = skip).
AVL_API int avl_dialog_backward(int B, int K, int D, int has_dialog, const float* goal, const float* const* params,
                                float* const* grads, const float* gout, float* dx_att, float* dd_emb, float* dgoal,
                                void* workspace, void* stream) {
  if (B < 0 || D != 256) return AVL_ERR_ARG;
  if (B == 0) return AVL_OK;
  if (!goal || !params || !gout || !workspace) return AVL_ERR_ARG;
  const int rows_cap = B * (K + 1);
  DlgBufs s;
  dlg_layout(static_cast<char*>(workspace), s, (size_t)B, (size_t)rows_cap, D, true);
  Launcher L{(cudaStream_t)stream};
  tf_backward(L, params, grads, s.tf, s.X0, s.off, s.total, rows_cap, B, D, goal, gout);
  float* gX0 = s.tf.GC;
  if (has_dialog) {
    lin_bwd_w(L, gX0, D, s.H1, D, gp(grads, DP_FUS2_W), D, gp(grads, DP_FUS2_B), rows_cap, D, D, s.total);
    lin_bwd_x(L, gX0, D, params[DP_FUS2_W], D, s.GH1, D, rows_cap, D, D, 0, s.total, s.tf.WT);
    relu_bwd(L, s.GH1, s.H1, s.total, rows_cap, D);
    lin_bwd_w(L, s.GH1, D, s.XIN, 2 * D, gp(grads, DP_FUS0_W), 2 * D, gp(grads, DP_FUS0_B), rows_cap, D, 2 * D, s.total);
    if (dx_att || dd_emb) {
      lin_bwd_x(L, s.GH1, D, params[DP_FUS0_W], 2 * D, s.GXIN, 2 * D, rows_cap, D, 2 * D, 0, s.total, s.tf.WT);
      AVL_LAUNCH(dlg_scatter_kernel, B, 256, 0, L.s, s.GXIN, 2 * D, s.off, D, dx_att, dd_emb);
      L.check();
    }
  } else if (dx_att) {
    AVL_LAUNCH(dlg_scatter_kernel, B, 256, 0, L.s, gX0, D, s.off, D, dx_att, (float*)nullptr);
    L.check();
  }
  if (dgoal) cudaMemcpyAsync(dgoal, s.tf.gB4, (size_t)B * D * sizeof(float), cudaMemcpyDeviceToDevice, L.s);
  return L.err;
}
