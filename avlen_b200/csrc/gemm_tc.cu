// Tensor-core GEMM / implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::tf32, fp32 operands read
// straight from shared memory, fp32 accumulation in TMEM), operands staged by a multi-stage cp.async
// pipeline into the canonical no-swizzle K-major UMMA layout, mbarrier-tracked MMA completion
// (tcgen05.commit), TMEM -> register epilogue (tcgen05.ld) with fused scale / bias / residual / ReLU.
//
//   C[m, n] = act( scale[n] * sum_k A(m, k) * B[n, k] + bias[n] + residual[m, n] )
//
// A is either a row-major matrix (dense linears of the SMT / heads) or the im2col view of an NHWC
// activation tensor (encoder convolutions, k = (r*KW + s)*C + ci, C % 4 == 0), gathered 16 bytes at a
// time by cp.async with zero-fill for padding / ragged edges.  B is [N][K] K-contiguous (nn.Linear
// weights as they are; conv weights packed once to (Cout, KH, KW, Cin)).
//
// Shared-memory operand layout (SWIZZLE_NONE, K-major): 16-byte chunk c = k/4 of row r lives at
//   c * plane_stride + r * 16      (8-row core matrices are contiguous: SBO = 128 B, LBO = plane_stride)
// plane_stride = rows*16 + 16 so that a warp's 16-byte cp.async writes spread over all banks.
// One MMA consumes K = 8 fp32 (two chunks).  M tile = 128 (one TMEM lane per row), N tile <= 256.
#include "nn_kernels.cuh"

#ifndef AVL_HOST_EMUL
namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;          // floats per k-tile = 8 chunks of 16 B
constexpr int TC_CHUNKS = TC_BK / 4;
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 24)) __trap();  // a lost arrival must fault, never hang the device
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// .ca variant: the im2col gather re-reads every input pixel KH*KW times from neighbouring rows of the same CTA
// tile; keeping those lines in L1 turns most of that traffic into L1 hits instead of L2 round trips.
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading byte offset,
// stride byte offset (all >> 4), version = 1 (Blackwell), layout type 0 = SWIZZLE_NONE.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, K-major A and B.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 2u << 7;                       // a_format = TF32
  d |= 2u << 10;                      // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17;      // n_dim
  d |= (uint32_t)(M >> 4) << 24;      // m_dim
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcArgs {
  const float* A;
  long long lda;     // dense: row stride of A in floats
  ConvGeom g;        // conv: geometry (C % 4 == 0)
  const float* B;    // [N][K]
  float* C;
  long long ldc;
  int M, N, K;
  int bn;            // N tile (multiple of 16, <= 256)
  int tmem_cols;     // power of two >= max(32, bn)
  const float* bias;
  const float* scale;
  const float* residual;
  long long ldr;
  int relu;
  const int* m_dev;
};

template <bool CONV, bool CA = false>
__global__ void __launch_bounds__(TC_THREADS) tc_gemm_kernel(TcArgs p) {
  AVL_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long bars[TC_STAGES + 1];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int m0 = blockIdx.x * TC_BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * p.bn;
  const int bn = p.bn;
  const int K = p.K;
  const int KT = (K + TC_BK - 1) / TC_BK;

  const uint32_t a_plane = TC_BM * 16 + 16;          // bytes
  const uint32_t b_plane = (uint32_t)bn * 16 + 16;
  const uint32_t a_stage = a_plane * TC_CHUNKS;
  const uint32_t b_stage = b_plane * TC_CHUNKS;
  const uint32_t stage_bytes = a_stage + b_stage;
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int i = 0; i <= TC_STAGES; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // ---- per-thread load coordinates: chunk column c = tid & 7 is fixed, rows (tid >> 3) + 16 j
  const int c = tid & 7;
  const int r_first = tid >> 3;
  const float* a_row[8];
  int a_ih[8], a_iw[8];
  bool a_ok[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int m = m0 + r_first + 16 * j;
    a_ok[j] = m < M;
    if (CONV) {
      int mm = a_ok[j] ? m : 0;
      int ow = mm % p.g.OW;
      int t = mm / p.g.OW;
      int oh = t % p.g.OH;
      int n = t / p.g.OH;
      a_row[j] = p.A + (long long)n * p.g.H * p.g.W * p.g.C;
      a_ih[j] = oh * p.g.stride - p.g.pad;
      a_iw[j] = ow * p.g.stride - p.g.pad;
    } else {
      a_row[j] = p.A + (long long)(a_ok[j] ? m : 0) * p.lda;
      a_ih[j] = a_iw[j] = 0;
    }
  }

  auto load_tile = [&](int kt, int slot) {
    const int k = kt * TC_BK + 4 * c;  // first element of this thread's chunk
    const bool k_ok = k < K;
    const uint32_t a_dst = smem_base + slot * stage_bytes + c * a_plane;
    int r = 0, s = 0, ci = 0;
    if (CONV && k_ok) {
      int tap = k / p.g.C;
      ci = k - tap * p.g.C;
      r = tap / p.g.KW;
      s = tap - r * p.g.KW;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int row = r_first + 16 * j;
      const float* src = p.A;
      uint32_t bytes = 0;
      if (a_ok[j] && k_ok) {
        if (CONV) {
          int ih = a_ih[j] + r, iw = a_iw[j] + s;
          if (ih >= 0 && ih < p.g.H && iw >= 0 && iw < p.g.W) {
            src = a_row[j] + ((long long)ih * p.g.W + iw) * p.g.C + ci;
            bytes = 16;
          }
        } else {
          src = a_row[j] + k;
          bytes = 16;
        }
      }
      if (CA) cp_async16_ca(a_dst + row * 16, src, bytes);
      else cp_async16(a_dst + row * 16, src, bytes);
    }
    const uint32_t b_dst = smem_base + slot * stage_bytes + a_stage + c * b_plane;
    for (int row = r_first; row < bn; row += 16) {
      const int n = n0 + row;
      const bool ok = (n < p.N) && k_ok;
      const float* src = ok ? p.B + (long long)n * K + k : p.B;
      cp_async16(b_dst + row * 16, src, ok ? 16u : 0u);
    }
  };

  const uint32_t idesc = umma_idesc_tf32(TC_BM, bn);

  for (int s = 0; s < TC_STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    const int nxt = kt + TC_STAGES - 1;
    if (nxt < KT) {
      const int slot = nxt % TC_STAGES;
      if (nxt >= TC_STAGES) mbar_wait(smem_u32(&bars[slot]), (uint32_t)((nxt / TC_STAGES - 1) & 1));
      load_tile(nxt, slot);
    }
    cp_async_commit();
    cp_async_wait<TC_STAGES - 1>();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const int slot = kt % TC_STAGES;
      const uint32_t a_base = smem_base + slot * stage_bytes;
      const uint32_t b_base = a_base + a_stage;
#pragma unroll
      for (int q = 0; q < TC_BK / 8; ++q) {
        uint64_t ad = umma_desc(a_base + 2 * q * a_plane, a_plane, 128);
        uint64_t bd = umma_desc(b_base + 2 * q * b_plane, b_plane, 128);
        umma_tf32(tmem_base, ad, bd, idesc, (kt > 0 || q > 0) ? 1u : 0u);
      }
      umma_commit(smem_u32(&bars[slot]));
      if (kt == KT - 1) umma_commit(smem_u32(&bars[TC_STAGES]));
    }
  }
  mbar_wait(smem_u32(&bars[TC_STAGES]), 0);
  tc_fence_after();

  // ---- epilogue: thread = one output row (TMEM lane), 16 columns at a time
  const int row = warp * 32 + lane;
  const int m = m0 + row;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < bn; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    if (m < M) {
      float* crow = p.C + (long long)m * p.ldc + n0 + c0;
      const float* rrow = p.residual ? p.residual + (long long)m * p.ldr + n0 + c0 : nullptr;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int n = n0 + c0 + j;
        if (n < p.N) {
          float x = __uint_as_float(v[j]);
          if (p.scale) x *= __ldg(p.scale + n);
          if (p.bias) x += __ldg(p.bias + n);
          if (rrow) x += rrow[j];
          if (p.relu) x = fmaxf(x, 0.f);
          crow[j] = x;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

static int g_tc_ca = 1;

static int pick_bn(int N) {
  int n16 = (N + 15) / 16 * 16;
  if (n16 <= 256) return n16;
  for (int bn = 256; bn >= 64; bn -= 16)
    if (n16 % bn == 0) return bn;
  return 256;
}

static int tc_launch(bool conv, TcArgs& p, cudaStream_t s) {
  p.bn = pick_bn(p.N);
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  size_t smem = (size_t)TC_STAGES * TC_CHUNKS * ((TC_BM * 16 + 16) + ((size_t)p.bn * 16 + 16));
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    AVL_CUDA_CHECK((cudaFuncSetAttribute(tc_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)));
    attr_set = true;
  }
  dim3 grid(avl_div_up(p.M, TC_BM), avl_div_up(p.N, p.bn));
  if (conv && g_tc_ca) tc_gemm_kernel<true, true><<<grid, TC_THREADS, smem, s>>>(p);
  else if (conv) tc_gemm_kernel<true><<<grid, TC_THREADS, smem, s>>>(p);
  else tc_gemm_kernel<false><<<grid, TC_THREADS, smem, s>>>(p);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

}  // namespace

// 1 (default): im2col gathers go through L1 (cp.async.ca); 0: L2 only (cp.async.cg).  Returns the old value.
AVL_API int avl_set_tc_conv_l1(int on) {
  int old = g_tc_ca;
  g_tc_ca = on ? 1 : 0;
  return old;
}

// Dense: C[M,N] = act(scale * A[M,K] B[N,K]^T + bias + residual).  A rows / B rows must be 16-byte aligned
// (lda % 4 == 0, K % 4 == 0).  m_dev: optional device-side row count (packed SMT rows).
AVL_API int avl_tc_gemm(const float* A, long long lda, const float* B, float* C, long long ldc, int M, int N, int K,
                        const float* scale, const float* bias, const float* residual, long long ldr, int relu,
                        const int* m_dev, void* stream) {
  if (M < 0 || N < 1 || K < 1) return AVL_ERR_ARG;
  if (M == 0) return AVL_OK;
  if (!A || !B || !C) return AVL_ERR_ARG;
  if ((K & 3) || (lda & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return AVL_ERR_UNSUPPORTED;
  TcArgs p = {};
  p.A = A; p.lda = lda; p.B = B; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.scale = scale; p.residual = residual; p.ldr = ldr; p.relu = relu; p.m_dev = m_dev;
  return tc_launch(false, p, (cudaStream_t)stream);
}

// NHWC convolution on the tensor cores.  w_packed: (Cout, KH, KW, C) (k = (r*KW + s)*C + ci), C % 4 == 0.
AVL_API int avl_tc_conv2d_fwd(const float* x, int N, int H, int W, int C, const float* w_packed, int Cout, int KH,
                              int KW, int stride, int pad, const float* scale, const float* bias,
                              const float* residual, long long ldr, int relu, float* y, long long ldy, void* stream) {
  if (N < 0 || H < 1 || W < 1 || C < 1 || Cout < 1 || KH < 1 || KW < 1 || stride < 1 || pad < 0) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !w_packed || !y) return AVL_ERR_ARG;
  if ((C & 3) || ((uintptr_t)x & 15) || ((uintptr_t)w_packed & 15)) return AVL_ERR_UNSUPPORTED;
  TcArgs p = {};
  p.g.N = N; p.g.H = H; p.g.W = W; p.g.C = C; p.g.KH = KH; p.g.KW = KW; p.g.stride = stride; p.g.pad = pad;
  p.g.OH = (H + 2 * pad - KH) / stride + 1;
  p.g.OW = (W + 2 * pad - KW) / stride + 1;
  if (p.g.OH < 1 || p.g.OW < 1) return AVL_ERR_ARG;
  long long M = (long long)N * p.g.OH * p.g.OW;
  if (M > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  p.A = x; p.B = w_packed; p.C = y; p.ldc = ldy; p.M = (int)M; p.N = Cout; p.K = KH * KW * C;
  p.bias = bias; p.scale = scale; p.residual = residual; p.ldr = ldr; p.relu = relu; p.m_dev = nullptr;
  return tc_launch(true, p, (cudaStream_t)stream);
}
#endif  // AVL_HOST_EMUL
