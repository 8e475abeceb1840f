// Dense GEMM on tcgen05 fed by TMA:  C[M, N] = act(A[M, K] B[N, K]^T + bias + residual), fp32 in HBM, TF32 MMA.
//
// The cp.async kernels (gemm_tc.cu) move every 16 bytes of an operand tile with a separate LSU request, which
// tops out at ~16 B/clk/SM on this part (profiles/: ~2 us per 48 KB k-tile whatever the pipeline depth or layout).
// Here one elected thread issues two cp.async.bulk.tensor (TMA) loads per k-tile — box 32 floats x 128 rows,
// SWIZZLE_128B, out-of-range rows / columns zero-filled by the unit — straight into the K-major SWIZZLE_128B
// tiles the MMA reads; completion is tracked by the stage's mbarrier transaction count.  Warp 0 lane 0: TMA
// producer; warp 4 lane 0: MMA issuer; warps 0-3: epilogue (TMEM lane quadrant = warp).
// Used for the dense layers that run on tensor cores: the CLIP text tower (row L), FC heads, level-2 SMT linears.
#include "tc_common.cuh"

#ifndef AVL_HOST_EMUL
#include <cuda.h>

namespace {

constexpr int TM_BM = 128, TM_BK = 32, TM_MAX_STAGES = 4, TM_THREADS = 160;

struct TmaArgs {
  float* C;
  long long ldc;
  int M, N, K, bn, tmem_cols;
  int stages;     // 2..4: sized so that two CTAs fit one SM (one CTA's epilogue overlaps the other's main loop)
  int vec_store;  // rows of C (and of the residual) are 16-byte aligned
  const float* bias;
  const float* residual;
  long long ldr;
  int relu;
  const int* m_dev;
};

__global__ void __launch_bounds__(TM_THREADS) tc_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, TmaArgs p) {
  AVL_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long bars[2 * TM_MAX_STAGES + 1];
  const int TM_STAGES = p.stages;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int m0 = blockIdx.x * TM_BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * p.bn;
  const int bn = p.bn;
  const int KT = (p.K + TM_BK - 1) / TM_BK;
  const uint32_t a_stage = TM_BM * 128u, b_stage = (uint32_t)bn * 128u, stage_bytes = a_stage + b_stage;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (TM_STAGES + s); };
  const uint32_t DONE = bar0 + 8u * (2 * TM_STAGES);
  if (tid == 0) {
    for (int i = 0; i < TM_STAGES; ++i) {
      mbar_init(FULL(i), 1);
      mbar_init(EMPTY(i), 1);
    }
    mbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TM_BM, bn);
      for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % TM_STAGES;
        mbar_wait(FULL(slot), (uint32_t)((kt / TM_STAGES) & 1));
        tc_fence_after();
        const uint32_t a_base = smem_base + slot * stage_bytes;
        const uint64_t ad0 = umma_desc_sw128(a_base), bd0 = umma_desc_sw128(a_base + a_stage);
#pragma unroll
        for (int q = 0; q < TM_BK / 8; ++q)
          umma_tf32(tmem_base, ad0 + 2u * q, bd0 + 2u * q, idesc, (kt > 0 || q > 0) ? 1u : 0u);
        umma_commit(EMPTY(slot));
      }
      umma_commit(DONE);
      for (int kt = max(0, KT - TM_STAGES); kt < KT; ++kt) mbar_wait(EMPTY(kt % TM_STAGES), (uint32_t)((kt / TM_STAGES) & 1));
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    return;
  }
  if (tid == 0) {  // TMA producer
    for (int kt = 0; kt < KT; ++kt) {
      const int slot = kt % TM_STAGES;
      if (kt >= TM_STAGES) mbar_wait(EMPTY(slot), (uint32_t)((kt / TM_STAGES - 1) & 1));
      const uint32_t a_dst = smem_base + slot * stage_bytes;
      mbar_arrive_expect_tx(FULL(slot), stage_bytes);
      tma_load_2d(a_dst, &tmA, kt * TM_BK, m0, FULL(slot));
      tma_load_2d(a_dst + a_stage, &tmB, kt * TM_BK, n0, FULL(slot));
    }
  }
  __syncwarp();
  mbar_wait(DONE, 0);
  tc_fence_after();
  // ---- epilogue: thread = one output row (TMEM lane), 16 columns at a time
  const int m = m0 + warp * 32 + lane;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < bn; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    if (m < M) {
      float* crow = p.C + (long long)m * p.ldc + n0 + c0;
      const float* rrow = p.residual ? p.residual + (long long)m * p.ldr + n0 + c0 : nullptr;
      if (p.vec_store && n0 + c0 + 16 <= p.N) {  // 16-byte stores: a 4-byte store per lane rewrites every sector 8 times
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 x = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
          if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
            x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
          }
          if (rrow) {
            const float4 r = *reinterpret_cast<const float4*>(rrow + j);
            x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
          }
          if (p.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
          *reinterpret_cast<float4*>(crow + j) = x;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c0 + j;
          if (n < p.N) {
            float x = __uint_as_float(v[j]);
            if (p.bias) x += __ldg(p.bias + n);
            if (rrow) x += rrow[j];
            if (p.relu) x = fmaxf(x, 0.f);
            crow[j] = x;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
}

// cuTensorMapEncodeTiled is a DRIVER entry point: resolved at run time through the runtime API so that the library
// carries no link-time dependency on libcuda (it must load, and export its symbols, on machines without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

bool make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int g_tma_on = 1;

}  // namespace

AVL_API int avl_set_tc_tma(int on) {
  avl_bump_config_epoch();
  int old = g_tma_on;
  g_tma_on = on ? 1 : 0;
  return old;
}

// Returns AVL_ERR_UNSUPPORTED (nothing launched) when the operands do not meet TMA's alignment rules (16-byte aligned
// bases, row strides that are multiples of 16 bytes) or the path is switched off; avl_tc_gemm then uses cp.async.
int avl_tc_gemm_tma_try(const float* A, long long lda, const float* B, float* C, long long ldc, int M, int N, int K,
                        const float* bias, const float* residual, long long ldr, int relu, const int* m_dev,
                        cudaStream_t stream) {
  if (!g_tma_on) return AVL_ERR_UNSUPPORTED;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 3) || (K & 3) || M < 1 || N < 1 || K < 1) return AVL_ERR_UNSUPPORTED;
  TmaArgs p = {};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.residual = residual; p.ldr = ldr; p.relu = relu;
  p.m_dev = m_dev;
  // N tile: the widest of 256 / 128 / 64 that still yields at least one CTA per SM (A tiles are re-read from L2)
  const int n16 = (N + 15) / 16 * 16;
  const int mtiles = avl_div_up(M, TM_BM), sms = avl_num_sms();
  p.bn = 64;
  for (int bn = 256; bn >= 64; bn >>= 1)
    if ((long long)mtiles * avl_div_up(N, bn) >= sms) { p.bn = bn; break; }
  if (p.bn > n16) p.bn = n16;
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap ta, tb;
  if (!make_map(&ta, A, M, K, lda, TM_BM) || !make_map(&tb, B, N, K, K, p.bn)) return AVL_ERR_UNSUPPORTED;
  const size_t stage = (TM_BM + (size_t)p.bn) * 128;
  p.stages = (int)((100 * 1024) / stage);
  if (p.stages > TM_MAX_STAGES) p.stages = TM_MAX_STAGES;
  if (p.stages < 2) p.stages = 2;
  p.vec_store = ((ldc & 3) == 0 && ((uintptr_t)C & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0) &&
                 (!residual || ((ldr & 3) == 0 && ((uintptr_t)residual & 15) == 0))) ? 1 : 0;
  const size_t smem = (size_t)p.stages * stage;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(avl_div_up(M, TM_BM), avl_div_up(N, p.bn));
  tc_gemm_tma_kernel<<<grid, TM_THREADS, smem, stream>>>(ta, tb, p);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
