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
#include <cuda_fp16.h>

namespace {

constexpr int TM_BM = 128, TM_BK = 32, TM_MAX_STAGES = 4, TM_THREADS = 160;

struct TmaArgs {
  float* C;
  long long ldc;
  int M, N, K, bn, tmem_cols;
  int stages;     // 2..4: sized so that two CTAs fit one SM (one CTA's epilogue overlaps the other's main loop)
  int vec_store;  // rows of C (and of the residual) are 16-byte aligned
  int st256;      // rows of C are 32-byte aligned: 32-byte stores
  const float* bias;
  const float* scale;
  const float* residual;
  long long ldr;
  int relu;
  const int* m_dev;
  // CONV: A is the im2col view of an NHWC tensor, loaded by TMA in im2col mode (one box = 128 output pixels x 32 channels
  // of one filter tap); k-tile kt = tap * cblocks + channel block
  int OH, OW, Cin, KW, conv_stride, pad, cblocks;
  // pixel shuffle (stride-2 data gradients): the N = 4 * ps_c output columns of row (n, oh, ow) are the ps_c channels of the
  // four pixels (2 oh + pa, 2 ow + pb), column block pa * 2 + pb; C then points at a (N, 2 OH, 2 OW, ps_c) tensor
  int ps_c;
  // F16 instantiation (fp16 A / B through TMA, kind::f16): C is fp32, or fp16 when out16; act 2 = QuickGELU x * sigmoid(1.702 x)
  int out16, act;
};

// cp.async.bulk.tensor im2col mode: {c, w, h, n} = channel offset and the input coordinates of the FIRST output pixel of
// the tile's tap (0, 0) (ow * stride - pad, oh * stride - pad: inside the map's pixel bounding box); the unit walks 128
// output pixels from there (W, then H, then N, stepping by the traversal strides of the map), adds the tap offsets and
// zero-fills whatever falls outside the tensor.
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const void* tmap, int c, int w, int h, int n,
                                                   unsigned short off_w, unsigned short off_h, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
      "{%7, %8};"
      ::"r"(dst), "l"(tmap), "r"(mbar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

template <bool CONV, bool F16 = false>
__global__ void __launch_bounds__(TM_THREADS) tc_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, TmaArgs p) {
  AVL_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long bars[2 * TM_MAX_STAGES + 1];
  const int TM_STAGES = p.stages;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M;
  if (p.m_dev) {
    avl_pdl_wait();  // the device-side row count may come from the kernel just before this one
    M = min(M, *p.m_dev);
  }
  const int m0 = blockIdx.x * TM_BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * p.bn;
  const int bn = p.bn;
  constexpr int BKE = F16 ? 64 : TM_BK;  // elements per 128-byte k-tile row
  const int KT = CONV ? p.K : (p.K + BKE - 1) / BKE;  // (CONV: the host passes the number of k-tiles)
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
  avl_pdl_wait();
  avl_pdl_trigger();

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = F16 ? umma_idesc_f16_kmajor(TM_BM, bn) : umma_idesc_tf32(TM_BM, bn);
      for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % TM_STAGES;
        mbar_wait(FULL(slot), (uint32_t)((kt / TM_STAGES) & 1));
        tc_fence_after();
        const uint32_t a_base = smem_base + slot * stage_bytes;
        const uint64_t ad0 = umma_desc_sw128(a_base), bd0 = umma_desc_sw128(a_base + a_stage);
#pragma unroll
        for (int q = 0; q < TM_BK / 8; ++q) {  // four MMAs per 128-byte row: K = 8 (tf32) or 16 (f16) each, 32 bytes apart
          if (F16) umma_f16(tmem_base, ad0 + 2u * q, bd0 + 2u * q, idesc, (kt > 0 || q > 0) ? 1u : 0u);
          else umma_tf32(tmem_base, ad0 + 2u * q, bd0 + 2u * q, idesc, (kt > 0 || q > 0) ? 1u : 0u);
        }
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
    int cw = 0, ch = 0, cn = 0;
    if (CONV) {
      const int ow = m0 % p.OW, t = m0 / p.OW;
      cw = ow * p.conv_stride - p.pad;
      ch = (t % p.OH) * p.conv_stride - p.pad;
      cn = t / p.OH;
    }
    int tap = 0, cb = 0;
    for (int kt = 0; kt < KT; ++kt) {
      const int slot = kt % TM_STAGES;
      if (kt >= TM_STAGES) mbar_wait(EMPTY(slot), (uint32_t)((kt / TM_STAGES - 1) & 1));
      const uint32_t a_dst = smem_base + slot * stage_bytes;
      mbar_arrive_expect_tx(FULL(slot), stage_bytes);
      if (CONV) {
        const int r = tap / p.KW, sx = tap - r * p.KW;
        tma_load_im2col_4d(a_dst, &tmA, cb * TM_BK, cw, ch, cn, (unsigned short)sx, (unsigned short)r, FULL(slot));
        tma_load_2d(a_dst + a_stage, &tmB, tap * p.Cin + cb * TM_BK, n0, FULL(slot));
        if (++cb == p.cblocks) { cb = 0; ++tap; }
      } else {
        tma_load_2d(a_dst, &tmA, kt * BKE, m0, FULL(slot));
        tma_load_2d(a_dst + a_stage, &tmB, kt * BKE, n0, FULL(slot));
      }
    }
  }
  __syncwarp();
  mbar_wait(DONE, 0);
  tc_fence_after();
  // ---- epilogue: thread = one output row (TMEM lane), 16 columns at a time
  const int m = m0 + warp * 32 + lane;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  long long ps_pix = 0;  // CONV pixel shuffle: pixel (2 oh, 2 ow) of the doubled grid
  if (CONV && p.ps_c) {
    const int mm = m < M ? m : 0;
    const int ow = mm % p.OW, t = mm / p.OW;
    ps_pix = ((long long)(t / p.OH) * (2 * p.OH) + 2 * (t % p.OH)) * (2 * p.OW) + 2 * ow;
  }
  for (int c0 = 0; c0 < bn; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    if (m < M) {
      float* crow = p.C + (long long)m * p.ldc + n0 + c0;
      if (CONV && p.ps_c) {  // (ps_c is a multiple of 16: a 16-column chunk never straddles two pixels)
        const int q = (n0 + c0) / p.ps_c, c = (n0 + c0) - q * p.ps_c;
        crow = p.C + (ps_pix + (long long)(q >> 1) * (2 * p.OW) + (q & 1)) * p.ldc + c;
      }
      const float* rrow = p.residual ? p.residual + (long long)m * p.ldr + n0 + c0 : nullptr;
      if (F16 && (p.out16 || p.act)) {  // (host: N % 16 == 0, rows 16-byte aligned) bias -> residual -> activation -> store
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = __uint_as_float(v[j]);
          if (p.bias) x += __ldg(p.bias + n0 + c0 + j);
          if (rrow) x += rrow[j];
          if (p.act == 2) x = x / (1.f + __expf(-1.702f * x));
          o[j] = x;
        }
        if (p.out16) {
          __half2 h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(o[2 * j], o[2 * j + 1]);
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.C) + (long long)m * p.ldc + n0 + c0);
          dst[0] = *reinterpret_cast<const uint4*>(&h[0]);
          dst[1] = *reinterpret_cast<const uint4*>(&h[4]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
      } else if (p.vec_store && n0 + c0 + 16 <= p.N) {  // 16-byte stores: a 4-byte store per lane rewrites every sector 8 times
        float4 xs[4];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 x = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
          if (p.scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + c0 + j));
            x.x *= sc.x; x.y *= sc.y; x.z *= sc.z; x.w *= sc.w;
          }
          if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
            x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
          }
          if (rrow) {
            const float4 r = *reinterpret_cast<const float4*>(rrow + j);
            x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
          }
          if (p.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
          xs[j >> 2] = x;
        }
        st_row16(crow, xs, p.st256 != 0);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c0 + j;
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

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeIm2colFn encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(ptr);
  }
  return fn;
}

// NHWC fp32 activations as a rank-4 im2col map (C, W, H, N): the pixel bounding box holds the input coordinates of tap
// (0, 0) of every output pixel — [-pad, size + pad - (K - 1)) per spatial dimension — traversed with the convolution's
// stride; one load = 128 pixels x 32 channels, SWIZZLE_128B (= the K-major UMMA operand tile).
bool make_im2col_map(CUtensorMap* map, const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                     int pad_hi = -1) {
  EncodeIm2colFn enc = encode_im2col();
  if (!enc) return false;
  if (pad_hi < 0) pad_hi = pad;  // (asymmetric padding: pad rows / columns before, pad_hi after)
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * sizeof(float), (cuuint64_t)W * C * sizeof(float), (cuuint64_t)H * W * C * sizeof(float)};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad_hi - (KW - 1), pad_hi - (KH - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, lower, upper,
                   (cuuint32_t)TM_BK, (cuuint32_t)TM_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int g_tma_on = 1;
int g_tma_conv_on = 1;
long long g_tma_conv_launches = 0;

}  // namespace

// 1 (default): convolutions with >= 16 input channels and enough output tiles to fill the GPU are fed by TMA in im2col
// mode (no per-16-byte gather instructions); 0: the cp.async im2col kernel everywhere.  Returns the old value.
AVL_API int avl_set_tc_conv_tma(int on) {
  avl_bump_config_epoch();
  int old = g_tma_conv_on;
  g_tma_conv_on = on ? 1 : 0;
  return old;
}

// The two tensor maps of a TMA-fed convolution (gemm_tc.cu builds them for its split-K kernel): im2col-mode map of the
// NHWC activations, tiled map of the packed weights with `bn` rows per box.  False: the driver entry point is missing or
// the shape is outside the maps' limits.
extern "C" bool avl_conv_tma_maps(CUtensorMap* ta, CUtensorMap* tb, const float* x, int N, int H, int W, int C,
                                  const float* w_packed, int Cout, int KH, int KW, int stride, int pad, int bn) {
  if (!g_tma_conv_on || KH != KW || C < 16 || (C & 3) || stride < 1 || stride > 8 || pad > 127 || KW - 1 - pad > 127 ||
      ((uintptr_t)x & 15) || ((uintptr_t)w_packed & 15) || H + 2 * pad - (KH - 1) < 1 || W + 2 * pad - (KW - 1) < 1)
    return false;
  return make_im2col_map(ta, x, N, H, W, C, KH, KW, stride, pad) && make_map(tb, w_packed, Cout, (long long)KH * KW * C,
                                                                             (long long)KH * KW * C, bn);
}

// Data gradient of a 3x3, stride-2, pad-1 convolution WITHOUT the zero-upsampled intermediate: dx[2a + pa, 2b + pb] only
// sees dy[a + u, b + v], u, v in {0, 1} (rows a / a + 1 through kernel rows 1 | 2, 0 depending on the parity), so it is ONE
// 2x2-tap stride-1 convolution of dy (no padding before, one row / column after) with 4 * Cin output columns — a column
// block per output parity — whose epilogue stores the four pixels of the doubled grid.  16 Cout Cin MACs per dy pixel
// instead of 36 on the upsampled tensor, and the 4x larger intermediate is never written or read.
// w2: [4 * Cin][4 * Cout] (row = (pa, pb, ci), column = (u, v, co)); dx: (N, 2 OH, 2 OW, Cin).
AVL_API int avl_tc_conv2d_dgrad_s2(const float* dy, int N, int OH, int OW, int Cout, const float* w2, int Cin, float* dx,
                                   void* stream) {
  if (N < 0 || OH < 1 || OW < 1 || Cout < 1 || Cin < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!dy || !w2 || !dx) return AVL_ERR_ARG;
  if (!g_tma_conv_on || Cout < 16 || (Cout & 3) || (Cin & 15) || ((uintptr_t)dy & 15) || ((uintptr_t)w2 & 15) ||
      ((uintptr_t)dx & 15))
    return AVL_ERR_UNSUPPORTED;
  const long long M = (long long)N * OH * OW;
  if (M > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  const int Ng = 4 * Cin, K = 4 * Cout;
  TmaArgs p = {};
  p.C = dx; p.ldc = Cin; p.M = (int)M; p.N = Ng; p.relu = 0; p.m_dev = nullptr;
  p.OH = OH; p.OW = OW; p.Cin = Cout; p.KW = 2; p.conv_stride = 1; p.pad = 0; p.ps_c = Cin;
  p.bn = Ng < 128 ? Ng : 128;
  const int mtiles = avl_div_up(M, TM_BM);
  if ((long long)mtiles * avl_div_up(Ng, p.bn) < 2LL * avl_num_sms()) return AVL_ERR_UNSUPPORTED;
  p.cblocks = avl_div_up(Cout, TM_BK);
  p.K = 4 * p.cblocks;
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap ta, tb;
  if (!make_im2col_map(&ta, dy, N, OH, OW, Cout, 2, 2, 1, 0, 1) || !make_map(&tb, w2, Ng, K, K, p.bn)) return AVL_ERR_UNSUPPORTED;
  const size_t stage = (TM_BM + (size_t)p.bn) * 128;
  p.stages = (int)((100 * 1024) / stage);
  if (p.stages > TM_MAX_STAGES) p.stages = TM_MAX_STAGES;
  if (p.stages < 2) p.stages = 2;
  p.vec_store = 1;
  p.st256 = avl_rows_32b(p.C, p.ldc, 4);
  const size_t smem = (size_t)p.stages * stage;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(mtiles, avl_div_up(Ng, p.bn));
  tc_gemm_tma_kernel<true><<<grid, TM_THREADS, smem, (cudaStream_t)stream>>>(ta, tb, p);
  AVL_LAUNCH_CHECK();
  ++g_tma_conv_launches;
  return AVL_OK;
}

// Number of convolutions this process has launched on the TMA im2col kernel (tests assert that the path is taken).
AVL_API long long avl_tc_conv_tma_count(void) { return g_tma_conv_launches; }
extern "C" void avl_tc_conv_tma_count_add() { ++g_tma_conv_launches; }

// Implicit-GEMM convolution fed by TMA (im2col mode).  AVL_ERR_UNSUPPORTED = not taken (avl_tc_conv2d_fwd falls through
// to the cp.async kernel): switched off, shape outside the map's limits, or too few output tiles (those go split-K).
int avl_tc_conv_tma_try(const float* x, int N, int H, int W, int C, const float* w_packed, int Cout, int KH, int KW,
                        int stride, int pad, const float* scale, const float* bias, const float* residual, long long ldr,
                        int relu, float* y, long long ldy, cudaStream_t stream) {
  if (!g_tma_conv_on) return AVL_ERR_UNSUPPORTED;
  if (KH != KW || C < 16 || (C & 3) || stride < 1 || stride > 8 || pad > 127 || KW - 1 - pad > 127 ||
      ((uintptr_t)x & 15) || ((uintptr_t)w_packed & 15))
    return AVL_ERR_UNSUPPORTED;
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  if (OH < 1 || OW < 1 || H + 2 * pad - (KH - 1) < 1 || W + 2 * pad - (KW - 1) < 1) return AVL_ERR_UNSUPPORTED;
  const long long M = (long long)N * OH * OW;
  if (M > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  const int K = KH * KW * C;
  TmaArgs p = {};
  p.C = y; p.ldc = ldy; p.M = (int)M; p.N = Cout; p.bias = bias; p.scale = scale; p.residual = residual; p.ldr = ldr;
  p.relu = relu; p.m_dev = nullptr;
  p.OH = OH; p.OW = OW; p.Cin = C; p.KW = KW; p.conv_stride = stride; p.pad = pad;
  const int n16 = (Cout + 15) / 16 * 16;
  const int mtiles = avl_div_up(M, TM_BM), sms = avl_num_sms();
  p.bn = n16 < 128 ? n16 : 128;
  if ((long long)mtiles * avl_div_up(Cout, p.bn) < 2LL * sms) return AVL_ERR_UNSUPPORTED;  // small grids: split-K kernel
  p.cblocks = avl_div_up(C, TM_BK);
  p.K = KH * KW * p.cblocks;  // number of k-tiles (the kernel's CONV loop bound)
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap ta, tb;
  if (!make_im2col_map(&ta, x, N, H, W, C, KH, KW, stride, pad) || !make_map(&tb, w_packed, Cout, K, K, p.bn))
    return AVL_ERR_UNSUPPORTED;
  const size_t stage = (TM_BM + (size_t)p.bn) * 128;
  p.stages = (int)((100 * 1024) / stage);
  if (p.stages > TM_MAX_STAGES) p.stages = TM_MAX_STAGES;
  if (p.stages < 2) p.stages = 2;
  p.vec_store = ((ldy & 3) == 0 && ((uintptr_t)y & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0) &&
                 (!scale || ((uintptr_t)scale & 15) == 0) &&
                 (!residual || ((ldr & 3) == 0 && ((uintptr_t)residual & 15) == 0))) ? 1 : 0;
  p.st256 = p.vec_store && avl_rows_32b(p.C, p.ldc, 4);
  const size_t smem = (size_t)p.stages * stage;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(mtiles, avl_div_up(Cout, p.bn));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  unsigned nat = 0;
  avl_pdl_attr(at, &nat);
  cfg.attrs = at;
  cfg.numAttrs = nat;
  AVL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_gemm_tma_kernel<true>, ta, tb, p));
  AVL_LAUNCH_CHECK();
  ++g_tma_conv_launches;
  return AVL_OK;
}

AVL_API int avl_set_tc_tma(int on) {
  avl_bump_config_epoch();
  int old = g_tma_on;
  g_tma_on = on ? 1 : 0;
  return old;
}

bool make_map16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp16 operands (kind::f16, fp32 accumulation): C[M, N] = act(A[M, K] W[N, K]^T + bias + residual), C fp32 or fp16.
// The CLIP text tower's linears (the reference runs that tower in fp16 on CUDA, policy.py:761).  N % 16 == 0, K % 8 == 0.
extern "C" int avl_tc_gemm_tma_f16(const void* A, long long lda, const void* W, void* C, long long ldc, int M, int N, int K,
                                   const float* bias, const float* residual, long long ldr, int act, int out16,
                                   const int* m_dev, cudaStream_t stream) {
  if (!g_tma_on || M < 1 || N < 1 || K < 1 || (N & 15) || (K & 7) || (lda & 7) || ((uintptr_t)A & 15) || ((uintptr_t)W & 15) ||
      ((uintptr_t)C & 15) || (ldc & (out16 ? 7 : 3)) || (bias && ((uintptr_t)bias & 15)) ||
      (residual && (((uintptr_t)residual & 15) || (ldr & 3))))
    return AVL_ERR_UNSUPPORTED;
  TmaArgs p = {};
  p.C = reinterpret_cast<float*>(C); p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.residual = residual; p.ldr = ldr;
  p.m_dev = m_dev; p.out16 = out16; p.act = act; p.vec_store = 1; p.st256 = 0;
  const int mtiles = avl_div_up(M, TM_BM), sms = avl_num_sms();
  p.bn = 64;
  for (int bn = 256; bn >= 64; bn >>= 1)
    if ((long long)mtiles * avl_div_up(N, bn) >= sms) { p.bn = bn; break; }
  if (m_dev && K >= 512 && mtiles <= 32) p.bn = 64;  // rollout sizes: a latency chain — narrow tiles, deep ring (see the fp32 entry)
  if (p.bn > N) p.bn = N;
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap ta, tb;
  if (!make_map16(&ta, A, M, K, lda, TM_BM) || !make_map16(&tb, W, N, K, K, p.bn)) return AVL_ERR_UNSUPPORTED;
  const size_t stage = (TM_BM + (size_t)p.bn) * 128;
  p.stages = (int)((100 * 1024) / stage);
  if (p.stages > TM_MAX_STAGES) p.stages = TM_MAX_STAGES;
  if (p.stages < 2) p.stages = 2;
  const size_t smem = (size_t)p.stages * stage;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK((cudaFuncSetAttribute(tc_gemm_tma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    attr_set = true;
  }
  dim3 grid(mtiles, avl_div_up(N, p.bn));
  AVL_LAUNCH_PDL((tc_gemm_tma_kernel<false, true>), grid, TM_THREADS, smem, stream, ta, tb, p);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
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
  // a device-side row count means M is an upper bound (packed transformer rows, the CLIP tower's distinct sequences):
  // at rollout sizes a few row tiles are live, so the launch is a latency chain of K / 32 k-tiles per CTA — narrow tiles (4
  // ring slots instead of 2, 4x the CTAs) keep more loads in flight; large problems keep the wide tiles
  if (m_dev && K >= 512 && mtiles <= 32) p.bn = 64;
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
  p.st256 = p.vec_store && avl_rows_32b(p.C, p.ldc, 4);
  const size_t smem = (size_t)p.stages * stage;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(avl_div_up(M, TM_BM), avl_div_up(N, p.bn));
  AVL_LAUNCH_PDL(tc_gemm_tma_kernel<false>, grid, TM_THREADS, smem, stream, ta, tb, p);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
