// Stride-1 "same" convolution (3x3 / 7x7, pad = K/2) on tcgen05 tensor cores WITHOUT im2col expansion.
//
// The generic implicit-GEMM kernel (gemm_tc.cu) gathers every input pixel KH*KW times into shared memory; at the
// shallow encoder layers (C = 4..32 channels, the bulk of custom_resnet18's bytes, smt_resnet.py:56-164) that
// 9x / 49x shared-memory fill — not HBM, not the tensor pipe — is what bounds it (profiles/r01_tc_conv_layer1*).
// Here a CTA stages a strip of input rows ONCE, zero halo included, as 16-byte channel-chunk planes
//     plane c (channels 4c..4c+3):  pixel (ir, ic) of the padded strip at  (ir * Wp + ic) * 16 bytes,  Wp = W + KW - 1
// and computes the outputs ON THE PADDED GRID: output q = oh * Wp + ow reads, for tap (r, s), padded pixel
// q + r * Wp + s.  Input and output share one pitch, so for ANY run of 128 consecutive q the A operand of tap (r, s)
// is the same plane shifted by (r * Wp + s) * 16 bytes: a plain K-major SWIZZLE_NONE UMMA descriptor (rows 16 bytes
// apart, SBO = 128, LBO = plane stride).  The KW - 1 outputs per row that fall on halo columns are computed and
// dropped (3 % at W = 64).  Per 128 outputs the tensor core runs KH * KW * C / 8 MMAs (M 128, N = Cout, K 8) straight
// from the staged strip.  The CTA is warp-specialised and persistent: loader warps cp.async strip i + 1 into the second
// strip buffer while one elected thread issues the MMAs of strip i and four epilogue warps drain the double-buffered
// TMEM accumulators (tcgen05.ld, scale / bias / residual / ReLU, NHWC store); the roles meet only at mbarriers.  C == 4 (conv1 on the channel-padded rgb / depth
// input) pairs horizontally adjacent taps in one K = 8 MMA by setting LBO = 16 bytes (the next pixel).
//
// HBM traffic = read x once (+ (KH-1)/R halo rows, L2 hits) + write y once: the algorithmic bytes of DESIGN.md §4.
//
// Pixel groups (G = 2 or 4, layers with Cout <= 32): an MMA with N = Cout = 16 costs as much as one with N = 64 —
// it is paced by fetching its 4 KB A operand from shared memory, not by its math.  So G horizontally adjacent
// output pixels are computed by ONE MMA row: row m of a tile is the pixel group g = q / G, the accumulator columns
// are (dx, cout), and the weight matrix of tap (r, s') — s' in [0, KW + G - 1) — holds W[r, s' - dx] in the column
// block dx (zero where s' - dx falls outside the kernel).  KH * (KW + G - 1) taps instead of G * KH * KW: 18 MMAs per
// 512 outputs instead of 36 at 3x3, G = 4.  For the A operand of tap (r, s') to be rows 16 bytes apart again, the
// padded strip is stored as G phase planes (pixel index P -> plane P % G, slot P / G) and the pitch Wp is a multiple of G.
// MEASURED (B200, batch 4800, layer1 3x3 16->16 @64x64): fp16 429 -> 559 us, TF32 863 -> 1178 us with G = 4 — half the
// MMAs but each N = 64 MMA costs ~2.6x an N = 16 one, so the premise (cost independent of N) does not hold; the path
// is parity-tested and left OFF (avl_set_tc_conv_halo_group).
//
// fp16 activation storage (template IN16 / OUT16): the widest tensors of the encoders (stem output and layer1 of
// custom_resnet18, 64x64x16 per frame) can be kept in HBM as fp16 — the same 10-bit mantissa the TF32 tensor core
// keeps of an fp32 operand, rounded to nearest instead of truncated.  A chunk plane then holds 8 channels per
// 16 bytes and one kind::f16 MMA covers K = 16 channels with the operand bytes of a K = 8 TF32 MMA: half the HBM
// bytes and half the MMAs (this kernel is paced by the tensor core's shared-memory operand fetch).  Accumulation
// stays fp32 in TMEM; weights are the TF32-rounded fp32 weights converted (exactly) to fp16.
#include <cuda_fp16.h>

#include "tc_common.cuh"

#ifndef AVL_HOST_EMUL
#include <cuda.h>
#include <string.h>
namespace {

// TMA variant of the strip loader: the NHWC tensor as a rank-5 tiled map (chunk elements, chunk, W, H, N); one load per
// 16-byte chunk plane brings the whole padded strip of that plane — box (16 bytes, 1, Wp, R + KH - 1, 1) starting at
// (w, h) = (-pad, oh0 - pad) — and the unit zero-fills the halo columns and the rows outside the image.  The landed box
// is exactly the plane layout the MMAs read (pixel (ir, ic) at (ir * Wp + ic) * 16 bytes).
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, int c4,
                                            uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// Round-to-nearest conversion that SATURATES to +-65504 instead of producing inf (fp16 has a 5-bit exponent; the
// GroupNorm that consumes the tensor flags saturated values, gn_cluster.cu g_f16_overflow).
__device__ __forceinline__ __half2 floats2half2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}


constexpr int HL_TILE = 128;
constexpr int HL_MAX_MMA = 256;   // MMAs per tile: KH * KW * C / 8 (C == 4: KH * ceil(KW / 2))
constexpr int HL_PARAM_MMA = 96;  // descriptor increments of up to this many MMAs travel in kernel-parameter (constant) space
constexpr int HL_EPI_WARPS = 4;    // warps 0-3: epilogue (TMEM lane quadrant = warp index)
constexpr int HL_MMA_WARP = 4;     // warp 4: one elected thread issues every tcgen05.mma
constexpr int HL_LOAD_WARPS = 4;   // warps 5-8: cp.async producers of the next strip
constexpr int HL_THREADS = 32 * (HL_EPI_WARPS + 1 + HL_LOAD_WARPS);
constexpr int HL_LOAD_THREADS = 32 * HL_LOAD_WARPS;

struct HaloArgs {
  const void* x;   // fp32, or fp16 when IN16
  const void* w;   // packed (Cout, KH, KW, C), same element type as x
  void* y;         // fp32, or fp16 when OUT16
  const float* bias;
  const float* scale;
  const float* residual;
  long long ldy, ldr;
  int relu;
  int vec_store;       // y rows are 16-byte aligned
  int st256;           // y rows (and every 16-channel group) are 32-byte aligned: 32-byte stores
  int N, H, W, C, Cout, KH, KW;
  int os, OH, OW;      // output stride (1, or 2: only the even positions of the stride-1 grid are stored) / output map
  int R;               // output rows per strip
  int Wp;              // padded pitch W + KW - 1
  int tiles;           // ceil(R * Wp / 128)
  int strips_per_img;  // ceil(H / R)
  int total_strips;
  int G;               // pixels per MMA row (1, 2 or 4)
  int g_shift;         // log2(G)
  int kwe;             // taps per kernel row: KW + G - 1
  uint32_t ppu;        // 16-byte slots per phase plane (in_plane = G * ppu * 16)
  int nc;              // 16-byte chunk planes of the input: C / 4 (fp32) or C / 8 (fp16)
  int kwp;             // C == 4: KW rounded up to even (weight planes per kernel row), else KW
  uint32_t in_plane;   // bytes per input chunk plane
  uint32_t w_plane;    // bytes per weight chunk plane (Cout * 16)
  int n_wplanes;
  int n_mma;           // MMAs per 128-output tile
  int tma;             // strips arrive by TMA (one thread, one load per chunk plane) instead of cp.async gathers
  int ncols;           // TMEM columns per accumulator buffer (Cout)
  int tmem_cols;       // allocation (power of two >= 32)
  uint32_t off_a[HL_PARAM_MMA], off_b[HL_PARAM_MMA];  // per-MMA descriptor increments (16-byte units), n_mma <= 96
};

// Warp-specialised, persistent over strips.  Three roles run their own loops over the same (strip, tile) sequence
// and meet only at mbarriers:
//   loaders  : wait in_empty[b]  -> cp.async the strip into buffer b -> in_full[b]
//   MMA      : wait in_full[b]; per tile: wait acc_empty[a] -> KH*KW*C/8 MMAs -> commit acc_full[a]; after the strip's
//              last tile: commit in_empty[b]
//   epilogue : wait acc_full[a] -> tcgen05.ld -> acc_empty[a] -> scale / bias / residual / ReLU -> NHWC store
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {  // kind::f16, fp16 A / B (format 0), fp32 accumulate
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
template <bool F16>
__device__ __forceinline__ void umma_halo_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  if (F16) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_tf32_elect(tmem_d, a_desc, b_desc, idesc, accumulate);
  }
}

template <bool IN16, bool OUT16>
__global__ void __launch_bounds__(HL_THREADS) tc_conv_halo_kernel(const __grid_constant__ HaloArgs p,
                                                                  const __grid_constant__ CUtensorMap tmX) {
  AVL_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long bars[8];  // in_full[2], in_empty[2], acc_full[2], acc_empty[2]
  __shared__ uint32_t tmem_base_smem;
  // per-MMA descriptor increments (16-byte units) for the (r, s, k-pair) sequence of one tile: building a descriptor
  // from scratch costs ~30 dependent single-thread instructions, which (not the tensor pipe) bounded the first
  // version of this kernel at ~240 cycles per MMA (profiles/r01_halo_conv_layer1_ncu_full.txt)
  __shared__ __align__(8) uint2 mma_off[HL_MAX_MMA];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t w_base = smem_u32(smem);
  const uint32_t in_bytes = (uint32_t)p.nc * p.in_plane;
  const uint32_t in_base0 = w_base + (uint32_t)p.n_wplanes * p.w_plane;
  const int pad = p.KH >> 1;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto IN_FULL = [&](int b) { return bar0 + 8u * b; };
  auto IN_EMPTY = [&](int b) { return bar0 + 8u * (2 + b); };
  auto ACC_FULL = [&](int a) { return bar0 + 8u * (4 + a); };
  auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (6 + a); };

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(IN_FULL(b), p.tma ? 1 : HL_LOAD_THREADS);
      mbar_init(IN_EMPTY(b), 1);
      mbar_init(ACC_FULL(b), 1);
      mbar_init(ACC_EMPTY(b), 32 * HL_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == HL_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- one-time shared-memory setup: zero everything (halo columns, tails and the odd-KW pad plane stay zero for
  // the life of the CTA), then the weights as chunk planes  plane(tap, c)[n] = w[n, tap, 4c..4c+3]
  {
    float4* z = reinterpret_cast<float4*>(smem);
    const uint32_t n16 = ((uint32_t)p.n_wplanes * p.w_plane + 2 * in_bytes) >> 4;
    for (uint32_t i = tid; i < n16; i += HL_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  {
    const int taps = p.KH * p.KW;
    const int per_n = taps * p.nc;  // 16-byte chunks per output channel
    const int total = p.Cout * per_n;
    for (int e = tid; e < total; e += HL_THREADS) {
      const int n = e / per_n;
      const int rest = e - n * per_n;  // tap * nc + c
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.w) + e);  // 16 bytes = 4 fp32 / 8 fp16 channels
      if (p.G > 1) {  // column block dx of tap (r, s + dx) holds W[r, s]
        const int tap = rest / p.nc, c = rest - tap * p.nc;
        const int r = tap / p.KW, sx = tap - r * p.KW;
        for (int dx = 0; dx < p.G; ++dx) {
          const int plane = (r * p.kwe + sx + dx) * p.nc + c;
          *reinterpret_cast<float4*>(smem + (size_t)plane * p.w_plane + (size_t)(dx * p.Cout + n) * 16) = v;
        }
        continue;
      }
      int plane = rest;
      if (p.nc == 1) {  // C == 4: planes indexed (r, s) with kwp planes per kernel row
        const int r = rest / p.KW;
        plane = r * p.kwp + (rest - r * p.KW);
      }
      *reinterpret_cast<float4*>(smem + (size_t)plane * p.w_plane + (size_t)n * 16) = v;
    }
  }
  for (int i = tid; i < p.n_mma; i += HL_THREADS) {
    uint32_t a_off, b_off;
    if (p.nc == 1) {
      const int half = p.kwp >> 1;
      const int r = i / half, s = (i - r * half) * 2;
      a_off = (uint32_t)(r * p.Wp + s);
      b_off = (uint32_t)(r * p.kwp + s) * (p.w_plane >> 4);
    } else {
      const int pairs = p.nc >> 1;
      const int tap = i / pairs, j = (i - tap * pairs) * 2;
      const int r = tap / p.KW, s = tap - r * p.KW;
      a_off = (uint32_t)(r * p.Wp + s) + (uint32_t)j * (p.in_plane >> 4);
      b_off = (uint32_t)(tap * p.nc + j) * (p.w_plane >> 4);
    }
    mma_off[i] = make_uint2(a_off, b_off);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // everything above touched only shared memory, TMEM and the (static) weights: under programmatic dependent launch it
  // overlaps the previous kernel; activations / residual / output are touched below this line only
  avl_pdl_wait();
  avl_pdl_trigger();

  if (warp >= HL_MMA_WARP + 1) {
    // ================================================================================ loaders
    const int lt = tid - 32 * (HL_MMA_WARP + 1);
    const int rows_in = p.R + p.KH - 1;
    const int per_row = p.W * p.nc;  // 16-byte chunks per input row
    const int items = rows_in * per_row;
    const int nc_shift = (p.nc & (p.nc - 1)) == 0 ? __ffs(p.nc) - 1 : -1;
    int it_strip = 0;
    if (p.tma) {
      if (lt == 0) {
        tma_prefetch_desc(&tmX);
        const uint32_t plane_bytes = (uint32_t)(rows_in * p.Wp) * 16u;
        for (int strip = blockIdx.x; strip < p.total_strips; strip += gridDim.x, ++it_strip) {
          const int b = it_strip & 1;
          mbar_wait(IN_EMPTY(b), (uint32_t)(((it_strip >> 1) & 1) ^ 1));
          const int n = strip / p.strips_per_img;
          const int oh0 = (strip - n * p.strips_per_img) * p.R;
          const uint32_t dst0 = in_base0 + (uint32_t)b * in_bytes;
          mbar_arrive_expect_tx(IN_FULL(b), plane_bytes * (uint32_t)p.nc);
          for (int c = 0; c < p.nc; ++c) tma_load_5d(dst0 + (uint32_t)c * p.in_plane, &tmX, 0, c, -pad, oh0 - pad, n, IN_FULL(b));
        }
      }
    } else
    for (int strip = blockIdx.x; strip < p.total_strips; strip += gridDim.x, ++it_strip) {
      const int b = it_strip & 1;
      mbar_wait(IN_EMPTY(b), (uint32_t)(((it_strip >> 1) & 1) ^ 1));  // first use of each buffer passes
      const int n = strip / p.strips_per_img;
      const int oh0 = (strip - n * p.strips_per_img) * p.R;
      constexpr int ESZ = IN16 ? 2 : 4;
      const unsigned char* xin = reinterpret_cast<const unsigned char*>(p.x) + (long long)n * p.H * p.W * p.C * ESZ;
      const uint32_t dst0 = in_base0 + (uint32_t)b * in_bytes;
      // (row, chunk-in-row) advance incrementally: two integer divisions per 16-byte chunk were most of this
      // kernel's issue slots once the fp16 variant halved its MMA count (profiles/r01_halo_conv_f16_layer1_ncu_full.txt)
      int ir = lt / per_row;
      int rem = lt - ir * per_row;  // iw * nc + c == 16-byte chunk index within the image row
      const long long row_bytes = (long long)p.W * p.C * ESZ;
      for (int it = lt; it < items; it += HL_LOAD_THREADS) {
        const int iw = nc_shift >= 0 ? (rem >> nc_shift) : rem / p.nc;
        const int c = rem - iw * p.nc;
        const int ih = oh0 - pad + ir;
        const bool ok = ih >= 0 && ih < p.H;
        const void* src = ok ? (const void*)(xin + ih * row_bytes + (long long)rem * 16) : p.x;
        const uint32_t P = (uint32_t)(ir * p.Wp + pad + iw);  // padded pixel index -> (phase plane, slot)
        const uint32_t slot = p.G > 1 ? (P & (uint32_t)(p.G - 1)) * p.ppu + (P >> p.g_shift) : P;
        cp_async16(dst0 + (uint32_t)c * p.in_plane + slot * 16, src, ok ? 16u : 0u);
        rem += HL_LOAD_THREADS;
        while (rem >= per_row) { rem -= per_row; ++ir; }
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(IN_FULL(b));
    }
  } else if (warp == HL_MMA_WARP) {
    // ================================================================================ MMA issuer
    // (whole warp in the loop, one elected lane issues: see umma_tf32_elect)
    {
      const uint32_t idesc = IN16 ? umma_idesc_f16(HL_TILE, p.ncols) : umma_idesc_tf32(HL_TILE, p.ncols);
      const uint64_t bd0 = umma_desc(w_base, p.w_plane, 128);
      int it_strip = 0, it_tile = 0;
      for (int strip = blockIdx.x; strip < p.total_strips; strip += gridDim.x, ++it_strip) {
        const int b = it_strip & 1;
        const uint32_t in_base = in_base0 + (uint32_t)b * in_bytes;
        mbar_wait(IN_FULL(b), (uint32_t)((it_strip >> 1) & 1));
        tc_fence_after();
        for (int t = 0; t < p.tiles; ++t, ++it_tile) {
          const int a = it_tile & 1;
          mbar_wait(ACC_EMPTY(a), (uint32_t)(((it_tile >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(a * p.ncols);
          const uint64_t ad0 = umma_desc(in_base + (uint32_t)t * HL_TILE * 16, p.nc == 1 ? 16u : p.in_plane, 128);
          if (p.n_mma <= HL_PARAM_MMA) {
            // increments read from constant (parameter) space: uniform loads, no shared-memory latency
#pragma unroll 6
            for (int i = 0; i < p.n_mma; ++i)
              umma_halo_elect<IN16>(d, ad0 + p.off_a[i], bd0 + p.off_b[i], idesc, i > 0 ? 1u : 0u);
          } else {
#pragma unroll 4
            for (int i = 0; i < p.n_mma; ++i) {
              const uint2 o = mma_off[i];
              umma_halo_elect<IN16>(d, ad0 + o.x, bd0 + o.y, idesc, i > 0 ? 1u : 0u);
            }
          }
          umma_commit_elect(ACC_FULL(a));
        }
        umma_commit_elect(IN_EMPTY(b));  // arrives once every MMA that reads this strip buffer has completed
      }
      // no commit may still be in flight towards this CTA's barriers when the CTA retires
      for (int b = 0; b < 2; ++b) {
        const int uses = (it_strip + 1 - b) >> 1;
        if (uses > 0) mbar_wait(IN_EMPTY(b), (uint32_t)((uses - 1) & 1));
      }
    }
  } else {
    // ================================================================================ epilogue
    int it_tile = 0;
    for (int strip = blockIdx.x; strip < p.total_strips; strip += gridDim.x) {
      const int n = strip / p.strips_per_img;
      const int oh0 = (strip - n * p.strips_per_img) * p.R;
      for (int t = 0; t < p.tiles; ++t, ++it_tile) {
        const int a = it_tile & 1;
        const int q = (t * HL_TILE + warp * 32 + lane) << p.g_shift;  // first pixel of this row's group
        const int ohl = q / p.Wp;
        const int ow0 = q - ohl * p.Wp;
        const int oh = oh0 + ohl;
        // stride 2 (pad = K / 2): output (oh2, ow2) is the stride-1 output at (2 oh2, 2 ow2) — the strip is convolved at
        // stride 1 on the tensor cores (which this HBM-bound kernel has to spare) and only the even positions leave
        const bool row_ok = ohl < p.R && oh < p.H && (p.os == 1 || !(oh & 1));
        const long long pix0 = p.os == 1 ? ((long long)n * p.H + oh) * p.W + ow0
                                         : ((long long)n * p.OH + (oh >> 1)) * p.OW + (ow0 >> 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * p.ncols);
        mbar_wait(ACC_FULL(a), (uint32_t)((it_tile >> 1) & 1));
        tc_fence_after();
        for (int cc = 0; cc < p.ncols; cc += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + cc, v);
          if (cc + 16 >= p.ncols) {  // accumulator fully read: hand the TMEM buffer back before the stores
            tc_fence_before();
            mbar_arrive(ACC_EMPTY(a));
          }
          const int dx = p.G > 1 ? cc / p.Cout : 0;  // accumulator columns are (dx, cout); Cout % 16 == 0
          const int c0 = cc - dx * p.Cout;
          const bool valid = row_ok && ow0 + dx < p.W && (p.os == 1 || !(ow0 & 1));
          const long long pix = pix0 + dx;
          if (valid) {
            float* yrow = reinterpret_cast<float*>(p.y) + pix * p.ldy + c0;  // (OUT16: recomputed below)
            const float* rrow = p.residual ? p.residual + pix * p.ldr + c0 : nullptr;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float acc = __uint_as_float(v[j]);
              if (p.scale) acc *= __ldg(p.scale + c0 + j);
              if (p.bias) acc += __ldg(p.bias + c0 + j);
              if (rrow) acc += rrow[j];
              if (p.relu) acc = fmaxf(acc, 0.f);
              o[j] = acc;
            }
            if (OUT16) {  // 16 channels = 32 bytes of fp16 (rows are 16-byte aligned: checked by the host)
              __half2 h[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) h[j] = floats2half2_sat(o[2 * j], o[2 * j + 1]);
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.y) + pix * p.ldy + c0);
              if (p.st256) {  // the pixel's 32 bytes as one full-sector store
                uint32_t u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = *reinterpret_cast<const uint32_t*>(&h[j]);
                st_global_256(dst, u);
              } else {
                dst[0] = *reinterpret_cast<const uint4*>(&h[0]);
                dst[1] = *reinterpret_cast<const uint4*>(&h[4]);
              }
            } else if (p.st256) {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                uint32_t u[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(o[j + i]);
                st_global_256(yrow + j, u);
              }
            } else if (p.vec_store) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(yrow + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) yrow[j] = o[j];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == HL_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn5 halo_encode_tiled() {
  static EncodeTiledFn5 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn5>(ptr);
  }
  return fn;
}
// x: (N, H, W, C) with `cpc` channels per 16-byte chunk -> dims (cpc, C / cpc, W, H, N), box (cpc, 1, Wp, rows_in, 1)
static bool halo_make_map(CUtensorMap* map, const void* x, bool in16, int N, int H, int W, int C, int Wp, int rows_in) {
  EncodeTiledFn5 enc = halo_encode_tiled();
  if (!enc || Wp > 256 || rows_in > 256) return false;
  const int cpc = in16 ? 8 : 4;
  const cuuint64_t esz = in16 ? 2 : 4;
  cuuint64_t dims[5] = {(cuuint64_t)cpc, (cuuint64_t)(C / cpc), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[4] = {16, (cuuint64_t)C * esz, (cuuint64_t)W * C * esz, (cuuint64_t)H * W * C * esz};
  cuuint32_t box[5] = {(cuuint32_t)cpc, 1, (cuuint32_t)Wp, (cuuint32_t)rows_in, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, in16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(x), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int g_halo_on = 1;
static int g_halo_small_grid_pct = 100;
static int g_halo_tma = 1;
static int g_halo_group = 0;  // pixel groups (G > 1): measured SLOWER on B200 (see header), kept as an opt-in
static int g_halo_rows = 0;  // 0: chosen per shape (see avl_tc_conv_halo_typed)
static int g_halo_stride2 = 1;  // stride-2 same-padded convolutions: stride-1 strip convolution, even positions stored

}  // namespace

// 1 (default): stride-2 odd-kernel convolutions with pad = K / 2 (the stage-entry convolutions of the ResNet-18s) run
// on the halo-strip kernel too (computed at stride 1, only the even positions are stored); 0: im2col kernel.  Returns old.
AVL_API int avl_set_tc_conv_halo_stride2(int on) {
  avl_bump_config_epoch();
  int old = g_halo_stride2;
  g_halo_stride2 = on ? 1 : 0;
  return old;
}

// 1 (default): stride-1 same-padded convolutions with few channels use the halo-strip kernel; 0: always the
// im2col-gather kernel.  rows > 0 sets the strip height, rows == 0 selects it per shape (default), rows < 0 keeps it.
// Returns the previous on/off state.
AVL_API int avl_set_tc_conv_halo(int on, int rows) {
  avl_bump_config_epoch();
  int old = g_halo_on;
  g_halo_on = on ? 1 : 0;
  if (rows >= 0) g_halo_rows = rows;  // 0: automatic
  return old;
}

// 1 (default): the halo-strip kernel's input strips arrive by TMA (rank-5 tiled map, zero-filled halo); 0: cp.async
// gathers by four loader warps.  Returns old.
AVL_API int avl_set_tc_conv_halo_tma(int on) {
  avl_bump_config_epoch();
  int old = g_halo_tma;
  g_halo_tma = on ? 1 : 0;
  return old;
}

// Grid cap of the halo-strip kernel at rollout batches (N * H <= 8192 rows), in CTAs per 100 SMs; 0: none.  Default 100:
// one CTA per SM walking over two or three strips instead of up to three CTAs per SM — the rollout phase is bound by SM
// time across its concurrent encoder chains (rollout 61.6k -> 62.8k env-steps/s; 50: 60.7k).
AVL_API int avl_set_tc_conv_halo_small_grid(int percent) {
  avl_bump_config_epoch();
  int old = g_halo_small_grid_pct;
  g_halo_small_grid_pct = percent < 0 ? 0 : percent;
  return old;
}

// 1: layers with Cout <= 32 compute 2 / 4 adjacent pixels per MMA row; 0 (default): one pixel per row.  Returns old.
AVL_API int avl_set_tc_conv_halo_group(int on) {
  avl_bump_config_epoch();
  int old = g_halo_group;
  g_halo_group = on ? 1 : 0;
  return old;
}

// Returns AVL_ERR_UNSUPPORTED (without launching) when the shape is outside what this kernel is built for; the
// caller (avl_tc_conv2d_fwd) then falls back to the im2col-gather kernel.  in16 / out16: x (and w) / y are fp16.
int avl_tc_conv_halo_typed(const void* x, int in16, int N, int H, int W, int C, const void* w_packed, int Cout, int KH,
                           int KW, int stride, int pad, const float* scale, const float* bias, const float* residual,
                           long long ldr, int relu, void* y, int out16, long long ldy, cudaStream_t stream) {
  if (!g_halo_on) return AVL_ERR_UNSUPPORTED;
  if ((stride != 1 && stride != 2) || KH != KW || !(KH & 1) || pad != KH / 2 || KH < 3) return AVL_ERR_UNSUPPORTED;
  if (stride == 2 && !g_halo_stride2) return AVL_ERR_UNSUPPORTED;
  // measured (profiles/r02_halo_stride2_bench.txt, fp32 / TF32 operands): the stride-1 strip wins for the stage-2 entry
  // (16 -> 32 @64x64: 1.06 -> 0.88 ms at batch 4800) and loses from 32 input channels on (0.47 -> 0.59 ms): 4x the MMAs
  if (stride == 2 && !in16 && C > 16) return AVL_ERR_UNSUPPORTED;
  if (in16) {
    if ((C % 16) || C > 128) return AVL_ERR_UNSUPPORTED;
  } else if (!(C == 4 || (C % 8 == 0 && C <= 64))) {
    return AVL_ERR_UNSUPPORTED;
  }
  if ((Cout % 16) || Cout > 128) return AVL_ERR_UNSUPPORTED;
  if (W < 16 || ((uintptr_t)x & 15) || ((uintptr_t)w_packed & 15)) return AVL_ERR_UNSUPPORTED;
  if (out16 && (residual || (ldy & 7) || ((uintptr_t)y & 15))) return AVL_ERR_UNSUPPORTED;
  const int cpc = in16 ? 8 : 4;  // channels per 16-byte chunk
  HaloArgs p = {};
  p.x = x; p.w = w_packed; p.y = y; p.bias = bias; p.scale = scale; p.residual = residual; p.ldy = ldy; p.ldr = ldr;
  p.relu = relu; p.vec_store = ((ldy & 3) == 0 && ((uintptr_t)y & 15) == 0) ? 1 : 0;
  p.st256 = avl_rows_32b(y, ldy, out16 ? 2 : 4);
  p.N = N; p.H = H; p.W = W; p.C = C; p.Cout = Cout; p.KH = KH; p.KW = KW;
  p.os = stride; p.OH = (H + 2 * pad - KH) / stride + 1; p.OW = (W + 2 * pad - KW) / stride + 1;
  const int nc_ = C / cpc;
  const bool c4 = !in16 && C == 4;
  // pixels per MMA row: fill N up to 64 accumulator columns when Cout is small (see the header); kept at 1 for the
  // C == 4 stem path, for wide kernels whose descriptor table would not fit the parameter space, and when disabled
  int G = 1;
  if (g_halo_group && !c4 && KH == 3 && stride == 1) {
    if (Cout * 4 <= 64 && (W % 4) == 0) G = 4;
    else if (Cout * 2 <= 64 && (W % 2) == 0) G = 2;
    if (KH * (KW + G - 1) * (nc_ / 2) > HL_PARAM_MMA) G = 1;
  }
  p.G = G;
  p.g_shift = G == 4 ? 2 : (G == 2 ? 1 : 0);
  p.kwe = KW + G - 1;
  p.Wp = (W + KW - 1 + G - 1) / G * G;
  const int gpr = p.Wp / G;  // MMA rows (pixel groups) per padded image row
  const size_t w_bytes = (size_t)(c4 ? KH * ((KW + 1) & ~1) : KH * p.kwe * nc_) * G * Cout * 16;
  auto plane_units = [&](int tiles) {  // 16-byte slots of one phase plane
    return (size_t)tiles * HL_TILE + (size_t)((KH - 1) * p.Wp) / G + (size_t)(KW - 1 + G - 1) / G + 8;
  };
  if (g_halo_rows > 0) {
    p.R = g_halo_rows < H ? g_halo_rows : H;
  } else {
    // strip height: outputs are computed on the padded grid in tiles of 128 rows (x G pixels), so the useful
    // fraction of the MMAs is H * W / (strips * tiles * 128 * G); pick the height that maximises it among those whose
    // double-buffered strip still leaves room for two CTAs per SM (one CTA alone cannot hide its strip turnaround)
    int best = 0;
    double best_eff = -1.0;
    for (int pass = 0; pass < 2 && best == 0; ++pass) {
      const size_t budget = pass == 0 ? 108 * 1024 : 200 * 1024;
      for (int R = 1; R <= H; ++R) {
        const int tiles = avl_div_up((long long)R * gpr, HL_TILE);
        const size_t in_plane = plane_units(tiles) * G * 16;
        if (w_bytes + 2 * nc_ * in_plane > budget) break;
        const double eff = (double)H * W / ((double)avl_div_up(H, R) * tiles * HL_TILE * G);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = R; }
      }
    }
    p.R = best > 0 ? best : (8 < H ? 8 : H);
  }
  p.tiles = avl_div_up((long long)p.R * gpr, HL_TILE);
  p.strips_per_img = avl_div_up(H, p.R);
  long long total = (long long)N * p.strips_per_img;
  if (total > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  p.total_strips = (int)total;
  p.nc = nc_;
  p.kwp = c4 ? ((KW + 1) & ~1) : KW;
  p.ppu = (uint32_t)plane_units(p.tiles);
  p.in_plane = p.ppu * (uint32_t)G * 16;
  if (G == 1) p.in_plane = (p.in_plane + 127u) & ~127u;  // TMA destinations are 128-byte aligned
  p.w_plane = (uint32_t)(G * Cout) * 16;
  p.n_wplanes = c4 ? KH * p.kwp : KH * p.kwe * p.nc;
  p.n_mma = c4 ? KH * (p.kwp / 2) : KH * p.kwe * (p.nc / 2);
  if (p.n_mma > HL_MAX_MMA) return AVL_ERR_UNSUPPORTED;
  if (p.n_mma <= HL_PARAM_MMA) {
    for (int i = 0; i < p.n_mma; ++i) {
      if (p.nc == 1) {
        const int half = p.kwp >> 1;
        const int r = i / half, s2 = (i - r * half) * 2;
        p.off_a[i] = (uint32_t)(r * p.Wp + s2);
        p.off_b[i] = (uint32_t)(r * p.kwp + s2) * (p.w_plane >> 4);
      } else {
        const int pairs = p.nc >> 1;
        const int tap = i / pairs, j = (i - tap * pairs) * 2;
        const int r = tap / p.kwe, s2 = tap - r * p.kwe;  // s2 in [0, KW + G - 1)
        // padded pixel G * g + r * Wp + s2 of output group g lives in phase plane s2 % G at slot g + r * Wp / G + s2 / G
        p.off_a[i] = (uint32_t)(s2 % G) * p.ppu + (uint32_t)((r * p.Wp) / G + s2 / G) + (uint32_t)j * (p.in_plane >> 4);
        p.off_b[i] = (uint32_t)(tap * p.nc + j) * (p.w_plane >> 4);
      }
    }
  }
  p.ncols = G * Cout;
  int cols = 32;
  while (cols < 2 * p.ncols) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)p.n_wplanes * p.w_plane + 2 * (size_t)p.nc * p.in_plane;  // double-buffered strip
  if (smem > 200 * 1024 || smem + (1 << 14) > (1u << 18)) return AVL_ERR_UNSUPPORTED;  // descriptor addresses: 18 bits
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_halo_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_halo_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_halo_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_halo_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  while (per_sm > 1 && per_sm * cols > 512) --per_sm;
  long long grid = (long long)avl_num_sms() * per_sm;
  if (grid > total) grid = total;
  if (g_halo_small_grid_pct > 0 && (long long)N * H <= 8192) {  // rollout batches: see avl_set_tc_conv_halo_small_grid
    const long long cap = (long long)avl_num_sms() * g_halo_small_grid_pct / 100;
    if (grid > cap) grid = cap > 1 ? cap : 1;
  }
  CUtensorMap tmx;
  memset(&tmx, 0, sizeof(tmx));
  p.tma = 0;
  if (g_halo_tma && G == 1 && ((p.n_wplanes * p.w_plane) & 127u) == 0 &&
      halo_make_map(&tmx, x, in16 != 0, N, H, W, C, p.Wp, p.R + KH - 1))
    p.tma = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(HL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  unsigned nat = 0;
  avl_pdl_attr(at, &nat);
  cfg.attrs = at;
  cfg.numAttrs = nat;
  auto kern = in16 ? (out16 ? tc_conv_halo_kernel<true, true> : tc_conv_halo_kernel<true, false>)
                   : (out16 ? tc_conv_halo_kernel<false, true> : tc_conv_halo_kernel<false, false>);
  AVL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p, tmx));
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

int avl_tc_conv_halo_try(const float* x, int N, int H, int W, int C, const float* w_packed, int Cout, int KH, int KW,
                         int stride, int pad, const float* scale, const float* bias, const float* residual,
                         long long ldr, int relu, float* y, long long ldy, cudaStream_t stream) {
  return avl_tc_conv_halo_typed(x, 0, N, H, W, C, w_packed, Cout, KH, KW, stride, pad, scale, bias, residual, ldr, relu, y,
                                0, ldy, stream);
}

// Test / bench entry of the fp16-storage variants (x / w fp16 when in16, y fp16 when out16); -2 when not covered.
AVL_API int avl_tc_conv_halo_f16(const void* x, int in16, int N, int H, int W, int C, const void* w_packed, int Cout,
                                 int KH, int KW, int pad, int relu, void* y, int out16, void* stream) {
  if (N < 0 || H < 1 || W < 1 || C < 1 || Cout < 1) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !w_packed || !y) return AVL_ERR_ARG;
  return avl_tc_conv_halo_typed(x, in16, N, H, W, C, w_packed, Cout, KH, KW, 1, pad, nullptr, nullptr, nullptr, 0, relu, y,
                                out16, Cout, (cudaStream_t)stream);
}
#endif  // AVL_HOST_EMUL
