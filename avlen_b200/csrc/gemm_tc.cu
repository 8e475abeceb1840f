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
#include "tc_common.cuh"
#ifndef AVL_HOST_EMUL
#include <cuda.h>
#endif

#ifndef AVL_HOST_EMUL
namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;          // floats per k-tile = 8 chunks of 16 B
constexpr int TC_CHUNKS = TC_BK / 4;
constexpr int TC_MAX_STAGES = 4;      // shared-memory ring (p.stages <= 4, sized so that two CTAs fit one SM)
constexpr int TC_INFLIGHT = 2;        // cp.async groups a loader thread keeps in flight before it signals a stage full
constexpr int TC_LOAD_THREADS = 128;  // warps 0-3: loaders, then the epilogue (TMEM lane quadrant = warp index)
constexpr int TC_THREADS = 160;       // warp 4: one elected thread issues every tcgen05.mma

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
  int stages;        // 3..4 ring slots
  int vec_store;     // rows of C / residual and the scale / bias vectors are 16-byte aligned
  int st256;         // rows of C are 32-byte aligned: 32-byte stores (tc_common.cuh st_row16)
  int swz;           // 1: SWIZZLE_128B K-major operand tiles (default); 0: SWIZZLE_NONE chunk planes
  int kt_per_split;  // k-tiles per blockIdx.z slice; splits > 1: raw partial sums are atomically added into C
  int splits;
  int cluster_red;   // splits > 1 only.  1: the blockIdx.z slices of one output tile form a thread-block cluster
                     // (1, 1, splits); partial tiles meet in the owners' shared memory over DSMEM and are summed in
                     // slice order (deterministic), epilogue applied in the same kernel — no zero / epilogue launches
  int splits_nz;     // slices that own at least one k-tile (the others only take part in the reduction)
  unsigned red_off;  // byte offset of the reduction buffer in dynamic shared memory (0: it aliases the operand ring)
  // TMA mode (template parameter TMA): operands arrive by cp.async.bulk.tensor — A through an im2col-mode tensor map
  // (one box = 128 output pixels x 32 channels of one filter tap), B through a tiled map of the packed weights; k-tile
  // kt = tap * cblocks + channel block; one elected thread issues both loads of a stage
  int kt_total, cblocks;
};

__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const void* tmap, int c, int w, int h, int n,
                                                   unsigned short off_w, unsigned short off_h, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
      "{%7, %8};"
      ::"r"(dst), "l"(tmap), "r"(mbar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

__device__ __forceinline__ void cluster_arrive_wait() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t local_addr, uint32_t rank, uint32_t a, uint32_t b, uint32_t c,
                                              uint32_t d) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <bool CONV, bool CA, bool TMA>
__device__ __forceinline__ void tc_gemm_body(const TcArgs& p, const CUtensorMap* tmA, const CUtensorMap* tmB) {
  AVL_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long bars[2 * TC_MAX_STAGES + 1];  // full[S], empty[S], done
  const int TC_STAGES = p.stages;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int m0 = blockIdx.x * TC_BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * p.bn;
  const int bn = p.bn;
  const int K = p.K;
  const int kt0 = blockIdx.z * p.kt_per_split;  // split-K: this CTA reduces k-tiles [kt0, kt0 + KT)
  const int KT = max(0, min((TMA ? p.kt_total : (K + TC_BK - 1) / TC_BK) - kt0, p.kt_per_split));
  const bool cred = p.cluster_red != 0;
  if (KT <= 0 && !cred) return;  // (a cluster member without k-tiles still owns output rows of the reduction)

  // SWIZZLE_NONE: chunk c of row r at c * plane + r * 16 (plane = rows * 16 + 16).
  // SWIZZLE_128B : one k-tile row is exactly one 128-byte swizzle row: chunk c of row r at r * 128 + ((c ^ (r & 7)) * 16),
  //                8-row atoms 1024 bytes apart (SBO), tiles 1024-byte aligned; a K = 8 step advances the start by 32 B.
  const bool swz = p.swz != 0;
  const uint32_t a_plane = swz ? 0u : TC_BM * 16 + 16;  // bytes
  const uint32_t b_plane = swz ? 0u : (uint32_t)bn * 16 + 16;
  const uint32_t a_stage = swz ? TC_BM * 128u : a_plane * TC_CHUNKS;
  const uint32_t b_stage = swz ? (uint32_t)bn * 128u : b_plane * TC_CHUNKS;
  const uint32_t stage_bytes = a_stage + b_stage;
  const uint32_t smem_base = smem_u32(smem);

  const uint32_t bar0 = smem_u32(&bars[0]);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
  const uint32_t DONE = bar0 + 8u * (2 * TC_STAGES);
  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(FULL(i), TMA ? 1 : TC_LOAD_THREADS);
      mbar_init(EMPTY(i), 1);
    }
    if (TMA) {
      tma_prefetch_desc(tmA);
      tma_prefetch_desc(tmB);
    }
    mbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
  avl_pdl_wait();     // barriers / TMEM / descriptor prefetch above overlap the previous kernel under programmatic launch
  avl_pdl_trigger();

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
    const uint32_t a_dst = smem_base + slot * stage_bytes + c * a_plane;  // SWIZZLE_NONE: plane base
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
      const uint32_t dst = swz ? a_dst + row * 128 + ((uint32_t)(c ^ (row & 7)) << 4) : a_dst + row * 16;
      if (CA) cp_async16_ca(dst, src, bytes);
      else cp_async16(dst, src, bytes);
    }
    const uint32_t b_dst = smem_base + slot * stage_bytes + a_stage + c * b_plane;
    for (int row = r_first; row < bn; row += 16) {
      const int n = n0 + row;
      const bool ok = (n < p.N) && k_ok;
      const float* src = ok ? p.B + (long long)n * K + k : p.B;
      cp_async16(swz ? b_dst + row * 128 + ((uint32_t)(c ^ (row & 7)) << 4) : b_dst + row * 16, src, ok ? 16u : 0u);
    }
  };

  if (warp == 4) {
    // ================================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, bn);
      for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % TC_STAGES;
        mbar_wait(FULL(slot), (uint32_t)((kt / TC_STAGES) & 1));
        tc_fence_after();
        const uint32_t a_base = smem_base + slot * stage_bytes;
        const uint64_t ad0 = swz ? umma_desc_sw128(a_base) : umma_desc(a_base, a_plane, 128);
        const uint64_t bd0 = swz ? umma_desc_sw128(a_base + a_stage) : umma_desc(a_base + a_stage, b_plane, 128);
        const uint64_t ainc = swz ? 2u : (uint64_t)((2 * a_plane) >> 4);  // per K = 8 step, in 16-byte units
        const uint64_t binc = swz ? 2u : (uint64_t)((2 * b_plane) >> 4);
#pragma unroll
        for (int q = 0; q < TC_BK / 8; ++q)
          umma_tf32(tmem_base, ad0 + q * ainc, bd0 + q * binc, idesc, (kt > 0 || q > 0) ? 1u : 0u);
        umma_commit(EMPTY(slot));  // the stage may be refilled once these MMAs have read it
      }
      if (KT > 0) umma_commit(DONE);
      // no commit may still be in flight towards this CTA's barriers when the CTA retires
      for (int kt = max(0, KT - TC_STAGES); kt < KT; ++kt) mbar_wait(EMPTY(kt % TC_STAGES), (uint32_t)((kt / TC_STAGES) & 1));
    }
    __syncwarp();
    if (cred) {  // every thread of the cluster takes part in both barriers of the split-K reduction
      cluster_arrive_wait();  // (1) every member has finished its main loop
      cluster_arrive_wait();  // (2) every partial tile has been delivered
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    return;
  }
  // ==================================================================================== loaders (warps 0-3)
  // Each thread keeps TC_INFLIGHT groups of cp.async in flight; a stage is signalled full by all 128 threads once
  // their own copies for it have landed (wait_group) and been fenced towards the async proxy.
  if (TMA) {
    if (tid == 0) {  // one thread keeps the whole ring in flight: loads are fire-and-forget
      const int ow = m0 % p.g.OW, t = m0 / p.g.OW;
      const int cw = ow * p.g.stride - p.g.pad, ch = (t % p.g.OH) * p.g.stride - p.g.pad, cn = t / p.g.OH;
      int tap = kt0 / p.cblocks, cb = kt0 - tap * p.cblocks;
      for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % TC_STAGES;
        if (kt >= TC_STAGES) mbar_wait(EMPTY(slot), (uint32_t)((kt / TC_STAGES - 1) & 1));
        const uint32_t a_dst = smem_base + slot * stage_bytes;
        mbar_arrive_expect_tx(FULL(slot), stage_bytes);
        const int r = tap / p.g.KW, sx = tap - r * p.g.KW;
        tma_load_im2col_4d(a_dst, tmA, cb * TC_BK, cw, ch, cn, (unsigned short)sx, (unsigned short)r, FULL(slot));
        tma_load_2d(a_dst + a_stage, tmB, tap * p.g.C + cb * TC_BK, n0, FULL(slot));
        if (++cb == p.cblocks) { cb = 0; ++tap; }
      }
    }
    __syncwarp();
  } else {
  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % TC_STAGES;
    if (kt >= TC_STAGES) mbar_wait(EMPTY(slot), (uint32_t)((kt / TC_STAGES - 1) & 1));
    load_tile(kt0 + kt, slot);
    cp_async_commit();
    if (kt >= TC_INFLIGHT) {
      cp_async_wait<TC_INFLIGHT>();
      fence_proxy_async();
      mbar_arrive(FULL((kt - TC_INFLIGHT) % TC_STAGES));
    }
  }
  cp_async_wait<0>();
  fence_proxy_async();
  for (int kt = max(0, KT - TC_INFLIGHT); kt < KT; ++kt) mbar_arrive(FULL(kt % TC_STAGES));
  }
  if (KT > 0) mbar_wait(DONE, 0);
  tc_fence_after();

  // ---- epilogue: thread = one output row (TMEM lane), 16 columns at a time
  const int row = warp * 32 + lane;
  const int m = m0 + row;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  if (cred) {
    // ---- split-K inside a cluster: rank r of the (1, 1, S) cluster reduced k-slice r of this output tile.  Output
    // rows are dealt out to the S members (128 / S rows each); every member sends each owner the rows it owns
    // (st.shared::cluster into slot [sender][row][col] of the owner's buffer), one cluster barrier, then each owner
    // adds the S_nz slots in slice order and applies the epilogue: fixed summation order, no atomics, no workspace.
    const int S = p.splits;
    const int rows_per = TC_BM / S;
    const uint32_t ldred = (uint32_t)bn + 4;  // floats; +4: the 16-byte stores of 32 lanes (32 rows) spread over the banks
    const uint32_t rank = cluster_ctarank();
    // The reduction buffer ALIASES the operand ring (one more CTA fits an SM): a peer may only write into it once this
    // CTA's last MMA has read its operands — every thread arrives after the DONE wait above (MMA warp: after its own
    // EMPTY waits), so once the barrier completes no ring of the cluster is in use any more.  The barrier also
    // guarantees that every peer has started executing, which distributed shared memory accesses require.
    cluster_arrive_wait();
    if (KT > 0) {
      const uint32_t owner = (uint32_t)(row / rows_per);
      const uint32_t dst = smem_base + p.red_off + ((rank * rows_per + (uint32_t)(row % rows_per)) * ldred) * 4u;
      for (int c0 = 0; c0 < bn; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4) st_cluster_v4(dst + (uint32_t)(c0 + j) * 4u, owner, v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    cluster_arrive_wait();
    const float* red = reinterpret_cast<const float*>(smem + p.red_off);
    const int nz = p.splits_nz;
    const int quads = bn >> 2;
    const bool vec = p.vec_store != 0;
    for (int idx = tid; idx < rows_per * quads; idx += TC_LOAD_THREADS) {
      const int rl = idx / quads, q = idx - rl * quads;
      const int mm = m0 + (int)rank * rows_per + rl;
      const int n = n0 + 4 * q;
      if (mm >= M || n >= p.N) continue;
      float4 x = *reinterpret_cast<const float4*>(red + (size_t)rl * ldred + 4 * q);
      for (int r = 1; r < nz; ++r) {
        const float4 y = *reinterpret_cast<const float4*>(red + ((size_t)r * rows_per + rl) * ldred + 4 * q);
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
      }
      float xv[4] = {x.x, x.y, x.z, x.w};
      float* crow = p.C + (long long)mm * p.ldc + n;
      const float* rrow = p.residual ? p.residual + (long long)mm * p.ldr + n : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < p.N) {
          if (p.scale) xv[j] *= __ldg(p.scale + n + j);
          if (p.bias) xv[j] += __ldg(p.bias + n + j);
          if (rrow) xv[j] += rrow[j];
          if (p.relu) xv[j] = fmaxf(xv[j], 0.f);
        }
      }
      if (vec && n + 4 <= p.N) {
        *reinterpret_cast<float4*>(crow) = make_float4(xv[0], xv[1], xv[2], xv[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) crow[j] = xv[j];
      }
    }
    tc_fence_before();
    __syncthreads();
    return;
  }
  for (int c0 = 0; c0 < bn; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + c0, v);
    if (m < M) {
      float* crow = p.C + (long long)m * p.ldc + n0 + c0;
      const float* rrow = p.residual ? p.residual + (long long)m * p.ldr + n0 + c0 : nullptr;
      if (p.vec_store && p.splits == 1 && n0 + c0 + 16 <= p.N) {
        // 16-byte stores: a 4-byte store per lane rewrites every 32-byte sector 8 times on its way to L2
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
          int n = n0 + c0 + j;
          if (n < p.N) {
            float x = __uint_as_float(v[j]);
            if (p.splits > 1) {
              atomicAdd(crow + j, x);
              continue;
            }
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
  __syncthreads();  // warp 4 deallocates TMEM after every epilogue warp has drained it
}

template <bool CONV, bool CA = false>
__global__ void __launch_bounds__(TC_THREADS) tc_gemm_kernel(TcArgs p) {
  tc_gemm_body<CONV, CA, false>(p, nullptr, nullptr);
}
// the same kernel fed by TMA (im2col-mode map for the activations, tiled map for the packed weights)
__global__ void __launch_bounds__(TC_THREADS) tc_conv_tma_splitk_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                        const __grid_constant__ CUtensorMap tmB, TcArgs p) {
  tc_gemm_body<true, false, true>(p, &tmA, &tmB);
}

static int g_tc_ca = 1;
static int g_tc_splitk = 1;
static int g_tc_splitk_cluster = 1;
static int g_tc_splitk_fill = 40;  // split-K aims at this many CTAs per 100 SMs (see avl_set_tc_splitk_fill)
static int g_tc_cluster16 = -1;  // -1: not probed yet; 0: clusters of 16 CTAs unavailable; 1: available
static int g_tc_swz = 1;
static int g_tc_stages = 0;  // 0: automatic; 3 / 4: forced ring depth (diagnostic)

__global__ void tc_zero_cols_kernel(float* y, long long ldy, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    y[(i / cols) * ldy + (i % cols)] = 0.f;
}
__global__ void tc_epilogue_cols_kernel(float* y, long long ldy, int rows, int cols, const float* __restrict__ scale,
                                        const float* __restrict__ bias, const float* __restrict__ residual,
                                        long long ldr, int relu) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const long long r = i / cols;
    float v = y[r * ldy + c];
    if (scale) v *= scale[c];
    if (bias) v += bias[c];
    if (residual) v += residual[r * ldr + c];
    if (relu) v = fmaxf(v, 0.f);
    y[r * ldy + c] = v;
  }
}

static int pick_bn(int N) {
  int n16 = (N + 15) / 16 * 16;
  if (n16 <= 256) return n16;
  for (int bn = 256; bn >= 64; bn -= 16)
    if (n16 % bn == 0) return bn;
  return 256;
}

extern "C" bool avl_conv_tma_maps(CUtensorMap* ta, CUtensorMap* tb, const float* x, int N, int H, int W, int C,
                                  const float* w_packed, int Cout, int KH, int KW, int stride, int pad, int bn);  // gemm_tma.cu
extern "C" void avl_tc_conv_tma_count_add();

// want_tma: a convolution whose operands may come by TMA (im2col-mode map); decided here once the N tile is known
static int tc_launch(bool conv, TcArgs& p, cudaStream_t s, bool want_tma = false) {
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AVL_CUDA_CHECK((cudaFuncSetAttribute(tc_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_tma_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaFuncSetAttribute(tc_conv_tma_splitk_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    // clusters of 16 CTAs for the longest reductions: opt-in per kernel, then ask the driver whether one fits
    bool ok16 = cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                cudaFuncSetAttribute(tc_gemm_kernel<true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (ok16) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1, 1, 16);
      cfg.blockDim = dim3(TC_THREADS);
      cfg.dynamicSmemBytes = 112 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 1;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 16;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      ok16 = cudaOccupancyMaxActiveClusters(&nclusters, tc_gemm_kernel<true, true>, &cfg) == cudaSuccess && nclusters >= 4;
    }
    cudaGetLastError();  // a failed probe must not leave an error behind
    g_tc_cluster16 = ok16 ? 1 : 0;
    attr_set = true;
  }
  p.bn = pick_bn(p.N);
  const int sms = avl_num_sms();
  const int mtiles = avl_div_up(p.M, TC_BM);
  want_tma = want_tma && conv && g_tc_swz && p.g.C >= 16 && p.g.KH == p.g.KW;
  p.cblocks = avl_div_up(p.g.C, TC_BK);
  p.kt_total = p.g.KH * p.g.KW * p.cblocks;
  const int KT = want_tma ? p.kt_total : avl_div_up(p.K, TC_BK);
  p.splits = 1;
  p.splits_nz = 1;
  p.cluster_red = 0;
  p.red_off = 0;
  p.kt_per_split = KT;
  p.swz = g_tc_swz;
  p.vec_store = ((p.ldc & 3) == 0 && ((uintptr_t)p.C & 15) == 0 && (!p.bias || ((uintptr_t)p.bias & 15) == 0) &&
                 (!p.scale || ((uintptr_t)p.scale & 15) == 0) &&
                 (!p.residual || ((p.ldr & 3) == 0 && ((uintptr_t)p.residual & 15) == 0))) ? 1 : 0;
  p.st256 = p.vec_store && avl_rows_32b(p.C, p.ldc, 4);
  if (g_tc_splitk && !p.m_dev && mtiles * avl_div_up(p.N, p.bn) * 2 <= sms && KT >= 8) {
    // few output tiles, long reduction (rollout-batch convolutions on small maps, belief-predictor layers): a
    // handful of CTAs would each stream the whole K extent through one SM's cp.async path.  Narrow the N tile and
    // split K over blockIdx.z so that ~2 CTAs per SM share the operand traffic; partial sums meet in C by atomics.
    if (p.bn > 64) p.bn = 64;
    const int ctas = mtiles * avl_div_up(p.N, p.bn);
    int splits = (g_tc_splitk_fill * sms / 100 + ctas - 1) / ctas;
    if (splits > KT / 4) splits = KT / 4;
    if (splits > 1 && g_tc_splitk_cluster && p.swz) {
      // the slices of one tile as one cluster (portable size <= 8, a divisor of the 128 tile rows)
      int S = splits >= 8 ? 8 : (splits >= 4 ? 4 : 2);
      if (splits >= 16 && g_tc_cluster16 > 0) S = 16;  // non-portable cluster size (opt-in attribute, probed once)
      p.kt_per_split = avl_div_up(KT, S);
      p.splits = S;
      p.splits_nz = avl_div_up(KT, p.kt_per_split);
      p.cluster_red = 1;
    } else if (splits > 1) {
      p.kt_per_split = avl_div_up(KT, splits);
      p.splits = avl_div_up(KT, p.kt_per_split);
      p.splits_nz = p.splits;
    }
  }
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage = p.swz ? (TC_BM + (size_t)p.bn) * 128 : (size_t)TC_CHUNKS * ((TC_BM * 16 + 16) + ((size_t)p.bn * 16 + 16));
  const size_t red_bytes = p.cluster_red ? (size_t)TC_BM * (p.bn + 4) * sizeof(float) : 0;
  // Ring depth: the loaders keep TC_INFLIGHT = 2 tiles in flight, so 3 slots suffice; a 4th slot never paid for itself
  // (tools/tc_conv_bench.py at update batch, TC_STAGES=0/3/4: 3 slots equal or faster on every layer, e.g. the
  // stride-2 entry convolutions 1.37 -> 1.06 ms, 0.59 -> 0.47 ms, 0.52 -> 0.34 ms) because the smaller footprint
  // admits one more resident CTA per SM — which hides the per-CTA fixed cost at update batch and lets the kernels of
  // the other streams co-reside at rollout batch.  Exception (measured on the rollout step, 40.4k vs 38.7k env-steps/s):
  // a single wave of CTAs with a long reduction is latency-bound and prefers the deeper ring.
  const long long n_ctas = (long long)mtiles * avl_div_up(p.N, p.bn) * p.splits;
  p.stages = 3;
  if (!p.cluster_red && KT > 6 && n_ctas < 2LL * sms) {
    p.stages = (int)((100 * 1024) / stage);
    if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
  }
  if (g_tc_stages >= 3 && g_tc_stages <= TC_MAX_STAGES && (size_t)g_tc_stages * stage <= 200 * 1024) p.stages = g_tc_stages;
  if (p.stages < 3) p.stages = 3;  // the loaders keep TC_INFLIGHT = 2 tiles in flight
  size_t smem = (size_t)p.stages * stage;
  p.red_off = 0;  // the reduction buffer aliases the operand ring (see the kernel)
  if (smem < red_bytes) smem = red_bytes;

  dim3 grid(mtiles, avl_div_up(p.N, p.bn), p.splits);
  const float *scale = p.scale, *bias = p.bias, *residual = p.residual;
  const int relu = p.relu;
  CUtensorMap ta, tb;
  // (TMA mode serves the single-launch forms: no split, or the in-cluster reduction; the atomic split-K fallback keeps
  // its helper kernels and the cp.async loaders)
  const bool tma = want_tma && (p.splits == 1 || p.cluster_red) &&
                   avl_conv_tma_maps(&ta, &tb, p.A, p.g.N, p.g.H, p.g.W, p.g.C, p.B, p.N, p.g.KH, p.g.KW, p.g.stride, p.g.pad, p.bn);
  if (want_tma && !tma && KT != avl_div_up(p.K, TC_BK)) {
    // the split was planned on per-tap k-tiles (channels not a multiple of 32): re-plan for the cp.async k-tiles
    return tc_launch(conv, p, s, false);
  }
  if (tma) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = (unsigned)(p.cluster_red ? p.splits : 1);
    unsigned nat = 1;
    avl_pdl_attr(at, &nat);
    cfg.attrs = at;
    cfg.numAttrs = nat;
    AVL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_conv_tma_splitk_kernel, ta, tb, p));
    AVL_LAUNCH_CHECK();
    avl_tc_conv_tma_count_add();
    return AVL_OK;
  }
  if (p.cluster_red) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = (unsigned)p.splits;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    auto kern = conv ? (g_tc_ca ? tc_gemm_kernel<true, true> : tc_gemm_kernel<true>) : tc_gemm_kernel<false>;
    AVL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
    AVL_LAUNCH_CHECK();
    return AVL_OK;
  }
  if (p.splits > 1) {
    long long tot = (long long)p.M * p.N;
    int g = (int)((tot + 255) / 256 > (long long)sms * 16 ? (long long)sms * 16 : (tot + 255) / 256);
    tc_zero_cols_kernel<<<g, 256, 0, s>>>(p.C, p.ldc, p.M, p.N);
    AVL_LAUNCH_CHECK();
    p.scale = p.bias = p.residual = nullptr;
    p.relu = 0;
  }
  if (conv && g_tc_ca) tc_gemm_kernel<true, true><<<grid, TC_THREADS, smem, s>>>(p);
  else if (conv) tc_gemm_kernel<true><<<grid, TC_THREADS, smem, s>>>(p);
  else tc_gemm_kernel<false><<<grid, TC_THREADS, smem, s>>>(p);
  AVL_LAUNCH_CHECK();
  if (p.splits > 1 && (scale || bias || residual || relu)) {
    long long tot = (long long)p.M * p.N;
    int g = (int)((tot + 255) / 256 > (long long)sms * 16 ? (long long)sms * 16 : (tot + 255) / 256);
    tc_epilogue_cols_kernel<<<g, 256, 0, s>>>(p.C, p.ldc, p.M, p.N, scale, bias, residual, p.ldr, relu);
    AVL_LAUNCH_CHECK();
  }
  return AVL_OK;
}

}  // namespace

// Shared-memory operand layout of the generic kernel: 1 (default) SWIZZLE_128B K-major tiles, 0 SWIZZLE_NONE planes.
AVL_API int avl_set_tc_swizzle(int on) {
  avl_bump_config_epoch();
  int old = g_tc_swz;
  g_tc_swz = on ? 1 : 0;
  return old;
}

// 1 (default): small-M / long-K problems are split over K (atomic partial sums); 0: never.  Returns the old value.
AVL_API int avl_set_tc_splitk(int on) {
  avl_bump_config_epoch();
  int old = g_tc_splitk;
  g_tc_splitk = on ? 1 : 0;
  return old;
}

// Ring depth of the generic kernel: 0 (default) automatic, 3 or 4 forced (diagnostic).  Returns the old value.
AVL_API int avl_set_tc_stages(int stages) {
  avl_bump_config_epoch();
  int old = g_tc_stages;
  g_tc_stages = stages;
  return old;
}

// Split-K aims at `percent` CTAs per 100 SMs.  Default 40: a rollout step runs four to eight encoder chains next to each
// other and is bound by SM time (shared-memory slots), not by one kernel's latency — measured on the whole step
// (bench.py, 64 envs): 200 (two CTAs per SM, what a kernel running alone prefers) 53.5k env-steps/s, 100 56.3k, 60 56.9k,
// 40 57.9k, 25 57.7k (final build: 50 62.6k, 40 63.7k, 30 63.7k); the isolated launch latencies stay within 10 %
// (tools/small_batch_conv_probe.py).  Returns old.
AVL_API int avl_set_tc_splitk_fill(int percent) {
  avl_bump_config_epoch();
  int old = g_tc_splitk_fill;
  if (percent >= 25 && percent <= 400) g_tc_splitk_fill = percent;
  return old;
}

// 1 (default): the k-slices of a split-K problem form a thread-block cluster and are reduced through distributed shared
// memory inside the kernel (deterministic, no helper launches); 0: atomic partial sums into a zeroed output + separate
// epilogue kernel.  Returns the old value.
AVL_API int avl_set_tc_splitk_cluster(int on) {
  avl_bump_config_epoch();
  int old = g_tc_splitk_cluster;
  g_tc_splitk_cluster = on ? 1 : 0;
  return old;
}

// 1 (default): im2col gathers go through L1 (cp.async.ca); 0: L2 only (cp.async.cg).  Returns the old value.
AVL_API int avl_set_tc_conv_l1(int on) {
  avl_bump_config_epoch();
  int old = g_tc_ca;
  g_tc_ca = on ? 1 : 0;
  return old;
}

int avl_tc_gemm_tma_try(const float* A, long long lda, const float* B, float* C, long long ldc, int M, int N, int K,
                        const float* bias, const float* residual, long long ldr, int relu, const int* m_dev,
                        cudaStream_t stream);  // gemm_tma.cu

// Dense: C[M,N] = act(scale * A[M,K] B[N,K]^T + bias + residual).  A rows / B rows must be 16-byte aligned
// (lda % 4 == 0, K % 4 == 0).  m_dev: optional device-side row count (packed SMT rows).
AVL_API int avl_tc_gemm(const float* A, long long lda, const float* B, float* C, long long ldc, int M, int N, int K,
                        const float* scale, const float* bias, const float* residual, long long ldr, int relu,
                        const int* m_dev, void* stream) {
  if (M < 0 || N < 1 || K < 1) return AVL_ERR_ARG;
  if (M == 0) return AVL_OK;
  if (!A || !B || !C) return AVL_ERR_ARG;
  if ((K & 3) || (lda & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return AVL_ERR_UNSUPPORTED;
  if (!scale && (m_dev || (long long)avl_div_up(M, TC_BM) * avl_div_up(N, 64) * 2 > avl_num_sms() ||
                 avl_div_up(K, TC_BK) < 8)) {
    // enough output tiles to fill the GPU without split-K (or split-K not applicable): the TMA-fed kernel
    int rc = avl_tc_gemm_tma_try(A, lda, B, C, ldc, M, N, K, bias, residual, ldr, relu, m_dev, (cudaStream_t)stream);
    if (rc != AVL_ERR_UNSUPPORTED) return rc;
  }
  TcArgs p = {};
  p.A = A; p.lda = lda; p.B = B; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.scale = scale; p.residual = residual; p.ldr = ldr; p.relu = relu; p.m_dev = m_dev;
  return tc_launch(false, p, (cudaStream_t)stream);
}

int avl_tc_conv_halo_try(const float* x, int N, int H, int W, int C, const float* w_packed, int Cout, int KH, int KW,
                         int stride, int pad, const float* scale, const float* bias, const float* residual,
                         long long ldr, int relu, float* y, long long ldy, cudaStream_t stream);  // conv_halo_tc.cu

int avl_tc_conv_tma_try(const float* x, int N, int H, int W, int C, const float* w_packed, int Cout, int KH, int KW,
                        int stride, int pad, const float* scale, const float* bias, const float* residual, long long ldr,
                        int relu, float* y, long long ldy, cudaStream_t stream);  // gemm_tma.cu

// NHWC convolution on the tensor cores.  w_packed: (Cout, KH, KW, C) (k = (r*KW + s)*C + ci), C % 4 == 0.
AVL_API int avl_tc_conv2d_fwd(const float* x, int N, int H, int W, int C, const float* w_packed, int Cout, int KH,
                              int KW, int stride, int pad, const float* scale, const float* bias,
                              const float* residual, long long ldr, int relu, float* y, long long ldy, void* stream) {
  if (N < 0 || H < 1 || W < 1 || C < 1 || Cout < 1 || KH < 1 || KW < 1 || stride < 1 || pad < 0) return AVL_ERR_ARG;
  if (N == 0) return AVL_OK;
  if (!x || !w_packed || !y) return AVL_ERR_ARG;
  if ((C & 3) || ((uintptr_t)x & 15) || ((uintptr_t)w_packed & 15)) return AVL_ERR_UNSUPPORTED;
  if (KH == H && KW == W && pad == 0) {
    // kernel covers the whole map (Linear after a flatten): a dense GEMM over the NHWC-flattened rows, no im2col
    return avl_tc_gemm(x, (long long)H * W * C, w_packed, y, ldy, N, Cout, H * W * C, scale, bias, residual, ldr, relu,
                       nullptr, stream);
  }
  {  // shallow stride-1 layers: halo-strip kernel (no im2col expansion); anything else falls through
    int rc = avl_tc_conv_halo_try(x, N, H, W, C, w_packed, Cout, KH, KW, stride, pad, scale, bias, residual, ldr, relu,
                                  y, ldy, (cudaStream_t)stream);
    if (rc != AVL_ERR_UNSUPPORTED) return rc;
  }
  {  // deep layers with enough output tiles: TMA in im2col mode (gemm_tma.cu)
    int rc = avl_tc_conv_tma_try(x, N, H, W, C, w_packed, Cout, KH, KW, stride, pad, scale, bias, residual, ldr, relu, y, ldy,
                                 (cudaStream_t)stream);
    if (rc != AVL_ERR_UNSUPPORTED) return rc;
  }
  TcArgs p = {};
  p.g.N = N; p.g.H = H; p.g.W = W; p.g.C = C; p.g.KH = KH; p.g.KW = KW; p.g.stride = stride; p.g.pad = pad;
  p.g.OH = (H + 2 * pad - KH) / stride + 1;
  p.g.OW = (W + 2 * pad - KW) / stride + 1;
  if (p.g.OH < 1 || p.g.OW < 1) return AVL_ERR_ARG;
  long long M = (long long)N * p.g.OH * p.g.OW;
  if (M > 2147483647LL) return AVL_ERR_UNSUPPORTED;
  p.A = x; p.B = w_packed; p.C = y; p.ldc = ldy; p.M = (int)M; p.N = Cout; p.K = KH * KW * C;
  p.bias = bias; p.scale = scale; p.residual = residual; p.ldr = ldr; p.relu = relu; p.m_dev = nullptr;
  return tc_launch(true, p, (cudaStream_t)stream, true);
}
#endif  // AVL_HOST_EMUL
