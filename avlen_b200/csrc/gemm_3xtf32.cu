// fp32-ACCURATE dense GEMM on tcgen05 (3xTF32):  C[M, N] = act(A[M, K] B[N, K]^T + bias + residual).
//
// The scene-memory transformer's linears must stay fp32-faithful (the reference runs torch.matmul in fp32; stated
// tolerance 1e-3), which kept them on the SIMT GEMM (~18 TFLOP/s) — by now 60 % of a PPO update.
// Here every operand is split into two TF32 numbers, x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi), and
//     A B^T  ~=  A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T          (dropped term: |A_lo B_lo| <= 2^-22 |A B|)
// is accumulated in fp32 in TMEM: three tensor-core MMAs per K = 8 slice, fp32-level accuracy (measured ~1e-6).
//   * B (the weights, <= 768 x 288) is split — and transposed when the backward pass needs W^T — by a tiny kernel
//     into a library-owned scratch; both parts arrive by TMA.
//   * A (activations, up to ~4e5 rows) arrives raw by TMA; four "splitter" warps rewrite the landed tile in shared
//     memory as hi (in place) and lo (second buffer) — elementwise, so the SWIZZLE_128B placement is untouched —
//     fence it towards the async proxy and hand it to the MMA warp.
// Warps 0-3: splitters; warps 4-7: epilogue; warp 8: MMA issue (elect.sync); warp 9: TMA producer (persistent CTAs).
#include "tc_common.cuh"

#ifndef AVL_HOST_EMUL
#include <cuda.h>

namespace {

constexpr int X3_BM = 128, X3_BK = 32, X3_MAX_STAGES = 3, X3_SPLITTERS = 128, X3_EPI = 128;
constexpr int X3_THREADS = X3_SPLITTERS + X3_EPI + 64;  // warps 0-3 split, 4-7 epilogue, 8 MMA, 9 TMA

struct X3Args {
  float* C;
  long long ldc;
  int M, N, K, bn, n_tiles, tmem_cols;
  int stages;  // ring depth: 2 (column tiles of 256: 96 KB per slot) or 3 (narrower tiles)
  const float* bias;
  const float* residual;
  long long ldr;
  int relu;
  int vec_store;
  int st256;  // rows of C are 32-byte aligned: 32-byte stores
  const int* m_dev;
};

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// PERSISTENT: one CTA per SM walks the output tiles (tile = blockIdx.x, + gridDim.x, ...); the TMA -> split -> MMA
// pipeline runs straight through tile boundaries and the accumulator is double-buffered in TMEM (2 x bn columns), so
// the epilogue of tile i (tcgen05.ld, bias / residual / ReLU, 16-byte stores) overlaps the main loop of tile i + 1.
// The first version (one tile per CTA, profiles/r01_gemm3x_ncu_full.txt) kept the tensor pipe 31 % busy: prologue,
// pipeline fill and the 128 KB epilogue of every tile were serial.
__global__ void __launch_bounds__(X3_THREADS) tc_gemm_3x_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmBhi,
                                                                const __grid_constant__ CUtensorMap tmBlo, X3Args p) {
  AVL_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long bars[3 * X3_MAX_STAGES + 4];  // full[S], split[S], empty[S], acc_full[2], acc_empty[2]
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int X3_STAGES = p.stages;
  const int bn = p.bn;
  const int KT = (p.K + X3_BK - 1) / X3_BK;
  const uint32_t a_tile = X3_BM * 128u, b_tile = (uint32_t)bn * 128u;
  const uint32_t stage_bytes = 2 * a_tile + 2 * b_tile;  // [A hi | A lo | B hi | B lo]
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto SPLIT = [&](int s) { return bar0 + 8u * (X3_STAGES + s); };
  auto EMPTY = [&](int s) { return bar0 + 8u * (2 * X3_STAGES + s); };
  auto ACC_FULL = [&](int a) { return bar0 + 8u * (3 * X3_STAGES + a); };
  auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (3 * X3_STAGES + 2 + a); };
  if (tid == 0) {
    for (int i = 0; i < X3_STAGES; ++i) {
      mbar_init(FULL(i), 1);
      mbar_init(SPLIT(i), X3_SPLITTERS);
      mbar_init(EMPTY(i), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(ACC_FULL(a), 1);
      mbar_init(ACC_EMPTY(a), X3_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBhi);
    tma_prefetch_desc(&tmBlo);
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // barriers, TMEM and descriptor prefetch above overlap the previous kernel under programmatic dependent launch; the
  // device-side row count and every operand are read below this line only
  avl_pdl_wait();
  avl_pdl_trigger();
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int total_tiles = ((M + X3_BM - 1) / X3_BM) * p.n_tiles;

  if (warp == 9) {
    // ================================================================================ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles) * X3_BM, n0 = (tile % p.n_tiles) * bn;
        for (int kt = 0; kt < KT; ++kt, ++it) {
          const int slot = it % X3_STAGES;
          if (it >= X3_STAGES) mbar_wait(EMPTY(slot), (uint32_t)((it / X3_STAGES - 1) & 1));
          const uint32_t base = smem_base + slot * stage_bytes;
          mbar_arrive_expect_tx(FULL(slot), a_tile + 2 * b_tile);
          tma_load_2d(base, &tmA, kt * X3_BK, m0, FULL(slot));
          tma_load_2d(base + 2 * a_tile, &tmBhi, kt * X3_BK, n0, FULL(slot));
          tma_load_2d(base + 2 * a_tile + b_tile, &tmBlo, kt * X3_BK, n0, FULL(slot));
        }
      }
    }
  } else if (warp == 8) {
    // ================================================================================ MMA issuer
    const uint32_t idesc = umma_idesc_tf32(X3_BM, bn);
    int it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti & 1;
      mbar_wait(ACC_EMPTY(acc), (uint32_t)(((ti >> 1) & 1) ^ 1));  // first use of each accumulator passes
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * bn);
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const int slot = it % X3_STAGES;
        mbar_wait(SPLIT(slot), (uint32_t)((it / X3_STAGES) & 1));
        tc_fence_after();
        const uint32_t base = smem_base + slot * stage_bytes;
        const uint64_t ahi = umma_desc_sw128(base), alo = umma_desc_sw128(base + a_tile);
        const uint64_t bhi = umma_desc_sw128(base + 2 * a_tile), blo = umma_desc_sw128(base + 2 * a_tile + b_tile);
#pragma unroll
        for (int q = 0; q < X3_BK / 8; ++q) {
          umma_tf32_elect(d, alo + 2u * q, bhi + 2u * q, idesc, (kt > 0 || q > 0) ? 1u : 0u);  // small terms first
          umma_tf32_elect(d, ahi + 2u * q, blo + 2u * q, idesc, 1u);
          umma_tf32_elect(d, ahi + 2u * q, bhi + 2u * q, idesc, 1u);
        }
        umma_commit_elect(EMPTY(slot));
      }
      umma_commit_elect(ACC_FULL(acc));
    }
    // no commit may still be in flight towards this CTA's barriers when the CTA retires
    for (int j = max(0, it - X3_STAGES); j < it; ++j) mbar_wait(EMPTY(j % X3_STAGES), (uint32_t)((j / X3_STAGES) & 1));
    for (int j = max(0, ti - 2); j < ti; ++j) mbar_wait(ACC_FULL(j & 1), (uint32_t)((j >> 1) & 1));
  } else if (warp < 4) {
    // ================================================================================ splitters (warps 0-3)
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const int slot = it % X3_STAGES;
        mbar_wait(FULL(slot), (uint32_t)((it / X3_STAGES) & 1));
        float4* hi = reinterpret_cast<float4*>(smem + (size_t)slot * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem + (size_t)slot * stage_bytes + a_tile);
#pragma unroll
        for (int i = 0; i < (int)(X3_BM * 128 / 16) / X3_SPLITTERS; ++i) {
          const int idx = tid + i * X3_SPLITTERS;
          const float4 v = hi[idx];
          float4 h, l;
          h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
          l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
          hi[idx] = h;
          lo[idx] = l;
        }
        fence_proxy_async();
        mbar_arrive(SPLIT(slot));
      }
    }
  } else {
    // ================================================================================ epilogue (warps 4-7)
    // thread = one output row (TMEM lane = 32 * (warp % 4) + lane), 16 columns at a time
    const int ew = warp - 4;
    int ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti & 1;
      const int m0 = (tile / p.n_tiles) * X3_BM, n0 = (tile % p.n_tiles) * bn;
      const int m = m0 + ew * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * bn);
      mbar_wait(ACC_FULL(acc), (uint32_t)((ti >> 1) & 1));
      tc_fence_after();
      for (int c0 = 0; c0 < bn; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        if (c0 + 16 >= bn) {  // accumulator fully read: hand it back before the stores
          tc_fence_before();
          mbar_arrive(ACC_EMPTY(acc));
        }
        if (m < M) {
          float* crow = p.C + (long long)m * p.ldc + n0 + c0;
          const float* rrow = p.residual ? p.residual + (long long)m * p.ldr + n0 + c0 : nullptr;
          if (p.vec_store && n0 + c0 + 16 <= p.N) {
            float4 xs[4];
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
              xs[j >> 2] = x;
            }
            st_row16(crow, xs, p.st256 != 0);
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// hi / lo TF32 parts of the (small) B operand; transpose != 0: B is given as [K][N] (ld = row stride) and the parts
// are written as [N][K]
__global__ void split_tf32_kernel(const float* __restrict__ src, long long ld, int N, int K, int transpose, float* hi,
                                  float* lo) {
  avl_pdl_wait();     // (the scratch this kernel overwrites may still be read by the previous GEMM)
  avl_pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i - n * K;
  const float x = transpose ? src[(long long)k * ld + n] : src[(long long)n * ld + k];
  const float h = rna_tf32(x);
  hi[i] = h;
  lo[i] = rna_tf32(x - h);
}


// =====================================================================================================================
// Weight gradient on tensor cores:  dW[N, K] += dY[rows, N]^T X[rows, K]   (reduction over the rows — up to ~7e5)
//
// Both operands have the REDUCTION index as their slow (row) index, i.e. they are "MN-major" for the tensor core.
// For 32-bit MN-major operands the tensor core accepts exactly one swizzled layout, SWIZZLE_128B with 32-byte atoms
// (UMMA layout type 1 = cute's Layout_MN_SW128_32B_Atom, Swizzle<2,5,2>: 128-byte line = 32 consecutive features of
// one row, 32-byte chunk index XOR (row & 3), 4 rows = one 512-byte K atom) — which is what a TMA box of 32 rows x
// 32 features lands with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  SBO = 512 bytes (next 4 rows), LBO = 4096 bytes (the
// next 32 features = the next box), one K = 8 slice = 1024 bytes.  So no transposition is needed anywhere: the
// instruction descriptor just flags A and B as MN-major.  The 3xTF32 split is elementwise and done in place like above; rows at
// or beyond the device-side row count are zeroed by the splitters (their content is stale).  The grid's z dimension
// splits the reduction; each CTA writes its partial tile to a workspace with plain stores and a second kernel sums
// the partials in a fixed order into dW (deterministic, no atomics).
constexpr int W3_BM = 128, W3_BK = 32, W3_STAGES = 2, W3_SPLIT_WARPS = 8, W3_THREADS = (W3_SPLIT_WARPS + 2) * 32;
constexpr uint32_t W3_BOX = 32 * 128;  // bytes of one 32 x 32 fp32 box

struct W3Args {
  float* ws;
  int rows, N, K, bn, tmem_cols;
  const int* rows_dev;
  uint32_t lbo, sbo;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

// k-tile range of split z when the reduction has `rows` rows (shared by the GEMM and the reduction kernel)
__device__ __forceinline__ void w3_range(int rows, int nsplit, int z, int& kt0, int& kt1, int& active) {
  const int kt_total = (rows + W3_BK - 1) / W3_BK;
  const int kps = max(1, (kt_total + nsplit - 1) / nsplit);
  active = (kt_total + kps - 1) / kps;
  kt0 = z * kps;
  kt1 = min(kt_total, kt0 + kps);
}

__global__ void __launch_bounds__(W3_THREADS) tc_wgrad_3x_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, W3Args p) {
  AVL_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long bars[3 * W3_STAGES + 1];  // full[S], split[S], empty[S], done
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int rows = p.rows;
  if (p.rows_dev) rows = min(rows, *p.rows_dev);
  int kt0, kt1, active;
  w3_range(rows, (int)gridDim.z, (int)blockIdx.z, kt0, kt1, active);
  if (kt0 >= kt1) return;  // the reduction kernel skips this split as well
  const int m0 = blockIdx.x * W3_BM, n0 = blockIdx.y * p.bn, bn = p.bn;
  const int boxes_a = min(W3_BM / 32, (p.N - m0 + 31) / 32);
  const int boxes_b_cap = (bn + 31) / 32;
  const int boxes_b = min(boxes_b_cap, (p.K - n0 + 31) / 32);
  const uint32_t a_tile = (W3_BM / 32) * W3_BOX, b_tile = (uint32_t)boxes_b_cap * W3_BOX;
  const uint32_t stage_bytes = 2 * a_tile + 2 * b_tile;  // [A hi | A lo | B hi | B lo]
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto SPLIT = [&](int s) { return bar0 + 8u * (W3_STAGES + s); };
  auto EMPTY = [&](int s) { return bar0 + 8u * (2 * W3_STAGES + s); };
  const uint32_t DONE = bar0 + 8u * (3 * W3_STAGES);
  if (tid == 0) {
    for (int i = 0; i < W3_STAGES; ++i) {
      mbar_init(FULL(i), 1);
      mbar_init(SPLIT(i), W3_SPLIT_WARPS * 32);
      mbar_init(EMPTY(i), 1);
    }
    mbar_init(DONE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  // boxes that lie entirely beyond the feature range are never loaded: they must read as zeros
  for (uint32_t i = tid; i < W3_STAGES * stage_bytes / 16; i += W3_THREADS)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  if (warp == W3_SPLIT_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int KT = kt1 - kt0;

  if (warp == W3_SPLIT_WARPS + 1) {
    // ================================================================================ TMA producer
    if (lane == 0) {
      for (int it = 0; it < KT; ++it) {
        const int slot = it % W3_STAGES;
        if (it >= W3_STAGES) mbar_wait(EMPTY(slot), (uint32_t)((it / W3_STAGES - 1) & 1));
        const uint32_t base = smem_base + slot * stage_bytes;
        const int r0 = (kt0 + it) * W3_BK;
        mbar_arrive_expect_tx(FULL(slot), (uint32_t)(boxes_a + boxes_b) * W3_BOX);
        for (int i = 0; i < boxes_a; ++i) tma_load_2d(base + i * W3_BOX, &tmA, m0 + 32 * i, r0, FULL(slot));
        for (int j = 0; j < boxes_b; ++j) tma_load_2d(base + 2 * a_tile + j * W3_BOX, &tmB, n0 + 32 * j, r0, FULL(slot));
      }
    }
    __syncthreads();  // matches the final barrier of the other roles
    return;
  }
  if (warp == W3_SPLIT_WARPS) {
    // ================================================================================ MMA issuer
    const uint32_t idesc = umma_idesc_tf32(W3_BM, bn) | (1u << 15) | (1u << 16);  // A and B MN-major
    for (int it = 0; it < KT; ++it) {
      const int slot = it % W3_STAGES;
      mbar_wait(SPLIT(slot), (uint32_t)((it / W3_STAGES) & 1));
      tc_fence_after();
      const uint32_t base = smem_base + slot * stage_bytes;
      const uint64_t ahi = umma_desc_mn_sw128_32b(base, p.lbo, p.sbo), alo = umma_desc_mn_sw128_32b(base + a_tile, p.lbo, p.sbo);
      const uint64_t bhi = umma_desc_mn_sw128_32b(base + 2 * a_tile, p.lbo, p.sbo);
      const uint64_t blo = umma_desc_mn_sw128_32b(base + 2 * a_tile + b_tile, p.lbo, p.sbo);
#pragma unroll
      for (int q = 0; q < W3_BK / 8; ++q) {  // one K = 8 slice = 8 rows = one 1024-byte atom per box
        const uint32_t adv = (uint32_t)q * (1024u >> 4);
        umma_tf32_elect(tmem_base, alo + adv, bhi + adv, idesc, (it > 0 || q > 0) ? 1u : 0u);  // small terms first
        umma_tf32_elect(tmem_base, ahi + adv, blo + adv, idesc, 1u);
        umma_tf32_elect(tmem_base, ahi + adv, bhi + adv, idesc, 1u);
      }
      umma_commit_elect(EMPTY(slot));
    }
    umma_commit_elect(DONE);
    for (int it = max(0, KT - W3_STAGES); it < KT; ++it)
      mbar_wait(EMPTY(it % W3_STAGES), (uint32_t)((it / W3_STAGES) & 1));
    tc_fence_before();
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    return;
  }
  // ==================================================================================== splitters (warps 0-7)
  const int srow = tid >> 3;  // row of this thread's 16-byte chunk inside every box (256 chunks per box)
  for (int it = 0; it < KT; ++it) {
    const int slot = it % W3_STAGES;
    mbar_wait(FULL(slot), (uint32_t)((it / W3_STAGES) & 1));
    const bool live = (kt0 + it) * W3_BK + srow < rows;
    unsigned char* st = smem + (size_t)slot * stage_bytes;
    auto split_boxes = [&](unsigned char* hi_base, unsigned char* lo_base, int nbox) {
      for (int i = 0; i < nbox; ++i) {
        float4* hp = reinterpret_cast<float4*>(hi_base + (size_t)i * W3_BOX) + tid;
        float4* lp = reinterpret_cast<float4*>(lo_base + (size_t)i * W3_BOX) + tid;
        float4 v = *hp;
        if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 h, l;
        h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
        l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
        *hp = h;
        *lp = l;
      }
    };
    split_boxes(st, st + a_tile, boxes_a);
    split_boxes(st + 2 * a_tile, st + 2 * a_tile + b_tile, boxes_b);
    fence_proxy_async();
    mbar_arrive(SPLIT(slot));
  }
  if (warp < 4) {
    mbar_wait(DONE, 0);
    tc_fence_after();
    // ---- epilogue: thread = one dW row (TMEM lane); the partial tile goes to ws[z][tile_m][tile_n][128][bn]
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* wrow = p.ws + ((((size_t)blockIdx.z * gridDim.x + blockIdx.x) * gridDim.y + blockIdx.y) * W3_BM + warp * 32 + lane) * bn;
    for (int c0 = 0; c0 < bn; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(wrow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
}

// dW[m][n] += sum over the active splits of the partial tiles, in split order
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* dW, long long lddw, int N, int K, int bn,
                                    int tiles_m, int tiles_n, int nsplit, int rows_cap, const int* rows_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  int rows = rows_cap;
  if (rows_dev) rows = min(rows, *rows_dev);
  int kt0, kt1, active;
  w3_range(rows, nsplit, 0, kt0, kt1, active);
  if (rows <= 0) return;
  const int m = i / K, n = i - m * K;
  const int bx = m / W3_BM, by = n / bn;
  const size_t tile = (size_t)W3_BM * bn;
  const float* src = ws + ((size_t)bx * tiles_n + by) * tile + (size_t)(m - bx * W3_BM) * bn + (n - by * bn);
  const size_t zstride = (size_t)tiles_m * tiles_n * tile;
  float acc = 0.f;
  for (int z = 0; z < active; ++z) acc += src[(size_t)z * zstride];
  dW[(long long)m * lddw + n] += acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled3() {
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
bool make_map3(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows,
               CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = encode_tiled3();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)X3_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr size_t X3_SCRATCH_FLOATS = 1u << 20;  // hi + lo of the B operand: N * K <= 512 Ki elements
// one scratch per caller stream: two scene-memory transformers (pi_q and pi_g of the AVLEN interactive step) may run
// concurrently on two streams, each splitting its own weights
struct StreamScratch {
  cudaStream_t stream;
  float* buf;
};
constexpr int X3_MAX_STREAMS = 8;
StreamScratch g_scratches[X3_MAX_STREAMS];
int g_num_scratches = 0;
int g_x3_on = 1;

int scratch_for(cudaStream_t s, float** out) {
  for (int i = 0; i < g_num_scratches; ++i)
    if (g_scratches[i].stream == s) { *out = g_scratches[i].buf; return AVL_OK; }
  if (g_num_scratches == X3_MAX_STREAMS) { *out = nullptr; return AVL_ERR_UNSUPPORTED; }  // caller falls back (fp32 SIMT)
  float* b = nullptr;
  AVL_CUDA_CHECK(cudaMalloc(&b, X3_SCRATCH_FLOATS * sizeof(float)));
  g_scratches[g_num_scratches].stream = s;
  g_scratches[g_num_scratches].buf = b;
  ++g_num_scratches;
  *out = b;
  return AVL_OK;
}

}  // namespace

AVL_API int avl_set_tc_3xtf32(int on) {
  avl_bump_config_epoch();
  int old = g_x3_on;
  g_x3_on = on ? 1 : 0;
  return old;
}

// b_transposed != 0: B is stored [K][N] with row stride ldb (the weight of a Linear seen from its backward pass).
// Returns AVL_ERR_UNSUPPORTED (nothing launched) when the shape / alignment is outside the kernel or the path is off.
AVL_API int avl_tc_gemm_3x(const float* A, long long lda, const float* B, long long ldb, int b_transposed, float* C,
                           long long ldc, int M, int N, int K, const float* bias, const float* residual, long long ldr,
                           int relu, const int* m_dev, void* stream) {
  if (!g_x3_on) return AVL_ERR_UNSUPPORTED;
  if (M < 1 || N < 1 || K < 1 || !A || !B || !C) return AVL_ERR_ARG;
  if (((uintptr_t)A & 15) || (lda & 3) || (K & 3) || (size_t)N * K * 2 > X3_SCRATCH_FLOATS) return AVL_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  float* scratch = nullptr;
  {
    const int src = scratch_for(s, &scratch);
    if (src) return src;
  }
  float* bhi = scratch;
  float* blo = scratch + (size_t)N * K;
  AVL_LAUNCH_PDL(split_tf32_kernel, avl_div_up((long long)N * K, 256), 256, 0, s, B, ldb, N, K, b_transposed, bhi, blo);
  AVL_LAUNCH_CHECK();
  X3Args p = {};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.bias = bias; p.residual = residual; p.ldr = ldr; p.relu = relu;
  p.m_dev = m_dev;
  const int n16 = (N + 15) / 16 * 16;
  p.bn = n16 < 256 ? n16 : 256;
  if (n16 > 256)
    for (int bn = 256; bn >= 64; bn -= 16)
      if (n16 % bn == 0) { p.bn = bn; break; }
  // few row tiles (rollout batch: <= 64 * 151 packed rows): narrower column tiles so that the tiles cover the SMs —
  // a 128 x 256 tile per CTA left 3/4 of the GPU idle and made every CTA pull the whole split weight (80 KB per k-tile)
  while (p.bn >= 128 && (p.bn % 32) == 0 && (long long)avl_div_up(M, X3_BM) * avl_div_up(N, p.bn) < avl_num_sms()) p.bn /= 2;
  int cols = 32;
  while (cols < 2 * p.bn) cols <<= 1;  // two accumulators
  p.tmem_cols = cols;
  p.n_tiles = avl_div_up(N, p.bn);
  p.vec_store = ((ldc & 3) == 0 && ((uintptr_t)C & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0) &&
                 (!residual || ((ldr & 3) == 0 && ((uintptr_t)residual & 15) == 0))) ? 1 : 0;
  p.st256 = p.vec_store && avl_rows_32b(C, ldc, 4);
  CUtensorMap ta, tbh, tbl;
  if (!make_map3(&ta, A, M, K, lda, X3_BM) || !make_map3(&tbh, bhi, N, K, K, p.bn) || !make_map3(&tbl, blo, N, K, K, p.bn))
    return AVL_ERR_UNSUPPORTED;
  // 2 slots of 96 KB with 256-column tiles; narrower tiles (rollout batch) afford a third slot, which matters there:
  // a tile is only 8-9 k-steps long, so the pipeline never reaches steady state with 2
  const size_t slot_bytes = (2 * X3_BM + 2 * (size_t)p.bn) * 128;
  p.stages = 3 * slot_bytes <= 200 * 1024 ? 3 : 2;
  const size_t smem = (size_t)p.stages * slot_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  long long tiles = (long long)avl_div_up(M, X3_BM) * p.n_tiles;
  const int grid = (int)(tiles < avl_num_sms() ? tiles : avl_num_sms());  // persistent: one CTA per SM
  AVL_LAUNCH_PDL(tc_gemm_3x_kernel, grid, X3_THREADS, smem, s, ta, tbh, tbl, p);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

namespace {
constexpr size_t W3_WS_FLOATS = 8u << 20;  // 32 MB of partial tiles
float* g_w3_ws = nullptr;
int g_w3_lbo = 4096, g_w3_sbo = 512;
}  // namespace

AVL_API int avl_set_wgrad_desc(int lbo_bytes, int sbo_bytes) {  // diagnostic: descriptor strides of the MN-major tiles
  g_w3_lbo = lbo_bytes;
  g_w3_sbo = sbo_bytes;
  return AVL_OK;
}

// dW[N, K] += dY[rows, N]^T X[rows, K], fp32-accurate (3xTF32) on tcgen05; rows_dev = optional device row count.
// Returns AVL_ERR_UNSUPPORTED (nothing launched) when the shape / alignment is outside the kernel or the path is off.
AVL_API int avl_tc_wgrad_3x(const float* dY, long long ldy, const float* X, long long ldx, float* dW, long long lddw,
                            int rows, int N, int K, const int* rows_dev, void* stream) {
  if (!g_x3_on) return AVL_ERR_UNSUPPORTED;
  if (rows < 1 || N < 1 || K < 1 || !dY || !X || !dW) return AVL_ERR_ARG;
  if (((uintptr_t)dY & 15) || ((uintptr_t)X & 15) || (ldy & 3) || (ldx & 3)) return AVL_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_w3_ws) AVL_CUDA_CHECK(cudaMalloc(&g_w3_ws, W3_WS_FLOATS * sizeof(float)));
  W3Args p = {};
  p.ws = g_w3_ws; p.rows = rows; p.N = N; p.K = K; p.rows_dev = rows_dev;
  p.lbo = (uint32_t)g_w3_lbo; p.sbo = (uint32_t)g_w3_sbo;
  const int n16 = (K + 15) / 16 * 16;
  p.bn = n16 < 256 ? n16 : 256;
  if (n16 > 256)
    for (int bn = 256; bn >= 64; bn -= 16)
      if (n16 % bn == 0) { p.bn = bn; break; }
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  const int tiles_m = avl_div_up(N, W3_BM), tiles_n = avl_div_up(K, p.bn);
  const size_t tile_floats = (size_t)tiles_m * tiles_n * W3_BM * p.bn;
  int splits = avl_num_sms() / (tiles_m * tiles_n);
  const int max_by_rows = avl_div_up(rows, 8 * W3_BK);
  if (splits > max_by_rows) splits = max_by_rows;
  if ((size_t)splits * tile_floats > W3_WS_FLOATS) splits = (int)(W3_WS_FLOATS / tile_floats);
  if (splits < 1) return AVL_ERR_UNSUPPORTED;
  CUtensorMap ta, tb;
  if (!make_map3(&ta, dY, rows, N, ldy, W3_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map3(&tb, X, rows, K, ldx, W3_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return AVL_ERR_UNSUPPORTED;
  const size_t smem = (size_t)W3_STAGES * (2 * (W3_BM / 32) + 2 * (size_t)((p.bn + 31) / 32)) * W3_BOX;
  static bool attr_set = false;
  if (!attr_set) {
    AVL_CUDA_CHECK(cudaFuncSetAttribute(tc_wgrad_3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(tiles_m, tiles_n, splits);
  tc_wgrad_3x_kernel<<<grid, W3_THREADS, smem, s>>>(ta, tb, p);
  AVL_LAUNCH_CHECK();
  wgrad_reduce_kernel<<<avl_div_up((long long)N * K, 256), 256, 0, s>>>(g_w3_ws, dW, lddw, N, K, p.bn, tiles_m, tiles_n,
                                                                       splits, rows, rows_dev);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
