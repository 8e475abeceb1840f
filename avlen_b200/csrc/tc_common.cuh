// tcgen05 / mbarrier / cp.async PTX wrappers shared by the tensor-core kernels (gemm_tc.cu, conv_halo_tc.cu).
#pragma once
#include "nn_kernels.cuh"

#ifndef AVL_HOST_EMUL
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
#ifndef AVL_MBAR_HINT_NS
#define AVL_MBAR_HINT_NS 20000u
#endif
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 24)) __trap();  // a lost arrival must fault, never hang the device
    // with a suspend-time hint the thread sleeps in hardware until the phase completes (or the hint expires) instead of
    // returning at once: the polling of waiting warps was a third of the instructions the halo-strip convolution issued
    // (profiles/r02_halo_f16_l1_ncu_full.txt) and competes with the working warps of the same scheduler
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(AVL_MBAR_HINT_NS)
        : "memory");
  } while (!done);
}
// 32-byte global store / load (sm_100: STG.256 / LDG.256): a thread that owns 32 contiguous bytes writes one full sector
// instead of two half-filled ones.  The address must be 32-byte aligned.
__device__ __forceinline__ void st_global_256(void* dst, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 16 consecutive floats of an output row (64-byte aligned when `wide`): two 32-byte stores, else four 16-byte stores
__device__ __forceinline__ void st_row16(float* dst, const float4 (&x)[4], bool wide) {
  if (wide) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t u[8] = {__float_as_uint(x[2 * h].x),     __float_as_uint(x[2 * h].y),     __float_as_uint(x[2 * h].z),
                             __float_as_uint(x[2 * h].w),     __float_as_uint(x[2 * h + 1].x), __float_as_uint(x[2 * h + 1].y),
                             __float_as_uint(x[2 * h + 1].z), __float_as_uint(x[2 * h + 1].w)};
      st_global_256(dst + 8 * h, u);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = x[q];
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
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
// K-major SWIZZLE_128B tile (rows of exactly 128 bytes, 8-row atoms of 1024 bytes, tile 1024-byte aligned):
// SBO = 1024, LBO unused, layout type 2 (cute::UMMA::LayoutType::SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// ---- TMA (cp.async.bulk.tensor) helpers: the tensor map lives in kernel parameter space (__grid_constant__)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t addr, uint32_t tx_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
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
// kind::f16 (fp16 A / B, format 0; fp32 accumulate), K-major operands: one MMA covers K = 16
__device__ __forceinline__ uint32_t umma_idesc_f16_kmajor(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-uniform variants: the WHOLE warp runs the issue loop (so the compiler keeps descriptors and loop counters in
// uniform registers and emits no per-instruction leader-election loop), and one elected lane — always the same one
// for a full-warp mask — executes the tcgen05 instruction.
__device__ __forceinline__ void umma_tf32_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t mbar_addr) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(mbar_addr)
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

}  // namespace
#endif  // AVL_HOST_EMUL
