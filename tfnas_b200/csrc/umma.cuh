// tcgen05 (5th-gen tensor core) building blocks for sm_100a: TMEM allocation, shared-memory matrix
// descriptors, kind::tf32 MMA issue, commit -> mbarrier, TMEM -> register loads, bulk (TMA) copies.
//
// Numerics: the 1e-3 parity bar needs fp32-equivalent products, so every fp32 operand x is
// split as x = hi + lo with hi = x & 0xFFFFE000 (exactly the 19 bits kind::tf32 consumes) and
// lo = x - hi (exact in fp32), and a product is issued as three MMAs: hi*hi + lo*hi + hi*lo.
//
// Shared-memory operand layouts (128-byte swizzle, fp32 containers, 32 elements = one 128 B row):
//   K-major  [rows][32 k]   : byte(r, k)  = r*128 + (((k>>2) ^ (r&7))<<4) + (k&3)*4          SBO = 1024 (8 rows)
//   MN-major [mn/32][k][32] : 32-bit MN-major operands must use the "128B swizzle with 32B base" mode
//                             (Swizzle<2,5,2>: the four 32 B units of a row are XOR-ed with k&3):
//                             byte(mn, k) = (mn>>5)*LBO + k*128 + ((((mn&31)>>3) ^ (k&3))<<5) + (mn&7)*4
//                             LBO = rows_k*128 (stride between 32-wide MN atoms), SBO = 512 (4 k rows)
// One kind::tf32 MMA consumes K = 8: K-major advances the start address by 32 B, MN-major by 1024 B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// same, but a thread that finds the phase incomplete is suspended by the hardware (up to `ns` nanoseconds per attempt)
// instead of spinning: waiting warps then stay off the issue / fetch path of the warps that do the work
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity, uint32_t ns) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}

// ---- proxies / fences -------------------------------------------------------------------------
// generic-proxy st.shared -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------
// one full warp; writes the TMEM base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ inline uint32_t tmem_cols(uint32_t n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// 32 lanes x 16 consecutive columns -> 16 registers per thread (thread = lane of its warp's quarter)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors ------------------------------------------------------------------------------
enum { SWIZZLE_NONE = 0, SWIZZLE_128B_BASE32B = 1, SWIZZLE_128B = 2 };
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
// kind::tf32, fp32 accumulate.  a_mn / b_mn: 1 = MN-major operand, 0 = K-major.
__host__ __device__ inline uint32_t idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;            // D format: F32
  d |= 2u << 7;            // A format: TF32
  d |= 2u << 10;           // B format: TF32
  d |= (a_mn & 1u) << 15;
  d |= (b_mn & 1u) << 16;
  d |= ((N >> 3) & 0x3Fu) << 17;
  d |= ((M >> 4) & 0x1Fu) << 24;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bulk copy global -> shared (TMA engine, contiguous bytes) -----------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- tf32 split + swizzled addressing -----------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}
// byte offset of the 16-byte chunk holding elements mn..mn+3 (mn % 4 == 0) of row k in an MN-major tile
__device__ __forceinline__ uint32_t mn_chunk_off(uint32_t mn, uint32_t k, uint32_t lbo) {
  return (mn >> 5) * lbo + k * 128u + ((((mn & 31u) >> 3) ^ (k & 3u)) << 5) + ((mn >> 2) & 1u) * 16u;
}
// byte offset of element (row r, k) in a K-major tile
__host__ __device__ inline uint32_t k_elem_off(uint32_t r, uint32_t k) {
  return r * 128u + ((((k >> 2) ^ (r & 7u))) << 4) + (k & 3u) * 4u;
}

}  // namespace umma
