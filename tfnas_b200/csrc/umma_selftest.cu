// Self-test of the tcgen05 building blocks (umma.cuh): C[n][p] = sum_k B[n][k] * A[k][p] with
//   A given as [K][M] fp32 (pixel-contiguous rows = MN-major operand, staged by the CTA's threads with
//     the hi/lo tf32 split, exactly like the production prologues),
//   B given as [N][K] fp32 (K-major), pre-split and pre-swizzled by a prep kernel and brought in with
//     one bulk (TMA) copy per K chunk, exactly like the production weight path,
//   three kind::tf32 MMAs per K=8 step (hi*hi + lo*hi + hi*lo), fp32 accumulation in TMEM,
//   epilogue tcgen05.ld 32x32b -> coalesced stores of C[n][p].
// Exported as tfnas_umma_selftest for tests/test_umma_gpu.py; `variant` perturbs descriptor choices
// while bringing the path up.
#include <stdio.h>
#include "kernels.h"
#include "umma.cuh"

using namespace umma;

#define ST_KC 32

// Wp layout: [kchunk][2 (hi, lo)][Npad rows * 128 B], K-major 128B-swizzled rows
__global__ void k_umma_prep_b(int N, int Npad, int K, const float* __restrict__ B, float* __restrict__ Wp) {
  const int kc = blockIdx.x;
  for (int i = threadIdx.x; i < Npad * ST_KC; i += blockDim.x) {
    int r = i / ST_KC, kk = i - r * ST_KC;
    int k = kc * ST_KC + kk;
    float x = (r < N && k < K) ? B[(size_t)r * K + k] : 0.f;
    float hi, lo;
    split_tf32(x, hi, lo);
    char* base = (char*)Wp + (size_t)kc * 2 * Npad * 128;
    *(float*)(base + k_elem_off(r, kk)) = hi;
    *(float*)(base + (size_t)Npad * 128 + k_elem_off(r, kk)) = lo;
  }
}

// one CTA per 128-pixel tile; Npad multiple of 16, <= 256
__global__ void __launch_bounds__(NT) k_umma_selftest(int M, int N, int Npad, int K, const float* __restrict__ A,
                                                       const float* __restrict__ Wp, float* __restrict__ C, int variant) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: A_hi 16 KB | A_lo 16 KB | B_hi Npad*128 | B_lo Npad*128 | barriers
  unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* a_hi = sm;
  unsigned char* a_lo = sm + 16384;
  unsigned char* b_hi = sm + 32768;
  unsigned char* b_lo = b_hi + Npad * 128;
  uint64_t* bar_b = (uint64_t*)(b_lo + Npad * 128);   // bulk copy landed
  uint64_t* bar_mma = bar_b + 1;                      // MMAs of this chunk retired
  uint32_t* tmem_slot = (uint32_t*)(bar_mma + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = blockIdx.x * 128;
  const uint32_t ncols = tmem_cols(Npad);
  if (tid == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = idesc_tf32(128, Npad, 1, 0);
  const uint32_t a_lbo = ST_KC * 128, a_sbo = 512, b_sbo = 1024;
  const int nchunks = (K + ST_KC - 1) / ST_KC;
  uint32_t phase = 0;
  for (int kc = 0; kc < nchunks; ++kc) {
    // weights: one bulk copy (hi and lo blocks are adjacent in Wp and in smem)
    if (tid == 0) {
      mbar_expect_tx(bar_b, 2 * Npad * 128);
      bulk_g2s(b_hi, (const char*)Wp + (size_t)kc * 2 * Npad * 128, 2 * Npad * 128, bar_b);
    }
    // activations: rows kk = warp, warp+8, ...; lane -> pixels 4*lane .. 4*lane+3 (one 16 B chunk)
#pragma unroll
    for (int i = 0; i < ST_KC / 8; ++i) {
      const int kk = warp + i * 8, k = kc * ST_KC + kk;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < K) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int p = p0 + lane * 4 + e;
          v[e] = p < M ? A[(size_t)k * M + p] : 0.f;
        }
      }
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
      const uint32_t off = mn_chunk_off(lane * 4, kk, a_lbo);
      *(float4*)(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
      *(float4*)(a_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_b, phase);
      tc_fence_after();
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
      for (int s = 0; s < ST_KC / 8; ++s) {
        const uint64_t dah = smem_desc(ah + s * 1024, a_lbo, a_sbo, SWIZZLE_128B_BASE32B);
        const uint64_t dal = smem_desc(al + s * 1024, a_lbo, a_sbo, SWIZZLE_128B_BASE32B);
        const uint64_t dbh = smem_desc(bh + s * 32, 0, b_sbo, SWIZZLE_128B);
        const uint64_t dbl = smem_desc(bl + s * 32, 0, b_sbo, SWIZZLE_128B);
        mma_tf32(tmem, dah, dbh, idesc, (kc | s) ? 1u : 0u);
        if (variant == 0) {
          mma_tf32(tmem, dal, dbh, idesc, 1u);
          mma_tf32(tmem, dah, dbl, idesc, 1u);
        }
      }
      mma_commit(bar_mma);
    }
    // everyone waits until the tensor core has consumed this chunk's smem before restaging
    mbar_wait(bar_mma, phase);
    tc_fence_after();
    phase ^= 1;
  }
  // epilogue: warp w reads lane quarter (w & 3), column half (w >> 2)
  const int q = warp & 3, half = warp >> 2;
  if (variant == 2) {   // bring-up: bypass the MMA result, write a known pattern with tcgen05.st
    if (half == 0) {
      for (int c = 0; c < Npad; ++c) {
        uint32_t val = __float_as_uint((float)(q * 32 + lane) + 1000.f * c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(q * 32) << 16) + c), "r"(val) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const int p = p0 + q * 32 + lane;
  const int cols_per_half = (Npad / 2 + 15) / 16 * 16;
  for (int c0 = half * cols_per_half; c0 < min(Npad, (half + 1) * cols_per_half); c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N && p < M) C[(size_t)(c0 + j) * M + p] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

extern "C" int tfnas_umma_selftest(int M, int N, int K, const float* A, const float* B, float* C, float* wp_scratch,
                                   size_t wp_bytes, int variant, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int Npad = (N + 15) / 16 * 16;
  if (Npad > 256 || N < 1 || M < 1 || K < 1) return TFNAS_E_INVALID;
  const int nchunks = (K + ST_KC - 1) / ST_KC;
  if (wp_bytes < (size_t)nchunks * 2 * Npad * 128) return TFNAS_E_WORKSPACE;
  k_umma_prep_b<<<nchunks, 256, 0, st>>>(N, Npad, K, B, wp_scratch);
  size_t smem = 1024 + 32768 + 2 * (size_t)Npad * 128 + 64;
  cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_umma_selftest<<<(M + 127) / 128, NT, smem, st>>>(M, N, Npad, K, A, wp_scratch, C, variant);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TFNAS_OK : TFNAS_E_CUDA;
}

// ---- microbenchmark: cycles per round of 12 kind::tf32 MMAs (one K=32 chunk of the hi/lo split) ----------------
//   mode 0: A MN-major (128B swizzle, 32B base) as in the pointwise kernels; mode 1: A K-major (128B swizzle) as in
//   the weight-gradient kernel.  Operand contents are irrelevant (zeros).  out[cta] = cycles per round.
__global__ void __launch_bounds__(128) k_umma_bench(int mode, int Npad, int rounds, long long* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_hi = sm;
  unsigned char* a_lo = sm + 16384;
  unsigned char* b_hi = sm + 32768;
  unsigned char* b_lo = b_hi + Npad * 128;
  uint64_t* bar_mma = (uint64_t*)(b_lo + Npad * 128);
  uint32_t* tmem_slot = (uint32_t*)(bar_mma + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32768 + 2 * Npad * 128) / 16; i += blockDim.x) ((float4*)sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t ncols = tmem_cols(Npad);
  if (tid == 0) { mbar_init(bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, ncols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(128, Npad, mode == 0 ? 1 : 0, 0);
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint64_t dah = mode == 0 ? smem_desc(ah + s * 1024, ST_KC * 128, 512, SWIZZLE_128B_BASE32B)
                                       : smem_desc(ah + s * 32, 16, 1024, SWIZZLE_128B);
        const uint64_t dal = mode == 0 ? smem_desc(al + s * 1024, ST_KC * 128, 512, SWIZZLE_128B_BASE32B)
                                       : smem_desc(al + s * 32, 16, 1024, SWIZZLE_128B);
        const uint64_t dbh = smem_desc(bh + s * 32, 16, 1024, SWIZZLE_128B);
        const uint64_t dbl = smem_desc(bl + s * 32, 16, 1024, SWIZZLE_128B);
        mma_tf32(tmem, dah, dbh, idesc, (r | s) ? 1u : 0u);
        mma_tf32(tmem, dal, dbh, idesc, 1u);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
      }
      mma_commit(bar_mma);
      mbar_wait(bar_mma, phase);
      phase ^= 1;
    }
    out[blockIdx.x] = (clock64() - t0) / rounds;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

// out: device [ctas] int64.  Returns 0 on launch success.
extern "C" int tfnas_umma_bench(int mode, int N, int rounds, int ctas, long long* out, void* stream) {
  const int Npad = (N + 15) / 16 * 16;
  if (Npad > 256 || N < 1) return TFNAS_E_INVALID;
  size_t smem = 1024 + 32768 + 2 * (size_t)Npad * 128 + 64;
  cudaFuncSetAttribute(k_umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_umma_bench<<<ctas, 128, smem, (cudaStream_t)stream>>>(mode, Npad, rounds, out);
  return cudaGetLastError() == cudaSuccess ? TFNAS_OK : TFNAS_E_CUDA;
}
