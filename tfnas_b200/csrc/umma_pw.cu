// tcgen05 versions of the four flattened-pixel 1x1-conv GEMMs (expand, project, dc, dx).
//
//   Out[o][p] = sum_k W[o][k] * In[k][p],   p = flattened (n,h,w) pixel, 128 pixels per CTA (MMA M = 128)
//
// * In rows are produced by each kernel's PROLOGUE (BN / activation / SE gate / BN-backward applied on
//   load), split into tf32 hi/lo and written by the CTA's 256 threads into the MN-major shared-memory
//   operand (128B swizzle, 32B base).  Global loads of chunk c+1 are issued before chunk c is emitted.
// * W is pre-split and pre-swizzled once per call by k_umma_prep_w into K-major 128B-swizzle blocks, and
//   brought in per K chunk with ONE bulk (TMA) copy by one thread.
// * Three kind::tf32 MMAs per K=8 step (hi*hi + lo*hi + hi*lo) accumulate in TMEM (fp32); two smem stages
//   so the tensor core works on chunk c while the threads stage chunk c+1.
// * EPILOGUE: each warp reads its TMEM lane quarter (thread = pixel, 16 channels per tcgen05.ld) and
//   applies the kernel's epilogue (BN statistics, SE partial sums, coalesced per-channel stores).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "kernels.h"
#include "pw.cuh"
#include "umma.cuh"

using namespace umma;

#define UM_KC 32                 // K rows per chunk (one 128 B swizzle row of the weights)
#define UM_NT 512                // threads per CTA of the pointwise tcgen05 kernels (16 warps, 2 K rows per warp per chunk)
#define UM_RW (UM_KC / (UM_NT / 32))   // K rows per warp per chunk
#define UM_STAGES 1              // one smem stage per CTA: 2-3 CTAs/SM overlap each other's load / MMA / epilogue phases

static inline void um_tile(int Nout, int& Nc, int& nN, int cap = 256) {
  nN = cdiv(Nout, cap);
  Nc = cdiv(cdiv(Nout, nN), 16) * 16;
}

// element (r, k) of the logical weight = src[r*ld_r + k*ld_k]; rows >= nrows / k >= K are zero
__global__ void __launch_bounds__(256) k_umma_prep_w(const float* __restrict__ src, int ld_r, int ld_k, int nrows, int K,
                                                      int Nc, int nN, int nK, float* __restrict__ dst) {
  const int kc = blockIdx.x, nc = blockIdx.y;
  char* base = (char*)dst + ((size_t)nc * nK + kc) * 2 * Nc * 128;
  for (int i = threadIdx.x; i < Nc * UM_KC; i += blockDim.x) {
    int r, kk;
    if (ld_k == 1) { r = i / UM_KC; kk = i - r * UM_KC; }      // coalesce along k
    else { kk = i / Nc; r = i - kk * Nc; }                      // coalesce along rows
    const int gr = nc * Nc + r, k = kc * UM_KC + kk;
    float x = (gr < nrows && k < K) ? src[(size_t)gr * ld_r + (size_t)k * ld_k] : 0.f;
    float hi, lo;
    split_tf32(x, hi, lo);
    *(float*)(base + k_elem_off(r, kk)) = hi;
    *(float*)(base + (size_t)Nc * 128 + k_elem_off(r, kk)) = lo;
  }
}

// -------------------------------------------------------------------------------------------------
// shared skeleton
// -------------------------------------------------------------------------------------------------
#define UM_MAXB 8                 // weight (B operand) slots: chunk c lives in slot c % nb, fetched nb-1 chunks ahead
struct UmSmem {
  unsigned char* a_hi[UM_STAGES];
  unsigned char* a_lo[UM_STAGES];
  unsigned char* b[UM_STAGES];      // first weight slot: hi block followed by lo block; slot j at b + j * 2*Nc*128
  uint64_t* bar_b;                  // [UM_MAXB]
  uint64_t* bar_mma;                // [UM_STAGES]
  uint32_t* tmem_slot;
  float2* cf;                       // [256] per-column epilogue coefficients (e.g. BN mean / rstd)
  unsigned char* ring;              // cp.async staging ring of raw input rows (thread-private 16 B slots)
  int nb;
};

// layout: [a_hi 16K | a_lo 16K | nb weight slots | barriers 128 B | cf 2 KB | ring]
__device__ __forceinline__ void um_carve(unsigned char* raw, int Nc, UmSmem& S, int nb = 1) {
  // align by OFFSET so the pointer keeps its shared address space (STS/LDS instead of generic ST/LD)
  unsigned char* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  const size_t wslot = (size_t)2 * Nc * 128;             // multiple of 4096 because Nc % 16 == 0
  S.a_hi[0] = sm;
  S.a_lo[0] = sm + 16384;
  S.b[0] = sm + 32768;
  unsigned char* tail = sm + 32768 + nb * wslot;
  S.bar_b = (uint64_t*)tail;
  S.bar_mma = S.bar_b + UM_MAXB;
  S.tmem_slot = (uint32_t*)(S.bar_mma + UM_STAGES);
  S.cf = (float2*)(tail + 128);
  S.ring = tail + 128 + 256 * sizeof(float2);
  S.nb = nb;
}
// one ring stage holds, for each of `ntens` input tensors, UM_RW rows x 4 pixels (16 B) per thread
__host__ __device__ inline size_t um_ring_stage_bytes(int ntens) { return (size_t)ntens * UM_RW * UM_NT * 16; }
static inline size_t um_smem_bytes(int Nc, size_t ring_bytes = 0, int nb = 1) {
  return 1024 + 32768 + (size_t)nb * 2 * Nc * 128 + 128 + 256 * sizeof(float2) + ring_bytes;
}

// One thread's share of a K chunk of the pixel operand: UM_RW rows x 4 pixels, already split into tf32 hi / lo.
// The prologue math fills it BEFORE the wait for the previous chunk's MMAs (so that math overlaps the tensor core's
// ~850-cycle issue -> commit -> mbarrier round trip); only the stores happen after the wait.
struct UmTile { float4 hi[UM_RW], lo[UM_RW]; };
__device__ __forceinline__ void um_split(UmTile& t, int i, const float (&v)[4]) {
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
  t.hi[i] = make_float4(h[0], h[1], h[2], h[3]);
  t.lo[i] = make_float4(l[0], l[1], l[2], l[3]);
}
// row kk = warp + i * (UM_NT / 32) of the chunk, 16 B chunk of pixels 4*lane .. 4*lane+3
__device__ __forceinline__ void um_store(const UmTile& t, unsigned char* a_hi, unsigned char* a_lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < UM_RW; ++i) {
    const uint32_t off = mn_chunk_off(lane * 4, warp + i * (UM_NT / 32), UM_KC * 128);
    *(float4*)(a_hi + off) = t.hi[i];
    *(float4*)(a_lo + off) = t.lo[i];
  }
}

// issue the 12 MMAs of one K chunk (one thread)
__device__ __forceinline__ void um_issue(const UmSmem& S, int s, int Nc, uint32_t tmem, uint32_t idesc, bool first, int bslot = 0) {
  const uint32_t ah = smem_u32(S.a_hi[s]), al = smem_u32(S.a_lo[s]);
  const uint32_t bh = smem_u32(S.b[0]) + bslot * 2 * Nc * 128, bl = bh + Nc * 128;
#pragma unroll
  for (int q = 0; q < UM_KC / 8; ++q) {
    const uint64_t dah = smem_desc(ah + q * 1024, UM_KC * 128, 512, SWIZZLE_128B_BASE32B);
    const uint64_t dal = smem_desc(al + q * 1024, UM_KC * 128, 512, SWIZZLE_128B_BASE32B);
    const uint64_t dbh = smem_desc(bh + q * 32, 16, 1024, SWIZZLE_128B);
    const uint64_t dbl = smem_desc(bl + q * 32, 16, 1024, SWIZZLE_128B);
    mma_tf32(tmem, dah, dbh, idesc, (first && q == 0) ? 0u : 1u);
    mma_tf32(tmem, dal, dbh, idesc, 1u);
    mma_tf32(tmem, dah, dbl, idesc, 1u);
  }
}

// Generic main loop.  F supplies:
//   int   nchunks()                                   K chunks this CTA walks
//   struct Regs                                             registers carried from load(c) to emit(c)
//   void  load(int c, Regs&)                                global loads of chunk c (rows warp+16i, 4 px / lane)
//   void  emit(int c, const Regs&, a_hi, a_lo)              prologue math + hi/lo split + st.shared
//   const void* wsrc(int c)                                 prepped weight block of chunk c (2*Nc*128 bytes)
// ---- optional in-kernel phase trace (debug; tfnas_debug_um_trace) -------------------------------------
// Thread 0 of each traced CTA accumulates clock64() deltas per phase: [0] total, [1] setup, [2] wait MMA,
// [3] weights TMA issue, [4] cp.async wait, [5] emit, [6] fence + barrier, [7] weights wait + MMA issue,
// [8] epilogue, [9] chunks, [10] kernel id.
#define UM_TRACE_SLOTS 16
__device__ unsigned long long* g_um_trace = nullptr;
__device__ int g_um_trace_n = 0;
#ifdef UM_TRACE
struct UmTrace {
  unsigned long long* out;
  long long t0, tstart;
  long long acc[UM_TRACE_SLOTS];
  __device__ __forceinline__ void begin() {
    out = nullptr;
    if (threadIdx.x == 0 && g_um_trace) {
      const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      if (cta < g_um_trace_n) out = g_um_trace + (size_t)cta * UM_TRACE_SLOTS;
    }
    if (out) {
#pragma unroll
      for (int i = 0; i < UM_TRACE_SLOTS; ++i) acc[i] = 0;
      t0 = tstart = clock64();
    }
  }
  __device__ __forceinline__ void mark(int slot) {
    if (out) { const long long t = clock64(); acc[slot] += t - t0; t0 = t; }
  }
  __device__ __forceinline__ void end(int kernel_id, int chunks) {
    if (out) {
      acc[0] = clock64() - tstart; acc[9] = chunks; acc[10] = kernel_id;
#pragma unroll
      for (int i = 0; i < UM_TRACE_SLOTS; ++i) out[i] = (unsigned long long)acc[i];
    }
  }
};
#else      // production build: the trace compiles to nothing
struct UmTrace {
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void end(int, int) {}
};
#endif

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// rg must already hold chunk 0's loads (f.load(0, rg) issued before um_setup so they fly during the TMEM allocation)
template <class F>
__device__ __forceinline__ void um_mainloop(F& f, const UmSmem& S, int Nc, uint32_t tmem, uint32_t idesc, typename F::Regs& rg) {
  const int warp = threadIdx.x >> 5;
  const int n = f.nchunks();
  uint32_t ph_b = 0, ph_m = 0;
  for (int c = 0; c < n; ++c) {
    UmTile t;
    f.compute(c, rg, t);                       // prologue math of chunk c while chunk c-1 is still on the tensor core
    if (c + 1 < n) f.load(c + 1, rg);          // next chunk's global loads in flight across the wait
    if (c > 0) {                               // the MMAs that read the operand tiles must have retired
      mbar_wait(&S.bar_mma[0], ph_m);
      ph_m ^= 1;
      tc_fence_after();
    }
    if (warp == 0) {
      if (elect_one()) {
        mbar_expect_tx(&S.bar_b[0], 2 * Nc * 128);
        bulk_g2s(S.b[0], f.wsrc(c), 2 * Nc * 128, &S.bar_b[0]);
      }
      __syncwarp();
    }
    um_store(t, S.a_hi[0], S.a_lo[0]);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        mbar_wait(&S.bar_b[0], ph_b);
        tc_fence_after();
        um_issue(S, 0, Nc, tmem, idesc, c == 0);
        mma_commit(&S.bar_mma[0]);
      }
      __syncwarp();
    }
    ph_b ^= 1;
  }
  if (n > 0) mbar_wait(&S.bar_mma[0], ph_m);
  tc_fence_after();
}

// ---- cp.async staging ring -----------------------------------------------------------------------
// The single-stage loop above keeps at most ONE K chunk of global loads in flight per CTA (they live in registers),
// which leaves the kernels latency-bound.  The ring variant parks the raw rows of the next RS-1 chunks in shared
// memory with cp.async (no registers): every thread copies, waits for and reads back only ITS OWN 16-byte slots, so
// no extra barrier is needed.  F supplies:
//   void issue(int c, unsigned char* slot)                  cp.async of chunk c's raw rows into ring stage `slot`
//   struct Regs; void consts(int c, Regs&)                  per-row constants of chunk c (registers, one chunk ahead)
//   void emit(int c, const unsigned char* slot, const Regs&, a_hi, a_lo)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;          // src-size 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool valid) {
  const uint32_t n = valid ? 4u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 4 pixels of row `ch` of an [N][C][HW] tensor -> this thread's 16 B ring slot (zero-filled when !rowok / invalid pixel)
__device__ __forceinline__ void ring_row(unsigned char* dst, const float* __restrict__ T, const float* __restrict__ Tb,
                                         const Px4& px, int C, int ch, int HW, bool rowok) {
  if (px.vec) {
    cp_async16(dst, rowok ? Tb + (size_t)ch * HW : T, rowok);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = rowok && px.v[e];
      cp_async4(dst + 4 * e, ok ? T + ((size_t)px.n[e] * C + ch) * HW + px.hw[e] : T, ok);
    }
  }
}

// first RS-1 chunks' copies and chunk 0's constants: called BEFORE um_setup so that they fly during the TMEM
// allocation / barrier initialisation (short-K CTAs spend a third of their life there otherwise)
template <int RS, class F>
__device__ __forceinline__ void um_ring_prefetch(F& f, const UmSmem& S, size_t stage_bytes, typename F::Regs& rg) {
  const int n = f.nchunks();
#pragma unroll
  for (int c = 0; c < RS - 1; ++c) {
    if (c < n) f.issue(c, S.ring + (size_t)c * stage_bytes);
    cp_async_commit();
  }
  if (n > 0) f.consts(0, rg);
}

template <int RS, class F>
__device__ __forceinline__ void um_mainloop_ring(F& f, const UmSmem& S, size_t stage_bytes, int Nc, uint32_t tmem, uint32_t idesc,
                                                 typename F::Regs& rg, UmTrace& tr) {
  const int warp = threadIdx.x >> 5;
  const int n = f.nchunks();
  const int nb = S.nb;
  const uint32_t wbytes = 2 * Nc * 128;
  if (warp == 0) {                          // weights of the first nb-1 chunks
    if (elect_one()) {
      for (int j = 0; j < nb - 1 && j < n; ++j) {
        mbar_expect_tx(&S.bar_b[j], wbytes);
        bulk_g2s(S.b[0] + (size_t)j * wbytes, f.wsrc(j), wbytes, &S.bar_b[j]);
      }
    }
    __syncwarp();
  }
  uint32_t ph_m = 0;
  int slot = 0, slot_issue = RS - 1;        // ring stage of chunk c / of chunk c + RS - 1
  int bslot = 0, bslot_issue = nb - 1;      // weight slot of chunk c / of chunk c + nb - 1
  uint32_t ph_b = 0;                        // parity of chunk c's weight barrier = (c / nb) & 1
  tr.mark(1);
  for (int c = 0; c < n; ++c) {
    if (c + RS - 1 < n) f.issue(c + RS - 1, S.ring + (size_t)slot_issue * stage_bytes);
    cp_async_commit();                      // one group per iteration (possibly empty) keeps the group count uniform
    cp_async_wait<RS - 1>();                // this thread's copies of chunk c have landed
    tr.mark(4);
    UmTile t;
    f.compute(c, S.ring + (size_t)slot * stage_bytes, rg, t);    // overlaps chunk c-1 on the tensor core
    if (c + 1 < n) f.consts(c + 1, rg);
    tr.mark(5);
    if (c > 0) {                            // the MMAs that read the operand tiles (and weight slot (c-1) % nb) must have retired
      mbar_wait(&S.bar_mma[0], ph_m);
      ph_m ^= 1;
      tc_fence_after();
    }
    tr.mark(2);
    if (warp == 0) {
      if (c + nb - 1 < n && elect_one()) {  // weights of chunk c + nb - 1 into the slot chunk c - 1 just released
        mbar_expect_tx(&S.bar_b[bslot_issue], wbytes);
        bulk_g2s(S.b[0] + (size_t)bslot_issue * wbytes, f.wsrc(c + nb - 1), wbytes, &S.bar_b[bslot_issue]);
      }
      __syncwarp();
    }
    tr.mark(3);
    um_store(t, S.a_hi[0], S.a_lo[0]);
    fence_proxy_async();
    __syncthreads();
    tr.mark(6);
    if (warp == 0) {
      if (elect_one()) {
        mbar_wait(&S.bar_b[bslot], ph_b);
        tc_fence_after();
        um_issue(S, 0, Nc, tmem, idesc, c == 0, bslot);
        mma_commit(&S.bar_mma[0]);
      }
      __syncwarp();
    }
    tr.mark(7);
    slot = slot + 1 == RS ? 0 : slot + 1;
    slot_issue = slot_issue + 1 == RS ? 0 : slot_issue + 1;
    if (++bslot == nb) { bslot = 0; ph_b ^= 1; }
    bslot_issue = bslot_issue + 1 == nb ? 0 : bslot_issue + 1;
  }
  if (n > 0) mbar_wait(&S.bar_mma[0], ph_m);
  tc_fence_after();
  tr.mark(2);
}

__device__ __forceinline__ uint32_t um_setup(const UmSmem& S, int Nc) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < UM_MAXB; ++s) mbar_init(&S.bar_b[s], 1);
    for (int s = 0; s < UM_STAGES; ++s) mbar_init(&S.bar_mma[s], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(S.tmem_slot, tmem_cols(Nc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *S.tmem_slot;
}
__device__ __forceinline__ void um_teardown(uint32_t tmem, int Nc) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(tmem, tmem_cols(Nc));
}

// per-thread pixel of the epilogue: warp w -> lane quarter (w & 3), thread = lane
struct EpiPx { int p, n, hw; bool v; };
__device__ __forceinline__ EpiPx epi_px(int tile0, int total, int HW) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  EpiPx e;
  e.p = tile0 + (warp & 3) * 32 + lane;
  e.v = e.p < total;
  e.n = e.v ? fast_div(e.p, HW, __frcp_rn((float)HW)) : 0;
  e.hw = e.v ? e.p - e.n * HW : 0;
  return e;
}
// column range of this warp: warps w and w+4 split the Nc columns in halves (multiples of 16)
__device__ __forceinline__ void epi_cols(int Nc, int& c_lo, int& c_hi) {
  const int parts = blockDim.x >> 7;                  // warps / 4
  const int part = (threadIdx.x >> 5) >> 2;
  const int h = ((Nc + parts - 1) / parts + 15) / 16 * 16;
  c_lo = min(Nc, part * h);
  c_hi = min(Nc, (part + 1) * h);
}
__device__ __forceinline__ uint32_t epi_taddr(uint32_t tmem, int col) {
  return tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) + (uint32_t)col;
}

// load 4 consecutive pixels of plane `ch` as float4 (vector when aligned, else masked scalars)
__device__ __forceinline__ float4 ld4(const float* __restrict__ T, const Px4& px, int C, int ch, int HW) {
  float d[4];
  load4(d, T, px, C, ch, HW);
  return make_float4(d[0], d[1], d[2], d[3]);
}

// -------------------------------------------------------------------------------------------------
// F1a: expand
// -------------------------------------------------------------------------------------------------
struct ExpandF {
  const Plan& P; const UmW& W; const float* x; const float* xb; Px4 px; int nc; int lane, warp;
  struct Regs { float4 a[UM_RW]; };
  __device__ int nchunks() const { return W.nK; }
  __device__ void load(int c, Regs& g) const {
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const int k = c * UM_KC + warp + i * (UM_NT / 32);
      if (k >= P.ic) g.a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      else if (px.vec) g.a[i] = *(const float4*)(xb + (size_t)k * P.HW);
      else g.a[i] = ld4(x, px, P.ic, k, P.HW);
    }
  }
  __device__ void compute(int c, const Regs& g, UmTile& t) const {
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const float v[4] = {g.a[i].x, g.a[i].y, g.a[i].z, g.a[i].w};
      um_split(t, i, v);
    }
  }
  __device__ const void* wsrc(int c) const { return (const char*)W.wp + ((size_t)nc * W.nK + c) * 2 * W.Nc * 128; }
};
__global__ void __launch_bounds__(UM_NT, 2) k_um_expand(Plan P, UmWAll WA, const float* __restrict__ x,
                                                   const float* __restrict__ bn1, float* __restrict__ UH) {
  extern __shared__ __align__(1024) unsigned char um_raw[];
  const int slot = blockIdx.z, nc = blockIdx.y;
  const UmW& W = WA.s[slot];
  if (nc >= W.nN) return;
  const Cand& cd = P.c[slot];
  UmSmem S;
  um_carve(um_raw, W.Nc, S);
  const int ncol = min(W.Nc, cd.mc - nc * W.Nc);        // valid output columns of this chunk
  const int cst0 = cd.coff + nc * W.Nc;                 // stacked channel of column 0
  for (int i = threadIdx.x; i < ncol; i += UM_NT) S.cf[i] = make_float2(bn1[cst0 + i], bn1[P.MC + cst0 + i]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ExpandF f{P, W, x, x, Px4(), nc, lane, warp};
  px_decomp(f.px, blockIdx.x * 128 + lane * 4, P.P, P.HW);
  f.xb = x + (size_t)f.px.n[0] * P.ic * P.HW + f.px.hw[0];
  ExpandF::Regs rg;
  if (f.nchunks() > 0) f.load(0, rg);
  const uint32_t tmem = um_setup(S, W.Nc);
  um_mainloop(f, S, W.Nc, tmem, idesc_tf32(128, W.Nc, 1, 0), rg);
  const EpiPx e = epi_px(blockIdx.x * 128, P.P, P.HW);
  int c_lo, c_hi;
  epi_cols(W.Nc, c_lo, c_hi);
  c_hi = min(c_hi, ncol);
  const size_t HW = (size_t)P.HW;
  float* ob = UH + ((size_t)e.n * P.MC + cst0) * HW + e.hw;
  for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
    float v[16];
    tmem_ld16(epi_taddr(tmem, c0), v);
    if (!e.v) continue;
    float* q = ob + (size_t)c0 * HW;
    const float2* cf = S.cf + c0;
    if (c0 + 16 <= c_hi) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { const float2 m = cf[j]; *q = (v[j] - m.x) * m.y; q += HW; }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < c_hi) { const float2 m = cf[j]; *q = (v[j] - m.x) * m.y; q += HW; }
    }
  }
  um_teardown(tmem, W.Nc);
}

// -------------------------------------------------------------------------------------------------
// F3: project
// -------------------------------------------------------------------------------------------------
template <int ACT>
struct ProjectF {
  const Plan& P; const UmW& W; const Cand& cd; const float* D; const float* Db; const float* bn2; const float* seg; Px4 px; int nc; int lane, warp;
  __device__ int nchunks() const { return W.nK; }
  // mu / r = BN2 mean / rstd of the row; gt = SE gate of the first pixel's image (1 without SE)
  struct Regs { float mu[UM_RW], r[UM_RW], gt[UM_RW]; };
  __device__ void issue(int c, unsigned char* slot) const {
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const int k = c * UM_KC + warp + i * (UM_NT / 32);
      ring_row(slot + ((size_t)i * UM_NT + threadIdx.x) * 16, D, Db, px, P.MC, cd.coff + k, P.HWo, k < cd.mc);
    }
  }
  __device__ void consts(int c, Regs& g) const {
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const int k = c * UM_KC + warp + i * (UM_NT / 32);
      if (k >= cd.mc) { g.mu[i] = g.r[i] = g.gt[i] = 0.f; continue; }
      const int cst = cd.coff + k;
      g.mu[i] = bn2[cst];
      g.r[i] = bn2[P.MC + cst];
      g.gt[i] = cd.se > 0 ? seg[(size_t)px.n[0] * P.MCse + cd.soff + k] : 1.f;
    }
  }
  __device__ void compute(int c, const unsigned char* slot, const Regs& g, UmTile& t) const {
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const int kk = warp + i * (UM_NT / 32), k = c * UM_KC + kk;
      const float4 a = *(const float4*)(slot + ((size_t)i * UM_NT + threadIdx.x) * 16);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < cd.mc) {
        const float mu = g.mu[i], r = g.r[i];
        const float d[4] = {a.x, a.y, a.z, a.w};
        if (px.vec) {        // 4 valid pixels of one image
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = act_f<ACT>((d[e] - mu) * r) * g.gt[i];
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float b = act_f<ACT>((d[e] - mu) * r);
            if (cd.se > 0) b *= seg[(size_t)px.n[e] * P.MCse + cd.soff + k];
            v[e] = px.v[e] ? b : 0.f;
          }
        }
      }
      um_split(t, i, v);
    }
  }
  __device__ const void* wsrc(int c) const { return (const char*)W.wp + ((size_t)nc * W.nK + c) * 2 * W.Nc * 128; }
};

template <int ACT, int RS>
__global__ void __launch_bounds__(UM_NT, 2) k_um_project(Plan P, UmWAll WA, const float* __restrict__ D,
                                                    const float* __restrict__ bn2, const float* __restrict__ seg,
                                                    float* __restrict__ Zb, double* __restrict__ st3, int nb) {
  extern __shared__ __align__(1024) unsigned char um_raw[];
  const int slot = blockIdx.z, nc = blockIdx.y;
  const UmW& W = WA.s[slot];
  if (nc >= W.nN) return;
  const Cand& cd = P.c[slot];
  UmTrace tr;
  tr.begin();
  UmSmem S;
  um_carve(um_raw, W.Nc, S, nb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ProjectF<ACT> f{P, W, cd, D, D, bn2, seg, Px4(), nc, lane, warp};
  px_decomp(f.px, blockIdx.x * 128 + lane * 4, P.Q, P.HWo);
  f.Db = D + (size_t)f.px.n[0] * P.MC * P.HWo + f.px.hw[0];
  typename ProjectF<ACT>::Regs rg;
  um_ring_prefetch<RS>(f, S, um_ring_stage_bytes(1), rg);
  const uint32_t tmem = um_setup(S, W.Nc);
  um_mainloop_ring<RS>(f, S, um_ring_stage_bytes(1), W.Nc, tmem, idesc_tf32(128, W.Nc, 1, 0), rg, tr);
  const EpiPx e = epi_px(blockIdx.x * 128, P.Q, P.HWo);
  const int oc = P.oc;
  const int ncol = min(W.Nc, oc - nc * W.Nc);
  int c_lo, c_hi;
  epi_cols(W.Nc, c_lo, c_hi);
  c_hi = min(c_hi, ncol);
  const size_t HWo = (size_t)P.HWo;
  const int o0 = slot * oc + nc * W.Nc;                  // Z / BN3 channel of column 0
  float* zb = Zb + ((size_t)e.n * P.na * oc + o0) * HWo + e.hw;
  // Rows of invalid pixels and weight rows past oc are zero in the operands, so their accumulators are exactly 0:
  // the BN3 sums need no masking, only the stores do.
  for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
    float v[16], q[16];
    tmem_ld16(epi_taddr(tmem, c0), v);
    if (e.v) {
      float* zp = zb + (size_t)c0 * HWo;
      if (c0 + 16 <= c_hi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { *zp = v[j]; zp += HWo; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < c_hi) *zp = v[j];
          zp += HWo;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) q[j] = v[j] * v[j];
    const float s1 = warp_sum16(v), s2 = warp_sum16(q);
    const int col = c0 + (lane & 15);
    if (lane < 16 && col < c_hi) {
      atomicAdd(&st3[2 * (o0 + col)], (double)s1);
      atomicAdd(&st3[2 * (o0 + col) + 1], (double)s2);
    }
  }
  tr.mark(8);
  tr.end(1, W.nK);
  um_teardown(tmem, W.Nc);
}

// -------------------------------------------------------------------------------------------------
// B2: dc = W3^T dz
// -------------------------------------------------------------------------------------------------
struct DcF {
  const Plan& P; const UmW& W; const float* G; const float* Zb; const float* Gb; const float* Zbb; const float4* cft; Px4 px; int nc, slot; int lane, warp;
  // a = G row, b = Z row; cft[o] = (A, B, C) in shared memory: dz = A*g + B*z + C  (BN3 backward folded, k_b2prep)
  struct Regs { float4 a[UM_RW], b[UM_RW]; };
  __device__ int nchunks() const { return W.nK; }
  __device__ void load(int c, Regs& g) const {
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const int o = c * UM_KC + warp + i * (UM_NT / 32);
      if (o < P.oc) {
        if (px.vec) {
          g.a[i] = *(const float4*)(Gb + (size_t)o * P.HWo);
          g.b[i] = *(const float4*)(Zbb + (size_t)o * P.HWo);
        } else {
          g.a[i] = ld4(G, px, P.oc, o, P.HWo);
          g.b[i] = ld4(Zb, px, P.na * P.oc, slot * P.oc + o, P.HWo);
        }
      } else {
        g.a[i] = g.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __device__ void compute(int c, const Regs& g, UmTile& t) const {
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const int o = c * UM_KC + warp + i * (UM_NT / 32);
      const float4 cf = o < P.oc ? cft[o] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float gg[4] = {g.a[i].x, g.a[i].y, g.a[i].z, g.a[i].w}, z[4] = {g.b[i].x, g.b[i].y, g.b[i].z, g.b[i].w};
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? fmaf(cf.x, gg[e], fmaf(cf.y, z[e], cf.z)) : 0.f;
      um_split(t, i, v);
    }
  }
  __device__ const void* wsrc(int c) const { return (const char*)W.wp + ((size_t)nc * W.nK + c) * 2 * W.Nc * 128; }
};

template <int ACT>
__global__ void __launch_bounds__(UM_NT, 2) k_um_dc(Plan P, UmWAll WA, const float* __restrict__ G, const float* __restrict__ Zb,
                                               const float4* __restrict__ dzc2,
                                               const float* __restrict__ D, const float* __restrict__ bn2,
                                               float* __restrict__ DC, float* __restrict__ dg, double* __restrict__ sD) {
  extern __shared__ __align__(1024) unsigned char um_raw[];
  const int slot = blockIdx.z, nc = blockIdx.y;
  const UmW& W = WA.s[slot];
  if (nc >= W.nN) return;
  const Cand& cd = P.c[slot];
  UmSmem S;
  um_carve(um_raw, W.Nc, S);
  const int ncol = min(W.Nc, cd.mc - nc * W.Nc);        // valid output columns (mid channels) of this chunk
  const int cst0 = cd.coff + nc * W.Nc;                 // stacked channel of column 0
  for (int i = threadIdx.x; i < ncol; i += UM_NT) S.cf[i] = make_float2(bn2[cst0 + i], bn2[P.MC + cst0 + i]);
  float4* cft = (float4*)S.ring;                        // [oc] BN3-backward coefficients of this slot
  for (int i = threadIdx.x; i < P.oc; i += UM_NT) cft[i] = dzc2[slot * P.oc + i];
  // BN2-backward sums of the CTA's four lane quarters meet in shared memory first: one global fp64 atomic per column
  // and CTA instead of four (same-address atomics serialise at ~9 ns each)
  double* sacc = (double*)(S.ring + (((size_t)P.oc * sizeof(float4) + 15) & ~(size_t)15));
  for (int i = threadIdx.x; i < 2 * W.Nc; i += UM_NT) sacc[i] = 0.0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DcF f{P, W, G, Zb, G, Zb, cft, Px4(), nc, slot, lane, warp};
  px_decomp(f.px, blockIdx.x * 128 + lane * 4, P.Q, P.HWo);
  f.Gb = G + (size_t)f.px.n[0] * P.oc * P.HWo + f.px.hw[0];
  f.Zbb = Zb + ((size_t)f.px.n[0] * P.na + slot) * P.oc * P.HWo + f.px.hw[0];
  DcF::Regs rg;
  if (f.nchunks() > 0) f.load(0, rg);
  const uint32_t tmem = um_setup(S, W.Nc);
  um_mainloop(f, S, W.Nc, tmem, idesc_tf32(128, W.Nc, 1, 0), rg);
  const EpiPx e = epi_px(blockIdx.x * 128, P.Q, P.HWo);
  const bool gated = cd.se > 0;
  // images covered by this warp's 32 consecutive pixels: at most two when HWo >= 32
  const int n_first = __shfl_sync(0xffffffffu, e.n, 0);
  int n_last = e.v ? e.n : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_last = max(n_last, __shfl_xor_sync(0xffffffffu, n_last, o));
  const bool two_img = __all_sync(0xffffffffu, (!e.v) || e.n == n_first || e.n == n_last);
  int c_lo, c_hi;
  epi_cols(W.Nc, c_lo, c_hi);
  c_hi = min(c_hi, ncol);
  const size_t HWo = (size_t)P.HWo;
  const size_t eoff = ((size_t)e.n * P.MC + cst0) * HWo + e.hw;
  const int soff0 = cd.soff + nc * W.Nc;                // SE-gated stacked channel of column 0
  // Rows of invalid pixels and weight rows past mc are zero in the operands, so their accumulators are exactly 0
  // and (with d loaded as 0) contribute 0 to every sum below: only loads and stores are masked.
  for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
    float v[16], d[16];
    const int nv = min(16, c_hi - c0);                  // valid columns of this group (uniform)
    {                                                   // D loads issued before the TMEM read so they overlap
      const float* dp = D + eoff + (size_t)c0 * HWo;
      if (nv == 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { d[j] = e.v ? *dp : 0.f; dp += HWo; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { d[j] = (e.v && j < nv) ? *dp : 0.f; dp += HWo; }
      }
    }
    tmem_ld16(epi_taddr(tmem, c0), v);
    const float2* cf = S.cf + c0;
    const int col = c0 + (lane & 15);
    if (e.v) {
      float* qp = DC + eoff + (size_t)c0 * HWo;
      if (gated) {
        // DC = dc (raw); v <- dc * act(BN2(d)) = this pixel's contribution to dL/dgate
        if (nv == 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 m = cf[j];
            *qp = v[j];
            qp += HWo;
            v[j] *= act_f<ACT>((d[j] - m.x) * m.y);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nv) {
              const float2 m = cf[j];
              *qp = v[j];
              v[j] *= act_f<ACT>((d[j] - m.x) * m.y);
            }
            qp += HWo;
          }
        }
      } else {
        // DC = dd-hat = dc * act'(d-hat); v <- dd-hat, d <- dd-hat * d-hat (BN2-backward sums)
        if (nv == 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 m = cf[j];
            const float dh = (d[j] - m.x) * m.y;
            const float o = v[j] * act_df<ACT>(dh);
            *qp = o;
            qp += HWo;
            v[j] = o;
            d[j] = o * dh;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nv) {
              const float2 m = cf[j];
              const float dh = (d[j] - m.x) * m.y;
              const float o = v[j] * act_df<ACT>(dh);
              *qp = o;
              v[j] = o;
              d[j] = o * dh;
            }
            qp += HWo;
          }
        }
      }
    }
    if (gated) {
      if (!two_img) {                    // tiny planes (HWo < 32): more than two images per warp
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (e.v && j < nv) atomicAdd(&dg[(size_t)e.n * P.MCse + soff0 + c0 + j], v[j]);
      } else if (n_last <= n_first) {    // the usual case: all 32 pixels belong to one image
        const float sa = warp_sum16(v);
        if (lane < 16 && col < c_hi) atomicAdd(&dg[(size_t)n_first * P.MCse + soff0 + col], sa);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) d[j] = (e.n == n_first) ? 0.f : v[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (e.n == n_first) ? v[j] : 0.f;
        const float sa = warp_sum16(v), sb = warp_sum16(d);
        if (lane < 16 && col < c_hi) {
          atomicAdd(&dg[(size_t)n_first * P.MCse + soff0 + col], sa);
          atomicAdd(&dg[(size_t)n_last * P.MCse + soff0 + col], sb);
        }
      }
    } else {
      const float s1 = warp_sum16(v), s2 = warp_sum16(d);
      if (lane < 16 && col < c_hi) {
        atomicAdd(&sacc[2 * col], (double)s1);
        atomicAdd(&sacc[2 * col + 1], (double)s2);
      }
    }
  }
  if (!gated) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * ncol; i += UM_NT) atomicAdd(&sD[2 * cst0 + i], sacc[i]);
  }
  um_teardown(tmem, W.Nc);
}

// -------------------------------------------------------------------------------------------------
// B3b: dx_main = sum_i W1_i^T (r1 * du-hat), K = stacked mid channels (per-candidate chunks of 32)
// -------------------------------------------------------------------------------------------------
template <int ACT>
struct DxF {
  const Plan& P; const UmW& W; const DxChunks& CH; const float* DA; const float* UH; const float* DAb; const float* UHb; double* sU; Px4 px;
  int ch0, ch1; int lane, warp;
  struct Regs {};
  __device__ int nchunks() const { return ch1 - ch0; }
  __device__ void locate(int c, int& slot, int& k0) const {
    // static indices only: a dynamically indexed by-value parameter would be copied to local memory
    const int g = ch0 + c;
    int f0 = 0;
    slot = 0;
#pragma unroll
    for (int s = 1; s < TFNAS_MAX_OPS; ++s)
      if (s < P.na && g >= CH.first[s]) { slot = s; f0 = CH.first[s]; }
    k0 = (g - f0) * UM_KC;
  }
  // ring stage: [DA rows | UH rows], each UM_RW x UM_NT x 16 B
  __device__ void issue(int c, unsigned char* slot_p) const {
    int slot, k0;
    locate(c, slot, k0);
    const Cand& cd = P.c[slot];
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      const int k = k0 + warp + i * (UM_NT / 32);
      const bool ok = k < cd.mc;
      const int cst = cd.coff + k;
      ring_row(slot_p + ((size_t)i * UM_NT + threadIdx.x) * 16, DA, DAb, px, P.MC, cst, P.HW, ok);
      ring_row(slot_p + ((size_t)(UM_RW + i) * UM_NT + threadIdx.x) * 16, UH, UHb, px, P.MC, cst, P.HW, ok);
    }
  }
  __device__ void consts(int, Regs&) const {}
  // Rows past the candidate's width and invalid pixels are zero-filled in the ring, so du = 0 * act'(0) = 0 there.
  // BN1's rstd is folded into the prepped weights (umma_prep_bwd), the operand is plain du-hat.
  __device__ void compute(int c, const unsigned char* slot_p, const Regs&, UmTile& t) const {
    int slot, k0;
    locate(c, slot, k0);
    const Cand& cd = P.c[slot];
    float st[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) st[i] = 0.f;
#pragma unroll
    for (int i = 0; i < UM_RW; ++i) {
      (void)0;
      const float4 a = *(const float4*)(slot_p + ((size_t)i * UM_NT + threadIdx.x) * 16);
      const float4 b = *(const float4*)(slot_p + ((size_t)(UM_RW + i) * UM_NT + threadIdx.x) * 16);
      const float da[4] = {a.x, a.y, a.z, a.w}, uh[4] = {b.x, b.y, b.z, b.w};
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = da[e] * act_df<ACT>(uh[e]);
        st[2 * i] += v[e];
        st[2 * i + 1] += v[e] * uh[e];
      }
      um_split(t, i, v);
    }
    // 2*UM_RW statistics (rows x {sum du, sum du*uh}) reduced together; lane l < 2*UM_RW ends up owning statistic l
    const float tot = warp_sum16(st);
    static_assert(UM_RW == 2, "statistic ownership below assumes two rows per warp per chunk");
    const int k = k0 + warp + ((lane & 2) ? UM_NT / 32 : 0);
    if (lane < 2 * UM_RW && k < cd.mc) atomicAdd(&sU[2 * (cd.coff + k) + (lane & 1)], (double)tot);
  }
  __device__ const void* wsrc(int c) const { return (const char*)W.wp + (size_t)(ch0 + c) * 2 * W.Nc * 128; }
};

template <int ACT, int RS>
__global__ void __launch_bounds__(UM_NT, 2) k_um_dx(Plan P, UmW W, DxChunks CH, int ksplit, const float* __restrict__ DA,
                                               const float* __restrict__ UH,
                                               float* __restrict__ dx, double* __restrict__ sU, int nb) {
  extern __shared__ __align__(1024) unsigned char um_raw[];
  UmTrace tr;
  tr.begin();
  UmSmem S;
  um_carve(um_raw, W.Nc, S, nb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch0 = (int)((long long)CH.total * blockIdx.y / ksplit), ch1 = (int)((long long)CH.total * (blockIdx.y + 1) / ksplit);
  DxF<ACT> f{P, W, CH, DA, UH, DA, UH, sU, Px4(), ch0, ch1, lane, warp};
  px_decomp(f.px, blockIdx.x * 128 + lane * 4, P.P, P.HW);
  {
    const size_t o = (size_t)f.px.n[0] * P.MC * P.HW + f.px.hw[0];
    f.DAb = DA + o;
    f.UHb = UH + o;
  }
  typename DxF<ACT>::Regs rg;
  um_ring_prefetch<RS>(f, S, um_ring_stage_bytes(2), rg);
  const uint32_t tmem = um_setup(S, W.Nc);
  um_mainloop_ring<RS>(f, S, um_ring_stage_bytes(2), W.Nc, tmem, idesc_tf32(128, W.Nc, 1, 0), rg, tr);
  const EpiPx e = epi_px(blockIdx.x * 128, P.P, P.HW);
  int c_lo, c_hi;
  epi_cols(W.Nc, c_lo, c_hi);
  c_hi = min(c_hi, P.ic);
  if (ch1 > ch0) {
    const size_t HW = (size_t)P.HW;
    float* ob = dx + (size_t)e.n * P.ic * HW + e.hw;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float v[16];
      tmem_ld16(epi_taddr(tmem, c0), v);
      if (!e.v) continue;
      float* q = ob + (size_t)c0 * HW;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (c0 + j < c_hi) {
          if (ksplit == 1) *q = v[j];
          else atomicAdd(q, v[j]);
        }
        q += HW;
      }
    }
  }
  tr.mark(8);
  tr.end(2, ch1 - ch0);
  um_teardown(tmem, W.Nc);
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
static int g_use_umma = -1;
int umma_enabled() {
  if (g_use_umma < 0) {
    const char* e = getenv("TFNAS_GEMM");
    g_use_umma = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return g_use_umma;
}

// bytes of prepped weights for a [Nout x K] GEMM
#define UM_EXPAND_NCAP (ws_enabled(0) ? 192 : 256)   // the persistent kernels keep two weight slots + two accumulators on chip
#define UM_DC_NCAP (ws_enabled(2) ? 128 : 256)
static size_t um_prep_bytes(int Nout, int K, int cap = 256) {
  int Nc, nN;
  um_tile(Nout, Nc, nN, cap);
  return (size_t)nN * cdiv(K, UM_KC) * 2 * Nc * 128;
}
size_t umma_fwd_prep_bytes(const Plan& P) {
  size_t b = 0;
  for (int s = 0; s < P.na; ++s) b += um_prep_bytes(P.c[s].mc, P.ic, UM_EXPAND_NCAP) + um_prep_bytes(P.oc, P.c[s].mc);
  return b + 1024;
}
size_t umma_bwd_prep_bytes(const Plan& P) {
  size_t b = 0;
  int chunks = 0;
  for (int s = 0; s < P.na; ++s) { b += um_prep_bytes(P.c[s].mc, P.oc, UM_DC_NCAP); chunks += cdiv(P.c[s].mc, UM_KC); }
  int Nc, nN;
  um_tile(P.ic, Nc, nN);
  return b + (size_t)chunks * 2 * Nc * 128 + 1024;
}

struct PrepJob { const float* src; const float* kscale; float* dst; int ld_r, ld_k, nrows, K, Nc, nN, nK; };   // kscale: optional per-k factor
#define PREP_MAXJ 64          // jobs per launch: 64 x 56 B of kernel parameters
struct PrepJobs { int n; PrepJob j[PREP_MAXJ]; };

// grid (max nK, max nN, jobs): all candidates' weights of one GEMM in ONE launch.  A CTA converts one (K chunk, N chunk) tile:
// coalesced loads along whichever index is contiguous in the source into a shared-memory tile, then one 16-byte chunk (4
// consecutive k of a row) per thread: tf32 hi / lo split and two vector stores into the swizzled K-major layout.
__global__ void __launch_bounds__(256) k_umma_prep_all(PrepJobs J) {
  const PrepJob& q = J.j[blockIdx.z];
  const int kc = blockIdx.x, nc = blockIdx.y;
  if (kc >= q.nK || nc >= q.nN) return;
  __shared__ float tile[UM_KC][256 + 1];
  char* base = (char*)q.dst + ((size_t)nc * q.nK + kc) * 2 * q.Nc * 128;
#pragma unroll 4
  for (int i = threadIdx.x; i < q.Nc * UM_KC; i += blockDim.x) {      // unrolled: four independent loads in flight
    int r, kk;
    if (q.ld_k == 1) { r = i / UM_KC; kk = i - r * UM_KC; }
    else { kk = i / q.Nc; r = i - kk * q.Nc; }
    const int gr = nc * q.Nc + r, k = kc * UM_KC + kk;
    float x = (gr < q.nrows && k < q.K) ? q.src[(size_t)gr * q.ld_r + (size_t)k * q.ld_k] : 0.f;
    if (q.kscale && k < q.K) x *= q.kscale[k];
    tile[kk][r] = x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < q.Nc * (UM_KC / 4); i += blockDim.x) {
    const int r = i / (UM_KC / 4), kk = (i - r * (UM_KC / 4)) * 4;
    float4 hi, lo;
    split_tf32(tile[kk][r], hi.x, lo.x);
    split_tf32(tile[kk + 1][r], hi.y, lo.y);
    split_tf32(tile[kk + 2][r], hi.z, lo.z);
    split_tf32(tile[kk + 3][r], hi.w, lo.w);
    *(float4*)(base + k_elem_off(r, kk)) = hi;
    *(float4*)(base + (size_t)q.Nc * 128 + k_elem_off(r, kk)) = lo;
  }
}

static void prep(PrepJobs& J, const float* src, int ld_r, int ld_k, int nrows, int K, UmW& W, float*& cursor,
                 int cap = 256, const float* kscale = nullptr) {
  um_tile(nrows, W.Nc, W.nN, cap);
  W.nK = cdiv(K, UM_KC);
  W.Nout = nrows;
  W.wp = cursor;
  PrepJob& q = J.j[J.n++];
  q.src = src; q.kscale = kscale; q.dst = cursor; q.ld_r = ld_r; q.ld_k = ld_k; q.nrows = nrows; q.K = K; q.Nc = W.Nc; q.nN = W.nN; q.nK = W.nK;
  cursor += (size_t)W.nN * W.nK * 2 * W.Nc * 32;
}
// Deferred prep: between umma_prep_batch_begin() and umma_prep_batch_flush() the per-MixedOP prep calls only compute their
// geometry and queue their jobs; the flush converts the weights of ALL queued MixedOPs in a few launches (the body executor
// knows every weight of a pass up front: 2 launches per sampled pass instead of 36 on its critical chain).
static thread_local std::vector<PrepJob>* g_prep_batch = nullptr;
static void prep_launch(const PrepJobs& J, cudaStream_t st);
void umma_prep_batch_begin() {
  static thread_local std::vector<PrepJob> batch;
  batch.clear();
  g_prep_batch = &batch;
}
void umma_prep_batch_flush(cudaStream_t st) {
  std::vector<PrepJob>* b = g_prep_batch;
  g_prep_batch = nullptr;
  if (!b) return;
  for (size_t i0 = 0; i0 < b->size(); i0 += PREP_MAXJ) {
    PrepJobs J;
    J.n = (int)min((size_t)PREP_MAXJ, b->size() - i0);
    for (int i = 0; i < J.n; ++i) J.j[i] = (*b)[i0 + i];
    prep_launch(J, st);
  }
  b->clear();
}
static void prep_launch(const PrepJobs& J, cudaStream_t st) {
  if (g_prep_batch) {
    for (int i = 0; i < J.n; ++i) g_prep_batch->push_back(J.j[i]);
    return;
  }
  int mk = 0, mn = 0;
  double bytes = 0;
  for (int i = 0; i < J.n; ++i) { mk = max(mk, J.j[i].nK); mn = max(mn, J.j[i].nN); bytes += 12.0 * J.j[i].nrows * J.j[i].K; }
  ProfScope ps("um_prep_w", bytes, 0, st);
  k_umma_prep_all<<<dim3(mk, mn, J.n), 256, 0, st>>>(J);
}

// all forward weights (expand W1, project W3) of the active candidates in one prep launch
void umma_prep_fwd(const Plan& P, float* prep_buf, UmWAll& WE, UmWAll& WP, cudaStream_t st) {
  PrepJobs J;
  J.n = 0;
  float* cur = prep_buf;
  for (int s = 0; s < P.na; ++s) prep(J, P.c[s].w1, P.ic, 1, P.c[s].mc, P.ic, WE.s[s], cur, UM_EXPAND_NCAP);
  for (int s = 0; s < P.na; ++s) prep(J, P.c[s].w3, P.c[s].mc, 1, P.oc, P.c[s].mc, WP.s[s], cur);
  prep_launch(J, st);
}

void umma_prep_project(const Plan& P, float* prep_buf, UmWAll& WP, cudaStream_t st) {
  PrepJobs J;
  J.n = 0;
  float* cur = prep_buf;
  for (int s = 0; s < P.na; ++s) prep(J, P.c[s].w3, P.c[s].mc, 1, P.oc, P.c[s].mc, WP.s[s], cur);
  prep_launch(J, st);
}

void umma_prep_dc(const Plan& P, float* prep_buf, UmWAll& WD, cudaStream_t st) {
  PrepJobs J;
  J.n = 0;
  float* cur = prep_buf;
  for (int s = 0; s < P.na; ++s) prep(J, P.c[s].w3, 1, P.c[s].mc, P.c[s].mc, P.oc, WD.s[s], cur, UM_DC_NCAP);
  prep_launch(J, st);
}

static void um_max(const Plan& P, const UmWAll& WA, int& maxN, int& maxNc) {
  maxN = 0; maxNc = 0;
  for (int s = 0; s < P.na; ++s) { maxN = max(maxN, WA.s[s].nN); maxNc = max(maxNc, WA.s[s].Nc); }
}

#define UM_SMEM_2CTA 115712   // (228 KB - 2 x 1 KB reserved) / 2

void umma_expand(const Plan& P, const UmWAll& WA, const float* x, const float* bn1, float* UH, cudaStream_t st) {
  if (ws_enabled(0) && ws_expand(P, WA, x, bn1, UH, st)) return;
  int maxN, maxNc;
  um_max(P, WA, maxN, maxNc);
  size_t smem = um_smem_bytes(maxNc);
  ensure_smem(k_um_expand, (size_t)(smem));
  ProfScope ps("expand", 4.0 * P.P * P.ic + 4.0 * P.P * P.MC + 4.0 * P.MC * P.ic, 2.0 * P.P * (double)P.MC * P.ic, st);
  k_um_expand<<<dim3(cdiv(P.P, 128), maxN, P.na), UM_NT, smem, st>>>(P, WA, x, bn1, UH);
}

#define UM_SMEM_1CTA 232448   // 227 KB opt-in maximum

// largest weight-slot count nb in [1, want] whose smem footprint stays within `limit` (0: not even one fits)
static int um_fit_nb(int Nc, size_t ring_bytes, int want, size_t limit) {
  for (int nb = want; nb >= 1; --nb)
    if (um_smem_bytes(Nc, ring_bytes, nb) <= limit) return nb;
  return 0;
}

template <int ACT, int RS>
static void launch_um_project(const Plan& P, const UmWAll& WA, dim3 grid, size_t smem, int nb, const float* D, const float* bn2,
                              const float* seg, float* Zb, double* st3, cudaStream_t st) {
  ensure_smem(k_um_project<ACT, RS>, (size_t)(smem));
  k_um_project<ACT, RS><<<grid, UM_NT, smem, st>>>(P, WA, D, bn2, seg, Zb, st3, nb);
}

void umma_project(const Plan& P, const UmWAll& WA, const float* D, const float* bn2, const float* seg, float* Zb,
                  double* st3, cudaStream_t st) {
  if (ws_enabled(1) && ws_project(P, WA, D, bn2, seg, Zb, st3, st)) return;
  int maxN, maxNc, maxK = 1;
  um_max(P, WA, maxN, maxNc);
  for (int s = 0; s < P.na; ++s) maxK = max(maxK, WA.s[s].nK);
  const int want = min(4, maxK);
  const size_t rs1 = um_ring_stage_bytes(1);
  // preference: two CTAs per SM with weights fetched >= 1 chunk ahead (nb >= 2) and 3, else 2 ring stages; else one CTA
  // per SM with a 5-stage ring; last resort two CTAs with a single weight slot
  int RS = 3, nb = um_fit_nb(maxNc, 3 * rs1, want, UM_SMEM_2CTA);
  if (nb < min(2, want)) {
    const int nb2 = um_fit_nb(maxNc, 2 * rs1, want, UM_SMEM_2CTA);
    if (nb2 >= min(2, want)) { RS = 2; nb = nb2; }
    else { RS = 5; nb = um_fit_nb(maxNc, 5 * rs1, want, UM_SMEM_1CTA); }
  }
  const size_t smem = um_smem_bytes(maxNc, RS * rs1, nb);
  dim3 grid(cdiv(P.Q, 128), maxN, P.na);
  ProfScope ps("project", 4.0 * P.Q * ((double)P.MC + (double)P.na * P.oc) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  const bool relu = P.act == TFNAS_ACT_RELU;
#define UM_PROJ(RS_) \
  (relu ? launch_um_project<TFNAS_ACT_RELU, RS_>(P, WA, grid, smem, nb, D, bn2, seg, Zb, st3, st) \
        : P.act == TFNAS_ACT_NONE ? launch_um_project<TFNAS_ACT_NONE, RS_>(P, WA, grid, smem, nb, D, bn2, seg, Zb, st3, st) \
        : launch_um_project<TFNAS_ACT_SWISH, RS_>(P, WA, grid, smem, nb, D, bn2, seg, Zb, st3, st))
  if (RS == 3) UM_PROJ(3);
  else if (RS == 2) UM_PROJ(2);
  else UM_PROJ(5);
#undef UM_PROJ
}

// all backward weights (dc: W3^T, dx: W1^T with per-candidate K chunks) in one prep launch
void umma_prep_bwd(const Plan& P, const float* bn1, float* prep_buf, UmWAll& WD, UmW& WX, DxChunks& CH, cudaStream_t st) {
  PrepJobs J;
  J.n = 0;
  float* cur = prep_buf;
  // logical weight (row = mid channel c, k = out channel o) = W3[o][c]
  for (int s = 0; s < P.na; ++s) prep(J, P.c[s].w3, 1, P.c[s].mc, P.c[s].mc, P.oc, WD.s[s], cur, UM_DC_NCAP);
  um_tile(P.ic, WX.Nc, WX.nN);    // ic <= 192 -> one N chunk
  WX.Nout = P.ic;
  WX.wp = cur;
  CH.first[0] = 0;
  for (int s = 0; s < P.na; ++s) {
    // logical weight (row = input channel k', k = mid channel c) = W1[c][k'], chunks never straddle candidates
    UmW Ws;
    // BN1's rstd (per mid channel = per k) is folded in here, so the dx prologue emits plain du-hat
    prep(J, P.c[s].w1, 1, P.ic, P.ic, P.c[s].mc, Ws, cur, 256, bn1 + P.MC + P.c[s].coff);
    CH.first[s + 1] = CH.first[s] + Ws.nK;
  }
  CH.total = CH.first[P.na];
  WX.nK = CH.total;
  prep_launch(J, st);
}

void umma_dc(const Plan& P, const UmWAll& WA, const float* G, const float* Zb, const float4* dzc2,
             const float* D, const float* bn2, float* DC, float* dg, double* sD, cudaStream_t st) {
  if (ws_enabled(2) && ws_dc(P, WA, G, Zb, dzc2, D, bn2, DC, dg, sD, st)) return;
  int maxN, maxNc;
  um_max(P, WA, maxN, maxNc);
  size_t smem = um_smem_bytes(maxNc, (size_t)P.oc * sizeof(float4) + 16 + (size_t)2 * maxNc * sizeof(double));     // + coefficient table + sums
  dim3 grid(cdiv(P.Q, 128), maxN, P.na);
  ProfScope ps("dc", 4.0 * P.Q * ((double)P.oc * (1 + P.na) + 2.0 * P.MC) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  if (P.act == TFNAS_ACT_RELU) {
    ensure_smem(k_um_dc<TFNAS_ACT_RELU>, (size_t)(smem));
    k_um_dc<TFNAS_ACT_RELU><<<grid, UM_NT, smem, st>>>(P, WA, G, Zb, dzc2, D, bn2, DC, dg, sD);
  } else if (P.act == TFNAS_ACT_NONE) {
    ensure_smem(k_um_dc<TFNAS_ACT_NONE>, (size_t)(smem));
    k_um_dc<TFNAS_ACT_NONE><<<grid, UM_NT, smem, st>>>(P, WA, G, Zb, dzc2, D, bn2, DC, dg, sD);
  } else {
    ensure_smem(k_um_dc<TFNAS_ACT_SWISH>, (size_t)(smem));
    k_um_dc<TFNAS_ACT_SWISH><<<grid, UM_NT, smem, st>>>(P, WA, G, Zb, dzc2, D, bn2, DC, dg, sD);
  }
}

template <int ACT, int RS>
static void launch_um_dx(const Plan& P, const UmW& W, const DxChunks& CH, dim3 grid, int ksplit, size_t smem, int nb,
                         const float* DA, const float* UH, float* dx, double* sU, cudaStream_t st) {
  ensure_smem(k_um_dx<ACT, RS>, (size_t)(smem));
  k_um_dx<ACT, RS><<<grid, UM_NT, smem, st>>>(P, W, CH, ksplit, DA, UH, dx, sU, nb);
}

void umma_dx(const Plan& P, const UmW& W, const DxChunks& CH, const float* DA, const float* UH, const float* bn1,
             float* dx, double* sU, cudaStream_t st) {
  (void)bn1;   // folded into the prepped weights
  if (ws_enabled(3) && ws_dx(P, W, CH, DA, UH, dx, sU, st)) return;
  const int tiles = cdiv(P.P, 128);
  const size_t rs2 = um_ring_stage_bytes(2);
  // 2 ring stages (64 KB) with two CTAs per SM when that fits, else one CTA per SM with 4 stages
  int nb = um_fit_nb(W.Nc, 2 * rs2, 4, UM_SMEM_2CTA);
  const bool two = nb >= 1;
  if (!two) nb = um_fit_nb(W.Nc, 4 * rs2, 4, UM_SMEM_1CTA);
  const size_t smem = um_smem_bytes(W.Nc, (two ? 2 : 4) * rs2, nb);
  int ksplit = max(1, min(CH.total, cdiv((two ? 2 : 1) * sm_count(), tiles)));
  if (ksplit > 1) cudaMemsetAsync(dx, 0, (size_t)P.P * P.ic * sizeof(float), st);
  dim3 grid(tiles, ksplit);
  ProfScope ps("dx", 4.0 * P.P * (2.0 * P.MC + P.ic) + 4.0 * P.MC * P.ic, 2.0 * P.P * (double)P.MC * P.ic, st);
  const bool relu = P.act == TFNAS_ACT_RELU;
  if (two) {
    if (relu) launch_um_dx<TFNAS_ACT_RELU, 2>(P, W, CH, grid, ksplit, smem, nb, DA, UH, dx, sU, st);
    else launch_um_dx<TFNAS_ACT_SWISH, 2>(P, W, CH, grid, ksplit, smem, nb, DA, UH, dx, sU, st);
  } else {
    if (relu) launch_um_dx<TFNAS_ACT_RELU, 4>(P, W, CH, grid, ksplit, smem, nb, DA, UH, dx, sU, st);
    else launch_um_dx<TFNAS_ACT_SWISH, 4>(P, W, CH, grid, ksplit, smem, nb, DA, UH, dx, sU, st);
  }
}

// -------------------------------------------------------------------------------------------------
// Weight gradients on the tensor cores (sampled w-step):  Out[a][b] += sum_p U[a][p] * V[b][p]
//   MODE 0 (W3): a = mid channel c (U = c-tilde = act(BN2(d))*gate), b = out channel o (V = dz)   -> dW3[o][c]
//   MODE 1 (W1): a = mid channel c (U = du-hat = DA*act'(UH)),        b = in  channel k (V = x)    -> SmatT[k][c]
//   MODE 2 (input covariance, F0): a, b = in channels, U = V = x - mean (means from the k_xsum pass)
//                 -> double accumulators  sum_p (x_a - m_a)(x_b - m_b)   (ic x ic)
// Both operands are K-major (pixels contiguous), staged by the CTA's threads with the tf32 split.
// grid (mc/128, N chunks, K splits over the pixel axis); fp32 atomics combine the K splits.
// -------------------------------------------------------------------------------------------------
struct WgArgs {
  const float* A0; const float* A1;   // MODE 0: D, -      MODE 1: DA, UH
  const float* B0; const float* B1;   // MODE 0: G, Z      MODE 1: x, -
  const float* bn2; const float* seg; const float* bn3; const float4* dzc;
  float* out;                         // MODE 0: dW3 [oc][mc]   MODE 1: SmatT [ic][mc]   MODE 2: double [ic + ic*ic]
  int Nc, nN;                         // N chunking of the b axis
  int hoist;                          // every tensor < 2^32 elements and HW >= 4: 32-bit row + pixel offsets (set by the launcher)
};

// NBR = B rows per thread (Nc <= 32 * NBR).  Software pipeline per 32-pixel K chunk: the raw rows of chunk c+1 are
// loaded into registers right after chunk c's operands are stored, so they fly during the barrier, the MMA issue and
// the tensor core's round trip; per-row constants (BN2 mean / rstd, BN3-backward coefficients) are hoisted.
// VEC (HW % 4 == 0): a thread's 4 pixels are one aligned 16-byte group of one image; every row is then ONE vector load at
// (per-row element offset, fixed) + (per-chunk pixel offset, shared by all rows) -- the general path recomputes
// ((n * C + ch) * HW + hw) per row and pixel, which made integer address arithmetic the bulk of the instruction stream.
template <int MODE, int ACT, int NBR, bool VEC>
__global__ void __launch_bounds__(NT) k_um_wgrad(Plan P, int slot, WgArgs g) {
  extern __shared__ __align__(1024) unsigned char um_raw[];
  unsigned char* sm = um_raw + ((1024u - (smem_u32(um_raw) & 1023u)) & 1023u);
  const int Nc = g.Nc;
  unsigned char* a_hi = sm;
  unsigned char* a_lo = sm + 16384;
  unsigned char* b_hi = sm + 32768;
  unsigned char* b_lo = b_hi + Nc * 128;
  uint64_t* bar_mma = (uint64_t*)(b_lo + Nc * 128);
  uint32_t* tmem_slot = (uint32_t*)(bar_mma + 1);
  const Cand& cd = P.c[slot];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * Nc;
  const int nb_total = MODE == 0 ? P.oc : P.ic;
  const int HWp = MODE == 0 ? P.HWo : P.HW;
  const int total = MODE == 0 ? P.Q : P.P;
  const int am = MODE == 2 ? P.ic : cd.mc;          // valid A rows
  const int acoff = MODE == 2 ? 0 : cd.coff;        // channel offset of A row 0 in its tensor
  const int CA = MODE == 2 ? P.ic : P.MC;           // channels of the A tensor
  const int nsplit = gridDim.z;
  const int p_lo = (int)((long long)total * blockIdx.z / nsplit) / 32 * 32;
  const int p_hi = (int)blockIdx.z + 1 < nsplit ? (int)((long long)total * (blockIdx.z + 1) / nsplit) / 32 * 32 : total;
  if (tid == 0) { mbar_init(bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols(Nc));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = idesc_tf32(128, Nc, 0, 0);
  const int q = tid & 7;               // 16 B chunk (4 pixels) of the 32-pixel K chunk owned by this thread
  const int r_base = tid >> 3;         // rows r_base + 32 i
  // ---- per-thread row constants ----
  int a_cst[4];                        // stacked mid channel of A row i, -1 past the candidate's width
  float a_mu[4], a_r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = m0 + r_base + 32 * i;
    a_cst[i] = c < am ? acoff + c : -1;
    a_mu[i] = a_r[i] = 0.f;
    if (MODE == 0 && a_cst[i] >= 0) { a_mu[i] = g.bn2[a_cst[i]]; a_r[i] = g.bn2[P.MC + a_cst[i]]; }
    if (MODE == 2 && a_cst[i] >= 0) a_mu[i] = (float)(((const double*)g.out)[a_cst[i]] / (double)P.P);   // channel mean
  }
  int b_row[NBR];                      // B channel of row j, -1 when absent
  float4 b_cf[MODE != 1 ? NBR : 1];    // MODE 0: dz = cf.x*g + cf.y*z + cf.z ; MODE 2: cf.x = channel mean
#pragma unroll
  for (int j = 0; j < NBR; ++j) {
    const int r = r_base + 32 * j, b = n0 + r;
    b_row[j] = (r < Nc && b < nb_total) ? b : -1;
    if (MODE == 0) b_cf[j] = b_row[j] >= 0 ? g.dzc[slot * P.oc + b] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 2) b_cf[j] = make_float4(b_row[j] >= 0 ? (float)(((const double*)g.out)[b] / (double)P.P) : 0.f, 0.f, 0.f, 0.f);
  }
  const bool vecshape = VEC || (HWp & 3) == 0;
  const bool gated = MODE == 0 && cd.se > 0;
  // VEC: element offsets of the rows inside their tensors (all tensors of this path hold < 2^32 elements)
  // (the scalar instantiation uses the same row offsets with four per-pixel offsets per chunk when g.hoist is set)
  constexpr bool OFFS = MODE != 2;
  uint32_t a_off[OFFS ? 4 : 1], a1_off[(OFFS && MODE == 1) ? 4 : 1], g_off[OFFS ? 4 : 1], b_off[OFFS ? NBR : 1], b1_off[(OFFS && MODE == 0) ? NBR : 1];
  const bool hoist = !VEC && OFFS && g.hoist;
  float rgt1[(!VEC && MODE == 0) ? 4 : 1];      // scalar path: SE gate of the image the quad's LAST pixel belongs to
  if (OFFS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cst = a_cst[i] >= 0 ? a_cst[i] : 0;
      a_off[i] = (uint32_t)cst * (uint32_t)HWp;
      if (MODE == 1) a1_off[i] = (uint32_t)cst * (uint32_t)HWp;
      g_off[i] = (uint32_t)(cd.soff + (cst - cd.coff));
    }
#pragma unroll
    for (int j = 0; j < NBR; ++j) {
      const int ch = b_row[j] >= 0 ? b_row[j] : 0;
      b_off[j] = (uint32_t)ch * (uint32_t)HWp;
      if (MODE == 0) b1_off[j] = (uint32_t)(slot * P.oc + ch) * (uint32_t)HWp;
    }
  }
  // ---- raw registers of one chunk ----
  float4 ra0[4], ra1[MODE == 1 ? 4 : 1], rb0[NBR], rb1[MODE == 0 ? NBR : 1];
  float rgt[4];                        // MODE 0: SE gate per A row (vector path: one image per 4 pixels)
  Px4 px;
  int p = p_lo + q * 4, n = 0, hw = 0;
  if (p < p_hi) { n = fast_div(p, HWp, __frcp_rn((float)HWp)); hw = p - n * HWp; }
  auto load_chunk = [&]() {
    if (VEC) {
      const bool ok = p < p_hi;        // p, p_hi are multiples of 4: all four pixels or none
      px.vec = ok;
#pragma unroll
      for (int e = 0; e < 4; ++e) { px.v[e] = ok; px.n[e] = ok ? n : 0; px.hw[e] = ok ? hw + e : 0; }
      const uint32_t nhw = ok ? (uint32_t)n * (uint32_t)HWp : 0u, hw0 = ok ? (uint32_t)hw : 0u;
      const uint32_t pa = nhw * (uint32_t)CA + hw0, pa1 = nhw * (uint32_t)P.MC + hw0;
      const uint32_t pb = nhw * (uint32_t)(MODE == 0 ? P.oc : P.ic) + hw0, pb1 = nhw * (uint32_t)(P.na * P.oc) + hw0;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool rok = ok && a_cst[i] >= 0;
        ra0[i] = rok ? *(const float4*)(g.A0 + (size_t)(a_off[i] + pa)) : z4;
        if (MODE == 1) ra1[i] = rok ? *(const float4*)(g.A1 + (size_t)(a1_off[i] + pa1)) : z4;
        rgt[i] = (gated && rok) ? g.seg[(size_t)((uint32_t)n * (uint32_t)P.MCse + g_off[i])] : 1.f;
      }
#pragma unroll
      for (int j = 0; j < NBR; ++j) {
        const bool rok = ok && b_row[j] >= 0;
        rb0[j] = rok ? *(const float4*)(g.B0 + (size_t)(b_off[j] + pb)) : z4;
        if (MODE == 0) rb1[j] = rok ? *(const float4*)(g.B1 + (size_t)(b1_off[j] + pb1)) : z4;
      }
      return;
    }
    if (OFFS && hoist) {               // general planes (e.g. 7x7): scalar loads at row offset + per-pixel offset
      px_decomp(px, p, p_hi, HWp);
      uint32_t pa[4], pa1[MODE == 1 ? 4 : 1], pb[4], pb1[MODE == 0 ? 4 : 1];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t nhw = (uint32_t)px.n[e] * (uint32_t)HWp, h = (uint32_t)px.hw[e];
        pa[e] = nhw * (uint32_t)CA + h;
        if (MODE == 1) pa1[e] = nhw * (uint32_t)P.MC + h;
        pb[e] = nhw * (uint32_t)(MODE == 0 ? P.oc : P.ic) + h;
        if (MODE == 0) pb1[e] = nhw * (uint32_t)(P.na * P.oc) + h;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool rok = a_cst[i] >= 0;
        float d[4], u[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          d[e] = (rok && px.v[e]) ? g.A0[(size_t)(a_off[i] + pa[e])] : 0.f;
          if (MODE == 1) u[e] = (rok && px.v[e]) ? g.A1[(size_t)(a1_off[i] + pa1[e])] : 0.f;
        }
        ra0[i] = make_float4(d[0], d[1], d[2], d[3]);
        if (MODE == 1) ra1[i] = make_float4(u[0], u[1], u[2], u[3]);
        rgt[i] = 1.f;
        if (MODE == 0) {
          rgt1[i] = 1.f;
          if (gated && rok) {
            rgt[i] = g.seg[(size_t)((uint32_t)px.n[0] * (uint32_t)P.MCse + g_off[i])];
            rgt1[i] = g.seg[(size_t)((uint32_t)px.n[3] * (uint32_t)P.MCse + g_off[i])];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NBR; ++j) {
        const bool rok = b_row[j] >= 0;
        float v0[4], v1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v0[e] = (rok && px.v[e]) ? g.B0[(size_t)(b_off[j] + pb[e])] : 0.f;
          if (MODE == 0) v1[e] = (rok && px.v[e]) ? g.B1[(size_t)(b1_off[j] + pb1[e])] : 0.f;
        }
        rb0[j] = make_float4(v0[0], v0[1], v0[2], v0[3]);
        if (MODE == 0) rb1[j] = make_float4(v1[0], v1[1], v1[2], v1[3]);
      }
      return;
    }
    if (vecshape) {                    // 4 valid pixels of one image (p, p_hi are multiples of 4)
      px.vec = p < p_hi;
#pragma unroll
      for (int e = 0; e < 4; ++e) { px.v[e] = px.vec; px.n[e] = px.vec ? n : 0; px.hw[e] = px.vec ? hw + e : 0; }
    } else {
      px_decomp(px, p, p_hi, HWp);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      if (a_cst[i] >= 0) load4(d, g.A0, px, CA, a_cst[i], HWp);
      ra0[i] = make_float4(d[0], d[1], d[2], d[3]);
      if (MODE == 1) {
        float u[4] = {0.f, 0.f, 0.f, 0.f};
        if (a_cst[i] >= 0) load4(u, g.A1, px, P.MC, a_cst[i], HWp);
        ra1[i] = make_float4(u[0], u[1], u[2], u[3]);
      }
      rgt[i] = 1.f;
      if (gated && a_cst[i] >= 0) rgt[i] = g.seg[(size_t)px.n[0] * P.MCse + cd.soff + (a_cst[i] - cd.coff)];
    }
#pragma unroll
    for (int j = 0; j < NBR; ++j) {
      float v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f};
      if (b_row[j] >= 0) {
        if (MODE == 0) {
          load4(v0, g.B0, px, P.oc, b_row[j], HWp);
          load4(v1, g.B1, px, P.na * P.oc, slot * P.oc + b_row[j], HWp);
        } else {
          load4(v0, g.B0, px, P.ic, b_row[j], HWp);
        }
      }
      rb0[j] = make_float4(v0[0], v0[1], v0[2], v0[3]);
      if (MODE == 0) rb1[j] = make_float4(v1[0], v1[1], v1[2], v1[3]);
    }
  };
  uint32_t phase = 0;
  bool first = true;
  if (p_lo < p_hi) load_chunk();
  for (int p0 = p_lo; p0 < p_hi; p0 += 32) {
    if (!first) {                       // previous chunk's MMAs must have consumed the tiles
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
    }
    // ---- A rows: mid channels m0 .. m0+127 ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r_base + 32 * i;
      const float d[4] = {ra0[i].x, ra0[i].y, ra0[i].z, ra0[i].w};
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (a_cst[i] >= 0) {
        if (MODE == 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float b = act_f<ACT>((d[e] - a_mu[i]) * a_r[i]);
            if (gated) {
              if (VEC || px.vec) b *= rgt[i];
              else if (!VEC && MODE == 0 && hoist) b *= px.n[e] == px.n[0] ? rgt[i] : rgt1[i];     // a quad spans <= 2 images
              else b *= g.seg[(size_t)px.n[e] * P.MCse + cd.soff + (a_cst[i] - cd.coff)];
            }
            v[e] = px.v[e] ? b : 0.f;
          }
        } else if (MODE == 1) {
          const float u[4] = {ra1[i].x, ra1[i].y, ra1[i].z, ra1[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? d[e] * act_df<ACT>(u[e]) : 0.f;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? d[e] - a_mu[i] : 0.f;      // centred
        }
      }
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
      const uint32_t off = r * 128u + ((uint32_t)(q ^ (r & 7)) << 4);
      *(float4*)(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
      *(float4*)(a_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
    // ---- B rows: n0 .. n0+Nc-1 ----
#pragma unroll
    for (int j = 0; j < NBR; ++j) {
      const int r = r_base + 32 * j;
      if (r < Nc) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (b_row[j] >= 0) {
          const float x0[4] = {rb0[j].x, rb0[j].y, rb0[j].z, rb0[j].w};
          if (MODE == 0) {
            const float z[4] = {rb1[j].x, rb1[j].y, rb1[j].z, rb1[j].w};
            const float4 cf = b_cf[j];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? fmaf(cf.x, x0[e], fmaf(cf.y, z[e], cf.z)) : 0.f;
          } else if (MODE == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = x0[e];      // invalid pixels were loaded as 0
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? x0[e] - b_cf[j].x : 0.f;
          }
        }
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
        const uint32_t off = r * 128u + ((uint32_t)(q ^ (r & 7)) << 4);
        *(float4*)(b_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *(float4*)(b_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
      }
    }
    // next chunk's raw rows: in flight across the barrier, the MMA issue and the tensor core's round trip
    p += 32;
    hw += 32;
    while (hw >= HWp) { hw -= HWp; ++n; }      // planes smaller than 32 pixels wrap more than once
    if (p0 + 32 < p_hi) load_chunk();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint64_t dah = smem_desc(ah + s * 32, 16, 1024, SWIZZLE_128B);
          const uint64_t dal = smem_desc(al + s * 32, 16, 1024, SWIZZLE_128B);
          const uint64_t dbh = smem_desc(bh + s * 32, 16, 1024, SWIZZLE_128B);
          const uint64_t dbl = smem_desc(bl + s * 32, 16, 1024, SWIZZLE_128B);
          mma_tf32(tmem, dah, dbh, idesc, (first && s == 0) ? 0u : 1u);
          mma_tf32(tmem, dal, dbh, idesc, 1u);
          mma_tf32(tmem, dah, dbl, idesc, 1u);
        }
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
    first = false;
  }
  if (!first) {
    mbar_wait(bar_mma, phase);
    tc_fence_after();
    // epilogue: lane = mid channel row, 16 b-columns per TMEM load; coalesced atomics along c
    const int c = m0 + (warp & 3) * 32 + lane;
    int c_lo, c_hi;
    epi_cols(Nc, c_lo, c_hi);
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float v[16];
      tmem_ld16(epi_taddr(tmem, c0), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int b = n0 + c0 + j;
        if (MODE == 2) {
          double* od = (double*)g.out;               // [ic] sums (input), then [ic][ic] centred second moments
          if (b < nb_total && c < am) atomicAdd(&od[P.ic + (size_t)c * P.ic + b], (double)v[j]);
        } else {
          if (b < nb_total && c < cd.mc) atomicAdd(&g.out[(size_t)b * cd.mc + c], v[j]);
        }
      }
    }
  }
  um_teardown(tmem, Nc);
}

template <int MODE, int ACT>
static void launch_um_wgrad(int nbr, dim3 grid, size_t smem, const Plan& P, int slot, const WgArgs& g, cudaStream_t st) {
  const int HWp = MODE == 0 ? P.HWo : P.HW;
  // vector rows: whole 16-byte groups per image, 32-bit element offsets (largest tensor: UH / DA with N * MC * HW elements)
  const bool small = (unsigned long long)P.N * P.MC * P.HW < (1ULL << 32) &&
                     (unsigned long long)P.N * P.na * P.oc * P.HWo < (1ULL << 32);
  const bool vec = MODE != 2 && (HWp & 3) == 0 && small;
  WgArgs gh = g;
  gh.hoist = (MODE == 0 && small && HWp >= 4) ? 1 : 0;      // (measured slower for the dW1 operands: 0.18 vs 0.146 ms at 7x7)
#define UM_WG(NBR_) do { \
    if (vec) { ensure_smem(k_um_wgrad<MODE, ACT, NBR_, true>, (size_t)(smem)); \
               k_um_wgrad<MODE, ACT, NBR_, true><<<grid, NT, smem, st>>>(P, slot, gh); } \
    else { ensure_smem(k_um_wgrad<MODE, ACT, NBR_, false>, (size_t)(smem)); \
           k_um_wgrad<MODE, ACT, NBR_, false><<<grid, NT, smem, st>>>(P, slot, gh); } } while (0)
  if (nbr <= 2) UM_WG(2);
  else if (nbr <= 4) UM_WG(4);
  else UM_WG(8);
#undef UM_WG
}

// dzc: the FOLDED BN3-backward coefficients (BwdScratch::dzc2): dz = A*g + B*z + C
void umma_wgrad(const Plan& P, int slot, int mode, const float* A0, const float* A1, const float* B0, const float* B1,
                const float* bn2, const float* seg, const float* bn3, const float4* dzc, float* out, cudaStream_t st) {
  const Cand& cd = P.c[slot];
  WgArgs g;
  g.A0 = A0; g.A1 = A1; g.B0 = B0; g.B1 = B1; g.bn2 = bn2; g.seg = seg; g.bn3 = bn3; g.dzc = dzc; g.out = out;
  const int nb = mode == 0 ? P.oc : P.ic;
  um_tile(nb, g.Nc, g.nN);
  const int total = mode == 0 ? P.Q : P.P;
  const int mt = cdiv(cd.mc, 128);
  // K splits over the pixel axis: enough CTAs for ~3 per SM, each at least 256 pixels long (the 14x14 / 7x7 stages have
  // only 25k / 6k pixels: with 1024-pixel splits a 1152-wide candidate ran on 54 CTAs)
  int nsplit = max(1, min(cdiv(total, 256), cdiv(3 * sm_count(), mt * g.nN)));
  // every split adds its partial tile with float atomics: keep (output elements x splits) bounded -- the head's
  // 1280 x 320 feature-mix gradient at 25 splits spent its time in 10 M atomics (cap: 6 M)
  // (cap swept on B200 with the vector-row kernels: dW1 2.56 / 2.31 / 2.09 / 2.06 ms per sampled pass at 1.5 / 3 / 6 / 12 M,
  //  dW3 flat from 3 M on; TFNAS_WG_ATOM overrides for A/B runs)
  static const long long atom_cap = [] { const char* e = getenv("TFNAS_WG_ATOM"); return e ? atoll(e) : 6000000LL; }();
  nsplit = max(1, min(nsplit, (int)(atom_cap / ((long long)cd.mc * nb))));
  size_t smem = 1024 + 32768 + (size_t)2 * g.Nc * 128 + 64;
  dim3 grid(mt, g.nN, nsplit);
  const bool relu = P.act == TFNAS_ACT_RELU;
  const int nbr = cdiv(g.Nc, 32);
  if (mode == 0) {
    ProfScope ps("wgrad_w3", 4.0 * P.Q * (2.0 * P.oc + cd.mc), 2.0 * P.Q * (double)P.oc * cd.mc, st);
    if (relu) launch_um_wgrad<0, TFNAS_ACT_RELU>(nbr, grid, smem, P, slot, g, st);
    else if (P.act == TFNAS_ACT_NONE) launch_um_wgrad<0, TFNAS_ACT_NONE>(nbr, grid, smem, P, slot, g, st);
    else launch_um_wgrad<0, TFNAS_ACT_SWISH>(nbr, grid, smem, P, slot, g, st);
  } else {
    ProfScope ps("wgrad_w1", 4.0 * P.P * (2.0 * cd.mc + P.ic), 2.0 * P.P * (double)P.ic * cd.mc, st);
    if (relu) launch_um_wgrad<1, TFNAS_ACT_RELU>(nbr, grid, smem, P, slot, g, st);
    else launch_um_wgrad<1, TFNAS_ACT_SWISH>(nbr, grid, smem, P, slot, g, st);
  }
}

// F0 on the tensor cores: acc = [sum_p x_k (ic, INPUT: from k_xsum) | sum_p (x_k - m_k)(x_l - m_l) (ic x ic, accumulated)]
// doubles.  Each CTA accumulates at most ~1.5k pixels in TMEM (fp32) before its partial goes to the double
// accumulators, so the rounding of the tensor core's fp32 accumulation stays ~1e-5 relative on the variances; the
// means (which decide ReLU gates) come from the fp64 k_xsum pass.
void umma_covariance(const Plan& P, const float* x, double* acc, cudaStream_t st) {
  WgArgs g;
  memset(&g, 0, sizeof(g));
  g.A0 = x; g.B0 = x; g.out = (float*)acc;
  um_tile(P.ic, g.Nc, g.nN);
  const int mt = cdiv(P.ic, 128);
  int nsplit = max(1, min(cdiv(P.P, 256), 1024));
  size_t smem = 1024 + 32768 + (size_t)2 * g.Nc * 128 + 64;
  dim3 grid(mt, g.nN, nsplit);
  ProfScope ps("xcov", 4.0 * P.P * P.ic, 2.0 * P.P * (double)P.ic * P.ic, st);
  launch_um_wgrad<2, TFNAS_ACT_RELU>(cdiv(g.Nc, 32), grid, smem, P, 0, g, st);
}

// debug: enable / disable the in-kernel phase trace (buf: device memory, n_ctas * UM_TRACE_SLOTS u64; nullptr disables)
extern "C" int tfnas_debug_um_trace(void* buf, int n_ctas) {
  unsigned long long* p = (unsigned long long*)buf;
  if (cudaMemcpyToSymbol(g_um_trace, &p, sizeof(p)) != cudaSuccess) return TFNAS_E_CUDA;
  if (cudaMemcpyToSymbol(g_um_trace_n, &n_ctas, sizeof(n_ctas)) != cudaSuccess) return TFNAS_E_CUDA;
  return TFNAS_OK;
}
