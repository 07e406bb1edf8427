// Persistent, warp-specialised tcgen05 versions of the four flattened-pixel 1x1-conv GEMMs (expand, project, dc, dx).
//
//   Out[o][p] = sum_k W[o][k] * In[k][p],   p = flattened (n,h,w) pixel, 128 pixels per tile (MMA M = 128)
//
// One CTA per SM walks a static list of work items (candidate slot, N chunk, 128-pixel tile).  The CTA's warps
// have fixed roles that meet only through mbarriers, so no phase of one tile ever waits for another phase:
//   NE warps      EPILOGUE   tcgen05.ld of the finished accumulator (lane quarter = warp & 3, column part = warp >> 2),
//                            BN statistics / SE partial sums / coalesced per-channel stores
//   NP warps      PRODUCERS  in G groups.  Group g owns every G-th K chunk (chunk j of the CTA's chunk stream goes to
//                            group j % G and operand stage j % S): its warps load the raw input rows of their NEXT chunk
//                            into registers right after handing over the current one, so the global-load latency hides
//                            behind the wait for the operand stage; then prologue math (BN / activation / SE gate /
//                            BN-backward on load), tf32 hi/lo split, st.shared into the stage (MN-major, 128B swizzle).
//                            A producer iteration is a long dependent instruction stream (measured ~3k cycles), so G
//                            chunks in flight are what keeps the tensor core fed.
//   1 warp        MMA        waits operand stage + weight slot; one elected lane issues the 12 kind::tf32 MMAs of the K
//                            chunk (hi*hi + lo*hi + hi*lo per K=8 step) and commits to the stage's / slot's "empty" barriers
//   1 warp        WEIGHTS    bulk (TMA) copies of the pre-split, pre-swizzled weight blocks, NB-1 chunks ahead
// The accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of tile j overlaps the main loop of
// tile j+1.  Per-CTA set-up (TMEM allocation, barrier init) is paid once per SM instead of once per tile.
//
// The numerics are those of umma_pw.cu (same operand split, same MMA order within a tile, same epilogue arithmetic).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "kernels.h"
#include "pw.cuh"
#include "umma.cuh"

using namespace umma;

#define WS_KC 32
#define WS_ACC_STRIDE 256                     // TMEM columns between the two accumulator buffers
// Warp split of a kernel: NE epilogue warps (multiple of 4), NP producer warps in G groups, then the MMA warp and the
// weight-copy warp.
template <int NE_, int NP_, int G_>
struct WsDim {
  static constexpr int NE = NE_, NP = NP_, G = G_;
  static constexpr int NPG = NP_ / G_;        // warps per producer group
  static constexpr int RW = WS_KC / NPG;      // K rows per producer warp per chunk
  static constexpr int MMA_WARP = NE_ + NP_, TMA_WARP = NE_ + NP_ + 1;
  static constexpr int NTHR = 32 * (NE_ + NP_ + 2);
  static_assert(NP_ % G_ == 0 && WS_KC % NPG == 0 && RW <= 8 && NE_ % 4 == 0, "bad warp split");
};
#define WS_MAXS 8
#define WS_MAXB 8
#define WS_SMEM_LIMIT 232448                  // 227 KB opt-in maximum

struct WsSched {
  int n_items, tiles_m, na;
  float inv_tiles;
  int first[TFNAS_MAX_OPS + 1];               // first item of each slot (items of a slot: N chunk major, tile minor)
};
struct WsCfg { int S, NB; uint32_t wslot, cf_bytes, acc_bytes; };

struct WsSmem {
  unsigned char *a, *w;
  uint64_t *full, *empty, *wfull, *wempty, *accfull, *accempty;
  uint32_t* tmem_slot;
  float2* cf;
  double* acc;             // per-CTA statistic accumulators (see "statistics" below)
};

// layout: [S operand stages x (hi 16K | lo 16K)] [NB weight slots] [barriers 512 B] [cf tables] [statistic accumulators]
//
// Statistics: the BN sums these kernels produce (BN3 in project, BN2-backward in dc, BN1-backward in dx) end in fp64
// global atomics on a handful of addresses.  Issued per tile they serialise at ~9 ns per same-address atomic (measured:
// project at 56x56 took 1.35 ms with them, 0.43 ms without), so a persistent CTA first accumulates in shared-memory
// doubles and flushes once per (candidate, N chunk) it visits: a few hundred global atomics per address instead of
// tens of thousands.
__device__ __forceinline__ void ws_carve(unsigned char* raw, const WsCfg& c, WsSmem& M) {
  unsigned char* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  M.a = sm;
  M.w = sm + (size_t)c.S * 32768;
  unsigned char* tail = M.w + (size_t)c.NB * c.wslot;
  M.full = (uint64_t*)tail;
  M.empty = M.full + WS_MAXS;
  M.wfull = M.empty + WS_MAXS;
  M.wempty = M.wfull + WS_MAXB;
  M.accfull = M.wempty + WS_MAXB;
  M.accempty = M.accfull + 2;
  M.tmem_slot = (uint32_t*)(M.accempty + 2);
  M.cf = (float2*)(tail + 512);
  M.acc = (double*)(tail + 512 + c.cf_bytes);
}
static size_t ws_smem_bytes(const WsCfg& c) {
  return 1024 + (size_t)c.S * 32768 + (size_t)c.NB * c.wslot + 512 + c.cf_bytes + c.acc_bytes;
}
// deepest pipeline that fits: operand stages S = G * SM (SM <= max_sm), weight slots NB (<= max_nb)
static bool ws_fit(int maxNc, int G, uint32_t cf_bytes, uint32_t acc_bytes, int max_sm, int max_nb, WsCfg& c) {
  static const int nbs[] = {8, 6, 4, 3, 2};
  c.wslot = (uint32_t)2 * maxNc * 128;
  c.cf_bytes = cf_bytes;
  c.acc_bytes = acc_bytes;
  for (int sm = max_sm; sm >= 1; --sm)
    for (int nb : nbs) {
      if (nb > max_nb) continue;
      c.S = G * sm; c.NB = nb;
      if (c.S <= WS_MAXS && ws_smem_bytes(c) <= WS_SMEM_LIMIT) return true;
    }
  c.S = G; c.NB = 1;
  return ws_smem_bytes(c) <= WS_SMEM_LIMIT;
}

__device__ __forceinline__ void ws_decode(const WsSched& Sc, int item, int& slot, int& nc, int& mt) {
  slot = 0;
  int f0 = 0;
#pragma unroll
  for (int s = 1; s < TFNAS_MAX_OPS; ++s)
    if (s < Sc.na && item >= Sc.first[s]) { slot = s; f0 = Sc.first[s]; }
  const int r = item - f0;
  nc = fast_div(r, Sc.tiles_m, Sc.inv_tiles);
  mt = r - nc * Sc.tiles_m;
}

// ---- small helpers -------------------------------------------------------------------------------
template <int NTHREADS>
__device__ __forceinline__ void ws_bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool ws_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// load 4 consecutive pixels of plane `ch` as float4 (vector when aligned, else masked scalars); zero when !rowok
__device__ __forceinline__ float4 ws_ld4(const float* __restrict__ T, const float* __restrict__ Tb, const Px4& px, int C, int ch,
                                         int HW, bool rowok) {
  if (!rowok) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (px.vec) return *(const float4*)(Tb + (size_t)ch * HW);
  float d[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) d[e] = px.v[e] ? T[((size_t)px.n[e] * C + ch) * HW + px.hw[e]] : 0.f;
  return make_float4(d[0], d[1], d[2], d[3]);
}

// split 4 values into tf32 hi / lo and store them as row kk of the operand stage (16 B chunk of pixels 4*lane..4*lane+3)
__device__ __forceinline__ void ws_emit_row(unsigned char* a_hi, int kk, int lane, const float (&v)[4]) {
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
  const uint32_t off = mn_chunk_off(lane * 4, kk, WS_KC * 128);
  *(float4*)(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
  *(float4*)(a_hi + 16384 + off) = make_float4(l[0], l[1], l[2], l[3]);
}

// the 12 MMAs of one K chunk (one thread): operand stage at `a` (hi | lo), weights at `b` (hi block, lo block Nc*128 later)
__device__ __forceinline__ void ws_issue(uint32_t a, uint32_t b, int Nc, uint32_t acc, uint32_t idesc, bool first) {
  // descriptor = constant fields | (shared address >> 4): stepping K only adds to the low word (addresses < 256 KB)
  const uint64_t ca = smem_desc(0, WS_KC * 128, 512, SWIZZLE_128B_BASE32B), cb = smem_desc(0, 16, 1024, SWIZZLE_128B);
  const uint64_t dah0 = ca | (a >> 4), dal0 = ca | ((a + 16384) >> 4);
  const uint64_t dbh0 = cb | (b >> 4), dbl0 = cb | ((b + Nc * 128) >> 4);
#pragma unroll
  for (int q = 0; q < WS_KC / 8; ++q) {
    const uint64_t dah = dah0 + q * 64, dal = dal0 + q * 64;       // + q * 1024 bytes
    const uint64_t dbh = dbh0 + q * 2, dbl = dbl0 + q * 2;         // + q * 32 bytes
    mma_tf32(acc, dah, dbh, idesc, (first && q == 0) ? 0u : 1u);
    mma_tf32(acc, dal, dbh, idesc, 1u);
    mma_tf32(acc, dah, dbl, idesc, 1u);
  }
}

// epilogue geometry of one thread: pixel of its TMEM lane, column range of its warp
struct WsEpi { int p, n, hw; bool v; int c_lo, c_hi; uint32_t taddr; };
template <class D>
__device__ __forceinline__ WsEpi ws_epi(int tile0, int total, int HW, int Nc, int ncol, uint32_t acc, int ew, int lane) {
  WsEpi e;
  e.p = tile0 + (ew & 3) * 32 + lane;
  e.v = e.p < total;
  e.n = e.v ? fast_div(e.p, HW, __frcp_rn((float)HW)) : 0;
  e.hw = e.v ? e.p - e.n * HW : 0;
  const int parts = D::NE / 4, part = ew >> 2;
  const int h = ((Nc + parts - 1) / parts + 15) / 16 * 16;
  e.c_lo = min(Nc, part * h);
  e.c_hi = min(min(Nc, (part + 1) * h), ncol);
  e.taddr = acc + ((uint32_t)((ew & 3) * 32) << 16);
  return e;
}

// ---- optional phase trace (debug build -DUM_TRACE; tfnas_debug_ws_trace) ------------------------------------
// Producer warp 0 / the MMA warp of each traced CTA accumulate clock64() deltas per phase.
//   producer slots: [0] iterations [1] advance + loads [2] - [3] - [4] wait empty [5] emit [6] fence [7] arrive
//   MMA slots:      [8] chunks [9] wait weights [10] wait operands [11] issue + commits [12] wait accumulator
#define WS_TRACE_SLOTS 16
__device__ unsigned long long* g_ws_trace = nullptr;
__device__ int g_ws_trace_n = 0;
#ifdef UM_TRACE
struct WsTrace {
  unsigned long long* out; long long t0; long long acc[8];
  __device__ __forceinline__ void begin(bool on) {
    out = (on && g_ws_trace && (int)blockIdx.x < g_ws_trace_n) ? g_ws_trace + (size_t)blockIdx.x * WS_TRACE_SLOTS : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0;
    if (out) t0 = clock64();
  }
  __device__ __forceinline__ void mark(int slot) { if (out) { const long long t = clock64(); acc[slot] += t - t0; t0 = t; } }
  __device__ __forceinline__ void count() { if (out) acc[0] += 1; }
  __device__ __forceinline__ void end(int base) {
    if (out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (base + i < WS_TRACE_SLOTS) out[base + i] = (unsigned long long)acc[i];
    }
  }
};
#else
struct WsTrace {
  __device__ __forceinline__ void begin(bool) {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void count() {}
  __device__ __forceinline__ void end(int) {}
};
#endif

// -------------------------------------------------------------------------------------------------
// skeleton.  T supplies
//   Args, Dim                                kernel arguments (by value, __grid_constant__), warp split
//   CF                                       bytes of per-column epilogue coefficient tables (0: none)
//   geom(A, Sc, item, nK, Nc, wb)            K chunks, MMA N and first prepped weight block of an item
//   Raw                                      registers holding one chunk of a warp's raw rows (+ its row constants)
//   Prod{A, wi, lane}: bind(Sc, item), nK, load(c, raw), emit(c, raw, stage)
//   epi_prep(A, Sc, item, cf, etid)          fill the per-column coefficient table (epilogue warps, before the barrier)
//   epi_run(A, Sc, item, acc, cf, ew, lane)  consume the accumulator
// -------------------------------------------------------------------------------------------------
template <class T, bool VEC>
__device__ __forceinline__ void ws_run(const typename T::Args& A, const WsSched& Sc, const WsCfg& cfg) {
  using D = typename T::Dim;
  extern __shared__ __align__(1024) unsigned char ws_raw[];
  WsSmem M;
  ws_carve(ws_raw, cfg, M);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < WS_MAXS; ++s) { mbar_init(&M.full[s], D::NPG); mbar_init(&M.empty[s], 1); }
    for (int s = 0; s < WS_MAXB; ++s) { mbar_init(&M.wfull[s], 1); mbar_init(&M.wempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&M.accfull[s], 1); mbar_init(&M.accempty[s], D::NE); }
    fence_barrier_init();
  }
  if (warp == D::MMA_WARP) tmem_alloc(M.tmem_slot, 512);
  for (uint32_t i = tid; i < cfg.acc_bytes / 8; i += D::NTHR) M.acc[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *M.tmem_slot;
  const int item0 = blockIdx.x, istep = gridDim.x;

  if (warp < D::NE) {
    // ------------------------------ epilogue ------------------------------
    int jj = 0, prev_item = -1, prev_key = -1;
    for (int item = item0; item < Sc.n_items; item += istep, ++jj) {
      const int ab = jj & 1;
      const uint32_t au = (uint32_t)jj >> 1;
      float2* cf = M.cf + ab * 256;
      if (T::EPI_ACC) {                     // statistics of the previous (candidate, N chunk) go out when the key changes
        const int key = T::epi_key(A, Sc, item);
        if (prev_item >= 0 && key != prev_key) T::epi_flush(A, Sc, prev_item, M.acc, tid);
        prev_item = item;
        prev_key = key;
      }
      T::epi_prep(A, Sc, item, cf, tid);
      if (T::CF) ws_bar_epi<D::NE * 32>();
      mbar_wait_suspend(&M.accfull[ab], au & 1, 2000);
      tc_fence_after();
      T::epi_run(A, Sc, item, tmem + ab * WS_ACC_STRIDE, cf, M.acc, warp, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&M.accempty[ab]);
    }
    if (T::EPI_ACC && prev_item >= 0) T::epi_flush(A, Sc, prev_item, M.acc, tid);
  } else if (warp < D::NE + D::NP) {
    // ------------------------------ producers ------------------------------
    const int pw = warp - D::NE, gq = pw / D::NPG, wi = pw % D::NPG;
    typename std::conditional<VEC, typename T::ProdV, typename T::Prod>::type f{A, wi, lane, M.acc};
    typename T::Raw raw;
    int item = item0, c = 0;
    bool valid = item < Sc.n_items;
    if (valid) f.bind(Sc, item);
    auto advance = [&]() {               // next chunk of the CTA's chunk stream
      if (valid && ++c == f.nK) {
        c = 0;
        item += istep;
        valid = item < Sc.n_items;
        if (valid) f.bind(Sc, item);
      }
    };
#pragma unroll 1
    for (int k = 0; k < gq; ++k) advance();
    if (valid) f.load(c, raw);
    const int SM = cfg.S / D::G;
    int sm = 0;
    uint32_t eph = 1;                    // parity to wait for on empty[s]: passes on the first use of each stage
    WsTrace tr;
    tr.begin(pw == 0 && lane == 0);
    while (valid) {
      const int s = gq + D::G * sm;
      tr.count();
      mbar_wait_suspend(&M.empty[s], eph, 1000);       // the MMAs that read this operand stage have retired
      tc_fence_after();
      tr.mark(4);
      f.emit(c, raw, M.a + (size_t)s * 32768);
      tr.mark(5);
      fence_proxy_async();
      tr.mark(6);
      __syncwarp();
      if (lane == 0) mbar_arrive(&M.full[s]);
      tr.mark(7);
#pragma unroll
      for (int k = 0; k < D::G; ++k) advance();
      if (valid) f.load(c, raw);         // in flight across the wait for the stage
      tr.mark(1);
      if (++sm == SM) { sm = 0; eph ^= 1; }
    }
    tr.end(0);
  } else if (warp == D::MMA_WARP) {
    // ------------------------------ MMA issuer (warp-uniform control flow, one elected lane issues) ------------
    int s = 0, rot = 0, jj = 0;
    uint32_t fph = 0, wpar = 0;              // parity of full[s]; bit b of wpar = parity of wfull[b]
    const uint32_t a0 = smem_u32(M.a), w0 = smem_u32(M.w);
    WsTrace tr;
    tr.begin(lane == 0);
    for (int item = item0; item < Sc.n_items; item += istep, ++jj) {
      int nK, Nc, key;
      const char* wb;
      T::geom(A, Sc, item, nK, Nc, wb, key);
      const bool resident = nK <= cfg.NB;    // every chunk of the item has its own weight slot (see the weight warp)
      const int ab = jj & 1;
      const uint32_t au = (uint32_t)jj >> 1;
      tr.mark(3);
      mbar_wait_suspend(&M.accempty[ab], (au & 1) ^ 1, 1000);     // the epilogue has drained this accumulator (passes for the first two)
      tc_fence_after();
      tr.mark(4);
      const uint32_t idesc = idesc_tf32(128, Nc, 1, 0);
      const uint32_t acc = tmem + ab * WS_ACC_STRIDE;
      for (int c = 0; c < nK; ++c) {
        tr.count();
        int b = c;
        if (!resident) { b = rot; rot = rot + 1 == cfg.NB ? 0 : rot + 1; }
        mbar_wait_suspend(&M.wfull[b], (wpar >> b) & 1, 1000);
        wpar ^= 1u << b;
        tr.mark(1);
        mbar_wait_suspend(&M.full[s], fph, 1000);
        tc_fence_after();
        tr.mark(2);
        if (ws_elect()) {
          ws_issue(a0 + s * 32768, w0 + b * cfg.wslot, Nc, acc, idesc, c == 0);
          mma_commit(&M.empty[s]);
          mma_commit(&M.wempty[b]);
        }
        __syncwarp();
        tr.mark(3);
        if (++s == cfg.S) { s = 0; fph ^= 1; }
      }
      if (ws_elect()) mma_commit(&M.accfull[ab]);
      __syncwarp();
    }
    tr.end(8);
  } else {
    // ------------------------------ weight copies ------------------------------
    // Items whose K chunks all fit the NB weight slots keep chunk c in slot c; consecutive items of a CTA nearly always
    // share (candidate, N chunk), and then the slots already hold the right blocks: no copy, just the hand-shake.
    // Longer items rotate through the slots as a ring.
    int rot = 0, cur_key = -1;
    uint32_t epar = 0;                         // bit b = number of uses of slot b so far, mod 2
    for (int item = item0; item < Sc.n_items; item += istep) {
      int nK, Nc, key;
      const char* wb;
      T::geom(A, Sc, item, nK, Nc, wb, key);
      const uint32_t bytes = (uint32_t)2 * Nc * 128;
      const bool resident = nK <= cfg.NB;
      const bool hit = resident && key == cur_key;
      for (int c = 0; c < nK; ++c) {
        int b = c;
        if (!resident) { b = rot; rot = rot + 1 == cfg.NB ? 0 : rot + 1; }
        mbar_wait_suspend(&M.wempty[b], ((epar >> b) & 1) ^ 1, 2000);     // passes on the first use of each slot
        epar ^= 1u << b;
        if (ws_elect()) {
          if (hit) {
            mbar_arrive(&M.wfull[b]);
          } else {
            mbar_expect_tx(&M.wfull[b], bytes);
            bulk_g2s(M.w + (size_t)b * cfg.wslot, wb + (size_t)c * bytes, bytes, &M.wfull[b]);
          }
        }
        __syncwarp();
      }
      cur_key = resident ? key : -1;
    }
  }
  tc_fence_before();
  __syncthreads();
  T::final_flush(A, M.acc, tid);           // producer-side accumulators (dx)
  if (warp == D::MMA_WARP) tmem_dealloc(tmem, 512);
}

// VEC mode (HW % 4 == 0, HW >= 128, pixel count a multiple of 128): every tile is full, a thread's 4 pixels lie in one
// image and are 16 B aligned, and a tile spans at most two images.  The producers then run straight-line code: rows
// past the K extent re-read the last valid row (clamped index) and are zeroed by a select.
static bool ws_vec_ok(int HW, int total) { return (HW % 4) == 0 && HW >= 128 && (total % 128) == 0; }
struct WsVecPx { int n, hw, img; };      // image / offset of the thread's first pixel; img = image index within the tile (0 / 1)
__device__ __forceinline__ WsVecPx ws_vec_px(int tile0, int lane, int HW) {
  const float r = __frcp_rn((float)HW);
  WsVecPx v;
  const int p = tile0 + lane * 4;
  v.n = fast_div(p, HW, r);
  v.hw = p - v.n * HW;
  v.img = v.n - fast_div(tile0, HW, r);
  return v;
}

// images covered by a 128-pixel tile: first image and count
__device__ __forceinline__ void ws_tile_images(int tile0, int total, int HW, int& n0, int& nimg) {
  const float r = __frcp_rn((float)HW);
  n0 = fast_div(min(tile0, total - 1), HW, r);
  nimg = fast_div(min(tile0 + 127, total - 1), HW, r) - n0 + 1;
}
__device__ __forceinline__ float4 ws_shfl4(const float4& v, int src) {
  return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src),
                     __shfl_sync(0xffffffffu, v.w, src));
}

// -------------------------------------------------------------------------------------------------
// F1a: expand      UH = BN1(W1 x)
// -------------------------------------------------------------------------------------------------
struct WsExpandArgs { Plan P; UmWAll WA; const float* x; const float* bn1; float* UH; };
template <class Dim_>
struct WsExpandT {
  using Args = WsExpandArgs;
  using Dim = Dim_;
  static constexpr bool EPI_ACC = false;
  static __device__ __forceinline__ int epi_key(const Args&, const WsSched&, int) { return 0; }
  static __device__ __forceinline__ void epi_flush(const Args&, const WsSched&, int, double*, int) {}
  static __device__ __forceinline__ void final_flush(const Args&, double*, int) {}
  static constexpr uint32_t CF = 2 * 256 * sizeof(float2);
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb, int& key) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const UmW& W = A.WA.s[slot];
    nK = W.nK; Nc = W.Nc;
    wb = (const char*)W.wp + (size_t)nc * W.nK * 2 * W.Nc * 128;
    key = slot * 64 + nc;
  }
  struct Raw { float4 a[Dim::RW]; };
  struct ProdV {
    const Args& A; int wi, lane; double* cta_acc;
    int nK; const float* xb;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int slot, nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      nK = A.WA.s[slot].nK;
      const WsVecPx v = ws_vec_px(mt * 128, lane, A.P.HW);
      xb = A.x + (size_t)v.n * A.P.ic * A.P.HW + v.hw;
    }
    __device__ __forceinline__ void load(int c, Raw& r) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int k = min(c * WS_KC + wi + i * Dim::NPG, A.P.ic - 1);
        r.a[i] = *(const float4*)(xb + (uint32_t)(k * A.P.HW));
      }
    }
    __device__ __forceinline__ void emit(int c, const Raw& r, unsigned char* a_hi) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const bool ok = c * WS_KC + wi + i * Dim::NPG < A.P.ic;
        const float v[4] = {ok ? r.a[i].x : 0.f, ok ? r.a[i].y : 0.f, ok ? r.a[i].z : 0.f, ok ? r.a[i].w : 0.f};
        ws_emit_row(a_hi, wi + i * Dim::NPG, lane, v);
      }
    }
  };
  struct Prod {
    const Args& A; int wi, lane; double* cta_acc;
    int nK; Px4 px; const float* xb;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int slot, nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      nK = A.WA.s[slot].nK;
      px_decomp(px, mt * 128 + lane * 4, A.P.P, A.P.HW);
      xb = A.x + (size_t)px.n[0] * A.P.ic * A.P.HW + px.hw[0];
    }
    __device__ __forceinline__ void load(int c, Raw& r) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int k = c * WS_KC + wi + i * Dim::NPG;
        r.a[i] = ws_ld4(A.x, xb, px, A.P.ic, k, A.P.HW, k < A.P.ic);
      }
    }
    __device__ __forceinline__ void emit(int, const Raw& r, unsigned char* a_hi) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const float v[4] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w};
        ws_emit_row(a_hi, wi + i * Dim::NPG, lane, v);
      }
    }
  };
  static __device__ __forceinline__ void epi_prep(const Args& A, const WsSched& Sc, int item, float2* cf, int etid) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    for (int i = etid; i < ncol; i += Dim::NE * 32) cf[i] = make_float2(A.bn1[cst0 + i], A.bn1[A.P.MC + cst0 + i]);
  }
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2* cft,
                                                 double* sacc, int ew, int lane) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    const WsEpi e = ws_epi<Dim>(mt * 128, A.P.P, A.P.HW, Nc, ncol, acc, ew, lane);
    const size_t HW = (size_t)A.P.HW;
    float* ob = A.UH + ((size_t)e.n * A.P.MC + cst0) * HW + e.hw;
    for (int c0 = e.c_lo; c0 < e.c_hi; c0 += 16) {
      float v[16];
      tmem_ld16(e.taddr + c0, v);
      if (!e.v) continue;
      float* q = ob + (size_t)c0 * HW;
      const float2* cf = cft + c0;
      if (c0 + 16 <= e.c_hi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float2 m = cf[j]; *q = (v[j] - m.x) * m.y; q += HW; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < e.c_hi) { const float2 m = cf[j]; *q = (v[j] - m.x) * m.y; q += HW; }
      }
    }
  }
};

// -------------------------------------------------------------------------------------------------
// F3: project      Z = W3 (act(BN2(D)) * gate), BN3 sums
// -------------------------------------------------------------------------------------------------
struct WsProjectArgs { Plan P; UmWAll WA; const float* D; const float* bn2; const float* seg; float* Zb; double* st3; int nostats; };
template <int ACT, class Dim_>
struct WsProjectT {
  using Args = WsProjectArgs;
  using Dim = Dim_;
  static constexpr bool EPI_ACC = true;       // BN3 sums: shared-memory doubles [2 * Nc], flushed per (candidate, N chunk)
  static constexpr uint32_t ACC_BYTES = 2 * 256 * sizeof(double);
  static __device__ __forceinline__ int epi_key(const Args& A, const WsSched& Sc, int item) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    return slot * 64 + nc;
  }
  static __device__ __forceinline__ void epi_flush(const Args& A, const WsSched& Sc, int item, double* sacc, int tid) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const int Nc = A.WA.s[slot].Nc, oc = A.P.oc;
    const int ncol = min(Nc, oc - nc * Nc);
    double* dst = A.st3 + 2 * (slot * oc + nc * Nc);
    ws_bar_epi<Dim::NE * 32>();                 // every epilogue warp has added its sums of the finished tiles
    for (int i = tid; i < 2 * ncol; i += Dim::NE * 32) {
      const double v = sacc[i];
      sacc[i] = 0.0;
      atomicAdd(&dst[i], v);
    }
    ws_bar_epi<Dim::NE * 32>();                 // zeroed before the next key's sums arrive
  }
  static __device__ __forceinline__ void final_flush(const Args&, double*, int) {}
  static constexpr uint32_t CF = 0;
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb, int& key) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const UmW& W = A.WA.s[slot];
    nK = W.nK; Nc = W.Nc;
    wb = (const char*)W.wp + (size_t)nc * W.nK * 2 * W.Nc * 128;
    key = slot * 64 + nc;
  }
  // cv: the row constants of the chunk, one per lane, handed out by shuffles: lanes 0..7 BN2 mean of rows 0..7,
  // lanes 8..15 rstd, lanes 16 + 2 i + m the SE gate of row i for the m-th image of the tile (tiles spanning more than
  // two images read the gate from global memory)
  struct Raw { float4 a[Dim::RW]; float cv; };
  struct ProdV {
    const Args& A; int wi, lane; double* cta_acc;
    int nK, mc, coff, soff, se, n0, img; const float* Db;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int slot, nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      const Cand& cd = A.P.c[slot];
      mc = cd.mc; coff = cd.coff; soff = cd.soff; se = cd.se;
      nK = A.WA.s[slot].nK;
      const WsVecPx v = ws_vec_px(mt * 128, lane, A.P.HWo);
      img = v.img;
      n0 = v.n - v.img;
      Db = A.D + ((size_t)v.n * A.P.MC + coff) * A.P.HWo + v.hw;
    }
    __device__ __forceinline__ void load(int c, Raw& r) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int k = min(c * WS_KC + wi + i * Dim::NPG, mc - 1);
        r.a[i] = *(const float4*)(Db + (uint32_t)(k * A.P.HWo));
      }
      // one constant per lane (clamped, always-valid addresses; unused lanes are never read back)
      const int i = lane < 16 ? (lane & 7) : ((lane - 16) >> 1);
      const int k = min(c * WS_KC + wi + min(i, Dim::RW - 1) * Dim::NPG, mc - 1);
      const float* src = lane < 16 ? A.bn2 + (lane < 8 ? 0 : A.P.MC) + coff + k
                                   : A.seg + (size_t)min(n0 + (lane & 1), A.P.N - 1) * A.P.MCse + soff + k;
      r.cv = (lane < 16 || se > 0) ? *src : 1.f;
    }
    __device__ __forceinline__ void emit(int c, const Raw& r, unsigned char* a_hi) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int kk = wi + i * Dim::NPG;
        const bool ok = c * WS_KC + kk < mc;
        const float mu = __shfl_sync(0xffffffffu, r.cv, i), rs = __shfl_sync(0xffffffffu, r.cv, 8 + i);
        const float gt = __shfl_sync(0xffffffffu, r.cv, 16 + 2 * i + img);
        const float d[4] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = ok ? act_f<ACT>((d[e] - mu) * rs) * gt : 0.f;
        ws_emit_row(a_hi, kk, lane, v);
      }
    }
  };
  struct Prod {
    const Args& A; int wi, lane; double* cta_acc;
    int nK, mc, coff, soff, se, n0, nimg; Px4 px; const float* Db;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int slot, nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      const Cand& cd = A.P.c[slot];
      mc = cd.mc; coff = cd.coff; soff = cd.soff; se = cd.se;
      nK = A.WA.s[slot].nK;
      px_decomp(px, mt * 128 + lane * 4, A.P.Q, A.P.HWo);
      Db = A.D + (size_t)px.n[0] * A.P.MC * A.P.HWo + px.hw[0];
      ws_tile_images(mt * 128, A.P.Q, A.P.HWo, n0, nimg);
    }
    __device__ __forceinline__ void load(int c, Raw& r) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int k = c * WS_KC + wi + i * Dim::NPG;
        r.a[i] = ws_ld4(A.D, Db, px, A.P.MC, coff + k, A.P.HWo, k < mc);
      }
      if (lane < 16) {
        const int i = lane & 7, k = c * WS_KC + wi + i * Dim::NPG;
        r.cv = (i < Dim::RW && k < mc) ? A.bn2[(lane < 8 ? 0 : A.P.MC) + coff + k] : 0.f;
      } else {
        const int i = (lane - 16) >> 1, m = lane & 1, k = c * WS_KC + wi + i * Dim::NPG;
        r.cv = (se > 0 && i < Dim::RW && k < mc && m < nimg) ? A.seg[(size_t)(n0 + m) * A.P.MCse + soff + k] : 1.f;
      }
    }
    __device__ __forceinline__ void emit(int c, const Raw& r, unsigned char* a_hi) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int kk = wi + i * Dim::NPG, k = c * WS_KC + kk;
        const float mu = __shfl_sync(0xffffffffu, r.cv, i), rs = __shfl_sync(0xffffffffu, r.cv, 8 + i);
        const float g0 = __shfl_sync(0xffffffffu, r.cv, 16 + 2 * i), g1 = __shfl_sync(0xffffffffu, r.cv, 17 + 2 * i);
        const float d[4] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float b = act_f<ACT>((d[e] - mu) * rs);
          if (se > 0) b *= nimg <= 2 ? (px.n[e] == n0 ? g0 : g1)
                                     : (px.v[e] && k < mc ? A.seg[(size_t)px.n[e] * A.P.MCse + soff + k] : 0.f);
          v[e] = (px.v[e] && k < mc) ? b : 0.f;
        }
        ws_emit_row(a_hi, kk, lane, v);
      }
    }
  };
  static __device__ __forceinline__ void epi_prep(const Args&, const WsSched&, int, float2*, int) {}
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2*,
                                                 double* sacc, int ew, int lane) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const int Nc = A.WA.s[slot].Nc, oc = A.P.oc;
    const int ncol = min(Nc, oc - nc * Nc);
    const WsEpi e = ws_epi<Dim>(mt * 128, A.P.Q, A.P.HWo, Nc, ncol, acc, ew, lane);
    const size_t HWo = (size_t)A.P.HWo;
    const int o0 = slot * oc + nc * Nc;                  // Z / BN3 channel of column 0
    float* zb = A.Zb + ((size_t)e.n * A.P.na * oc + o0) * HWo + e.hw;
    // Rows of invalid pixels and weight rows past oc are zero in the operands, so their accumulators are exactly 0:
    // the BN3 sums need no masking, only the stores do.
    for (int c0 = e.c_lo; c0 < e.c_hi; c0 += 16) {
      float v[16], q[16];
      tmem_ld16(e.taddr + c0, v);
      if (e.v) {
        float* zp = zb + (size_t)c0 * HWo;
        if (c0 + 16 <= e.c_hi) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { *zp = v[j]; zp += HWo; }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (c0 + j < e.c_hi) *zp = v[j];
            zp += HWo;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) q[j] = v[j] * v[j];
      const float s1 = warp_sum16(v), s2 = warp_sum16(q);
      const int col = c0 + (lane & 15);
      if (lane < 16 && col < e.c_hi && !A.nostats) {
        atomicAdd(&sacc[2 * col], (double)s1);
        atomicAdd(&sacc[2 * col + 1], (double)s2);
      }
    }
  }
};

// -------------------------------------------------------------------------------------------------
// B2: dc = W3^T dz   (dz = A*g + B*z + C on load), epilogue dd-hat / SE partial sums
// -------------------------------------------------------------------------------------------------
struct WsDcArgs {
  Plan P; UmWAll WA; const float* G; const float* Zb; const float4* dzc2; const float* D; const float* bn2;
  float* DC; float* dg; double* sD;
};
template <int ACT, class Dim_>
struct WsDcT {
  using Args = WsDcArgs;
  using Dim = Dim_;
  static constexpr bool EPI_ACC = true;       // BN2-backward sums: shared-memory doubles [2 * Nc], flushed per (candidate, N chunk)
  static constexpr uint32_t ACC_BYTES = 2 * 256 * sizeof(double);
  static __device__ __forceinline__ int epi_key(const Args& A, const WsSched& Sc, int item) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    return slot * 64 + nc;
  }
  static __device__ __forceinline__ void epi_flush(const Args& A, const WsSched& Sc, int item, double* sacc, int tid) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc);
    double* dst = A.sD + 2 * (cd.coff + nc * Nc);
    ws_bar_epi<Dim::NE * 32>();
    if (cd.se == 0) {                            // gated candidates send their sums to dg instead
      for (int i = tid; i < 2 * ncol; i += Dim::NE * 32) {
        const double v = sacc[i];
        sacc[i] = 0.0;
        atomicAdd(&dst[i], v);
      }
    }
    ws_bar_epi<Dim::NE * 32>();
  }
  static __device__ __forceinline__ void final_flush(const Args&, double*, int) {}
  static constexpr uint32_t CF = 2 * 256 * sizeof(float2);
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb, int& key) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const UmW& W = A.WA.s[slot];
    nK = W.nK; Nc = W.Nc;
    wb = (const char*)W.wp + (size_t)nc * W.nK * 2 * W.Nc * 128;
    key = slot * 64 + nc;
  }
  // cf: lane i < RW holds (A, B, C, -) of row i
  struct Raw { float4 a[Dim::RW], b[Dim::RW]; float4 cf; };
  struct Prod;
  typedef Prod ProdV;
  struct Prod {
    const Args& A; int wi, lane; double* cta_acc;
    int nK, slot; Px4 px; const float* Gb; const float* Zbb;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      nK = A.WA.s[slot].nK;
      px_decomp(px, mt * 128 + lane * 4, A.P.Q, A.P.HWo);
      Gb = A.G + (size_t)px.n[0] * A.P.oc * A.P.HWo + px.hw[0];
      Zbb = A.Zb + (size_t)px.n[0] * A.P.na * A.P.oc * A.P.HWo + px.hw[0];
    }
    __device__ __forceinline__ void load(int c, Raw& r) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int o = c * WS_KC + wi + i * Dim::NPG;
        r.a[i] = ws_ld4(A.G, Gb, px, A.P.oc, o, A.P.HWo, o < A.P.oc);
        r.b[i] = ws_ld4(A.Zb, Zbb, px, A.P.na * A.P.oc, slot * A.P.oc + o, A.P.HWo, o < A.P.oc);
      }
      const int o = c * WS_KC + wi + (lane & 7) * Dim::NPG;
      r.cf = ((lane & 7) < Dim::RW && lane < 8 && o < A.P.oc) ? A.dzc2[slot * A.P.oc + o] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void emit(int, const Raw& r, unsigned char* a_hi) const {
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const float4 cf = ws_shfl4(r.cf, i);              // zero past oc
        const float gg[4] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w}, z[4] = {r.b[i].x, r.b[i].y, r.b[i].z, r.b[i].w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? fmaf(cf.x, gg[e], fmaf(cf.y, z[e], cf.z)) : 0.f;
        ws_emit_row(a_hi, wi + i * Dim::NPG, lane, v);
      }
    }
  };
  static __device__ __forceinline__ void epi_prep(const Args& A, const WsSched& Sc, int item, float2* cf, int etid) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    for (int i = etid; i < ncol; i += Dim::NE * 32) cf[i] = make_float2(A.bn2[cst0 + i], A.bn2[A.P.MC + cst0 + i]);
  }
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2* cft,
                                                 double* sacc, int ew, int lane) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const Plan& P = A.P;
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    const WsEpi e = ws_epi<Dim>(mt * 128, P.Q, P.HWo, Nc, ncol, acc, ew, lane);
    const bool gated = cd.se > 0;
    // images covered by this warp's 32 consecutive pixels: at most two when HWo >= 32
    const int n_first = __shfl_sync(0xffffffffu, e.n, 0);
    int n_last = e.v ? e.n : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_last = max(n_last, __shfl_xor_sync(0xffffffffu, n_last, o));
    const bool two_img = __all_sync(0xffffffffu, (!e.v) || e.n == n_first || e.n == n_last);
    const size_t HWo = (size_t)P.HWo;
    const size_t eoff = ((size_t)e.n * P.MC + cst0) * HWo + e.hw;
    const int soff0 = cd.soff + nc * Nc;                // SE-gated stacked channel of column 0
    // Rows of invalid pixels and weight rows past mc are zero in the operands, so their accumulators are exactly 0
    // and (with d loaded as 0) contribute 0 to every sum below: only loads and stores are masked.
    for (int c0 = e.c_lo; c0 < e.c_hi; c0 += 16) {
      float v[16], d[16];
      const int nv = min(16, e.c_hi - c0);                // valid columns of this group (uniform)
      {                                                   // D loads issued before the TMEM read so they overlap
        const float* dp = A.D + eoff + (size_t)c0 * HWo;
        if (nv == 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { d[j] = e.v ? *dp : 0.f; dp += HWo; }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { d[j] = (e.v && j < nv) ? *dp : 0.f; dp += HWo; }
        }
      }
      tmem_ld16(e.taddr + c0, v);
      const float2* cf = cft + c0;
      const int col = c0 + (lane & 15);
      if (e.v) {
        float* qp = A.DC + eoff + (size_t)c0 * HWo;
        if (gated) {
          // DC = dc (raw); v <- dc * act(BN2(d)) = this pixel's contribution to dL/dgate
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nv) {
              const float2 m = cf[j];
              *qp = v[j];
              v[j] *= act_f<ACT>((d[j] - m.x) * m.y);
            }
            qp += HWo;
          }
        } else {
          // DC = dd-hat = dc * act'(d-hat); v <- dd-hat, d <- dd-hat * d-hat (BN2-backward sums)
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
      if (gated) {
        if (!two_img) {                    // tiny planes (HWo < 32): more than two images per warp
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (e.v && j < nv) atomicAdd(&A.dg[(size_t)e.n * P.MCse + soff0 + c0 + j], v[j]);
        } else if (n_last <= n_first) {    // the usual case: all 32 pixels belong to one image
          const float sa = warp_sum16(v);
          if (lane < 16 && col < e.c_hi) atomicAdd(&A.dg[(size_t)n_first * P.MCse + soff0 + col], sa);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) d[j] = (e.n == n_first) ? 0.f : v[j];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (e.n == n_first) ? v[j] : 0.f;
          const float sa = warp_sum16(v), sb = warp_sum16(d);
          if (lane < 16 && col < e.c_hi) {
            atomicAdd(&A.dg[(size_t)n_first * P.MCse + soff0 + col], sa);
            atomicAdd(&A.dg[(size_t)n_last * P.MCse + soff0 + col], sb);
          }
        }
      } else {
        const float s1 = warp_sum16(v), s2 = warp_sum16(d);
        if (lane < 16 && col < e.c_hi) {
          atomicAdd(&sacc[2 * col], (double)s1);
          atomicAdd(&sacc[2 * col + 1], (double)s2);
        }
      }
    }
  }
};

// -------------------------------------------------------------------------------------------------
// B3b: dx_main = sum_i W1_i^T (r1 * du-hat), K = stacked mid channels (per-candidate chunks of 32), split over items
// -------------------------------------------------------------------------------------------------
struct WsDxArgs { Plan P; UmW W; DxChunks CH; int ksplit; const float* DA; const float* UH; float* dx; double* sU; int acc_rows; };
template <int ACT, class Dim_>
struct WsDxT {
  using Args = WsDxArgs;
  using Dim = Dim_;
  static constexpr bool EPI_ACC = false;
  static __device__ __forceinline__ int epi_key(const Args&, const WsSched&, int) { return 0; }
  static __device__ __forceinline__ void epi_flush(const Args&, const WsSched&, int, double*, int) {}
  // BN1-backward sums are produced by the PRODUCERS (one pair per mid channel and tile).  Without a K split every item
  // covers all ΣMC rows, so the CTA accumulates them in shared-memory doubles [2 * MC] and flushes once at the end
  // (A.acc_rows = MC when that fits, 0 = straight global atomics: the split shapes have few tiles per channel anyway).
  static __device__ __forceinline__ void final_flush(const Args& A, double* sacc, int tid) {
    for (int i = tid; i < 2 * A.acc_rows; i += Dim::NTHR) {
      const double v = sacc[i];
      if (v != 0.0) atomicAdd(&A.sU[i], v);
    }
  }
  static constexpr uint32_t CF = 0;
  // item = tile + tiles_m * ksplit part
  static __device__ __forceinline__ void part(const Args& A, const WsSched& Sc, int item, int& mt, int& ch0, int& ch1) {
    const int ks = fast_div(item, Sc.tiles_m, Sc.inv_tiles);
    mt = item - ks * Sc.tiles_m;
    ch0 = (int)((long long)A.CH.total * ks / A.ksplit);
    ch1 = (int)((long long)A.CH.total * (ks + 1) / A.ksplit);
  }
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb, int& key) {
    int mt, ch0, ch1;
    part(A, Sc, item, mt, ch0, ch1);
    nK = ch1 - ch0; Nc = A.W.Nc;
    wb = (const char*)A.W.wp + (size_t)ch0 * 2 * A.W.Nc * 128;
    key = ch0;
  }
  struct Raw { float4 a[Dim::RW], b[Dim::RW]; };
  struct ProdV {
    const Args& A; int wi, lane; double* cta_acc;
    int nK, ch0; const float* DAb; ptrdiff_t uh_minus_da;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int mt, ch1;
      part(A, Sc, item, mt, ch0, ch1);
      nK = ch1 - ch0;
      const WsVecPx v = ws_vec_px(mt * 128, lane, A.P.HW);
      DAb = A.DA + (size_t)v.n * A.P.MC * A.P.HW + v.hw;
      uh_minus_da = A.UH - A.DA;
    }
    __device__ __forceinline__ void locate(int c, int& mc, int& coff, int& k0) const {
      const int g = ch0 + c;
      int f0 = 0, slot = 0;
#pragma unroll
      for (int s = 1; s < TFNAS_MAX_OPS; ++s)
        if (s < A.P.na && g >= A.CH.first[s]) { slot = s; f0 = A.CH.first[s]; }
      k0 = (g - f0) * WS_KC;
      mc = A.P.c[slot].mc;
      coff = A.P.c[slot].coff;
    }
    __device__ __forceinline__ void load(int c, Raw& r) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int k = coff + min(k0 + wi + i * Dim::NPG, mc - 1);
        const float* p = DAb + (uint32_t)(k * A.P.HW);
        r.a[i] = *(const float4*)p;
        r.b[i] = *(const float4*)(p + uh_minus_da);
      }
    }
    __device__ __forceinline__ void emit(int c, const Raw& r, unsigned char* a_hi) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
      float sacc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) sacc[i] = 0.f;
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const bool ok = k0 + wi + i * Dim::NPG < mc;
        const float da[4] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w}, uh[4] = {r.b[i].x, r.b[i].y, r.b[i].z, r.b[i].w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = ok ? da[e] * act_df<ACT>(uh[e]) : 0.f;
          sacc[2 * i] += v[e];
          sacc[2 * i + 1] += v[e] * uh[e];
        }
        ws_emit_row(a_hi, wi + i * Dim::NPG, lane, v);
      }
      const float tot = warp_sum16(sacc);
      const int k = k0 + wi + (lane >> 1) * Dim::NPG;
      if (lane < 2 * Dim::RW && k < mc) {
        if (A.acc_rows) atomicAdd(&cta_acc[2 * (coff + k) + (lane & 1)], (double)tot);
        else atomicAdd(&A.sU[2 * (coff + k) + (lane & 1)], (double)tot);
      }
    }
  };
  struct Prod {
    const Args& A; int wi, lane; double* cta_acc;
    int nK, ch0; Px4 px; const float* DAb; const float* UHb;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int mt, ch1;
      part(A, Sc, item, mt, ch0, ch1);
      nK = ch1 - ch0;
      px_decomp(px, mt * 128 + lane * 4, A.P.P, A.P.HW);
      const size_t o = (size_t)px.n[0] * A.P.MC * A.P.HW + px.hw[0];
      DAb = A.DA + o;
      UHb = A.UH + o;
    }
    __device__ __forceinline__ void locate(int c, int& mc, int& coff, int& k0) const {
      const int g = ch0 + c;
      int f0 = 0, slot = 0;
#pragma unroll
      for (int s = 1; s < TFNAS_MAX_OPS; ++s)
        if (s < A.P.na && g >= A.CH.first[s]) { slot = s; f0 = A.CH.first[s]; }
      k0 = (g - f0) * WS_KC;
      mc = A.P.c[slot].mc;
      coff = A.P.c[slot].coff;
    }
    __device__ __forceinline__ void load(int c, Raw& r) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const int k = k0 + wi + i * Dim::NPG;
        r.a[i] = ws_ld4(A.DA, DAb, px, A.P.MC, coff + k, A.P.HW, k < mc);
        r.b[i] = ws_ld4(A.UH, UHb, px, A.P.MC, coff + k, A.P.HW, k < mc);
      }
    }
    // Rows past the candidate's width and invalid pixels are loaded as zeros, so du = 0 * act'(0) = 0 there.
    // BN1's rstd is folded into the prepped weights (umma_prep_bwd), the operand is plain du-hat.
    __device__ __forceinline__ void emit(int c, const Raw& r, unsigned char* a_hi) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
      float sacc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) sacc[i] = 0.f;
#pragma unroll
      for (int i = 0; i < Dim::RW; ++i) {
        const float da[4] = {r.a[i].x, r.a[i].y, r.a[i].z, r.a[i].w}, uh[4] = {r.b[i].x, r.b[i].y, r.b[i].z, r.b[i].w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = da[e] * act_df<ACT>(uh[e]);
          sacc[2 * i] += v[e];
          sacc[2 * i + 1] += v[e] * uh[e];
        }
        ws_emit_row(a_hi, wi + i * Dim::NPG, lane, v);
      }
      // 2*RW statistics (rows x {sum du, sum du*uh}) reduced together; lane l < 2*RW ends up owning statistic l
      static_assert(2 * Dim::RW <= 16, "statistics must fit the 16-value warp reduction");
      const float tot = warp_sum16(sacc);
      const int k = k0 + wi + (lane >> 1) * Dim::NPG;
      if (lane < 2 * Dim::RW && k < mc) {
        if (A.acc_rows) atomicAdd(&cta_acc[2 * (coff + k) + (lane & 1)], (double)tot);
        else atomicAdd(&A.sU[2 * (coff + k) + (lane & 1)], (double)tot);
      }
    }
  };
  static __device__ __forceinline__ void epi_prep(const Args&, const WsSched&, int, float2*, int) {}
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2*,
                                                 double* sacc, int ew, int lane) {
    int mt, ch0, ch1;
    part(A, Sc, item, mt, ch0, ch1);
    const WsEpi e = ws_epi<Dim>(mt * 128, A.P.P, A.P.HW, A.W.Nc, A.P.ic, acc, ew, lane);
    const size_t HW = (size_t)A.P.HW;
    float* ob = A.dx + (size_t)e.n * A.P.ic * HW + e.hw;
    for (int c0 = e.c_lo; c0 < e.c_hi; c0 += 16) {
      float v[16];
      tmem_ld16(e.taddr + c0, v);
      if (!e.v) continue;
      float* q = ob + (size_t)c0 * HW;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (c0 + j < e.c_hi) {
          if (A.ksplit == 1) *q = v[j];
          else atomicAdd(q, v[j]);
        }
        q += HW;
      }
    }
  }
};

// -------------------------------------------------------------------------------------------------
// kernels + host side
// -------------------------------------------------------------------------------------------------
typedef WsDim<16, 8, 2> DimExpand;            // epilogue-bound (stores N = mc columns), K = ic is 1..6 chunks
typedef WsDim<4, 16, 4> DimProject;           // prologue-bound: 4 chunks in flight, 8 rows per warp
typedef WsDim<16, 8, 1> DimDc;                // epilogue-bound (loads D, stores DC); two tensors per chunk
typedef WsDim<4, 16, 2> DimDx;                // prologue-bound, two tensors per chunk: 2 chunks in flight, 4 rows per warp

template <bool VEC>
__global__ void __launch_bounds__(DimExpand::NTHR, 1) k_ws_expand(const __grid_constant__ WsExpandArgs A, const __grid_constant__ WsSched Sc,
                                                                const __grid_constant__ WsCfg cfg) {
  ws_run<WsExpandT<DimExpand>, VEC>(A, Sc, cfg);
}
template <int ACT, bool VEC>
__global__ void __launch_bounds__(DimProject::NTHR, 1) k_ws_project(const __grid_constant__ WsProjectArgs A, const __grid_constant__ WsSched Sc,
                                                                  const __grid_constant__ WsCfg cfg) {
  ws_run<WsProjectT<ACT, DimProject>, VEC>(A, Sc, cfg);
}
template <int ACT, bool VEC>
__global__ void __launch_bounds__(DimDc::NTHR, 1) k_ws_dc(const __grid_constant__ WsDcArgs A, const __grid_constant__ WsSched Sc,
                                                        const __grid_constant__ WsCfg cfg) {
  ws_run<WsDcT<ACT, DimDc>, VEC>(A, Sc, cfg);
}
template <int ACT, bool VEC>
__global__ void __launch_bounds__(DimDx::NTHR, 1) k_ws_dx(const __grid_constant__ WsDxArgs A, const __grid_constant__ WsSched Sc,
                                                        const __grid_constant__ WsCfg cfg) {
  ws_run<WsDxT<ACT, DimDx>, VEC>(A, Sc, cfg);
}

#define WS_LAUNCH_V(KERN, ACT_, NTHR_) do { \
    if (vec) { ensure_smem(KERN<ACT_, true>, smem); KERN<ACT_, true><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } \
    else { ensure_smem(KERN<ACT_, false>, smem); KERN<ACT_, false><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } } while (0)
#define WS_LAUNCH(KERN, NTHR_) do { \
    if (P.act == TFNAS_ACT_RELU) WS_LAUNCH_V(KERN, TFNAS_ACT_RELU, NTHR_); \
    else WS_LAUNCH_V(KERN, TFNAS_ACT_SWISH, NTHR_); } while (0)
// + the identity activation (head: feature-mix conv on the project kernel)
#define WS_LAUNCH3(KERN, NTHR_) do { \
    if (P.act == TFNAS_ACT_NONE) WS_LAUNCH_V(KERN, TFNAS_ACT_NONE, NTHR_); \
    else WS_LAUNCH(KERN, NTHR_); } while (0)

// TFNAS_WS: comma-separated subset of {expand,project,dc,dx} run by the persistent kernels ("none" disables, "all" = all
// four).  Default: expand, project, dx -- dc is bound by its epilogue (D loads, DC stores, statistics), which the
// many-CTA kernel of umma_pw.cu overlaps better.
static int g_ws_mask = -1;
int ws_enabled(int which) {
  if (g_ws_mask < 0) {
    const char* e = getenv("TFNAS_WS");
    if (!e) g_ws_mask = 1 | 2 | 8;
    else {
      g_ws_mask = 0;
      if (strstr(e, "expand")) g_ws_mask |= 1;
      if (strstr(e, "project")) g_ws_mask |= 2;
      if (strstr(e, "dc")) g_ws_mask |= 4;
      if (strstr(e, "dx")) g_ws_mask |= 8;
      if (strstr(e, "all")) g_ws_mask = 15;
    }
  }
  return (g_ws_mask >> which) & 1;
}

// items of the per-slot GEMMs: slot major, then N chunk, then pixel tile
static void ws_sched_slots(const Plan& P, const UmWAll& WA, int tiles, WsSched& Sc, int& maxNc) {
  Sc.tiles_m = tiles;
  Sc.inv_tiles = 1.f / (float)tiles;
  Sc.na = P.na;
  Sc.first[0] = 0;
  maxNc = 0;
  for (int s = 0; s < TFNAS_MAX_OPS; ++s) {
    const int nN = s < P.na ? WA.s[s].nN : 0;
    Sc.first[s + 1] = Sc.first[s] + nN * tiles;
    if (s < P.na) maxNc = max(maxNc, WA.s[s].Nc);
  }
  Sc.n_items = Sc.first[TFNAS_MAX_OPS];
}

bool ws_expand(const Plan& P, const UmWAll& WA, const float* x, const float* bn1, float* UH, cudaStream_t st) {
  WsSched Sc;
  int maxNc;
  const int tiles = cdiv(P.P, 128);
  if ((long long)tiles * TFNAS_MAX_OPS * 64 >= (1 << 23)) return false;
  ws_sched_slots(P, WA, tiles, Sc, maxNc);
  WsCfg cfg;
  if (maxNc > 256 || !ws_fit(maxNc, DimExpand::G, WsExpandT<DimExpand>::CF, 0, 2, 6, cfg)) return false;
  const size_t smem = ws_smem_bytes(cfg);
  WsExpandArgs A{P, WA, x, bn1, UH};
  ProfScope ps("expand", 4.0 * P.P * P.ic + 4.0 * P.P * P.MC + 4.0 * P.MC * P.ic, 2.0 * P.P * (double)P.MC * P.ic, st);
  const int grid = min(Sc.n_items, sm_count());
  if (ws_vec_ok(P.HW, P.P)) { ensure_smem(k_ws_expand<true>, smem); k_ws_expand<true><<<grid, DimExpand::NTHR, smem, st>>>(A, Sc, cfg); }
  else { ensure_smem(k_ws_expand<false>, smem); k_ws_expand<false><<<grid, DimExpand::NTHR, smem, st>>>(A, Sc, cfg); }
  return true;
}

bool ws_project(const Plan& P, const UmWAll& WA, const float* D, const float* bn2, const float* seg, float* Zb, double* st3,
                cudaStream_t st) {
  WsSched Sc;
  int maxNc;
  const int tiles = cdiv(P.Q, 128);
  if ((long long)tiles * TFNAS_MAX_OPS * 64 >= (1 << 23)) return false;
  ws_sched_slots(P, WA, tiles, Sc, maxNc);
  WsCfg cfg;
  if (maxNc > 256 || !ws_fit(maxNc, DimProject::G, 0, 2 * 256 * sizeof(double), 1, 4, cfg)) return false;
  const size_t smem = ws_smem_bytes(cfg);
#ifdef UM_TRACE      // debug (trace) build only: the product library has no switch that changes results
  static const int nostats = getenv("TFNAS_DEBUG_NOSTATS") ? 1 : 0;     // timing experiment only: results are wrong
#else
  const int nostats = 0;
#endif
  WsProjectArgs A{P, WA, D, bn2, seg, Zb, st3, nostats};
  ProfScope ps("project", 4.0 * P.Q * ((double)P.MC + (double)P.na * P.oc) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  const int grid = min(Sc.n_items, sm_count());
  const bool vec = ws_vec_ok(P.HWo, P.Q);
  WS_LAUNCH3(k_ws_project, DimProject::NTHR);
  return true;
}

bool ws_dc(const Plan& P, const UmWAll& WA, const float* G, const float* Zb, const float4* dzc2, const float* D,
           const float* bn2, float* DC, float* dg, double* sD, cudaStream_t st) {
  WsSched Sc;
  int maxNc;
  const int tiles = cdiv(P.Q, 128);
  if ((long long)tiles * TFNAS_MAX_OPS * 64 >= (1 << 23)) return false;
  ws_sched_slots(P, WA, tiles, Sc, maxNc);
  WsCfg cfg;
  if (maxNc > 256 || !ws_fit(maxNc, DimDc::G, 2 * 256 * sizeof(float2), 2 * 256 * sizeof(double), 2, 4, cfg)) return false;
  const size_t smem = ws_smem_bytes(cfg);
  WsDcArgs A{P, WA, G, Zb, dzc2, D, bn2, DC, dg, sD};
  ProfScope ps("dc", 4.0 * P.Q * ((double)P.oc * (1 + P.na) + 2.0 * P.MC) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  const int grid = min(Sc.n_items, sm_count());
  const bool vec = false;
  WS_LAUNCH(k_ws_dc, DimDc::NTHR);
  return true;
}

bool ws_dx(const Plan& P, const UmW& W, const DxChunks& CH, const float* DA, const float* UH, float* dx, double* sU,
           cudaStream_t st) {
  const int tiles = cdiv(P.P, 128);
  const int sms = sm_count();
  // split the stacked K axis until there are about three items per SM (each at least 4 chunks long)
  int ksplit = 1;
  if (tiles < 3 * sms) ksplit = max(1, min(max(1, CH.total / 4), (3 * sms) / tiles));
  if ((long long)tiles * ksplit >= (1 << 23)) return false;
  WsCfg cfg;
  // per-CTA accumulation of the BN1-backward sums when the K axis is not split and 2 * MC doubles fit next to the pipeline
  // (off by default: measured 5-18 % slower on B200 -- shared-memory fp64 adds are CAS loops, and with 2 * MC distinct
  //  addresses the global atomics of this kernel do not serialise the way the BN3 / BN2-backward sums did)
  static const bool dx_acc = getenv("TFNAS_DX_ACC") && strcmp(getenv("TFNAS_DX_ACC"), "1") == 0;
  int acc_rows = (dx_acc && ksplit == 1 && P.MC <= 2048) ? P.MC : 0;
  if (W.Nc > 256) return false;
  if (!acc_rows || !ws_fit(W.Nc, DimDx::G, 0, (uint32_t)(2 * acc_rows * sizeof(double)), 2, 4, cfg) || cfg.NB < 2) {
    acc_rows = 0;
    if (!ws_fit(W.Nc, DimDx::G, 0, 0, 2, 4, cfg)) return false;
  }
  WsSched Sc;
  memset(&Sc, 0, sizeof(Sc));
  Sc.tiles_m = tiles;
  Sc.inv_tiles = 1.f / (float)tiles;
  Sc.na = P.na;
  Sc.n_items = tiles * ksplit;
  if (ksplit > 1) cudaMemsetAsync(dx, 0, (size_t)P.P * P.ic * sizeof(float), st);
  const size_t smem = ws_smem_bytes(cfg);
  WsDxArgs A{P, W, CH, ksplit, DA, UH, dx, sU, acc_rows};
  ProfScope ps("dx", 4.0 * P.P * (2.0 * P.MC + P.ic) + 4.0 * P.MC * P.ic, 2.0 * P.P * (double)P.MC * P.ic, st);
  const int grid = min(Sc.n_items, sms);
  const bool vec = ws_vec_ok(P.HW, P.P);
  WS_LAUNCH(k_ws_dx, DimDx::NTHR);
  return true;
}

// debug: enable / disable the phase trace of the persistent kernels (buf: device memory, n_ctas * WS_TRACE_SLOTS u64)
extern "C" int tfnas_debug_ws_trace(void* buf, int n_ctas) {
  unsigned long long* p = (unsigned long long*)buf;
  if (cudaMemcpyToSymbol(g_ws_trace, &p, sizeof(p)) != cudaSuccess) return TFNAS_E_CUDA;
  if (cudaMemcpyToSymbol(g_ws_trace_n, &n_ctas, sizeof(n_ctas)) != cudaSuccess) return TFNAS_E_CUDA;
  return TFNAS_OK;
}

// test helper (host only, no device needed): the pipeline configuration the persistent kernel `which` (0 expand, 1 project,
// 2 dc, 3 dx) would run an N chunk of `max_nc` columns with -> out5 = {operand stages S, weight slots NB, dynamic shared
// memory bytes, producer groups G, threads per CTA}; returns TFNAS_E_UNSUPPORTED when nothing fits (the caller would fall
// back to the per-tile kernels)
extern "C" int tfnas_debug_ws_config(int which, int max_nc, uint32_t* out5) {
  if (!out5 || max_nc < 16 || (max_nc & 15)) return TFNAS_E_INVALID;
  WsCfg c;
  bool ok = false;
  int G = 0, nthr = 0;
  switch (which) {
    case 0: G = DimExpand::G; nthr = DimExpand::NTHR; ok = ws_fit(max_nc, G, WsExpandT<DimExpand>::CF, 0, 2, 6, c); break;
    case 1: G = DimProject::G; nthr = DimProject::NTHR; ok = ws_fit(max_nc, G, 0, 2 * 256 * sizeof(double), 1, 4, c); break;
    case 2: G = DimDc::G; nthr = DimDc::NTHR; ok = ws_fit(max_nc, G, 2 * 256 * sizeof(float2), 2 * 256 * sizeof(double), 2, 4, c); break;
    case 3: G = DimDx::G; nthr = DimDx::NTHR; ok = ws_fit(max_nc, G, 0, 0, 2, 4, c); break;
    default: return TFNAS_E_INVALID;
  }
  if (!ok || max_nc > 256) return TFNAS_E_UNSUPPORTED;
  out5[0] = (uint32_t)c.S; out5[1] = (uint32_t)c.NB; out5[2] = (uint32_t)ws_smem_bytes(c); out5[3] = (uint32_t)G; out5[4] = (uint32_t)nthr;
  return TFNAS_OK;
}
