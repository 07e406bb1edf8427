// Persistent, warp-specialised tcgen05 versions of the four flattened-pixel 1x1-conv GEMMs (expand, project, dc, dx).
//
//   Out[o][p] = sum_k W[o][k] * In[k][p],   p = flattened (n,h,w) pixel, 128 pixels per tile (MMA M = 128)
//
// One CTA per SM walks a static list of work items (candidate slot, N chunk, 128-pixel tile).  The CTA's warps
// have fixed roles that meet only through mbarriers, so no phase of one tile ever waits for another phase:
//   NE warps      EPILOGUE   tcgen05.ld of the finished accumulator (lane quarter = warp & 3, column part = warp >> 2),
//                            BN statistics / SE partial sums / coalesced per-channel stores
//   NP warps      PRODUCERS  cp.async of the raw input rows (and their per-row constants) RS-1 K chunks ahead into a
//                            thread-private ring; prologue math (BN / activation / SE gate / BN-backward on load),
//                            tf32 hi/lo split, st.shared into one of S operand stages (MN-major, 128B swizzle)
//   1 warp        MMA        one thread: waits operand stage + weight slot, issues the 12 kind::tf32 MMAs of the K chunk
//                            (hi*hi + lo*hi + hi*lo per K=8 step), commits to the stage's / slot's "empty" barriers
//   1 warp        WEIGHTS    one thread: bulk (TMA) copies of the pre-split, pre-swizzled weight blocks, NB-1 chunks ahead
// (NE, NP) = (8, 16) for the prologue-bound kernels (project, dx), (16, 8) for the epilogue-bound ones (expand, dc):
// the prologue / epilogue instruction streams are latency-bound per warp, so each side gets the warps it can use.
// The accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of tile j overlaps the main loop of
// tile j+1.  Per-CTA set-up (TMEM allocation, barrier init) is paid once per SM instead of once per tile.
//
// The numerics are those of umma_pw.cu (same operand split, same MMA order, same epilogue arithmetic).
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
#define WS_ROWS_BYTES(ntens) ((ntens) * WS_KC * 128 * 4)      // one K chunk of raw rows: 32 rows x 128 pixels (16 KB) per tensor
#define WS_WCONST 128                         // bytes of per-warp row constants in one ring stage
// Warp split of a kernel: NE epilogue warps (multiple of 4: lane quarter = warp & 3, column part = warp >> 2),
// NP producer warps (divides 32), then the MMA warp and the weight-copy warp.
template <int NE_, int NP_>
struct WsDim {
  static constexpr int NE = NE_, NP = NP_;
  static constexpr int RW = WS_KC / NP_;      // K rows per producer warp per chunk
  static constexpr int NPT = NP_ * 32;
  static constexpr int MMA_WARP = NE_ + NP_, TMA_WARP = NE_ + NP_ + 1;
  static constexpr int NTHR = 32 * (NE_ + NP_ + 2);
};
typedef WsDim<8, 16> DimProd;                 // prologue-bound kernels (project, dx)
typedef WsDim<16, 8> DimEpi;                  // epilogue-bound kernels (expand, dc)
#define WS_MAXS 4
#define WS_MAXB 4
#define WS_SMEM_LIMIT 232448                  // 227 KB opt-in maximum

struct WsSched {
  int n_items, tiles_m, na;
  float inv_tiles;
  int first[TFNAS_MAX_OPS + 1];               // first item of each slot (items of a slot: N chunk major, tile minor)
};
struct WsCfg { int S, NB, RS; uint32_t wslot, stage_bytes, cf_bytes; };

struct WsSmem {
  unsigned char *a, *w, *ring;
  uint64_t *full, *empty, *wfull, *wempty, *accfull, *accempty;
  uint64_t* rbar;          // bulk mode: [RS][NP] "rows landed" barriers, one per (ring stage, producer warp)
  uint32_t* tmem_slot;
  float2* cf;
};

// layout: [S operand stages x (hi 16K | lo 16K)] [NB weight slots] [barriers 1 KB] [cf tables] [ring]
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
  M.rbar = (uint64_t*)(tail + 256);           // 4 x 16 x 8 B
  M.cf = (float2*)(tail + 1024);
  M.ring = tail + 1024 + c.cf_bytes;
}
static size_t ws_smem_bytes(const WsCfg& c) {
  return 1024 + (size_t)c.S * 32768 + (size_t)c.NB * c.wslot + 1024 + c.cf_bytes + (size_t)c.RS * c.stage_bytes;
}
// deepest pipeline that fits: operand stages S, weight slots NB, ring stages RS
static bool ws_fit(int maxNc, uint32_t stage_bytes, uint32_t cf_bytes, WsCfg& c) {
  static const int pref[][3] = {{3, 4, 4}, {3, 3, 4}, {3, 3, 3}, {2, 3, 3}, {2, 2, 3}, {2, 2, 2}, {2, 1, 2}};
  c.wslot = (uint32_t)2 * maxNc * 128;
  c.stage_bytes = stage_bytes;
  c.cf_bytes = cf_bytes;
  for (auto& p : pref) {
    c.S = p[0]; c.NB = p[1]; c.RS = p[2];
    if (ws_smem_bytes(c) <= WS_SMEM_LIMIT) return true;
  }
  return false;
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
__device__ __forceinline__ void ws_cp16(void* dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;          // src-size 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ws_cp4(void* dst, const void* src, bool valid) {
  const uint32_t n = valid ? 4u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ws_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ws_cp_wait(int pending) {      // wait_group needs an immediate; pending is CTA-uniform
  if (pending >= 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
  else if (pending == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
  else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
  else asm volatile("cp.async.wait_group 0;" ::: "memory");
}
template <int NTHREADS>
__device__ __forceinline__ void ws_bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// 4 pixels of row `ch` of an [N][C][HW] tensor -> this thread's 16 B ring slot (zero-filled when !rowok / invalid pixel)
__device__ __forceinline__ void ws_ring_row(unsigned char* dst, const float* __restrict__ T, const float* __restrict__ Tb,
                                            const Px4& px, int C, int ch, int HW, bool rowok) {
  if (px.vec) {
    ws_cp16(dst, rowok ? Tb + (size_t)ch * HW : T, rowok);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = rowok && px.v[e];
      ws_cp4(dst + 4 * e, ok ? T + ((size_t)px.n[e] * C + ch) * HW + px.hw[e] : T, ok);
    }
  }
}

template <class D>
struct WsTile { float4 hi[D::RW], lo[D::RW]; };
template <class D>
__device__ __forceinline__ void ws_split(WsTile<D>& t, int i, const float (&v)[4]) {
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
  t.hi[i] = make_float4(h[0], h[1], h[2], h[3]);
  t.lo[i] = make_float4(l[0], l[1], l[2], l[3]);
}
// row kk = pw + i * NP of the chunk, 16 B chunk of pixels 4*lane .. 4*lane+3
template <class D>
__device__ __forceinline__ void ws_store(const WsTile<D>& t, unsigned char* a_hi, int pw, int lane) {
#pragma unroll
  for (int i = 0; i < D::RW; ++i) {
    const uint32_t off = mn_chunk_off(lane * 4, pw + i * D::NP, WS_KC * 128);
    *(float4*)(a_hi + off) = t.hi[i];
    *(float4*)(a_hi + 16384 + off) = t.lo[i];
  }
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

// -------------------------------------------------------------------------------------------------
// skeleton.  T supplies
//   Args                                     kernel arguments (by value, __grid_constant__)
//   NTENS                                    raw tensors staged per K chunk (ring stage = NTENS x 16 KB + constants)
//   geom(A, Sc, item, nK, Nc, wb)            K chunks, MMA N and first prepped weight block of an item
//   Prod{A, ptid, pw, lane}: bind(Sc, item), nK, issue(c, stage), compute(c, stage, tile)
//   epi_prep(A, Sc, item, cf, etid)          fill the per-column coefficient table (epilogue warps, before the barrier)
//   epi_run(A, Sc, item, acc, cf, ew, lane)  consume the accumulator
// -------------------------------------------------------------------------------------------------
// ---- optional phase trace (debug build -DUM_TRACE; tfnas_debug_ws_trace) ------------------------------------
// Producer warp 0 / the MMA thread of each traced CTA accumulate clock64() deltas per phase.
//   producer slots: [0] iterations [1] issue [2] wait data [3] compute [4] wait empty [5] store [6] fence [7] arrive
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

template <class T, bool BULK>
__device__ __forceinline__ void ws_run(const typename T::Args& A, const WsSched& Sc, const WsCfg& cfg) {
  using D = typename T::Dim;
  extern __shared__ __align__(1024) unsigned char ws_raw[];
  WsSmem M;
  ws_carve(ws_raw, cfg, M);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < WS_MAXS; ++s) { mbar_init(&M.full[s], D::NP); mbar_init(&M.empty[s], 1); }
    for (int s = 0; s < WS_MAXB; ++s) { mbar_init(&M.wfull[s], 1); mbar_init(&M.wempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&M.accfull[s], 1); mbar_init(&M.accempty[s], D::NE); }
    for (int s = 0; s < WS_MAXS * D::NP; ++s) mbar_init(&M.rbar[s], 1);
    fence_barrier_init();
  }
  if (warp == D::MMA_WARP) tmem_alloc(M.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *M.tmem_slot;
  const int item0 = blockIdx.x, istep = gridDim.x;

  if (warp < D::NE) {
    // ------------------------------ epilogue ------------------------------
    int jj = 0;
    for (int item = item0; item < Sc.n_items; item += istep, ++jj) {
      const int ab = jj & 1;
      const uint32_t au = (uint32_t)jj >> 1;
      float2* cf = M.cf + ab * 256;
      T::epi_prep(A, Sc, item, cf, tid);
      if (T::CF) ws_bar_epi<D::NE * 32>();
      mbar_wait(&M.accfull[ab], au & 1);
      tc_fence_after();
      T::epi_run(A, Sc, item, tmem + ab * WS_ACC_STRIDE, cf, warp, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&M.accempty[ab]);
    }
  } else if (warp < D::NE + D::NP) {
    // ------------------------------ producers ------------------------------
    const int ptid = tid - D::NE * 32, pw = ptid >> 5;
    typename std::conditional<BULK, typename T::ProdB, typename T::Prod>::type fi{A, ptid, pw, lane}, fc{A, ptid, pw, lane};
    int item_i = item0, ci = 0;
    bool vi = item_i < Sc.n_items;
    if (vi) fi.bind(Sc, item_i);
    // fetch the next chunk of this warp's rows into ring stage `stg`.  Generic mode: per-thread cp.async into private
    // slots.  Bulk mode: a few lanes issue bulk (TMA) copies of whole 512 B rows, completion on the warp's own mbarrier.
    auto issue_next = [&](int stg) {
      if (vi) {
        if (BULK) fi.issue(ci, M.ring + (size_t)stg * cfg.stage_bytes, &M.rbar[stg * D::NP + pw]);
        else fi.issue(ci, M.ring + (size_t)stg * cfg.stage_bytes, nullptr);
        if (++ci == fi.nK) {
          ci = 0;
          item_i += istep;
          vi = item_i < Sc.n_items;
          if (vi) fi.bind(Sc, item_i);
        }
      }
      if (!BULK || T::WCONST) ws_cp_commit();  // one group per call (possibly empty) keeps the group count uniform
    };
    const int RS = cfg.RS, S = cfg.S;
    for (int k = 0; k < RS - 1; ++k) issue_next(k);
    WsTrace tr;
    tr.begin(pw == 0 && lane == 0);
    int rs = 0, rs_issue = RS - 1;             // ring stage of the chunk being emitted / being fetched
    uint32_t rph = 0;                          // parity of the ring barriers of stage rs
    int s = 0;
    uint32_t eph = 1;                          // parity to wait for on empty[s]: passes on the first use of each stage
    for (int item = item0; item < Sc.n_items; item += istep) {
      fc.bind(Sc, item);
      const int n = fc.nK;
      for (int c = 0; c < n; ++c) {
        __syncwarp();                          // every lane is done with the ring stage about to be refilled
        tr.count();
        tr.mark(7);
        issue_next(rs_issue);
        tr.mark(1);
        if (!BULK || T::WCONST) ws_cp_wait(RS - 1);      // this thread's cp.async copies of the current chunk have landed
        if (BULK) mbar_wait(&M.rbar[rs * D::NP + pw], rph);
        __syncwarp();                          // ... and so have the warp-shared constants fetched by the other lanes
        tr.mark(2);
        WsTile<D> t;
        fc.compute(c, M.ring + (size_t)rs * cfg.stage_bytes, t);
        tr.mark(3);
        mbar_wait(&M.empty[s], eph);           // the MMAs that read this operand stage have retired
        tc_fence_after();
        tr.mark(4);
        ws_store<D>(t, M.a + (size_t)s * 32768, pw, lane);
        tr.mark(5);
        fence_proxy_async();
        tr.mark(6);
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.full[s]);
        if (++rs == RS) { rs = 0; rph ^= 1; }
        rs_issue = rs_issue + 1 == RS ? 0 : rs_issue + 1;
        if (++s == S) { s = 0; eph ^= 1; }
      }
    }
    if (!BULK || T::WCONST) ws_cp_wait(0);
    tr.end(0);
  } else if (warp == D::MMA_WARP) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int s = 0, b = 0, jj = 0;
      uint32_t fph = 0, wph = 0;               // parities of full[s] / wfull[b]
      const uint32_t a0 = smem_u32(M.a), w0 = smem_u32(M.w);
      WsTrace tr;
      tr.begin(true);
      for (int item = item0; item < Sc.n_items; item += istep, ++jj) {
        int nK, Nc;
        const char* wb;
        T::geom(A, Sc, item, nK, Nc, wb);
        const int ab = jj & 1;
        const uint32_t au = (uint32_t)jj >> 1;
        tr.mark(3);
        mbar_wait(&M.accempty[ab], (au & 1) ^ 1);     // the epilogue has drained this accumulator (passes for the first two)
        tc_fence_after();
        tr.mark(4);
        const uint32_t idesc = idesc_tf32(128, Nc, 1, 0);
        const uint32_t acc = tmem + ab * WS_ACC_STRIDE;
        for (int c = 0; c < nK; ++c) {
          tr.count();
          mbar_wait(&M.wfull[b], wph);
          tr.mark(1);
          mbar_wait(&M.full[s], fph);
          tc_fence_after();
          tr.mark(2);
          ws_issue(a0 + s * 32768, w0 + b * cfg.wslot, Nc, acc, idesc, c == 0);
          mma_commit(&M.empty[s]);
          mma_commit(&M.wempty[b]);
          tr.mark(3);
          if (++s == cfg.S) { s = 0; fph ^= 1; }
          if (++b == cfg.NB) { b = 0; wph ^= 1; }
        }
        mma_commit(&M.accfull[ab]);
      }
      tr.end(8);
    }
    __syncwarp();
  } else {
    // ------------------------------ weight copies ------------------------------
    if (lane == 0) {
      int b = 0;
      uint32_t eph = 1;
      for (int item = item0; item < Sc.n_items; item += istep) {
        int nK, Nc;
        const char* wb;
        T::geom(A, Sc, item, nK, Nc, wb);
        const uint32_t bytes = (uint32_t)2 * Nc * 128;
        for (int c = 0; c < nK; ++c) {
          mbar_wait(&M.wempty[b], eph);
          mbar_expect_tx(&M.wfull[b], bytes);
          bulk_g2s(M.w + (size_t)b * cfg.wslot, wb + (size_t)c * bytes, bytes, &M.wfull[b]);
          if (++b == cfg.NB) { b = 0; eph ^= 1; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == D::MMA_WARP) tmem_dealloc(tmem, 512);
}


// Bulk mode (HW % 4 == 0, HW >= 128, pixel count a multiple of 128): a 128-pixel tile is one contiguous 512 B run per
// channel row, or two runs when it straddles an image boundary (len1 pixels in image n0, the rest at the start of n0+1).
struct WsSeg { int n0, hw0, len1; };
__device__ __forceinline__ WsSeg ws_seg(int tile0, int HW) {
  WsSeg g;
  g.n0 = fast_div(tile0, HW, __frcp_rn((float)HW));
  g.hw0 = tile0 - g.n0 * HW;
  g.len1 = min(128, HW - g.hw0);
  return g;
}
// lane `l` of the issuing group copies run (l & 1) of row kk: src0 / src1 = start of the row's first / second run
__device__ __forceinline__ void ws_bulk_row(unsigned char* row_dst, const float* src0, const float* src1, int seg, int len1,
                                            bool rowok, uint64_t* bar) {
  const int nbytes = seg ? (128 - len1) * 4 : len1 * 4;
  if (rowok && nbytes > 0) bulk_g2s(row_dst + (seg ? len1 * 4 : 0), seg ? src1 : src0, (uint32_t)nbytes, bar);
}
static bool ws_bulk_ok(int HW, int total) { return (HW % 4) == 0 && HW >= 128 && (total % 128) == 0; }

// images covered by a 128-pixel tile: first image and count
__device__ __forceinline__ void ws_tile_images(int tile0, int total, int HW, int& n0, int& nimg) {
  const float r = __frcp_rn((float)HW);
  n0 = fast_div(min(tile0, total - 1), HW, r);
  nimg = fast_div(min(tile0 + 127, total - 1), HW, r) - n0 + 1;
}

// -------------------------------------------------------------------------------------------------
// F1a: expand      UH = BN1(W1 x)
// -------------------------------------------------------------------------------------------------
struct WsExpandArgs { Plan P; UmWAll WA; const float* x; const float* bn1; float* UH; };
struct WsExpandT {
  using Args = WsExpandArgs;
  using Dim = DimEpi;
  static constexpr int NTENS = 1;
  static constexpr uint32_t STAGE = WS_ROWS_BYTES(1);
  static constexpr uint32_t CF = 2 * 256 * sizeof(float2);
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const UmW& W = A.WA.s[slot];
    nK = W.nK; Nc = W.Nc;
    wb = (const char*)W.wp + (size_t)nc * W.nK * 2 * W.Nc * 128;
  }
  struct Prod {
    const Args& A; int ptid, pw, lane;
    int nK; Px4 px; const float* xb;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int slot, nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      nK = A.WA.s[slot].nK;
      px_decomp(px, mt * 128 + lane * 4, A.P.P, A.P.HW);
      xb = A.x + (size_t)px.n[0] * A.P.ic * A.P.HW + px.hw[0];
    }
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t*) const {
#pragma unroll
      for (int i = 0; i < DimEpi::RW; ++i) {
        const int k = c * WS_KC + pw + i * DimEpi::NP;
        ws_ring_row(st + ((size_t)i * DimEpi::NPT + ptid) * 16, A.x, xb, px, A.P.ic, k, A.P.HW, k < A.P.ic);
      }
    }
    __device__ __forceinline__ void compute(int, const unsigned char* st, WsTile<DimEpi>& t) const {
#pragma unroll
      for (int i = 0; i < DimEpi::RW; ++i) {
        const float4 a = *(const float4*)(st + ((size_t)i * DimEpi::NPT + ptid) * 16);
        const float v[4] = {a.x, a.y, a.z, a.w};
        ws_split<DimEpi>(t, i, v);
      }
    }
  };
  struct ProdB {
    const Args& A; int ptid, pw, lane;
    int nK, len1; size_t off0, off1;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int slot, nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      nK = A.WA.s[slot].nK;
      const WsSeg g = ws_seg(mt * 128, A.P.HW);
      len1 = g.len1;
      off0 = (size_t)g.n0 * A.P.ic * A.P.HW + g.hw0;
      off1 = (size_t)(g.n0 + 1) * A.P.ic * A.P.HW;
    }
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t* bar) const {
      if (lane < 2 * DimEpi::RW) {
        const int kk = pw + (lane >> 1) * DimEpi::NP, k = c * WS_KC + kk;
        const size_t ro = (size_t)k * A.P.HW;
        ws_bulk_row(st + kk * 512, A.x + off0 + ro, A.x + off1 + ro, lane & 1, len1, k < A.P.ic, bar);
      }
      if (lane == 0) {
        int nv = 0;
#pragma unroll
        for (int i = 0; i < DimEpi::RW; ++i) nv += (c * WS_KC + pw + i * DimEpi::NP < A.P.ic) ? 1 : 0;
        mbar_expect_tx(bar, nv * 512);
      }
    }
    __device__ __forceinline__ void compute(int c, const unsigned char* st, WsTile<DimEpi>& t) const {
#pragma unroll
      for (int i = 0; i < DimEpi::RW; ++i) {
        const int kk = pw + i * DimEpi::NP;
        const float4 a = *(const float4*)(st + kk * 512 + lane * 16);
        const bool ok = c * WS_KC + kk < A.P.ic;
        const float v[4] = {ok ? a.x : 0.f, ok ? a.y : 0.f, ok ? a.z : 0.f, ok ? a.w : 0.f};
        ws_split<DimEpi>(t, i, v);
      }
    }
  };
  static constexpr bool WCONST = false;
  static __device__ __forceinline__ void epi_prep(const Args& A, const WsSched& Sc, int item, float2* cf, int etid) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    for (int i = etid; i < ncol; i += DimEpi::NE * 32) cf[i] = make_float2(A.bn1[cst0 + i], A.bn1[A.P.MC + cst0 + i]);
  }
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2* cft,
                                                 int ew, int lane) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    const WsEpi e = ws_epi<DimEpi>(mt * 128, A.P.P, A.P.HW, Nc, ncol, acc, ew, lane);
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
struct WsProjectArgs { Plan P; UmWAll WA; const float* D; const float* bn2; const float* seg; float* Zb; double* st3; };
template <int ACT>
struct WsProjectT {
  using Args = WsProjectArgs;
  using Dim = DimProd;
  static constexpr int NTENS = 1;
  static constexpr uint32_t STAGE = WS_ROWS_BYTES(1) + DimProd::NP * WS_WCONST;
  static constexpr uint32_t CF = 0;
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const UmW& W = A.WA.s[slot];
    nK = W.nK; Nc = W.Nc;
    wb = (const char*)W.wp + (size_t)nc * W.nK * 2 * W.Nc * 128;
  }
  // per-warp constants of a ring stage (floats): [0..3] BN2 mean of rows 0..3, [4..7] rstd, [8 + 4 i + m] SE gate of
  // row i for the m-th image of the tile (tiles spanning more than 4 images read the gate from global memory)
  struct Prod {
    const Args& A; int ptid, pw, lane;
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
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t*) const {
#pragma unroll
      for (int i = 0; i < DimProd::RW; ++i) {
        const int k = c * WS_KC + pw + i * DimProd::NP;
        ws_ring_row(st + ((size_t)i * DimProd::NPT + ptid) * 16, A.D, Db, px, A.P.MC, coff + k, A.P.HWo, k < mc);
      }
      float* wc = (float*)(st + WS_ROWS_BYTES(1) + pw * WS_WCONST);
      if (lane < 8) {
        const int i = lane & 3, k = c * WS_KC + pw + i * DimProd::NP;
        const bool ok = i < DimProd::RW && k < mc;
        ws_cp4(wc + lane, A.bn2 + (lane < 4 ? 0 : A.P.MC) + coff + (ok ? k : 0), ok);
      } else if (lane < 8 + 4 * DimProd::RW && se > 0 && nimg <= 4) {
        const int i = (lane - 8) >> 2, m = (lane - 8) & 3;
        const int k = c * WS_KC + pw + i * DimProd::NP;
        const bool ok = k < mc && m < nimg;
        ws_cp4(wc + lane, ok ? A.seg + (size_t)(n0 + m) * A.P.MCse + soff + k : A.seg, ok);
      }
    }
    __device__ __forceinline__ void compute(int c, const unsigned char* st, WsTile<DimProd>& t) const {
      const float* wc = (const float*)(st + WS_ROWS_BYTES(1) + pw * WS_WCONST);
#pragma unroll
      for (int i = 0; i < DimProd::RW; ++i) {
        const int k = c * WS_KC + pw + i * DimProd::NP;
        const float4 a = *(const float4*)(st + ((size_t)i * DimProd::NPT + ptid) * 16);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < mc) {
          const float mu = wc[i], r = wc[4 + i];
          const float d[4] = {a.x, a.y, a.z, a.w};
          if (px.vec) {        // 4 valid pixels of one image
            const float gt = se > 0 ? (nimg <= 4 ? wc[8 + 4 * i + (px.n[0] - n0)] : A.seg[(size_t)px.n[0] * A.P.MCse + soff + k]) : 1.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = act_f<ACT>((d[e] - mu) * r) * gt;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float b = act_f<ACT>((d[e] - mu) * r);
              if (se > 0 && px.v[e])
                b *= nimg <= 4 ? wc[8 + 4 * i + (px.n[e] - n0)] : A.seg[(size_t)px.n[e] * A.P.MCse + soff + k];
              v[e] = px.v[e] ? b : 0.f;
            }
          }
        }
        ws_split<DimProd>(t, i, v);
      }
    }
  };
  struct ProdB {
    const Args& A; int ptid, pw, lane;
    int nK, mc, coff, soff, se, n0, len1; size_t off0, off1;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int slot, nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      const Cand& cd = A.P.c[slot];
      mc = cd.mc; coff = cd.coff; soff = cd.soff; se = cd.se;
      nK = A.WA.s[slot].nK;
      const WsSeg g = ws_seg(mt * 128, A.P.HWo);
      n0 = g.n0; len1 = g.len1;
      off0 = ((size_t)g.n0 * A.P.MC + coff) * A.P.HWo + g.hw0;
      off1 = ((size_t)(g.n0 + 1) * A.P.MC + coff) * A.P.HWo;
    }
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t* bar) const {
      if (lane < 2 * DimProd::RW) {
        const int kk = pw + (lane >> 1) * DimProd::NP, k = c * WS_KC + kk;
        const size_t ro = (size_t)k * A.P.HWo;
        ws_bulk_row(st + kk * 512, A.D + off0 + ro, A.D + off1 + ro, lane & 1, len1, k < mc, bar);
      }
      if (lane == 0) {
        int nv = 0;
#pragma unroll
        for (int i = 0; i < DimProd::RW; ++i) nv += (c * WS_KC + pw + i * DimProd::NP < mc) ? 1 : 0;
        mbar_expect_tx(bar, nv * 512);
      }
      // per-warp constants: [0..3] BN2 mean, [4..7] rstd, [8 + 4 i + m] SE gate of row i for image n0 + m (m = 0, 1)
      float* wc = (float*)(st + WS_ROWS_BYTES(1) + pw * WS_WCONST);
      if (lane < 8) {
        const int i = lane & 3, k = c * WS_KC + pw + i * DimProd::NP;
        const bool ok = i < DimProd::RW && k < mc;
        ws_cp4(wc + lane, A.bn2 + (lane < 4 ? 0 : A.P.MC) + coff + (ok ? k : 0), ok);
      } else if (lane < 8 + 4 * DimProd::RW && se > 0) {
        const int i = (lane - 8) >> 2, m = (lane - 8) & 3;
        const int k = c * WS_KC + pw + i * DimProd::NP;
        const bool ok = k < mc && m < 2 && (m == 0 || len1 < 128);
        ws_cp4(wc + lane, ok ? A.seg + (size_t)(n0 + m) * A.P.MCse + soff + k : A.seg, ok);
      }
    }
    __device__ __forceinline__ void compute(int c, const unsigned char* st, WsTile<DimProd>& t) const {
      const float* wc = (const float*)(st + WS_ROWS_BYTES(1) + pw * WS_WCONST);
      const int img = lane * 4 >= len1 ? 1 : 0;       // the thread's 4 pixels lie in one image (HW % 4 == 0)
#pragma unroll
      for (int i = 0; i < DimProd::RW; ++i) {
        const int kk = pw + i * DimProd::NP;
        const bool ok = c * WS_KC + kk < mc;
        const float4 a = *(const float4*)(st + kk * 512 + lane * 16);
        const float mu = wc[i], r = wc[4 + i];
        const float gt = se > 0 ? wc[8 + 4 * i + img] : 1.f;
        const float d[4] = {a.x, a.y, a.z, a.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = ok ? act_f<ACT>((d[e] - mu) * r) * gt : 0.f;   // rows past mc were not fetched
        ws_split<DimProd>(t, i, v);
      }
    }
  };
  static constexpr bool WCONST = true;
  static __device__ __forceinline__ void epi_prep(const Args&, const WsSched&, int, float2*, int) {}
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2*,
                                                 int ew, int lane) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const int Nc = A.WA.s[slot].Nc, oc = A.P.oc;
    const int ncol = min(Nc, oc - nc * Nc);
    const WsEpi e = ws_epi<DimProd>(mt * 128, A.P.Q, A.P.HWo, Nc, ncol, acc, ew, lane);
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
      if (lane < 16 && col < e.c_hi) {
        atomicAdd(&A.st3[2 * (o0 + col)], (double)s1);
        atomicAdd(&A.st3[2 * (o0 + col) + 1], (double)s2);
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
template <int ACT>
struct WsDcT {
  using Args = WsDcArgs;
  using Dim = DimEpi;
  static constexpr int NTENS = 2;
  static constexpr uint32_t STAGE = WS_ROWS_BYTES(2) + DimEpi::NP * WS_WCONST;
  static constexpr uint32_t CF = 2 * 256 * sizeof(float2);
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const UmW& W = A.WA.s[slot];
    nK = W.nK; Nc = W.Nc;
    wb = (const char*)W.wp + (size_t)nc * W.nK * 2 * W.Nc * 128;
  }
  // per-warp constants: float4 (A, B, C, -) of rows 0..3
  struct Prod {
    const Args& A; int ptid, pw, lane;
    int nK, slot; Px4 px; const float* Gb; const float* Zbb;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      nK = A.WA.s[slot].nK;
      px_decomp(px, mt * 128 + lane * 4, A.P.Q, A.P.HWo);
      Gb = A.G + (size_t)px.n[0] * A.P.oc * A.P.HWo + px.hw[0];
      Zbb = A.Zb + ((size_t)px.n[0] * A.P.na + slot) * A.P.oc * A.P.HWo + px.hw[0];
    }
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t*) const {
#pragma unroll
      for (int i = 0; i < DimEpi::RW; ++i) {
        const int o = c * WS_KC + pw + i * DimEpi::NP;
        const bool ok = o < A.P.oc;
        ws_ring_row(st + ((size_t)i * DimEpi::NPT + ptid) * 16, A.G, Gb, px, A.P.oc, o, A.P.HWo, ok);
        // Zbb already points at this slot's first channel; the scalar path indexes the full [N][na*oc] tensor
        if (px.vec) ws_cp16(st + ((size_t)(DimEpi::RW + i) * DimEpi::NPT + ptid) * 16, ok ? Zbb + (size_t)o * A.P.HWo : A.Zb, ok);
        else ws_ring_row(st + ((size_t)(DimEpi::RW + i) * DimEpi::NPT + ptid) * 16, A.Zb, A.Zb, px, A.P.na * A.P.oc, slot * A.P.oc + o,
                         A.P.HWo, ok);
      }
      if (lane < DimEpi::RW) {
        const int o = c * WS_KC + pw + lane * DimEpi::NP;
        const bool ok = o < A.P.oc;
        ws_cp16(st + WS_ROWS_BYTES(2) + pw * WS_WCONST + lane * 16, A.dzc2 + (ok ? slot * A.P.oc + o : 0), ok);
      }
    }
    __device__ __forceinline__ void compute(int, const unsigned char* st, WsTile<DimEpi>& t) const {
      const float4* wc = (const float4*)(st + WS_ROWS_BYTES(2) + pw * WS_WCONST);
#pragma unroll
      for (int i = 0; i < DimEpi::RW; ++i) {
        const float4 cf = wc[i];                 // zero past oc
        const float4 a = *(const float4*)(st + ((size_t)i * DimEpi::NPT + ptid) * 16);
        const float4 b = *(const float4*)(st + ((size_t)(DimEpi::RW + i) * DimEpi::NPT + ptid) * 16);
        const float gg[4] = {a.x, a.y, a.z, a.w}, z[4] = {b.x, b.y, b.z, b.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? fmaf(cf.x, gg[e], fmaf(cf.y, z[e], cf.z)) : 0.f;
        ws_split<DimEpi>(t, i, v);
      }
    }
  };
  struct ProdB {
    const Args& A; int ptid, pw, lane;
    int nK, slot, len1; size_t g0, g1, z0, z1;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int nc, mt;
      ws_decode(Sc, item, slot, nc, mt);
      nK = A.WA.s[slot].nK;
      const WsSeg g = ws_seg(mt * 128, A.P.HWo);
      len1 = g.len1;
      const size_t HWo = A.P.HWo;
      g0 = (size_t)g.n0 * A.P.oc * HWo + g.hw0;
      g1 = (size_t)(g.n0 + 1) * A.P.oc * HWo;
      z0 = ((size_t)g.n0 * A.P.na + slot) * A.P.oc * HWo + g.hw0;
      z1 = ((size_t)(g.n0 + 1) * A.P.na + slot) * A.P.oc * HWo;
    }
    // ring stage: [G rows 16 KB | Z rows 16 KB | constants]
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t* bar) const {
      if (lane < 4 * DimEpi::RW) {
        const int tsel = lane / (2 * DimEpi::RW), r = lane % (2 * DimEpi::RW);
        const int kk = pw + (r >> 1) * DimEpi::NP, o = c * WS_KC + kk;
        const size_t ro = (size_t)o * A.P.HWo;
        const float* T = tsel ? A.Zb : A.G;
        ws_bulk_row(st + tsel * 16384 + kk * 512, T + (tsel ? z0 : g0) + ro, T + (tsel ? z1 : g1) + ro, r & 1, len1, o < A.P.oc, bar);
      }
      if (lane == 0) {
        int nv = 0;
#pragma unroll
        for (int i = 0; i < DimEpi::RW; ++i) nv += (c * WS_KC + pw + i * DimEpi::NP < A.P.oc) ? 1 : 0;
        mbar_expect_tx(bar, nv * 1024);
      }
      if (lane < DimEpi::RW) {
        const int o = c * WS_KC + pw + lane * DimEpi::NP;
        const bool ok = o < A.P.oc;
        ws_cp16(st + WS_ROWS_BYTES(2) + pw * WS_WCONST + lane * 16, A.dzc2 + (ok ? slot * A.P.oc + o : 0), ok);
      }
    }
    __device__ __forceinline__ void compute(int c, const unsigned char* st, WsTile<DimEpi>& t) const {
      const float4* wc = (const float4*)(st + WS_ROWS_BYTES(2) + pw * WS_WCONST);
#pragma unroll
      for (int i = 0; i < DimEpi::RW; ++i) {
        const int kk = pw + i * DimEpi::NP;
        const bool ok = c * WS_KC + kk < A.P.oc;
        const float4 cf = wc[i];
        const float4 a = *(const float4*)(st + kk * 512 + lane * 16);
        const float4 b = *(const float4*)(st + 16384 + kk * 512 + lane * 16);
        const float gg[4] = {a.x, a.y, a.z, a.w}, z[4] = {b.x, b.y, b.z, b.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = ok ? fmaf(cf.x, gg[e], fmaf(cf.y, z[e], cf.z)) : 0.f;
        ws_split<DimEpi>(t, i, v);
      }
    }
  };
  static constexpr bool WCONST = true;
  static __device__ __forceinline__ void epi_prep(const Args& A, const WsSched& Sc, int item, float2* cf, int etid) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    for (int i = etid; i < ncol; i += DimEpi::NE * 32) cf[i] = make_float2(A.bn2[cst0 + i], A.bn2[A.P.MC + cst0 + i]);
  }
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2* cft,
                                                 int ew, int lane) {
    int slot, nc, mt;
    ws_decode(Sc, item, slot, nc, mt);
    const Cand& cd = A.P.c[slot];
    const Plan& P = A.P;
    const int Nc = A.WA.s[slot].Nc;
    const int ncol = min(Nc, cd.mc - nc * Nc), cst0 = cd.coff + nc * Nc;
    const WsEpi e = ws_epi<DimEpi>(mt * 128, P.Q, P.HWo, Nc, ncol, acc, ew, lane);
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
          atomicAdd(&A.sD[2 * (cst0 + col)], (double)s1);
          atomicAdd(&A.sD[2 * (cst0 + col) + 1], (double)s2);
        }
      }
    }
  }
};

// -------------------------------------------------------------------------------------------------
// B3b: dx_main = sum_i W1_i^T (r1 * du-hat), K = stacked mid channels (per-candidate chunks of 32), split over items
// -------------------------------------------------------------------------------------------------
struct WsDxArgs { Plan P; UmW W; DxChunks CH; int ksplit; const float* DA; const float* UH; float* dx; double* sU; };
template <int ACT>
struct WsDxT {
  using Args = WsDxArgs;
  using Dim = DimProd;
  static constexpr int NTENS = 2;
  static constexpr uint32_t STAGE = WS_ROWS_BYTES(2);
  static constexpr uint32_t CF = 0;
  // item = tile + tiles_m * ksplit part
  static __device__ __forceinline__ void part(const Args& A, const WsSched& Sc, int item, int& mt, int& ch0, int& ch1) {
    const int ks = fast_div(item, Sc.tiles_m, Sc.inv_tiles);
    mt = item - ks * Sc.tiles_m;
    ch0 = (int)((long long)A.CH.total * ks / A.ksplit);
    ch1 = (int)((long long)A.CH.total * (ks + 1) / A.ksplit);
  }
  static __device__ __forceinline__ void geom(const Args& A, const WsSched& Sc, int item, int& nK, int& Nc, const char*& wb) {
    int mt, ch0, ch1;
    part(A, Sc, item, mt, ch0, ch1);
    nK = ch1 - ch0; Nc = A.W.Nc;
    wb = (const char*)A.W.wp + (size_t)ch0 * 2 * A.W.Nc * 128;
  }
  struct Prod {
    const Args& A; int ptid, pw, lane;
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
    // ring stage: [DA rows | UH rows]
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t*) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
#pragma unroll
      for (int i = 0; i < DimProd::RW; ++i) {
        const int k = k0 + pw + i * DimProd::NP;
        const bool ok = k < mc;
        const int cst = coff + k;
        ws_ring_row(st + ((size_t)i * DimProd::NPT + ptid) * 16, A.DA, DAb, px, A.P.MC, cst, A.P.HW, ok);
        ws_ring_row(st + ((size_t)(DimProd::RW + i) * DimProd::NPT + ptid) * 16, A.UH, UHb, px, A.P.MC, cst, A.P.HW, ok);
      }
    }
    // Rows past the candidate's width and invalid pixels are zero-filled in the ring, so du = 0 * act'(0) = 0 there.
    // BN1's rstd is folded into the prepped weights (umma_prep_bwd), the operand is plain du-hat.
    __device__ __forceinline__ void compute(int c, const unsigned char* st, WsTile<DimProd>& t) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
      float sacc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) sacc[i] = 0.f;
#pragma unroll
      for (int i = 0; i < DimProd::RW; ++i) {
        const float4 a = *(const float4*)(st + ((size_t)i * DimProd::NPT + ptid) * 16);
        const float4 b = *(const float4*)(st + ((size_t)(DimProd::RW + i) * DimProd::NPT + ptid) * 16);
        const float da[4] = {a.x, a.y, a.z, a.w}, uh[4] = {b.x, b.y, b.z, b.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = da[e] * act_df<ACT>(uh[e]);
          sacc[2 * i] += v[e];
          sacc[2 * i + 1] += v[e] * uh[e];
        }
        ws_split<DimProd>(t, i, v);
      }
      // 2*DimProd::RW statistics (rows x {sum du, sum du*uh}) reduced together; lane l < 2*DimProd::RW ends up owning statistic l
      static_assert(2 * DimProd::RW <= 16, "statistics must fit the 16-value warp reduction");
      const float tot = warp_sum16(sacc);
      const int k = k0 + pw + (lane >> 1) * DimProd::NP;
      if (lane < 2 * DimProd::RW && k < mc) atomicAdd(&A.sU[2 * (coff + k) + (lane & 1)], (double)tot);
    }
  };
  struct ProdB {
    const Args& A; int ptid, pw, lane;
    int nK, ch0, len1; size_t off0, off1;
    __device__ __forceinline__ void bind(const WsSched& Sc, int item) {
      int mt, ch1;
      part(A, Sc, item, mt, ch0, ch1);
      nK = ch1 - ch0;
      const WsSeg g = ws_seg(mt * 128, A.P.HW);
      len1 = g.len1;
      off0 = (size_t)g.n0 * A.P.MC * A.P.HW + g.hw0;
      off1 = (size_t)(g.n0 + 1) * A.P.MC * A.P.HW;
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
    // ring stage: [DA rows 16 KB | UH rows 16 KB]
    __device__ __forceinline__ void issue(int c, unsigned char* st, uint64_t* bar) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
      if (lane < 4 * DimProd::RW) {
        const int tsel = lane / (2 * DimProd::RW), r = lane % (2 * DimProd::RW);
        const int kk = pw + (r >> 1) * DimProd::NP, k = k0 + kk;
        const size_t ro = (size_t)(coff + k) * A.P.HW;
        const float* T = tsel ? A.UH : A.DA;
        ws_bulk_row(st + tsel * 16384 + kk * 512, T + off0 + ro, T + off1 + ro, r & 1, len1, k < mc, bar);
      }
      if (lane == 0) {
        int nv = 0;
#pragma unroll
        for (int i = 0; i < DimProd::RW; ++i) nv += (k0 + pw + i * DimProd::NP < mc) ? 1 : 0;
        mbar_expect_tx(bar, nv * 1024);
      }
    }
    __device__ __forceinline__ void compute(int c, const unsigned char* st, WsTile<DimProd>& t) const {
      int mc, coff, k0;
      locate(c, mc, coff, k0);
      float sacc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) sacc[i] = 0.f;
#pragma unroll
      for (int i = 0; i < DimProd::RW; ++i) {
        const int kk = pw + i * DimProd::NP;
        const bool ok = k0 + kk < mc;                 // rows past the candidate's width were not fetched
        const float4 a = *(const float4*)(st + kk * 512 + lane * 16);
        const float4 b = *(const float4*)(st + 16384 + kk * 512 + lane * 16);
        const float da[4] = {a.x, a.y, a.z, a.w}, uh[4] = {b.x, b.y, b.z, b.w};
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float u = ok ? uh[e] : 0.f;
          v[e] = ok ? da[e] * act_df<ACT>(u) : 0.f;
          sacc[2 * i] += v[e];
          sacc[2 * i + 1] += v[e] * u;
        }
        ws_split<DimProd>(t, i, v);
      }
      const float tot = warp_sum16(sacc);
      const int k = k0 + pw + (lane >> 1) * DimProd::NP;
      if (lane < 2 * DimProd::RW && k < mc) atomicAdd(&A.sU[2 * (coff + k) + (lane & 1)], (double)tot);
    }
  };
  static constexpr bool WCONST = false;
  static __device__ __forceinline__ void epi_prep(const Args&, const WsSched&, int, float2*, int) {}
  static __device__ __forceinline__ void epi_run(const Args& A, const WsSched& Sc, int item, uint32_t acc, const float2*,
                                                 int ew, int lane) {
    int mt, ch0, ch1;
    part(A, Sc, item, mt, ch0, ch1);
    const WsEpi e = ws_epi<DimProd>(mt * 128, A.P.P, A.P.HW, A.W.Nc, A.P.ic, acc, ew, lane);
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
template <bool BULK>
__global__ void __launch_bounds__(DimEpi::NTHR, 1) k_ws_expand(const __grid_constant__ WsExpandArgs A, const __grid_constant__ WsSched Sc,
                                                             const __grid_constant__ WsCfg cfg) {
  ws_run<WsExpandT, BULK>(A, Sc, cfg);
}
template <int ACT, bool BULK>
__global__ void __launch_bounds__(DimProd::NTHR, 1) k_ws_project(const __grid_constant__ WsProjectArgs A, const __grid_constant__ WsSched Sc,
                                                               const __grid_constant__ WsCfg cfg) {
  ws_run<WsProjectT<ACT>, BULK>(A, Sc, cfg);
}
template <int ACT, bool BULK>
__global__ void __launch_bounds__(DimEpi::NTHR, 1) k_ws_dc(const __grid_constant__ WsDcArgs A, const __grid_constant__ WsSched Sc,
                                                         const __grid_constant__ WsCfg cfg) {
  ws_run<WsDcT<ACT>, BULK>(A, Sc, cfg);
}
template <int ACT, bool BULK>
__global__ void __launch_bounds__(DimProd::NTHR, 1) k_ws_dx(const __grid_constant__ WsDxArgs A, const __grid_constant__ WsSched Sc,
                                                          const __grid_constant__ WsCfg cfg) {
  ws_run<WsDxT<ACT>, BULK>(A, Sc, cfg);
}

// launch helper: pick the activation / bulk instantiation
#define WS_LAUNCH1(KERN, BULK_, NTHR_) do { \
    if (BULK_) { ensure_smem(KERN<true>, smem); KERN<true><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } \
    else { ensure_smem(KERN<false>, smem); KERN<false><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } } while (0)
#define WS_LAUNCH2(KERN, RELU_, BULK_, NTHR_) do { \
    if (RELU_) { \
      if (BULK_) { ensure_smem(KERN<TFNAS_ACT_RELU, true>, smem); KERN<TFNAS_ACT_RELU, true><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } \
      else { ensure_smem(KERN<TFNAS_ACT_RELU, false>, smem); KERN<TFNAS_ACT_RELU, false><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } \
    } else { \
      if (BULK_) { ensure_smem(KERN<TFNAS_ACT_SWISH, true>, smem); KERN<TFNAS_ACT_SWISH, true><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } \
      else { ensure_smem(KERN<TFNAS_ACT_SWISH, false>, smem); KERN<TFNAS_ACT_SWISH, false><<<grid, NTHR_, smem, st>>>(A, Sc, cfg); } \
    } } while (0)

// TFNAS_WS: comma-separated subset of {expand,project,dc,dx} run by the persistent kernels; "0"/"none" disables, unset = all
static int g_ws_mask = -1;
int ws_enabled(int which) {
  if (g_ws_mask < 0) {
    const char* e = getenv("TFNAS_WS");
    if (!e) g_ws_mask = 15;
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

// TFNAS_WS_BULK=0 forces the per-thread cp.async loads (A/B debugging)
static bool ws_bulk_enabled() {
  static const bool on = !(getenv("TFNAS_WS_BULK") && strcmp(getenv("TFNAS_WS_BULK"), "0") == 0);
  return on;
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
  if (maxNc > 256 || !ws_fit(maxNc, WsExpandT::STAGE, WsExpandT::CF, cfg)) return false;
  const size_t smem = ws_smem_bytes(cfg);
  WsExpandArgs A{P, WA, x, bn1, UH};
  ProfScope ps("expand", 4.0 * P.P * P.ic + 4.0 * P.P * P.MC + 4.0 * P.MC * P.ic, 2.0 * P.P * (double)P.MC * P.ic, st);
  const int grid = min(Sc.n_items, sm_count());
  const bool bulk = ws_bulk_enabled() && ws_bulk_ok(P.HW, P.P);
  WS_LAUNCH1(k_ws_expand, bulk, DimEpi::NTHR);
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
  if (maxNc > 256 || !ws_fit(maxNc, WsProjectT<0>::STAGE, WsProjectT<0>::CF, cfg)) return false;
  const size_t smem = ws_smem_bytes(cfg);
  WsProjectArgs A{P, WA, D, bn2, seg, Zb, st3};
  ProfScope ps("project", 4.0 * P.Q * ((double)P.MC + (double)P.na * P.oc) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  const int grid = min(Sc.n_items, sm_count());
  const bool bulk = ws_bulk_enabled() && ws_bulk_ok(P.HWo, P.Q);
  WS_LAUNCH2(k_ws_project, P.act == TFNAS_ACT_RELU, bulk, DimProd::NTHR);
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
  if (maxNc > 256 || !ws_fit(maxNc, WsDcT<0>::STAGE, WsDcT<0>::CF, cfg)) return false;
  const size_t smem = ws_smem_bytes(cfg);
  WsDcArgs A{P, WA, G, Zb, dzc2, D, bn2, DC, dg, sD};
  ProfScope ps("dc", 4.0 * P.Q * ((double)P.oc * (1 + P.na) + 2.0 * P.MC) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  const int grid = min(Sc.n_items, sm_count());
  const bool bulk = ws_bulk_enabled() && ws_bulk_ok(P.HWo, P.Q);
  WS_LAUNCH2(k_ws_dc, P.act == TFNAS_ACT_RELU, bulk, DimEpi::NTHR);
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
  if (W.Nc > 256 || !ws_fit(W.Nc, WsDxT<0>::STAGE, WsDxT<0>::CF, cfg)) return false;
  WsSched Sc;
  memset(&Sc, 0, sizeof(Sc));
  Sc.tiles_m = tiles;
  Sc.inv_tiles = 1.f / (float)tiles;
  Sc.na = P.na;
  Sc.n_items = tiles * ksplit;
  if (ksplit > 1) cudaMemsetAsync(dx, 0, (size_t)P.P * P.ic * sizeof(float), st);
  const size_t smem = ws_smem_bytes(cfg);
  WsDxArgs A{P, W, CH, ksplit, DA, UH, dx, sU};
  ProfScope ps("dx", 4.0 * P.P * (2.0 * P.MC + P.ic) + 4.0 * P.MC * P.ic, 2.0 * P.P * (double)P.MC * P.ic, st);
  const int grid = min(Sc.n_items, sms);
  const bool bulk = ws_bulk_enabled() && ws_bulk_ok(P.HW, P.P);
  WS_LAUNCH2(k_ws_dx, P.act == TFNAS_ACT_RELU, bulk, DimProd::NTHR);
  return true;
}

// debug: enable / disable the phase trace of the persistent kernels (buf: device memory, n_ctas * WS_TRACE_SLOTS u64)
extern "C" int tfnas_debug_ws_trace(void* buf, int n_ctas) {
  unsigned long long* p = (unsigned long long*)buf;
  if (cudaMemcpyToSymbol(g_ws_trace, &p, sizeof(p)) != cudaSuccess) return TFNAS_E_CUDA;
  if (cudaMemcpyToSymbol(g_ws_trace_n, &n_ctas, sizeof(n_ctas)) != cudaSuccess) return TFNAS_E_CUDA;
  return TFNAS_OK;
}
