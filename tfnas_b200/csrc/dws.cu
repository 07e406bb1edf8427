// Register sliding-window depthwise kernels: stride 1 (14 of the 18 MixedOPs) and, further down, stride 2 (even planes).
//
// A lane owns VW consecutive columns of one (image, channel) plane and walks the plane's rows top to bottom:
//   * every input element is loaded from global memory exactly once (coalesced VW-wide loads, next row prefetched),
//     the prologue (activation / BN2-backward affine) is applied once per element,
//   * the horizontal halo comes from the neighbouring lanes by warp shuffles,
//   * the vertical reuse lives in a ring of KS per-output-row accumulators in registers, so an input row is
//     multiplied into the KS output rows it touches and a finished output row is stored once.
// No shared memory, no barriers.  A 32-lane warp is split into 32/Lpad segments of Lpad = pow2 >= W/VW lanes, one
// plane per segment (W = 56 -> 2 planes per warp, 28 / 14 / 7 -> 4 planes per warp).
//
//   forward  (F1b):  D  = DW (*) act(UH),  BN2 sums                       (models/layers.py:547-548)
//   backward (B3a):  DA = DW^T (*) dd,  dd = r2 (dd-hat - m1 - d-hat m2) applied on load; optional dDW
#include <stdlib.h>
#include "kernels.h"

#ifndef DWS_PF
#define DWS_PF 4        // input rows in flight per lane in the fixed-plane instantiations
#endif

struct DwsWork {
  int n;             // channel segments (candidates with this kernel size)
  int slot[4];
  int cstart[5];     // cumulative channel counts; cstart[n] = channels in the group
  float* gw[4];      // dDW destinations (weight-grad mode)
};

template <int VW>
__device__ __forceinline__ void ldv(float (&v)[VW], const float* __restrict__ p) {
  if (VW == 4) { const float4 t = *(const float4*)p; v[0] = t.x; v[1] = t.y; v[2 % VW] = t.z; v[3 % VW] = t.w; }
  else if (VW == 2) { const float2 t = *(const float2*)p; v[0] = t.x; v[1 % VW] = t.y; }
  else v[0] = *p;
}
template <int VW>
__device__ __forceinline__ void stv(float* __restrict__ p, const float (&v)[VW]) {
  if (VW == 4) *(float4*)p = make_float4(v[0], v[1], v[2 % VW], v[3 % VW]);
  else if (VW == 2) *(float2*)p = make_float2(v[0], v[1 % VW]);
  else *p = v[0];
}

// window[i] = value at column offset i - pad relative to this lane's first column; own values in the middle, the
// 2*pad halo values from the neighbouring lanes of the segment (zero outside the plane).  All 32 lanes execute.
template <int KS, int VW>
__device__ __forceinline__ void dws_window(float (&w)[VW + KS - 1], const float (&v)[VW], int lane, int li, int L) {
  constexpr int pad = KS / 2;
#pragma unroll
  for (int j = 0; j < VW; ++j) w[pad + j] = v[j];
#pragma unroll
  for (int d = 1; d <= pad; ++d) {
    {   // column offset -d : lane li - q, element e
      const int q = (d + VW - 1) / VW, e = q * VW - d;
      const float t = __shfl_sync(0xffffffffu, v[e], lane - q);
      w[pad - d] = (li - q >= 0) ? t : 0.f;
    }
    {   // column offset VW - 1 + d : lane li + q, element e
      const int o = VW - 1 + d, q = o / VW, e = o - q * VW;
      const float t = __shfl_sync(0xffffffffu, v[e], lane + q);
      w[pad + VW - 1 + d] = (li + q < L) ? t : 0.f;
    }
  }
}

// plane -> (image, channel) and the owning candidate
struct DwsPlane { bool ok; int n, e, cl, cst; };
__device__ __forceinline__ DwsPlane dws_plane(const Plan& P, const DwsWork& Wk, int lpl) {
  const int ppw = 32 >> lpl;
  const int lane = threadIdx.x & 31;
  const long long wg = ((long long)blockIdx.x * NT + threadIdx.x) >> 5;
  const long long plane = wg * ppw + (lane >> lpl);
  const int Cg = Wk.cstart[Wk.n];
  DwsPlane r;
  r.ok = plane < (long long)P.N * Cg;
  const int pl = r.ok ? (int)plane : 0;
  r.n = pl / Cg;
  const int g = pl - r.n * Cg;
  r.e = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < Wk.n && g >= Wk.cstart[i]) r.e = i;
  int c0 = 0, slot = Wk.slot[0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (r.e == i) { c0 = Wk.cstart[i]; slot = Wk.slot[i]; }
  r.cl = g - c0;
  r.cst = P.c[slot].coff + r.cl;
  r.e = slot;          // candidate slot from here on
  return r;
}

// HH > 0: the plane is HH x HH, known at compile time (the 14x14 and 7x7 stages, 11 of the 14 stride-1 MixedOPs): the row
// loop unrolls completely, every row / tap-row predicate and row offset folds to a constant -- the generic kernel spends
// ~40 % of its instructions on them (ISETP / BRA / IMAD), and these kernels are issue-bound.
template <int KS, int VW, int ACT, int HH>
__global__ void __launch_bounds__(NT) k_dws_fwd(Plan P, DwsWork Wk, int lpl_rt, const float* __restrict__ UH,
                                                 float* __restrict__ D, double* __restrict__ st2) {
  constexpr int pad = KS / 2;
  constexpr int LH = HH > 0 ? HH / VW : 0;
  const int lpl = HH > 0 ? (LH <= 1 ? 0 : LH <= 2 ? 1 : LH <= 4 ? 2 : LH <= 8 ? 3 : LH <= 16 ? 4 : 5) : lpl_rt;
  const int lane = threadIdx.x & 31, li = lane & ((1 << lpl) - 1);
  const int H = HH > 0 ? HH : P.H, W = HH > 0 ? HH : P.W, L = W / VW;
  const DwsPlane pl = dws_plane(P, Wk, lpl);
  const bool active = pl.ok && li < L;
  const Cand& cd = P.c[pl.e];
  float wr[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) wr[i] = active ? cd.dw[(size_t)pl.cl * KS * KS + i] : 0.f;
  const size_t base = ((size_t)pl.n * P.MC + pl.cst) * H * W + (size_t)li * VW;
  const float* src = UH + base;
  float* dst = D + base;
  float acc[KS][VW];
#pragma unroll
  for (int a = 0; a < KS; ++a)
#pragma unroll
    for (int j = 0; j < VW; ++j) acc[a][j] = 0.f;
  float s1 = 0.f, s2 = 0.f;
  // input rows in flight per lane: one ahead in the generic kernel; DWS_PF ahead when the rows are unrolled (the ring index
  // is then static) -- on the small planes these kernels are bound by bytes in flight, not by issue slots
  constexpr int PF = HH > 0 ? DWS_PF : 1;
  float cur[PF][VW];
#pragma unroll
  for (int a = 0; a < PF; ++a) {
#pragma unroll
    for (int j = 0; j < VW; ++j) cur[a][j] = 0.f;
    if (active && a < H) ldv<VW>(cur[a], src + (size_t)a * W);
  }
#pragma unroll (HH > 0 ? 16 : 1)
  for (int r0 = 0; r0 < H + pad; r0 += KS) {
#pragma unroll
    for (int u = 0; u < KS; ++u) {
      const int r = r0 + u;
      if (r < H + pad) {
        float v[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = (active && r < H) ? act_f<ACT>(cur[r % PF][j]) : 0.f;
        if (active && r + PF < H) ldv<VW>(cur[r % PF], src + (size_t)(r + PF) * W);     // row r + PF in flight
        float win[VW + KS - 1];
        dws_window<KS, VW>(win, v, lane, li, L);
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          const int o = r + pad - ky;                 // output row this input row feeds through tap row ky
          if (o >= 0 && o < H) {
            const int sl = (u + pad - ky + KS) % KS;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
#pragma unroll
              for (int j = 0; j < VW; ++j) acc[sl][j] += wr[ky * KS + kx] * win[j + kx];
          }
        }
        const int o = r - pad;                        // complete after input row r
        if (o >= 0) {
          const int sl = (u - pad + KS) % KS;
          if (active) stv<VW>(dst + (size_t)o * W, acc[sl]);
#pragma unroll
          for (int j = 0; j < VW; ++j) { s1 += acc[sl][j]; s2 += acc[sl][j] * acc[sl][j]; acc[sl][j] = 0.f; }
        }
      }
    }
  }
  // BN2 sums of the plane: reduce over the segment's lanes (inactive lanes hold zeros)
  for (int o = (1 << lpl) >> 1; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (pl.ok && li == 0) {
    atomicAdd(&st2[2 * pl.cst], (double)s1);
    atomicAdd(&st2[2 * pl.cst + 1], (double)s2);
  }
}

// Transposed depthwise (stride 1 => a correlation with the flipped filter) over dd = ca*ddh + cb*d + cc.
// WG: dDW[ky][kx] += sum dd[o][x] * a[o+ky-pad][x+kx-pad] with a = act(UH), accumulated per lane in registers.
template <int KS, int VW, int ACT, bool WG, int HH>
__global__ void __launch_bounds__(NT) k_dws_bwd(Plan P, DwsWork Wk, int lpl_rt, const float* __restrict__ DC,
                                                 const float* __restrict__ D, const float* __restrict__ bn2,
                                                 const double* __restrict__ sD, const float* __restrict__ UH,
                                                 float* __restrict__ DA) {
  constexpr int pad = KS / 2;
  constexpr int LH = HH > 0 ? HH / VW : 0;
  const int lpl = HH > 0 ? (LH <= 1 ? 0 : LH <= 2 ? 1 : LH <= 4 ? 2 : LH <= 8 ? 3 : LH <= 16 ? 4 : 5) : lpl_rt;
  const int lane = threadIdx.x & 31, li = lane & ((1 << lpl) - 1);
  const int H = HH > 0 ? HH : P.H, W = HH > 0 ? HH : P.W, L = W / VW;
  const DwsPlane pl = dws_plane(P, Wk, lpl);
  const bool active = pl.ok && li < L;
  const Cand& cd = P.c[pl.e];
  float wf[KS * KS];                     // flipped filter
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) wf[i] = active ? cd.dw[(size_t)pl.cl * KS * KS + (KS * KS - 1 - i)] : 0.f;
  float ca = 0.f, cb = 0.f, cc = 0.f;    // dd = r*(ddh - m1 - (d-mu)*r*m2) = ca*ddh + cb*d + cc
  if (active) {
    const double invQ = 1.0 / (double)P.Q;
    const float mu = bn2[pl.cst], r = bn2[P.MC + pl.cst];
    const float m1 = (float)(sD[2 * pl.cst] * invQ), m2 = (float)(sD[2 * pl.cst + 1] * invQ);
    ca = r; cb = -r * r * m2; cc = r * (mu * r * m2 - m1);
  }
  const size_t base = ((size_t)pl.n * P.MC + pl.cst) * H * W + (size_t)li * VW;
  const float* s0 = DC + base;
  const float* s1p = D + base;
  const float* ua = UH + base;
  float* dst = DA + base;
  float acc[KS][VW];
  float ar[WG ? KS : 1][VW];             // ring of activated UH rows o mod KS (weight-grad mode)
  float gacc[WG ? KS * KS : 1];
#pragma unroll
  for (int a = 0; a < KS; ++a)
#pragma unroll
    for (int j = 0; j < VW; ++j) acc[a][j] = 0.f;
  if (WG) {
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) gacc[i] = 0.f;
#pragma unroll
    for (int a = 0; a < KS; ++a)
#pragma unroll
      for (int j = 0; j < VW; ++j) ar[a][j] = 0.f;
    // rows 0 .. pad-1 of a are needed from the first iteration on
#pragma unroll
    for (int a = 0; a < pad; ++a) {
      if (active && a < H) {
        float t[VW];
        ldv<VW>(t, ua + (size_t)a * W);
#pragma unroll
        for (int j = 0; j < VW; ++j) ar[a][j] = act_f<ACT>(t[j]);
      }
    }
  }
  constexpr int PF = HH > 0 ? DWS_PF : 1;       // rows of DC and D in flight per lane (see k_dws_fwd)
  float c0[PF][VW], c1[PF][VW], an[VW];
#pragma unroll
  for (int j = 0; j < VW; ++j) an[j] = 0.f;
#pragma unroll
  for (int a = 0; a < PF; ++a) {
#pragma unroll
    for (int j = 0; j < VW; ++j) c0[a][j] = c1[a][j] = 0.f;
    if (active && a < H) { ldv<VW>(c0[a], s0 + (size_t)a * W); ldv<VW>(c1[a], s1p + (size_t)a * W); }
  }
  if (WG && active && pad < H) ldv<VW>(an, ua + (size_t)pad * W);       // a row `pad`, committed at iteration 0
#pragma unroll (HH > 0 ? 16 : 1)
  for (int r0 = 0; r0 < H + pad; r0 += KS) {
#pragma unroll
    for (int u = 0; u < KS; ++u) {
      const int r = r0 + u;
      if (r < H + pad) {
        float v[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = (active && r < H) ? fmaf(ca, c0[r % PF][j], fmaf(cb, c1[r % PF][j], cc)) : 0.f;
        if (active && r + PF < H) {
          ldv<VW>(c0[r % PF], s0 + (size_t)(r + PF) * W);
          ldv<VW>(c1[r % PF], s1p + (size_t)(r + PF) * W);
        }
        if (WG) {
          // commit a row r+pad (loaded one iteration ago) into its ring slot, prefetch row r+pad+1
          const int sl = (u + pad) % KS;
#pragma unroll
          for (int j = 0; j < VW; ++j) ar[sl][j] = (active && r + pad < H) ? act_f<ACT>(an[j]) : 0.f;
          if (active && r + pad + 1 < H) ldv<VW>(an, ua + (size_t)(r + pad + 1) * W);
        }
        float win[VW + KS - 1];
        dws_window<KS, VW>(win, v, lane, li, L);
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          const int o = r + pad - ky;                 // da row fed by dd row r through flipped tap row ky
          if (o >= 0 && o < H) {
            const int sl = (u + pad - ky + KS) % KS;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
              for (int j = 0; j < VW; ++j) acc[sl][j] += wf[ky * KS + kx] * win[j + kx];
              if (WG) {
                float t = 0.f;
#pragma unroll
                for (int j = 0; j < VW; ++j) t += win[j + kx] * ar[sl][j];
                gacc[KS * KS - 1 - (ky * KS + kx)] += t;
              }
            }
          }
        }
        const int o = r - pad;
        if (o >= 0) {
          const int sl = (u - pad + KS) % KS;
          if (active) stv<VW>(dst + (size_t)o * W, acc[sl]);
#pragma unroll
          for (int j = 0; j < VW; ++j) acc[sl][j] = 0.f;
        }
      }
    }
  }
  if (WG) {
    float* gp = nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < Wk.n && Wk.slot[i] == pl.e) gp = Wk.gw[i];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) {
      float t = gacc[i];
      for (int o = (1 << lpl) >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (pl.ok && li == 0 && gp) atomicAdd(&gp[(size_t)pl.cl * KS * KS + i], t);
    }
  }
}


// ---- stride 2 ------------------------------------------------------------------------------------------------------
// Same idea on the stride-2 MixedOPs (even H, W): a lane owns VW consecutive INPUT columns (VO = VW/2 output columns)
// of one plane.  Forward walks the input rows: row r feeds output row o through tap row ky = r + pad - 2o, so at most
// RS = (KS+1)/2 output rows are live; rows are processed in unrolled groups of 2*RS so that every ring slot and the
// parity tests are compile-time.  Backward walks the dd (output-plane) rows: row oy feeds the KS input-plane rows
// 2oy - pad .. 2oy + pad and completes two of them; ring of KS+1 accumulator rows, groups of KS+1 dd rows.
// Input rows are prefetched a whole group ahead (static ring index).
template <int KS, int VW, int ACT>
__global__ void __launch_bounds__(NT) k_dws2_fwd(Plan P, DwsWork Wk, int lpl, const float* __restrict__ UH,
                                                  float* __restrict__ D, double* __restrict__ st2) {
  constexpr int pad = KS / 2, VO = VW / 2, RS = (KS + 1) / 2, G = 2 * RS;
  constexpr int PF = KS == 5 ? G / 2 : G;              // input rows in flight per lane (5x5: 97 registers with a full group)
  const int lane = threadIdx.x & 31, li = lane & ((1 << lpl) - 1);
  const int H = P.H, W = P.W, Ho = P.Ho, Wo = P.Wo, L = W / VW;
  const DwsPlane pl = dws_plane(P, Wk, lpl);
  const bool active = pl.ok && li < L;
  const Cand& cd = P.c[pl.e];
  float wr[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) wr[i] = active ? cd.dw[(size_t)pl.cl * KS * KS + i] : 0.f;
  const size_t plane = (size_t)pl.n * P.MC + pl.cst;
  const float* src = UH + plane * H * W + (size_t)li * VW;
  float* dst = D + plane * Ho * Wo + (size_t)li * VO;
  float acc[RS][VO];
#pragma unroll
  for (int a = 0; a < RS; ++a)
#pragma unroll
    for (int j = 0; j < VO; ++j) acc[a][j] = 0.f;
  float s1 = 0.f, s2 = 0.f;
  float cur[PF][VW];
#pragma unroll
  for (int a = 0; a < PF; ++a) {
#pragma unroll
    for (int j = 0; j < VW; ++j) cur[a][j] = 0.f;
    if (active && a < H) ldv<VW>(cur[a], src + (size_t)a * W);
  }
  for (int r0 = 0; r0 < H + pad; r0 += G) {
    const int ob = r0 >> 1;                            // multiple of RS: ring slots below are compile-time
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int r = r0 + u;
      if (r < H + pad) {
        float v[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = (active && r < H) ? act_f<ACT>(cur[u % PF][j]) : 0.f;
        if (active && r + PF < H) ldv<VW>(cur[u % PF], src + (size_t)(r + PF) * W);
        float win[VW + KS - 1];
        dws_window<KS, VW>(win, v, lane, li, L);
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          constexpr int dummy = 0; (void)dummy;
          const int t = u + pad - ky;                  // 2 * (o - ob)
          if ((t & 1) == 0) {
            const int o = ob + t / 2;
            if (o >= 0 && o < Ho) {
              const int sl = ((t / 2) % RS + RS) % RS;
#pragma unroll
              for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                for (int j = 0; j < VO; ++j) acc[sl][j] += wr[ky * KS + kx] * win[2 * j + kx];
            }
          }
        }
        if (((u - pad) & 1) == 0) {                    // output row (r - pad) / 2 is complete after input row r
          const int o = ob + (u - pad) / 2;
          if (o >= 0 && o < Ho) {
            const int sl = (((u - pad) / 2) % RS + RS) % RS;
            if (active) stv<VO>(dst + (size_t)o * Wo, acc[sl]);
#pragma unroll
            for (int j = 0; j < VO; ++j) { s1 += acc[sl][j]; s2 += acc[sl][j] * acc[sl][j]; acc[sl][j] = 0.f; }
          }
        }
      }
    }
  }
  for (int o = (1 << lpl) >> 1; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (pl.ok && li == 0) {
    atomicAdd(&st2[2 * pl.cst], (double)s1);
    atomicAdd(&st2[2 * pl.cst + 1], (double)s2);
  }
}

// DA[iy][ix] = sum w[ky][kx] dd[oy][ox] over 2oy + ky - pad = iy, 2ox + kx - pad = ix;  dd = ca*ddh + cb*d + cc on load
template <int KS, int VW, int ACT>
__global__ void __launch_bounds__(NT, 3) k_dws2_bwd(Plan P, DwsWork Wk, int lpl, const float* __restrict__ DC,
                                                  const float* __restrict__ D, const float* __restrict__ bn2,
                                                  const double* __restrict__ sD, float* __restrict__ DA) {
  constexpr int pad = KS / 2, VO = VW / 2, RB = KS + 1;
  constexpr int PF = KS == 5 ? RB / 2 : RB;            // dd rows in flight per lane
  const int lane = threadIdx.x & 31, li = lane & ((1 << lpl) - 1);
  const int H = P.H, W = P.W, Ho = P.Ho, Wo = P.Wo, L = W / VW;
  const DwsPlane pl = dws_plane(P, Wk, lpl);
  const bool active = pl.ok && li < L;
  const Cand& cd = P.c[pl.e];
  float wr[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) wr[i] = active ? cd.dw[(size_t)pl.cl * KS * KS + i] : 0.f;
  float ca = 0.f, cb = 0.f, cc = 0.f;
  if (active) {
    const double invQ = 1.0 / (double)P.Q;
    const float mu = bn2[pl.cst], r = bn2[P.MC + pl.cst];
    const float m1 = (float)(sD[2 * pl.cst] * invQ), m2 = (float)(sD[2 * pl.cst + 1] * invQ);
    ca = r; cb = -r * r * m2; cc = r * (mu * r * m2 - m1);
  }
  const size_t plane = (size_t)pl.n * P.MC + pl.cst;
  const float* s0 = DC + plane * Ho * Wo + (size_t)li * VO;
  const float* s1p = D + plane * Ho * Wo + (size_t)li * VO;
  float* dst = DA + plane * H * W + (size_t)li * VW;
  float acc[RB][VW];
#pragma unroll
  for (int a = 0; a < RB; ++a)
#pragma unroll
    for (int j = 0; j < VW; ++j) acc[a][j] = 0.f;
  float c0[PF][VO], c1[PF][VO];
#pragma unroll
  for (int a = 0; a < PF; ++a) {
#pragma unroll
    for (int j = 0; j < VO; ++j) c0[a][j] = c1[a][j] = 0.f;
    if (active && a < Ho) { ldv<VO>(c0[a], s0 + (size_t)a * Wo); ldv<VO>(c1[a], s1p + (size_t)a * Wo); }
  }
  for (int g0 = 0; g0 < Ho + 1; g0 += RB) {            // 2 * g0 is a multiple of RB: ring slots are compile-time
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int oy = g0 + u;
      if (oy < Ho + 1) {                                // oy == Ho: a zero row that flushes the last `pad` input rows
        float v[VO];
#pragma unroll
        for (int j = 0; j < VO; ++j) v[j] = (active && oy < Ho) ? fmaf(ca, c0[u % PF][j], fmaf(cb, c1[u % PF][j], cc)) : 0.f;
        if (active && oy + PF < Ho) {
          ldv<VO>(c0[u % PF], s0 + (size_t)(oy + PF) * Wo);
          ldv<VO>(c1[u % PF], s1p + (size_t)(oy + PF) * Wo);
        }
        float win[VO + 2];                              // win[i] = dd column (first own column) + i - 1
        dws_window<3, VO>(win, v, lane, li, L);
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          const int iy = 2 * oy + ky - pad;
          if (iy >= 0 && iy < H) {
            const int sl = ((2 * u + ky - pad) % RB + RB) % RB;
#pragma unroll
            for (int j = 0; j < VW; ++j)
#pragma unroll
              for (int kx = 0; kx < KS; ++kx) {
                const int t = j + pad - kx;             // 2 * (ox - first own dd column)
                if ((t & 1) == 0) acc[sl][j] += wr[ky * KS + kx] * win[t / 2 + 1];
              }
          }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {                   // input rows 2oy - pad, 2oy - pad + 1 are complete
          const int iy = 2 * oy - pad + e;
          if (iy >= 0 && iy < H) {
            const int sl = ((2 * u - pad + e) % RB + RB) % RB;
            if (active) stv<VW>(dst + (size_t)iy * W, acc[sl]);
#pragma unroll
            for (int j = 0; j < VW; ++j) acc[sl][j] = 0.f;
          }
        }
      }
    }
  }
}


// Depthwise weight gradient for stride 2 (sampled passes): dDW[ky][kx] += sum dd[oy][ox] * a[2oy+ky-pad][2ox+kx-pad],
// a = act(UH).  Walks the INPUT rows like the forward kernel (activation and horizontal window once per row); the dd
// rows an input row meets (oy = (r + pad - ky) / 2) sit in a ring of RS rows of the lane's own VO columns, each
// committed when first needed from loads issued two input rows earlier.  Runs beside k_dws2_bwd (weight-gradient
// side stream): it only reads DC / D / UH.
template <int KS, int VW, int ACT>
__global__ void __launch_bounds__(NT) k_dws2_wg(Plan P, DwsWork Wk, int lpl, const float* __restrict__ DC,
                                                 const float* __restrict__ D, const float* __restrict__ bn2,
                                                 const double* __restrict__ sD, const float* __restrict__ UH) {
  constexpr int pad = KS / 2, VO = VW / 2, RS = (KS + 1) / 2, G = 2 * RS;
  const int lane = threadIdx.x & 31, li = lane & ((1 << lpl) - 1);
  const int H = P.H, W = P.W, Ho = P.Ho, Wo = P.Wo, L = W / VW;
  const DwsPlane pl = dws_plane(P, Wk, lpl);
  const bool active = pl.ok && li < L;
  float ca = 0.f, cb = 0.f, cc = 0.f;
  if (active) {
    const double invQ = 1.0 / (double)P.Q;
    const float mu = bn2[pl.cst], r = bn2[P.MC + pl.cst];
    const float m1 = (float)(sD[2 * pl.cst] * invQ), m2 = (float)(sD[2 * pl.cst + 1] * invQ);
    ca = r; cb = -r * r * m2; cc = r * (mu * r * m2 - m1);
  }
  const size_t plane = (size_t)pl.n * P.MC + pl.cst;
  const float* src = UH + plane * H * W + (size_t)li * VW;
  const float* s0 = DC + plane * Ho * Wo + (size_t)li * VO;
  const float* s1p = D + plane * Ho * Wo + (size_t)li * VO;
  float gacc[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) gacc[i] = 0.f;
  float ddr[RS][VO], n0[VO], n1[VO];
#pragma unroll
  for (int a = 0; a < RS; ++a)
#pragma unroll
    for (int j = 0; j < VO; ++j) ddr[a][j] = 0.f;
#pragma unroll
  for (int j = 0; j < VO; ++j) n0[j] = n1[j] = 0.f;
  if (active) {                                        // dd row 0 is needed from input row 0 on; stage row 1
    ldv<VO>(n0, s0); ldv<VO>(n1, s1p);
#pragma unroll
    for (int j = 0; j < VO; ++j) ddr[0][j] = fmaf(ca, n0[j], fmaf(cb, n1[j], cc));
    if (1 < Ho) { ldv<VO>(n0, s0 + Wo); ldv<VO>(n1, s1p + Wo); }
  }
  float cur[G][VW];
#pragma unroll
  for (int a = 0; a < G; ++a) {
#pragma unroll
    for (int j = 0; j < VW; ++j) cur[a][j] = 0.f;
    if (active && a < H) ldv<VW>(cur[a], src + (size_t)a * W);
  }
  for (int r0 = 0; r0 < H; r0 += G) {
    const int ob = r0 >> 1;                            // multiple of RS
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int r = r0 + u;
      if (r < H) {
        if (((u + pad) & 1) == 0) {                    // dd row (r + pad) / 2 >= 1 enters the ring at input row r
          const int on = ob + (u + pad) / 2;
          const int sl = (((u + pad) / 2) % RS + RS) % RS;
          if (on >= 1) {
#pragma unroll
            for (int j = 0; j < VO; ++j) ddr[sl][j] = (active && on < Ho) ? fmaf(ca, n0[j], fmaf(cb, n1[j], cc)) : 0.f;
            if (active && on + 1 < Ho) { ldv<VO>(n0, s0 + (size_t)(on + 1) * Wo); ldv<VO>(n1, s1p + (size_t)(on + 1) * Wo); }
          }
        }
        float v[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = active ? act_f<ACT>(cur[u][j]) : 0.f;
        if (active && r + G < H) ldv<VW>(cur[u], src + (size_t)(r + G) * W);
        float win[VW + KS - 1];
        dws_window<KS, VW>(win, v, lane, li, L);
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          const int t = u + pad - ky;
          if ((t & 1) == 0) {
            const int o = ob + t / 2;
            if (o >= 0 && o < Ho) {
              const int sl = ((t / 2) % RS + RS) % RS;
#pragma unroll
              for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                for (int j = 0; j < VO; ++j) gacc[ky * KS + kx] += ddr[sl][j] * win[2 * j + kx];
            }
          }
        }
      }
    }
  }
  float* gp = nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < Wk.n && Wk.slot[i] == pl.e) gp = Wk.gw[i];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) {
    float t = gacc[i];
    for (int o = (1 << lpl) >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (pl.ok && li == 0 && gp) atomicAdd(&gp[(size_t)pl.cl * KS * KS + i], t);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
static bool dws_setup(const Plan& P, int KS, const TfnasCandPtrs* dweights, DwsWork& w, int& VW, int& lpl, double& mck) {
  w.n = 0;
  w.cstart[0] = 0;
  mck = 0;
  for (int s = 0; s < P.na; ++s) {
    if (P.c[s].k != KS) continue;
    if (w.n == 4) return false;
    w.slot[w.n] = s;
    w.cstart[w.n + 1] = w.cstart[w.n] + P.c[s].mc;
    w.gw[w.n] = dweights ? dweights[P.c[s].id].dw : nullptr;
    mck += P.c[s].mc;
    ++w.n;
  }
  for (int i = w.n; i < 4; ++i) { w.slot[i] = -1; w.cstart[i + 1] = w.cstart[w.n]; w.gw[i] = nullptr; }
  VW = (P.W & 3) == 0 ? 4 : (P.W & 1) == 0 ? 2 : 1;
  const int L = P.W / VW;
  lpl = 0;
  while ((1 << lpl) < L) ++lpl;
  return L <= 32;
}

// TFNAS_DWS_S2=0 sends the stride-2 MixedOPs back to the shared-memory tile kernels of fwd.cu / bwd.cu (A/B timing)
bool dws_supported(const Plan& P) {
  if (P.stride == 2) {
    static const bool on = [] { const char* e = getenv("TFNAS_DWS_S2"); return !(e && e[0] == '0'); }();
    if (!on || (P.H & 1) || (P.W & 1)) return false;
  } else if (P.stride != 1) {
    return false;
  }
  const int VW = (P.W & 3) == 0 ? 4 : (P.W & 1) == 0 ? 2 : 1;
  int n3 = 0, n5 = 0;
  for (int s = 0; s < P.na; ++s) (P.c[s].k == 3 ? n3 : n5)++;
  return P.W / VW <= 32 && n3 <= 4 && n5 <= 4 && (long long)P.N * P.MC < (1LL << 23);
}

template <int KS, int VW, int HH>
static void dws_fwd_launch(const Plan& P, const DwsWork& w, int lpl, dim3 grid, const float* UH, float* D, double* st2,
                           cudaStream_t st) {
  if (P.act == TFNAS_ACT_RELU) k_dws_fwd<KS, VW, TFNAS_ACT_RELU, HH><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
  else k_dws_fwd<KS, VW, TFNAS_ACT_SWISH, HH><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
}

// compile-time plane sizes (TFNAS_DWS_FIXED=0 keeps every shape on the generic kernels, for A/B timing)
static int dws_fixed_plane(const Plan& P, int VW) {
  static const bool on = [] { const char* e = getenv("TFNAS_DWS_FIXED"); return !(e && e[0] == '0'); }();
  if (!on || P.H != P.W) return 0;
  if (P.H == 14 && VW == 2) return 14;
  if (P.H == 7 && VW == 1) return 7;
  static const bool no28 = [] { const char* e = getenv("TFNAS_DWS_NO28"); return e && e[0] == '1'; }();
  if (P.H == 28 && VW == 4 && !no28) return 28;
  return 0;
}

template <int KS>
static void dws_fwd_ks(const Plan& P, const float* UH, float* D, double* st2, cudaStream_t st) {
  DwsWork w;
  int VW, lpl;
  double mck;
  if (!dws_setup(P, KS, nullptr, w, VW, lpl, mck) || !w.n) return;
  const long long planes = (long long)P.N * w.cstart[w.n];
  const long long warps = (planes + (32 >> lpl) - 1) / (32 >> lpl);
  dim3 grid((unsigned)((warps + NT / 32 - 1) / (NT / 32)));
  ProfScope ps(KS == 3 ? "dw_fwd_k3" : "dw_fwd_k5", 4.0 * mck * ((double)P.P + P.Q), 2.0 * KS * KS * mck * P.Q, st);
  if (P.stride == 2) {
    const bool relu = P.act == TFNAS_ACT_RELU;
    if (VW == 4) {
      if (relu) k_dws2_fwd<KS, 4, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
      else k_dws2_fwd<KS, 4, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
    } else {
      if (relu) k_dws2_fwd<KS, 2, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
      else k_dws2_fwd<KS, 2, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
    }
    return;
  }
  const int hh = dws_fixed_plane(P, VW);
  if (hh == 28) dws_fwd_launch<KS, 4, 28>(P, w, lpl, grid, UH, D, st2, st);
  else if (hh == 14) dws_fwd_launch<KS, 2, 14>(P, w, lpl, grid, UH, D, st2, st);
  else if (hh == 7) dws_fwd_launch<KS, 1, 7>(P, w, lpl, grid, UH, D, st2, st);
  else if (VW == 4) dws_fwd_launch<KS, 4, 0>(P, w, lpl, grid, UH, D, st2, st);
  else if (VW == 2) dws_fwd_launch<KS, 2, 0>(P, w, lpl, grid, UH, D, st2, st);
  else dws_fwd_launch<KS, 1, 0>(P, w, lpl, grid, UH, D, st2, st);
}

void launch_dws_fwd(const Plan& P, const float* UH, float* D, double* st2, cudaStream_t st) {
  dws_fwd_ks<3>(P, UH, D, st2, st);
  dws_fwd_ks<5>(P, UH, D, st2, st);
}

template <int KS, int VW, bool WG, int HH>
static void dws_bwd_launch(const Plan& P, const DwsWork& w, int lpl, dim3 grid, const float* DC, const float* D,
                           const float* bn2, const double* sD, const float* UH, float* DA, cudaStream_t st) {
  if (P.act == TFNAS_ACT_RELU) k_dws_bwd<KS, VW, TFNAS_ACT_RELU, WG, HH><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, UH, DA);
  else k_dws_bwd<KS, VW, TFNAS_ACT_SWISH, WG, HH><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, UH, DA);
}

template <int KS>
static void dws_bwd_ks(const Plan& P, const float* DC, const float* D, const float* bn2, const double* sD,
                       const float* UH, float* DA, const TfnasCandPtrs* dweights, cudaStream_t st, cudaStream_t wg_st) {
  DwsWork w;
  int VW, lpl;
  double mck;
  if (!dws_setup(P, KS, dweights, w, VW, lpl, mck) || !w.n) return;
  const long long planes = (long long)P.N * w.cstart[w.n];
  const long long warps = (planes + (32 >> lpl) - 1) / (32 >> lpl);
  dim3 grid((unsigned)((warps + NT / 32 - 1) / (NT / 32)));
  ProfScope ps(KS == 3 ? "dw_bwd_k3" : "dw_bwd_k5", 4.0 * mck * (2.0 * P.Q + (dweights ? 2.0 : 1.0) * P.P),
               2.0 * KS * KS * mck * P.Q * (dweights ? 2 : 1), st);
  if (P.stride == 2) {                                 // DA on the caller's stream, dDW beside it on wg_st
    const bool relu = P.act == TFNAS_ACT_RELU;
    if (VW == 4) {
      if (relu) k_dws2_bwd<KS, 4, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, DA);
      else k_dws2_bwd<KS, 4, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, DA);
    } else {
      if (relu) k_dws2_bwd<KS, 2, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, DA);
      else k_dws2_bwd<KS, 2, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, DA);
    }
    if (dweights) {
      count_launch(1);
      if (VW == 4) {
        if (relu) k_dws2_wg<KS, 4, TFNAS_ACT_RELU><<<grid, NT, 0, wg_st>>>(P, w, lpl, DC, D, bn2, sD, UH);
        else k_dws2_wg<KS, 4, TFNAS_ACT_SWISH><<<grid, NT, 0, wg_st>>>(P, w, lpl, DC, D, bn2, sD, UH);
      } else {
        if (relu) k_dws2_wg<KS, 2, TFNAS_ACT_RELU><<<grid, NT, 0, wg_st>>>(P, w, lpl, DC, D, bn2, sD, UH);
        else k_dws2_wg<KS, 2, TFNAS_ACT_SWISH><<<grid, NT, 0, wg_st>>>(P, w, lpl, DC, D, bn2, sD, UH);
      }
    }
    return;
  }
#define DWS_B(VW_, HH_) do { \
    if (dweights) dws_bwd_launch<KS, VW_, true, HH_>(P, w, lpl, grid, DC, D, bn2, sD, UH, DA, st); \
    else dws_bwd_launch<KS, VW_, false, HH_>(P, w, lpl, grid, DC, D, bn2, sD, UH, DA, st); } while (0)
  // with the weight gradient in the kernel (sampled passes) the unrolled 28 / 14 planes need 125-255 registers and measured
  // slower than the generic kernel (28x28: 0.178 vs 0.129 ms); only the 7x7 plane keeps its fixed instantiation there
  int hh = dws_fixed_plane(P, VW);
  if (dweights && hh != 7) hh = 0;
  if (hh == 28) dws_bwd_launch<KS, 4, false, 28>(P, w, lpl, grid, DC, D, bn2, sD, UH, DA, st);
  else if (hh == 14) dws_bwd_launch<KS, 2, false, 14>(P, w, lpl, grid, DC, D, bn2, sD, UH, DA, st);
  else if (hh == 7) DWS_B(1, 7);
  else if (VW == 4) DWS_B(4, 0);
  else if (VW == 2) DWS_B(2, 0);
  else DWS_B(1, 0);
#undef DWS_B
}

// dweights != nullptr: the dDW buffers must have been zeroed by the caller (atomic accumulation)
// wg_st: stream of the stride-2 depthwise weight-gradient kernel (ordered after DC / sD are final; may equal st)
void launch_dws_bwd(const Plan& P, const float* DC, const float* D, const float* bn2, const double* sD, const float* UH,
                    float* DA, const TfnasCandPtrs* dweights, cudaStream_t st, cudaStream_t wg_st) {
  dws_bwd_ks<3>(P, DC, D, bn2, sD, UH, DA, dweights, st, wg_st);
  dws_bwd_ks<5>(P, DC, D, bn2, sD, UH, DA, dweights, st, wg_st);
}
