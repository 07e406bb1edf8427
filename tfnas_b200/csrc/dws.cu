// Register sliding-window depthwise kernels for the stride-1 MixedOPs (14 of the 18 blocks).
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
#include "kernels.h"

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

template <int KS, int VW, int ACT>
__global__ void __launch_bounds__(NT) k_dws_fwd(Plan P, DwsWork Wk, int lpl, const float* __restrict__ UH,
                                                 float* __restrict__ D, double* __restrict__ st2) {
  constexpr int pad = KS / 2;
  const int lane = threadIdx.x & 31, li = lane & ((1 << lpl) - 1);
  const int H = P.H, W = P.W, L = W / VW;
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
  float cur[VW];
#pragma unroll
  for (int j = 0; j < VW; ++j) cur[j] = 0.f;
  if (active) ldv<VW>(cur, src);
  for (int r0 = 0; r0 < H + pad; r0 += KS) {
#pragma unroll
    for (int u = 0; u < KS; ++u) {
      const int r = r0 + u;
      if (r < H + pad) {
        float v[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = (active && r < H) ? act_f<ACT>(cur[j]) : 0.f;
        if (active && r + 1 < H) ldv<VW>(cur, src + (size_t)(r + 1) * W);     // next row in flight
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
template <int KS, int VW, int ACT, bool WG>
__global__ void __launch_bounds__(NT) k_dws_bwd(Plan P, DwsWork Wk, int lpl, const float* __restrict__ DC,
                                                 const float* __restrict__ D, const float* __restrict__ bn2,
                                                 const double* __restrict__ sD, const float* __restrict__ UH,
                                                 float* __restrict__ DA) {
  constexpr int pad = KS / 2;
  const int lane = threadIdx.x & 31, li = lane & ((1 << lpl) - 1);
  const int H = P.H, W = P.W, L = W / VW;
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
  float c0[VW], c1[VW], an[VW];
#pragma unroll
  for (int j = 0; j < VW; ++j) c0[j] = c1[j] = an[j] = 0.f;
  if (active) { ldv<VW>(c0, s0); ldv<VW>(c1, s1p); }
  if (WG && active && pad < H) ldv<VW>(an, ua + (size_t)pad * W);       // a row `pad`, committed at iteration 0
  for (int r0 = 0; r0 < H + pad; r0 += KS) {
#pragma unroll
    for (int u = 0; u < KS; ++u) {
      const int r = r0 + u;
      if (r < H + pad) {
        float v[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) v[j] = (active && r < H) ? fmaf(ca, c0[j], fmaf(cb, c1[j], cc)) : 0.f;
        if (active && r + 1 < H) { ldv<VW>(c0, s0 + (size_t)(r + 1) * W); ldv<VW>(c1, s1p + (size_t)(r + 1) * W); }
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

bool dws_supported(const Plan& P) {
  if (P.stride != 1) return false;
  const int VW = (P.W & 3) == 0 ? 4 : (P.W & 1) == 0 ? 2 : 1;
  int n3 = 0, n5 = 0;
  for (int s = 0; s < P.na; ++s) (P.c[s].k == 3 ? n3 : n5)++;
  return P.W / VW <= 32 && n3 <= 4 && n5 <= 4 && (long long)P.N * P.MC < (1LL << 23);
}

template <int KS, int VW>
static void dws_fwd_launch(const Plan& P, const DwsWork& w, int lpl, dim3 grid, const float* UH, float* D, double* st2,
                           cudaStream_t st) {
  if (P.act == TFNAS_ACT_RELU) k_dws_fwd<KS, VW, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
  else k_dws_fwd<KS, VW, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, w, lpl, UH, D, st2);
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
  if (VW == 4) dws_fwd_launch<KS, 4>(P, w, lpl, grid, UH, D, st2, st);
  else if (VW == 2) dws_fwd_launch<KS, 2>(P, w, lpl, grid, UH, D, st2, st);
  else dws_fwd_launch<KS, 1>(P, w, lpl, grid, UH, D, st2, st);
}

void launch_dws_fwd(const Plan& P, const float* UH, float* D, double* st2, cudaStream_t st) {
  dws_fwd_ks<3>(P, UH, D, st2, st);
  dws_fwd_ks<5>(P, UH, D, st2, st);
}

template <int KS, int VW, bool WG>
static void dws_bwd_launch(const Plan& P, const DwsWork& w, int lpl, dim3 grid, const float* DC, const float* D,
                           const float* bn2, const double* sD, const float* UH, float* DA, cudaStream_t st) {
  if (P.act == TFNAS_ACT_RELU) k_dws_bwd<KS, VW, TFNAS_ACT_RELU, WG><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, UH, DA);
  else k_dws_bwd<KS, VW, TFNAS_ACT_SWISH, WG><<<grid, NT, 0, st>>>(P, w, lpl, DC, D, bn2, sD, UH, DA);
}

template <int KS>
static void dws_bwd_ks(const Plan& P, const float* DC, const float* D, const float* bn2, const double* sD,
                       const float* UH, float* DA, const TfnasCandPtrs* dweights, cudaStream_t st) {
  DwsWork w;
  int VW, lpl;
  double mck;
  if (!dws_setup(P, KS, dweights, w, VW, lpl, mck) || !w.n) return;
  const long long planes = (long long)P.N * w.cstart[w.n];
  const long long warps = (planes + (32 >> lpl) - 1) / (32 >> lpl);
  dim3 grid((unsigned)((warps + NT / 32 - 1) / (NT / 32)));
  ProfScope ps(KS == 3 ? "dw_bwd_k3" : "dw_bwd_k5", 4.0 * mck * (2.0 * P.Q + (dweights ? 2.0 : 1.0) * P.P),
               2.0 * KS * KS * mck * P.Q * (dweights ? 2 : 1), st);
#define DWS_B(VW_) do { \
    if (dweights) dws_bwd_launch<KS, VW_, true>(P, w, lpl, grid, DC, D, bn2, sD, UH, DA, st); \
    else dws_bwd_launch<KS, VW_, false>(P, w, lpl, grid, DC, D, bn2, sD, UH, DA, st); } while (0)
  if (VW == 4) DWS_B(4);
  else if (VW == 2) DWS_B(2);
  else DWS_B(1);
#undef DWS_B
}

// dweights != nullptr: the dDW buffers must have been zeroed by the caller (atomic accumulation)
void launch_dws_bwd(const Plan& P, const float* DC, const float* D, const float* bn2, const double* sD, const float* UH,
                    float* DA, const TfnasCandPtrs* dweights, cudaStream_t st) {
  dws_bwd_ks<3>(P, DC, D, bn2, sD, UH, DA, dweights, st);
  dws_bwd_ks<5>(P, DC, D, bn2, sD, UH, DA, dweights, st);
}
