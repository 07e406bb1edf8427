// Backward phases B1..B4 of one MixedOP call (see DESIGN.md "Kernels"; math: SURVEY.md appendix C,
// validated on CPU in oracle/fused_math.py).
//
//  B1  k_b1, k_b2prep        BN3-backward sums  sG[c] = sum G, sGY_i[c] = sum G*yhat_i ; dL/dw_i
//  B2  k_dc<TC,ACT>          dz on load from (G, Z); dc = W3^T dz (K = oc); non-SE: dd-hat = dc*act'(d-hat),
//                            BN2-backward sums; SE: dg[n,c] = sum_hw dc*b
//      k_se_bwd<ACT>         SE FCs backward -> dp ;  k_b2b<ACT>: SE candidates' dd-hat + BN2-backward sums
//  B3a k_dw_bwd<KS,S,ACT,WG> dd on load from (DC, D); transposed depthwise -> DA (+ depthwise weight grad)
//  B3b k_dx<TK,ACT>          du-hat = DA*act'(UH) on load, BN1-backward sums, dx_main = sum_i W1_i^T (r1 du-hat)
//  B4  k_b4prep, k_dxfin<TK> BN1 backward folded into an ic x ic correction: dx = dx_main - cvec - Mm (x - mu_x) (+G)
//      k_alpha_grad          softmax/Gumbel Jacobian -> dL/dlog_alpha
//  weight-grad mode (sampled w-step): k_wgrad<MODE>, k_w1fin, k_se_wgrad
#include "kernels.h"
#include "pw.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string.h>

// ----------------------------------------------------------------------------------------------
// B1
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_b1(Plan P, const float* __restrict__ G, const float* __restrict__ Zb,
                                            const float* __restrict__ bn3, double* __restrict__ sG,
                                            double* __restrict__ sGY) {
  const int c = blockIdx.x, oc = P.oc, na = P.na, HWo = P.HWo;
  const int nsplit = gridDim.y;
  // split boundaries on multiples of 4 so the vector path never straddles them
  const int i0 = (int)((long long)(P.Q >> 2) * blockIdx.y / nsplit) << 2;
  const int i1 = (int)blockIdx.y + 1 == nsplit ? P.Q : ((int)((long long)(P.Q >> 2) * (blockIdx.y + 1) / nsplit) << 2);
  float mu[TFNAS_MAX_OPS], r[TFNAS_MAX_OPS], acc[TFNAS_MAX_OPS], sg = 0.f;
#pragma unroll
  for (int s = 0; s < TFNAS_MAX_OPS; ++s) {
    acc[s] = 0.f;
    mu[s] = s < na ? bn3[s * oc + c] : 0.f;
    r[s] = s < na ? bn3[na * oc + s * oc + c] : 0.f;
  }
  const float inv_hwo = 1.f / (float)HWo;
  if ((HWo & 3) == 0) {
    for (int i = (i0 >> 2) + threadIdx.x; i < (i1 >> 2); i += NT) {
      const int e0 = i << 2;
      const int n = fast_div(e0, HWo, inv_hwo), hw = e0 - n * HWo;
      const float4 g = *(const float4*)(G + ((size_t)n * oc + c) * HWo + hw);
      sg += g.x + g.y + g.z + g.w;
#pragma unroll
      for (int s = 0; s < TFNAS_MAX_OPS; ++s)
        if (s < na) {
          const float4 z = *(const float4*)(Zb + ((size_t)(n * na + s) * oc + c) * HWo + hw);
          acc[s] += (g.x * (z.x - mu[s]) + g.y * (z.y - mu[s]) + g.z * (z.z - mu[s]) + g.w * (z.w - mu[s])) * r[s];
        }
    }
  } else {
    for (int i = i0 + threadIdx.x; i < i1; i += NT) {
      const int n = fast_div(i, HWo, inv_hwo), hw = i - n * HWo;
      const float g = G[((size_t)n * oc + c) * HWo + hw];
      sg += g;
#pragma unroll
      for (int s = 0; s < TFNAS_MAX_OPS; ++s)
        if (s < na) acc[s] += g * (Zb[((size_t)(n * na + s) * oc + c) * HWo + hw] - mu[s]) * r[s];
    }
  }
  __shared__ double red[NT / 32][TFNAS_MAX_OPS + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v = warp_sum_d((double)sg);
  if (lane == 0) red[warp][TFNAS_MAX_OPS] = v;
#pragma unroll
  for (int s = 0; s < TFNAS_MAX_OPS; ++s) {
    double t = warp_sum_d((double)acc[s]);
    if (lane == 0) red[warp][s] = t;
  }
  __syncthreads();
  if (threadIdx.x <= TFNAS_MAX_OPS) {
    double t = 0;
    for (int w = 0; w < NT / 32; ++w) t += red[w][threadIdx.x];
    if (threadIdx.x == TFNAS_MAX_OPS) atomicAdd(&sG[c], t);
    else if ((int)threadIdx.x < na) atomicAdd(&sGY[threadIdx.x * oc + c], t);
  }
}

// dzc[slot*oc + o] = {A = w_i*r3, B = sG/Q, C = sGY/Q, 0}; dmix[id] = sum_c sGY
__global__ void k_b2prep(Plan P, const float* __restrict__ mixw, const float* __restrict__ bn3,
                         const double* __restrict__ sG, const double* __restrict__ sGY, float4* __restrict__ dzc,
                         float4* __restrict__ dzc2, float* __restrict__ dmix) {
  const int oc = P.oc, na = P.na;
  const double invQ = 1.0 / (double)P.Q;
  for (int i = threadIdx.x; i < na * oc; i += blockDim.x) {
    int s = i / oc, o = i - s * oc;
    float w = mixw[P.c[s].id];
    const double r3 = bn3[na * oc + i], mu3 = bn3[i];
    const double A = (double)w * r3, B = sG[o] * invQ, C = sGY[i] * invQ;
    dzc[i] = make_float4(w * bn3[na * oc + i], (float)B, (float)C, 0.f);
    // dz = A*(g - B - (z - mu3)*r3*C) = A*g + (-A*C*r3)*z + A*(C*r3*mu3 - B)
    dzc2[i] = make_float4((float)A, (float)(-A * C * r3), (float)(A * (C * r3 * mu3 - B)), 0.f);
  }
  if (threadIdx.x < TFNAS_MAX_OPS) dmix[threadIdx.x] = 0.f;
  __syncthreads();
  if ((int)threadIdx.x < na) {
    double t = 0;
    for (int o = 0; o < oc; ++o) t += sGY[threadIdx.x * oc + o];
    dmix[P.c[threadIdx.x].id] = (float)t;
  }
}

// ----------------------------------------------------------------------------------------------
// B2: dc = W3^T dz
// ----------------------------------------------------------------------------------------------
template <int TC, int ACT>
__global__ void __launch_bounds__(NT) k_dc(Plan P, const float* __restrict__ G, const float* __restrict__ Zb,
                                            const float* __restrict__ bn3, const float4* __restrict__ dzc,
                                            const float* __restrict__ D, const float* __restrict__ bn2,
                                            float* __restrict__ DC, float* __restrict__ dg, double* __restrict__ sD) {
  __shared__ __align__(16) float ins[PW_KC * PW_LDP];
  __shared__ __align__(16) float ws[PW_KC * (8 * TC + 4)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = blockIdx.z;
  const Cand& cd = P.c[slot];
  const int mc = cd.mc, oc = P.oc, na = P.na;
  const int c0 = blockIdx.y * 8 * TC;
  if (c0 >= mc) return;
  const int no = min(8 * TC, mc - c0);
  Px4 px;
  px_decomp(px, blockIdx.x * PW_TPX + lane * 4, P.Q, P.HWo);
  float acc[TC][4];
#pragma unroll
  for (int j = 0; j < TC; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  for (int k0 = 0; k0 < oc; k0 += PW_KC) {
    const int nk = min(PW_KC, oc - k0);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PW_KC / 8; ++i) {
      const int kk = warp + i * 8, o = k0 + kk;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (kk < nk) {
        const float4 cf = dzc[slot * oc + o];
        const float mu3 = bn3[slot * oc + o], r3 = bn3[na * oc + slot * oc + o];
        float g[4], z[4];
        load4(g, G, px, oc, o, P.HWo);
        load4(z, Zb, px, na * oc, slot * oc + o, P.HWo);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = px.v[e] ? cf.x * (g[e] - cf.y - (z[e] - mu3) * r3 * cf.z) : 0.f;
      }
      *(float4*)(ins + kk * PW_LDP + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    stage_w_n<TC>(ws, cd.w3, mc, c0, no, k0, nk);
    __syncthreads();
    if (warp * TC < no) pw_mma<TC>(acc, ins, ws, lane, warp);
  }
  if (warp * TC >= no) return;
  const bool gated = cd.se > 0;
  // are all lanes of this warp inside one image? (then the SE partial sum is one atomic per warp)
  const int n_first = __shfl_sync(0xffffffffu, px.n[0], 0);
  const bool one_img = __all_sync(0xffffffffu, (!px.v[0]) || (px.n[0] == n_first && px.n[3] == n_first && px.v[3]));
#pragma unroll
  for (int j = 0; j < TC; ++j) {
    const int c = c0 + warp * TC + j;   // warp-uniform
    if (c >= mc) continue;
    const int cst = cd.coff + c;
    const float mu = bn2[cst], r = bn2[P.MC + cst];
    float d[4];
    load4(d, D, px, P.MC, cst, P.HWo);
    if (gated) {
      store4(DC, acc[j], px, P.MC, cst, P.HWo);
      float part[4], tot = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        part[e] = px.v[e] ? acc[j][e] * act_f<ACT>((d[e] - mu) * r) : 0.f;
        tot += part[e];
      }
      if (one_img) {
        tot = warp_sum(tot);
        if (lane == 0) atomicAdd(&dg[(size_t)n_first * P.MCse + cd.soff + c], tot);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (px.v[e]) atomicAdd(&dg[(size_t)px.n[e] * P.MCse + cd.soff + c], part[e]);
      }
    } else {
      float o[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float dh = (d[e] - mu) * r;
        o[e] = px.v[e] ? acc[j][e] * act_df<ACT>(dh) : 0.f;
        s1 += o[e];
        s2 += o[e] * dh;
      }
      store4(DC, o, px, P.MC, cst, P.HWo);
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        atomicAdd(&sD[2 * cst], (double)s1);
        atomicAdd(&sD[2 * cst + 1], (double)s2);
      }
    }
  }
}

// SE backward, FC by FC (tiled small GEMMs, see fc_tile_g):
//   de = dg * g (1-g)                         (on load; optionally saved for the weight grads)
//   du[n][j] = sum_c We[c][j] de[n][c]                  k_se_bwd1  grid (N/32, se/64, na*ksplit)  split-K, atomics
//   dt = act'(t) * du                                   applied by the consumers (k_se_bwd2, k_se_wgrad)
//   dp[n][c] = sum_j Wr[j][c] dt[n][j]  -> dg in place  k_se_bwd2  grid (N/32, mc/64, na)
template <int ACT>
__global__ void __launch_bounds__(NT) k_se_bwd1(Plan P, int ksplit, const float* __restrict__ seg,
                                                 const float* __restrict__ dg, float* __restrict__ sede,
                                                 float* __restrict__ sedu) {
  const int slot = blockIdx.z / ksplit, part = blockIdx.z - slot * ksplit;
  const Cand& cd = P.c[slot];
  if (cd.se == 0 || (int)blockIdx.y * FC_TO >= cd.se) return;
  const int chunks = (cd.mc + FC_KC - 1) / FC_KC;
  const int ks = min(ksplit, chunks);               // this candidate's split: every part < ks owns at least one chunk
  if (part >= ks) return;
  const int k0 = (chunks * part / ks) * FC_KC, k1 = min(cd.mc, (chunks * (part + 1) / ks) * FC_KC);
  const bool keep = sede != nullptr && blockIdx.y == 0;
  fc_tile<true>(P.N, cd.se, cd.mc, cd.ew,
                [&](int n, int k) {
                  const size_t gi = (size_t)n * P.MCse + cd.soff + k;
                  const float g = seg[gi];
                  const float de = dg[gi] * g * (1.f - g);
                  if (keep) sede[gi] = de;
                  return de;
                },
                [&](int n, int o, float a) { atomicAdd(&sedu[(size_t)n * P.SEH + cd.hoff + o], a); },
                k0, k1);
}
template <int ACT>
__global__ void __launch_bounds__(NT) k_se_bwd2(Plan P, const float* __restrict__ sedu, const float* __restrict__ set,
                                                 float* __restrict__ dg) {
  const Cand& cd = P.c[blockIdx.z];
  if (cd.se == 0 || (int)blockIdx.y * FC_TO >= cd.mc) return;
  fc_tile<true>(P.N, cd.mc, cd.se, cd.rw,
                [&](int n, int k) {
                  const size_t ti = (size_t)n * P.SEH + cd.hoff + k;
                  return sedu[ti] * act_df<ACT>(set[ti]);
                },
                [&](int n, int o, float a) { dg[(size_t)n * P.MCse + cd.soff + o] = a; });
}

// SE candidates: db = dc*g + dp/HWo ; dd-hat = db*act'(d-hat) in place ; BN2-backward sums.  warp per plane.
template <int ACT>
__global__ void __launch_bounds__(NT) k_b2b(Plan P, const float* __restrict__ D, const float* __restrict__ bn2,
                                             const float* __restrict__ seg, const float* __restrict__ dp,
                                             float* __restrict__ DC, double* __restrict__ sD) {
  const int widx = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  if (widx >= P.MCse) return;
  int s = -1;
  for (int i = 0; i < P.na; ++i)
    if (P.c[i].se > 0 && widx >= P.c[i].soff && widx < P.c[i].soff + P.c[i].mc) s = i;
  const int cst = P.c[s].coff + (widx - P.c[s].soff);
  const float mu = bn2[cst], r = bn2[P.MC + cst];
  const float g = seg[(size_t)n * P.MCse + widx];
  const float dpn = dp[(size_t)n * P.MCse + widx] / (float)P.HWo;
  const size_t base = ((size_t)n * P.MC + cst) * P.HWo;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < P.HWo; i += 32) {
    float dh = (D[base + i] - mu) * r;
    float v = (DC[base + i] * g + dpn) * act_df<ACT>(dh);
    DC[base + i] = v;
    s1 += v;
    s2 += v * dh;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    atomicAdd(&sD[2 * cst], (double)s1);
    atomicAdd(&sD[2 * cst + 1], (double)s2);
  }
}

// ----------------------------------------------------------------------------------------------
// B3a: transposed depthwise
// ----------------------------------------------------------------------------------------------

static DwCfg dwb_config(const Plan& P, int KS) {
  // tile over INPUT rows; the dd tile holds the output rows those input rows touch, zero-haloed
  DwCfg c;
  const int S = P.stride, pad = KS / 2;
  int CPB = 1;
  while (CPB < 32 && CPB * P.HW < 2048) CPB <<= 1;
  c.CPB = CPB;
  c.WP = dw_wp(P.Wo, pad);
  for (int tiles = 1;; ++tiles) {
    int R = cdiv(P.H, tiles);
    int IR = S == 1 ? R + KS - 1 : (R + KS - 2) / S + 2;
    size_t smem = ((size_t)CPB * IR * c.WP + 16) * 4;
    if (smem <= 40 * 1024 || R == 1) {
      c.R = R; c.IR = IR; c.tiles = cdiv(P.H, R); c.smem = smem;
      break;
    }
  }
  return c;
}

template <int KS, int S, int ACT, bool WG>
__global__ void __launch_bounds__(NT) k_dw_bwd(Plan P, DwWork Wk, DwCfg cfg, const float* __restrict__ DC,
                                                const float* __restrict__ D, const float* __restrict__ bn2,
                                                const double* __restrict__ sD, const float* __restrict__ UH,
                                                float* __restrict__ DA, DwGrads gw) {
  extern __shared__ __align__(16) float dds[];   // [CPB][IR][WP] (+16 slack)
  constexpr int pad = KS / 2;
  const int tid = threadIdx.x, n = blockIdx.z;
  int e = 0;
  while (e + 1 < Wk.n && (int)blockIdx.y >= Wk.gstart[e + 1]) ++e;
  const Cand& cd = P.c[Wk.slot[e]];
  const int CPB = cfg.CPB, IR = cfg.IR, WP = cfg.WP;
  const int cbase = ((int)blockIdx.y - Wk.gstart[e]) * CPB;
  const int nc = min(CPB, cd.mc - cbase);
  const int H = P.H, W = P.W, Ho = P.Ho, Wo = P.Wo;
  const int r0 = blockIdx.x * cfg.R, r1 = min(H, r0 + cfg.R);
  // output rows touching input rows [r0, r1): oy*S - pad + ky = r
  const int t0 = r0 + pad - KS + 1;
  const int oy_lo = max(0, t0 > 0 ? (t0 + S - 1) / S : 0);
  const int oy_hi = min(Ho - 1, (r1 - 1 + pad) / S);
  const int row0 = S == 1 ? r0 - pad : oy_lo;       // dd row held by tile row 0
  const double invQ = 1.0 / (double)P.Q;
  for (int i = tid; i < (CPB * IR * WP + 16) / 4; i += NT) ((float4*)dds)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // stage dd = r2 (dd-hat - m1 - d-hat m2) for the output rows this tile touches.  The BN2-backward
  // coefficients of the CTA's channels go to smem first: dd = ca*ddh + cb*d + cc.
  __shared__ float coef[3][32];
  if (tid < CPB) {
    float ca = 0.f, cb = 0.f, cc = 0.f;
    if (tid < nc) {
      const int cst = cd.coff + cbase + tid;
      const float mu = bn2[cst], r = bn2[P.MC + cst];
      const float m1 = (float)(sD[2 * cst] * invQ), m2 = (float)(sD[2 * cst + 1] * invQ);
      // r*(ddh - m1 - (d-mu)*r*m2) = r*ddh - r*r*m2*d + r*(mu*r*m2 - m1)
      ca = r; cb = -r * r * m2; cc = r * (mu * r * m2 - m1);
    }
    coef[0][tid] = ca; coef[1][tid] = cb; coef[2][tid] = cc;
  }
  __syncthreads();
  {
    const int vo_lo = max(row0, 0), vo_hi = min(row0 + IR - 1, oy_hi);
    if (vo_hi >= vo_lo) {
      const size_t off = (((size_t)n * P.MC + cd.coff + cbase) * Ho + vo_lo) * Wo;
      const float* dcp = DC + off;
      const ptrdiff_t d_minus_dc = D - DC;      // both tensors share the indexing
      // the functor receives the DC value; it fetches the matching D value itself (same address + delta)
      stage_planes2(dds, dcp, d_minus_dc, (size_t)Ho * Wo, nc, vo_hi - vo_lo + 1, Wo, IR, WP, vo_lo - row0, pad,
                    [&](float g, float d, int c) { return coef[0][c] * g + coef[1][c] * d + coef[2][c]; });
    }
  }
  __syncthreads();
  const int TPC = NT / CPB;
  const int cl = tid / TPC, jl = tid - cl * TPC;
  float gacc[WG ? KS * KS : 1];
#pragma unroll
  for (int i = 0; i < (WG ? KS * KS : 1); ++i) gacc[i] = 0.f;
  if (cl < nc) {
    const int cst = cd.coff + cbase + cl;
    float wr[KS * KS];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) wr[i] = cd.dw[(size_t)(cbase + cl) * KS * KS + i];
    const float* db = dds + (size_t)cl * IR * WP;
    const size_t pbase = (((size_t)n * P.MC + cst) * H + r0) * W;
    const int gpr = (W + 3) >> 2;
    const int ngroups = (r1 - r0) * gpr;
    const float inv_gpr = 1.f / (float)gpr;
    for (int g = jl; g < ngroups; g += TPC) {
      const int rl = fast_div(g, gpr, inv_gpr), col0 = (g - rl * gpr) * 4;
      const int r = r0 + rl;
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      if (WG) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col0 + j < W) a[j] = act_f<ACT>(UH[pbase + (size_t)rl * W + col0 + j]);
      }
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (S == 1) {
        // da[r][col] = sum_{ky',kx'} w[KS-1-ky'][KS-1-kx'] dd[r-pad+ky'][col-pad+kx'] : a correlation with the
        // flipped filter over the zero-haloed dd tile (tile row 0 <-> dd row r0-pad, tile col pad <-> dd col 0)
#pragma unroll
        for (int kyp = 0; kyp < KS; ++kyp) {
          const float* rp = db + (size_t)(rl + kyp) * WP + col0;
          float4 t0v = *(const float4*)rp, t1v = *(const float4*)(rp + 4);
          const float v[8] = {t0v.x, t0v.y, t0v.z, t0v.w, t1v.x, t1v.y, t1v.z, t1v.w};
#pragma unroll
          for (int kxp = 0; kxp < KS; ++kxp) {
            const int tap = (KS - 1 - kyp) * KS + (KS - 1 - kxp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o[j] += wr[tap] * v[j + kxp];
              if (WG) gacc[tap] += v[j + kxp] * a[j];
            }
          }
        }
      } else {
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          const int ty = r + pad - ky;
          if (ty < 0 || (ty & 1)) continue;
          const int oy = ty >> 1;
          if (oy < oy_lo || oy > oy_hi) continue;
          const float* drow = db + (size_t)(oy - oy_lo) * WP + pad + (col0 >> 1);
          const float v[4] = {drow[-1], drow[0], drow[1], drow[2]};
#pragma unroll
          for (int kx = 0; kx < KS; ++kx)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              constexpr int dummy = 0;
              (void)dummy;
              const int t = j + pad - kx;          // compile-time after unrolling
              if ((t & 1) == 0) {
                const float dd = v[(t >> 1) + 1];
                o[j] += wr[ky * KS + kx] * dd;
                if (WG) gacc[ky * KS + kx] += dd * a[j];
              }
            }
        }
      }
      float* q = DA + pbase + (size_t)rl * W + col0;
      if ((W & 3) == 0) {
        *(float4*)q = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col0 + j < W) q[j] = o[j];
      }
    }
  }
  if (WG) {
    float* gp = gw.p[e];
    const int GW = TPC < 32 ? TPC : 32;
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) {
      float t = group_sum(gacc[i], GW);
      if ((jl & (GW - 1)) == 0 && cl < nc) atomicAdd(&gp[(size_t)(cbase + cl) * KS * KS + i], t);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// B3b: dx_main = sum_i W1_i^T (r1 * du-hat), K = stacked MC
// ----------------------------------------------------------------------------------------------
template <int TK, int ACT>
__global__ void __launch_bounds__(NT) k_dx(Plan P, OcTile T, int ksplit, const float* __restrict__ DA,
                                            const float* __restrict__ UH, const float* __restrict__ bn1,
                                            float* __restrict__ dx, double* __restrict__ sU) {
  __shared__ __align__(16) float ins[PW_KC * PW_LDP];
  __shared__ __align__(16) float ws[PW_KC * (8 * TK + 4)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ic = P.ic;
  Px4 px;
  px_decomp(px, blockIdx.x * PW_TPX + lane * 4, P.P, P.HW);
  float acc[TK][4];
#pragma unroll
  for (int j = 0; j < TK; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  // this CTA's share of the stacked K chunks
  int total_chunks = 0;
  for (int s = 0; s < P.na; ++s) total_chunks += (P.c[s].mc + PW_KC - 1) / PW_KC;
  const int ch0 = (int)((long long)total_chunks * blockIdx.y / ksplit);
  const int ch1 = (int)((long long)total_chunks * (blockIdx.y + 1) / ksplit);
  int chunk = 0;
  for (int s = 0; s < P.na; ++s) {
    const Cand& cd = P.c[s];
    for (int k0 = 0; k0 < cd.mc; k0 += PW_KC, ++chunk) {
      if (chunk < ch0 || chunk >= ch1) continue;
      const int nk = min(PW_KC, cd.mc - k0);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < PW_KC / 8; ++i) {
        const int kk = warp + i * 8;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (kk < nk) {   // warp-uniform
          const int cst = cd.coff + k0 + kk;
          const float r1 = bn1[P.MC + cst];
          float da[4], uh[4];
          load4(da, DA, px, P.MC, cst, P.HW);
          load4(uh, UH, px, P.MC, cst, P.HW);
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float du = px.v[e] ? da[e] * act_df<ACT>(uh[e]) : 0.f;
            s1 += du;
            s2 += du * uh[e];
            v[e] = du * r1;
          }
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          if (lane == 0) {
            atomicAdd(&sU[2 * cst], (double)s1);
            atomicAdd(&sU[2 * cst + 1], (double)s2);
          }
        }
        *(float4*)(ins + kk * PW_LDP + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
      stage_w_n<TK>(ws, cd.w1, ic, 0, ic, k0, nk);
      __syncthreads();
      if (warp < T.ng) pw_mma<TK>(acc, ins, ws, lane, warp);
    }
  }
  if (warp < T.ng) {
#pragma unroll
    for (int j = 0; j < TK; ++j) {
      const int k = warp * TK + j;
      if (k < ic) {
        if (ksplit == 1) store4(dx, acc[j], px, ic, k, P.HW);
        else atomic_add4(dx, acc[j], px, ic, k, P.HW);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// B4: fold the BN1 backward into an ic x ic correction
// ----------------------------------------------------------------------------------------------
// Mm[k][k'] = sum_c W1[c][k] a2_c W1[c][k'] as a tiled CUDA-core GEMM over the stacked channels:
// grid (ic/64, ic/64, c-splits); 64x64 tile, 4x4 per thread, 32 channels per smem chunk; float atomics into Mm.
#define B4_T 64
#define B4_KC 32
// The per-channel coefficients a1 = r1 m1, a2 = r1^2 m2 (m = BN1-backward means, from the sums of the dx GEMM) are formed
// here per chunk (they were a launch of their own, k_b4coef).
__global__ void __launch_bounds__(NT) k_b4mm(Plan P, const float* __restrict__ bn1, const double* __restrict__ sU, double invP,
                                              float* __restrict__ Mm, float* __restrict__ cvec) {
  __shared__ __align__(16) float As[B4_KC][B4_T + 4];   // a2_c * W1[c][k0 + .]
  __shared__ __align__(16) float Bs[B4_KC][B4_T + 4];   // W1[c][kp0 + .]
  __shared__ float a1s[B4_KC], a2s[B4_KC];              // a1_c, a2_c of the chunk (CTAs of the first tile row also form cvec)
  float cv = 0.f;                                       // cvec[kp0 + tid] partial (tid < B4_T, blockIdx.x == 0)
  const int ic = P.ic, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * B4_T, kp0 = blockIdx.y * B4_T;
  const int c_lo = (int)((long long)P.MC * blockIdx.z / gridDim.z), c_hi = (int)((long long)P.MC * (blockIdx.z + 1) / gridDim.z);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int cb = c_lo; cb < c_hi; cb += B4_KC) {
    __syncthreads();
    if (tid < B4_KC) {
      const int c = cb + tid;
      float a1 = 0.f, a2 = 0.f;
      if (c < c_hi) {
        const double r1 = (double)bn1[P.MC + c];
        a1 = (float)(r1 * sU[2 * c] * invP);
        a2 = (float)(r1 * r1 * sU[2 * c + 1] * invP);
      }
      a1s[tid] = a1;
      a2s[tid] = a2;
    }
    __syncthreads();
    for (int i = tid; i < B4_KC * B4_T; i += NT) {
      const int cc = i / B4_T, kk = i - cc * B4_T;
      const int cst = cb + cc;
      float wa = 0.f, wb = 0.f;
      if (cst < c_hi) {
        int s = 0;
        while (s + 1 < P.na && cst >= P.c[s + 1].coff) ++s;
        const float* w = P.c[s].w1 + (size_t)(cst - P.c[s].coff) * ic;
        if (k0 + kk < ic) wa = w[k0 + kk] * a2s[cc];
        if (kp0 + kk < ic) wb = w[kp0 + kk];
      }
      As[cc][kk] = wa;
      Bs[cc][kk] = wb;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < B4_T) {          // cvec[k'] = sum_c W1[c][k'] a1_c (was a separate launch)
#pragma unroll 8
      for (int cc = 0; cc < B4_KC; ++cc) cv += Bs[cc][tid] * a1s[cc];
    }
#pragma unroll 8
    for (int cc = 0; cc < B4_KC; ++cc) {
      const float4 a = *(const float4*)&As[cc][ty * 4];
      const float4 bq = *(const float4*)&Bs[cc][tx * 4];
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + ty * 4 + i, kp = kp0 + tx * 4 + j;
      if (k < ic && kp < ic) atomicAdd(&Mm[k * ic + kp], acc[i][j]);
    }
  if (blockIdx.x == 0 && tid < B4_T && kp0 + tid < ic) atomicAdd(&cvec[kp0 + tid], cv);
}

// dx = dx_main - Mm x - (cvec - Mm mu_x) (+ G); the per-channel constant is formed in the prologue
template <int TK>
__global__ void __launch_bounds__(NT) k_dxfin(Plan P, OcTile T, const float* __restrict__ x,
                                               const float* __restrict__ Mm, const float* __restrict__ cvec2,
                                               const double* __restrict__ xmom,
                                               const float* __restrict__ G, float* __restrict__ dx) {
  __shared__ __align__(16) float ins[PW_KC * PW_LDP];
  __shared__ __align__(16) float ws[PW_KC * (8 * TK + 4)];
  __shared__ float cvs[8 * TK];       // cvec[k] - sum_k' Mm[k][k'] mu_x[k'] for this CTA's channels (was k_b4fin)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ic = P.ic;
  const int o0 = blockIdx.y * T.occ;          // output-channel chunk of this CTA (small planes: several CTAs per pixel tile)
  if (tid < 8 * TK) {
    const int k = o0 + tid;
    float c = 0.f;
    if (k < ic && tid < T.occ) {
      double t = 0.0;
      for (int kp = 0; kp < ic; ++kp) t += (double)Mm[k * ic + kp] * xmom[kp];
      c = (float)((double)cvec2[k] - t);
    }
    cvs[tid] = c;
  }
  Px4 px;
  px_decomp(px, blockIdx.x * PW_TPX + lane * 4, P.P, P.HW);
  float acc[TK][4];
#pragma unroll
  for (int j = 0; j < TK; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  for (int k0 = 0; k0 < ic; k0 += PW_KC) {
    const int nk = min(PW_KC, ic - k0);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PW_KC / 8; ++i) {
      const int kk = warp + i * 8;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      if (kk < nk) load4(d, x, px, ic, k0 + kk, P.HW);
      *(float4*)(ins + kk * PW_LDP + lane * 4) = make_float4(d[0], d[1], d[2], d[3]);
    }
    stage_w_t<TK>(ws, Mm, ic, o0, min(T.occ, ic - o0), k0, nk);
    __syncthreads();
    if (warp < T.ng) pw_mma<TK>(acc, ins, ws, lane, warp);
  }
  if (warp < T.ng) {
#pragma unroll
    for (int j = 0; j < TK; ++j) {
      const int k = o0 + warp * TK + j;
      if (k < ic && warp * TK + j < T.occ) {
        float m[4], o[4], g[4] = {0.f, 0.f, 0.f, 0.f};
        load4(m, dx, px, ic, k, P.HW);
        if (P.residual) load4(g, G, px, ic, k, P.HW);   // residual => oc == ic, HWo == HW
        const float cv = cvs[warp * TK + j];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = m[e] - acc[j][e] - cv + g[e];
        store4(dx, o, px, ic, k, P.HW);
      }
    }
  }
}

__global__ void k_alpha_grad(int num_ops, const float* __restrict__ mixw, const float* __restrict__ lat,
                             const float* __restrict__ dmix, const float* __restrict__ dlat, float T,
                             float* __restrict__ dalpha) {
  if (threadIdx.x != 0) return;
  float dl = dlat ? *dlat : 0.f;
  float dw[TFNAS_MAX_OPS], dot = 0.f;
  for (int i = 0; i < num_ops; ++i) {
    dw[i] = dmix[i] + dl * lat[i];
    dot += dw[i] * mixw[i];
  }
  for (int i = 0; i < num_ops; ++i) dalpha[i] = mixw[i] * (dw[i] - dot) / T;
}

// ----------------------------------------------------------------------------------------------
// weight gradients (sampled w-step: one active candidate)
// ----------------------------------------------------------------------------------------------
// Out[a][b] += sum_p U[a][p] * V[b][p].  MODE 0 (W3): a = oc (U = dz), b = mc (V = c), pixels Q.
//                                         MODE 1 (W1): a = mc (U = du-hat), b = ic (V = x), pixels P -> Smat.
#define WG_T 64
#define WG_KP 32
#define WG_LD (WG_T + 4)
template <int MODE, int ACT>
__global__ void __launch_bounds__(NT) k_wgrad(Plan P, int slot, const float* __restrict__ A0,
                                               const float* __restrict__ A1, const float* __restrict__ B0,
                                               const float* __restrict__ bnA, const float4* __restrict__ dzc,
                                               const float* __restrict__ bnB, const float* __restrict__ seg,
                                               float* __restrict__ Out) {
  __shared__ __align__(16) float Us[WG_KP * WG_LD];
  __shared__ __align__(16) float Vs[WG_KP * WG_LD];
  const Cand& cd = P.c[slot];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int na_ = MODE == 0 ? P.oc : cd.mc;    // rows of Out
  const int nb_ = MODE == 0 ? cd.mc : P.ic;    // cols of Out
  const int HWp = MODE == 0 ? P.HWo : P.HW;
  const int total = MODE == 0 ? P.Q : P.P;
  const int a0 = blockIdx.x * WG_T, b0 = blockIdx.y * WG_T;
  const int nsplit = gridDim.z;
  int p_lo = (int)((long long)total * blockIdx.z / nsplit), p_hi = (int)((long long)total * (blockIdx.z + 1) / nsplit);
  p_lo = p_lo / WG_KP * WG_KP;
  if ((int)blockIdx.z + 1 < nsplit) p_hi = p_hi / WG_KP * WG_KP;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int pp = tid & 31, rr = tid >> 5;   // staging: 32 pixels x 8 rows per pass
  for (int p0 = p_lo; p0 < p_hi; p0 += WG_KP) {
    const int p = p0 + pp;
    const bool pv = p < p_hi;
    const int n = pv ? p / HWp : 0, hw = pv ? p - n * HWp : 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < WG_T / 8; ++i) {
      const int row = rr + i * 8;
      float u = 0.f, v = 0.f;
      const int a = a0 + row, b = b0 + row;
      if (MODE == 0) {
        if (pv && a < na_) {   // dz[o][p]
          const float4 cf = dzc[slot * P.oc + a];
          const float mu3 = bnA[slot * P.oc + a], r3 = bnA[P.na * P.oc + slot * P.oc + a];
          float g = A0[((size_t)n * P.oc + a) * HWp + hw];
          float z = A1[((size_t)(n * P.na + slot) * P.oc + a) * HWp + hw];
          u = cf.x * (g - cf.y - (z - mu3) * r3 * cf.z);
        }
        if (pv && b < nb_) {   // c[c][p] = act(BN2(d)) * gate
          const int cst = cd.coff + b;
          float d = B0[((size_t)n * P.MC + cst) * HWp + hw];
          v = act_f<ACT>((d - bnB[cst]) * bnB[P.MC + cst]);
          if (cd.se > 0) v *= seg[(size_t)n * P.MCse + cd.soff + b];
        }
      } else {
        if (pv && a < na_) {   // du-hat[c][p] = DA * act'(UH)
          const int cst = cd.coff + a;
          const size_t idx = ((size_t)n * P.MC + cst) * HWp + hw;
          u = A0[idx] * act_df<ACT>(A1[idx]);
        }
        if (pv && b < nb_) v = B0[((size_t)n * P.ic + b) * HWp + hw];
      }
      Us[pp * WG_LD + row] = u;
      Vs[pp * WG_LD + row] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int q = 0; q < WG_KP; ++q) {
      float4 ua = *(const float4*)(Us + q * WG_LD + ty * 4);
      float4 vb = *(const float4*)(Vs + q * WG_LD + tx * 4);
      const float uu[4] = {ua.x, ua.y, ua.z, ua.w}, vv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += uu[i] * vv[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int a = a0 + ty * 4 + i, b = b0 + tx * 4 + j;
      if (a < na_ && b < nb_) atomicAdd(&Out[(size_t)a * nb_ + b], acc[i][j]);
    }
}

// dW1[c][k] = r1 ( S[c][k] - m1 P mu_x[k] - m2 r1 P (W1 cov)[c][k] ).  8 channels per CTA, thread = input channel k
#define W1F_CH 8
__global__ void __launch_bounds__(NT) k_w1fin(Plan P, int slot, const float* __restrict__ Smat, int smat_t,
                                               const float* __restrict__ bn1, const double* __restrict__ sU,
                                               const double* __restrict__ xmom, float* __restrict__ gw1) {
  extern __shared__ float wsm[];     // [W1F_CH][ic]
  const Cand& cd = P.c[slot];
  const int ic = P.ic, tid = threadIdx.x;
  const int c0 = blockIdx.x * W1F_CH;
  for (int i = tid; i < W1F_CH * ic; i += NT) {
    const int cc = i / ic, k = i - cc * ic;
    wsm[i] = c0 + cc < cd.mc ? cd.w1[(size_t)(c0 + cc) * ic + k] : 0.f;
  }
  __syncthreads();
  const double* mean = xmom;
  const double* cov = xmom + ic;
  for (int k = tid; k < ic; k += NT) {
    double wc[W1F_CH];
#pragma unroll
    for (int c = 0; c < W1F_CH; ++c) wc[c] = 0.0;
    for (int kp = 0; kp < ic; ++kp) {
      const double cv = cov[kp * ic + k];
#pragma unroll
      for (int c = 0; c < W1F_CH; ++c) wc[c] += (double)wsm[c * ic + kp] * cv;
    }
#pragma unroll
    for (int c = 0; c < W1F_CH; ++c) {
      const int cl = c0 + c;
      if (cl < cd.mc) {
        const int cst = cd.coff + cl;
        const double r1 = (double)bn1[P.MC + cst];
        const double s1 = sU[2 * cst], s2 = sU[2 * cst + 1];   // = m1*P, m2*P
        const double smv = (double)(smat_t ? Smat[(size_t)k * cd.mc + cl] : Smat[(size_t)cl * ic + k]);
        gw1[(size_t)cl * ic + k] = (float)(r1 * (smv - s1 * mean[k] - s2 * r1 * wc[c]));
      }
    }
  }
}

// SE weight grads: two small GEMMs over the batch axis (K = N images) plus the bias column sums.
//   blockIdx.z == 0:  dWe[c][j] = sum_n de[n][c] * act(t[n][j])      rows c (mc), outputs j (se);  dbe[c] = sum_n de[n][c]
//   blockIdx.z == 1:  dWr[j][c] = sum_n dt[n][j] * p[n][c]           rows j (se), outputs c (mc);  dbr[j] = sum_n dt[n][j]
// with dt = du * act'(t).  grid (max(mc, se)/32, max(mc, se)/64, 2)
template <int ACT>
__global__ void __launch_bounds__(NT) k_se_wgrad(Plan P, int slot, const float* __restrict__ sede,
                                                  const float* __restrict__ sedu, const float* __restrict__ sep,
                                                  const float* __restrict__ set, TfnasCandPtrs gw) {
  const Cand& cd = P.c[slot];
  const int mc = cd.mc, se = cd.se, N = P.N;
  __shared__ __align__(16) FcSmem sm;
  if (blockIdx.z == 0) {
    if ((int)blockIdx.x * FC_TN >= mc || (int)blockIdx.y * FC_TO >= se) return;
    fc_tile_g<true, true>(sm, mc, se, 0, N,
                          [&](int c, int n) { return sede[(size_t)n * P.MCse + cd.soff + c]; },
                          [&](int j, int n) { return act_f<ACT>(set[(size_t)n * P.SEH + cd.hoff + j]); },
                          [&](int c, int j, float a) { gw.se_ew[(size_t)c * se + j] = a; });
    if (blockIdx.y == 0) {
      const int c = blockIdx.x * FC_TN + threadIdx.x;
      if ((int)threadIdx.x < FC_TN && c < mc) {
        float a = 0.f;
        for (int n = 0; n < N; ++n) a += sede[(size_t)n * P.MCse + cd.soff + c];
        gw.se_eb[c] = a;
      }
    }
  } else {
    if ((int)blockIdx.x * FC_TN >= se || (int)blockIdx.y * FC_TO >= mc) return;
    auto dt = [&](int j, int n) {
      const size_t ti = (size_t)n * P.SEH + cd.hoff + j;
      return sedu[ti] * act_df<ACT>(set[ti]);
    };
    fc_tile_g<true, true>(sm, se, mc, 0, N, dt,
                          [&](int c, int n) { return sep[(size_t)n * P.MCse + cd.soff + c]; },
                          [&](int j, int c, float a) { gw.se_rw[(size_t)j * mc + c] = a; });
    if (blockIdx.y == 0) {
      const int j = blockIdx.x * FC_TN + threadIdx.x;
      if ((int)threadIdx.x < FC_TN && j < se) {
        float a = 0.f;
        for (int n = 0; n < N; ++n) a += dt(j, n);
        gw.se_rb[j] = a;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
template <int TC>
static void launch_dc(const Plan& P, int maxmc, const float* G, const float* Zb, const float* bn3, const float4* dzc,
                      const float* D, const float* bn2, float* DC, float* dg, double* sD, cudaStream_t st) {
  dim3 grid(cdiv(P.Q, PW_TPX), cdiv(maxmc, 8 * TC), P.na);
  ProfScope ps("dc", 4.0 * P.Q * ((double)P.oc * (1 + P.na) + 2.0 * P.MC) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  if (P.act == TFNAS_ACT_RELU)
    k_dc<TC, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, G, Zb, bn3, dzc, D, bn2, DC, dg, sD);
  else
    k_dc<TC, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, G, Zb, bn3, dzc, D, bn2, DC, dg, sD);
}

template <int KS, int S>
static void launch_dw_bwd(const Plan& P, const float* DC, const float* D, const float* bn2, const double* sD,
                          const float* UH, float* DA, const TfnasCandPtrs* dweights, cudaStream_t st) {
  DwCfg cfg = dwb_config(P, KS);
  DwWork w;
  w.n = 0;
  w.gstart[0] = 0;
  DwGrads gw;
  memset(&gw, 0, sizeof(gw));
  double mck = 0;
  for (int s = 0; s < P.na; ++s) {
    if (P.c[s].k != KS) continue;
    w.slot[w.n] = s;
    w.gstart[w.n + 1] = w.gstart[w.n] + cdiv(P.c[s].mc, cfg.CPB);
    if (dweights && w.n < 4) gw.p[w.n] = dweights[P.c[s].id].dw;
    mck += P.c[s].mc;
    ++w.n;
  }
  if (!w.n) return;
  dim3 grid(cfg.tiles, w.gstart[w.n], P.N);
  const bool relu = P.act == TFNAS_ACT_RELU;
  ProfScope ps(KS == 3 ? "dw_bwd_k3" : "dw_bwd_k5", 4.0 * mck * (2.0 * P.Q + (dweights ? 2.0 : 1.0) * P.P),
               2.0 * KS * KS * mck * P.Q * (dweights ? 2 : 1), st);
  if (dweights) {
    auto kern = relu ? k_dw_bwd<KS, S, TFNAS_ACT_RELU, true> : k_dw_bwd<KS, S, TFNAS_ACT_SWISH, true>;
    ensure_smem(kern, (size_t)(cfg.smem));
    kern<<<grid, NT, cfg.smem, st>>>(P, w, cfg, DC, D, bn2, sD, UH, DA, gw);
  } else {
    auto kern = relu ? k_dw_bwd<KS, S, TFNAS_ACT_RELU, false> : k_dw_bwd<KS, S, TFNAS_ACT_SWISH, false>;
    ensure_smem(kern, (size_t)(cfg.smem));
    kern<<<grid, NT, cfg.smem, st>>>(P, w, cfg, DC, D, bn2, sD, UH, DA, gw);
  }
}

template <int TK>
static void launch_dx(const Plan& P, OcTile T, int ksplit, const float* DA, const float* UH, const float* bn1,
                      float* dx, double* sU, cudaStream_t st) {
  dim3 grid(cdiv(P.P, PW_TPX), ksplit);
  ProfScope ps("dx", 4.0 * P.P * (2.0 * P.MC + P.ic) + 4.0 * P.MC * P.ic, 2.0 * P.P * (double)P.MC * P.ic, st);
  if (P.act == TFNAS_ACT_RELU)
    k_dx<TK, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, T, ksplit, DA, UH, bn1, dx, sU);
  else
    k_dx<TK, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, T, ksplit, DA, UH, bn1, dx, sU);
}

template <int TK>
static void launch_dxfin(const Plan& P, OcTile T, const float* x, const float* Mm, const float* cvec2, const double* xmom,
                         const float* G, float* dx, cudaStream_t st) {
  ProfScope ps("dxfin", 4.0 * P.P * P.ic * (3.0 + (P.residual ? 1 : 0)), 2.0 * P.P * (double)P.ic * P.ic, st);
  k_dxfin<TK><<<dim3(cdiv(P.P, PW_TPX), T.nchunk), NT, 0, st>>>(P, T, x, Mm, cvec2, xmom, G, dx);
}

// ---- side stream for the weight-gradient GEMMs ---------------------------------------------------------------------
// dW3 / dW1 only feed the caller (nothing on the dgrad chain reads them), and a sampled single-candidate pass leaves most
// SMs idle, so they are forked onto a library-owned stream per caller stream and joined before the call's last kernels:
// every launch of the call is still ordered before whatever the caller enqueues next on ITS stream.
// On by default since round 2 (TFNAS_SIDE_STREAM=0 disables): with the launch sequence issued from C++ (body executor) the
// weight-gradient GEMMs of a sampled pass overlap the dx chain instead of sitting on its critical path: 2288 -> 2322
// images/s (profiles/bench_r2*.json).  Round 1, with Python issuing every launch, had measured no gain (1843 vs 1907).
#include <map>
#include <mutex>
struct SideStream { cudaStream_t s; cudaEvent_t fork1, fork2, join; };
static int g_side_stream = -1;      // -1: not decided yet (TFNAS_SIDE_STREAM, default on); tfnas_config_side_stream overrides
extern "C" int tfnas_config_side_stream(int on) {
  const int prev = g_side_stream;
  g_side_stream = on ? 1 : 0;
  return prev;
}
static SideStream* side_stream_for(cudaStream_t main) {
  if (g_side_stream < 0) g_side_stream = !(getenv("TFNAS_SIDE_STREAM") && strcmp(getenv("TFNAS_SIDE_STREAM"), "0") == 0) ? 1 : 0;
  if (!g_side_stream) return nullptr;
  static std::map<std::pair<int, cudaStream_t>, SideStream> streams;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_pair(dev, main);
  auto it = streams.find(key);
  if (it == streams.end()) {
    if (streams.size() >= 64) return nullptr;
    SideStream sd;
    // lowest priority (on current drivers that IS the default, 0): a caller that wants its dx critical path (dozens of small
    // dependent kernels) scheduled ahead of the weight-gradient tiles runs the pass on a higher-priority stream, as
    // search_loop.w_step does (torch.cuda.Stream(priority=-1))
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&sd.s, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaEventCreateWithFlags(&sd.fork1, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&sd.fork2, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming);
    it = streams.insert({key, sd}).first;
  }
  return &it->second;
}

void launch_backward(const Plan& P, const float* x, const float* dout, const float* dlat, float T,
                     int alpha_mode, const char* saved, const SavedLayout& L, const BwdScratch& S,
                     float* dx, float* dlog_alphas, const TfnasCandPtrs* dweights, cudaStream_t st, int stop_at_da,
                     const PreppedBwd* pre) {
  const double* xmom = (const double*)(saved + L.xmom);
  const float* bn1 = (const float*)(saved + L.bn1);
  const float* bn2 = (const float*)(saved + L.bn2);
  const float* bn3 = (const float*)(saved + L.bn3);
  const float* mixw = (const float*)(saved + L.mixw);
  const float* latsave = (const float*)(saved + L.lat);
  const float* sep = (const float*)(saved + L.sep);
  const float* set = (const float*)(saved + L.set);
  const float* seg = (const float*)(saved + L.seg);
  const float* UH = (const float*)(saved + L.UH);
  const float* D = (const float*)(saved + L.D);
  const float* Zb = (const float*)(saved + L.Z);
  const bool relu = P.act == TFNAS_ACT_RELU;
  const int ic = P.ic, oc = P.oc;
  // zero the accumulators: [sG | sGY | sD | sU] doubles are contiguous, then dg floats
  cudaMemsetAsync(S.sG, 0, S.zero_bytes, st);       // sums, cvec2, Mm, sedt, dg (api.cu::bwd_scratch)
  // B1
  {
    int nsplit = max(1, min(cdiv(P.Q, 1024), cdiv(4 * sm_count(), oc)));
    { ProfScope ps("b1", 4.0 * P.Q * oc * (1 + P.na), 3.0 * P.Q * oc * P.na, st);
      k_b1<<<dim3(oc, nsplit), NT, 0, st>>>(P, dout, Zb, bn3, S.sG, S.sGY); }
    { ProfScope ps("b2prep", 32.0 * P.na * oc, 0, st);
      k_b2prep<<<1, 256, 0, st>>>(P, mixw, bn3, S.sG, S.sGY, S.dzc, S.dzc2, S.dmix); }
  }
  const float4* dzc = S.dzc;
  if (!dx && !stop_at_da) {   // input needs no gradient (first MixedOP of the alpha step): only dL/dlog_alpha
    if (alpha_mode && dlog_alphas) {
      ProfScope ps("alpha_grad", 128, 0, st);
      k_alpha_grad<<<1, 32, 0, st>>>(P.num_ops, mixw, latsave, S.dmix, dlat, T, dlog_alphas);
    }
    return;
  }
  // B2
  UmWAll WD;
  UmW WX;
  DxChunks CHX;
  if (umma_enabled()) {
    // stop_at_da (second stem: no expand conv in front of the depthwise stage): only the dc weights exist
    if (pre) { WD = pre->WD; WX = pre->WX; CHX = pre->CH; }
    else if (stop_at_da) umma_prep_dc(P, S.umprep, WD, st);
    else umma_prep_bwd(P, bn1, S.umprep, WD, WX, CHX, st);
    umma_dc(P, WD, dout, Zb, S.dzc2, D, bn2, S.DC, S.dg, S.sD, st);
  } else {
    int maxmc = 0;
    for (int s = 0; s < P.na; ++s) maxmc = max(maxmc, P.c[s].mc);
    if (maxmc > 64) launch_dc<16>(P, maxmc, dout, Zb, bn3, dzc, D, bn2, S.DC, S.dg, S.sD, st);
    else launch_dc<8>(P, maxmc, dout, Zb, bn3, dzc, D, bn2, S.DC, S.dg, S.sD, st);
  }
  SideStream* side = (dweights && umma_enabled() && !stop_at_da) ? side_stream_for(st) : nullptr;
  const cudaStream_t wst = side ? side->s : st;      // stream of the weight-gradient GEMMs
  if (dweights) {   // dW3 = sum_p dz c^T  (reads D, dout, Z, bn2, seg, dzc2: none of them is written again in this call)
    if (side) { cudaEventRecord(side->fork1, st); cudaStreamWaitEvent(wst, side->fork1, 0); }
    for (int s = 0; s < P.na; ++s) {
      const Cand& cd = P.c[s];
      float* gw3 = dweights[cd.id].w3;
      cudaMemsetAsync(gw3, 0, (size_t)oc * cd.mc * sizeof(float), umma_enabled() ? wst : st);
      if (umma_enabled()) {
        umma_wgrad(P, s, 0, D, nullptr, dout, Zb, bn2, seg, bn3, S.dzc2, gw3, wst);
      } else {
        int nsplit = max(1, min(cdiv(P.Q, 2048), cdiv(6 * sm_count(), cdiv(oc, WG_T) * cdiv(cd.mc, WG_T))));
        dim3 grid(cdiv(oc, WG_T), cdiv(cd.mc, WG_T), nsplit);
        ProfScope ps("wgrad_w3", 4.0 * P.Q * (2.0 * oc + cd.mc), 2.0 * P.Q * (double)oc * cd.mc, st);
        if (relu) k_wgrad<0, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, s, dout, Zb, D, bn3, dzc, bn2, seg, gw3);
        else k_wgrad<0, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, s, dout, Zb, D, bn3, dzc, bn2, seg, gw3);
      }
    }
  }
  if (stop_at_da == 2) return;     // head (feature-mix conv): S.DC = dL/d(input), dW3 written; no depthwise stage behind it
  // SE backward + B2b
  if (P.MCse > 0) {
    int maxmc = 0, maxse = 0;
    double fcw = 0;
    for (int s = 0; s < P.na; ++s)
      if (P.c[s].se > 0) { maxmc = max(maxmc, P.c[s].mc); maxse = max(maxse, P.c[s].se); fcw += 2.0 * P.c[s].mc * P.c[s].se; }
    float* sede = dweights ? S.sede : nullptr;
    dim3 gpl(cdiv(P.MCse * 32, NT), P.N);
    { ProfScope ps("se_bwd", 4.0 * fcw + 12.0 * P.N * P.MCse, 2.0 * P.N * fcw, st);
      const int ksplit = max(1, min(8, cdiv(maxmc, 2 * FC_KC)));
      dim3 g1(cdiv(P.N, FC_TN), cdiv(maxse, FC_TO), P.na * ksplit);
      if (relu) k_se_bwd1<TFNAS_ACT_RELU><<<g1, NT, 0, st>>>(P, ksplit, seg, S.dg, sede, S.sedt);
      else k_se_bwd1<TFNAS_ACT_SWISH><<<g1, NT, 0, st>>>(P, ksplit, seg, S.dg, sede, S.sedt);
      dim3 g2(cdiv(P.N, FC_TN), cdiv(maxmc, FC_TO), P.na);
      if (relu) k_se_bwd2<TFNAS_ACT_RELU><<<g2, NT, 0, st>>>(P, S.sedt, set, S.dg);
      else k_se_bwd2<TFNAS_ACT_SWISH><<<g2, NT, 0, st>>>(P, S.sedt, set, S.dg);
      count_launch(1); }
    { ProfScope ps("b2b", 12.0 * P.Q * P.MCse, 8.0 * P.Q * P.MCse, st);
      if (relu) k_b2b<TFNAS_ACT_RELU><<<gpl, NT, 0, st>>>(P, D, bn2, seg, S.dg, S.DC, S.sD);
      else k_b2b<TFNAS_ACT_SWISH><<<gpl, NT, 0, st>>>(P, D, bn2, seg, S.dg, S.DC, S.sD); }
    if (dweights) {
      // the SE weight gradients only read what se_bwd left (sede, sedt) and saved state: beside b2b, on the side stream
      if (side) { cudaEventRecord(side->fork2, st); cudaStreamWaitEvent(wst, side->fork2, 0); }
      for (int s = 0; s < P.na; ++s) {
        const Cand& cd = P.c[s];
        if (!cd.se) continue;
        const int big = max(cd.mc, cd.se);
        dim3 gwg(cdiv(big, FC_TN), cdiv(big, FC_TO), 2);
        ProfScope ps("se_wgrad", 8.0 * cd.mc * cd.se, 4.0 * P.N * cd.mc * cd.se, wst);
        if (relu) k_se_wgrad<TFNAS_ACT_RELU><<<gwg, NT, 0, wst>>>(P, s, S.sede, S.sedt, sep, set, dweights[cd.id]);
        else k_se_wgrad<TFNAS_ACT_SWISH><<<gwg, NT, 0, wst>>>(P, s, S.sede, S.sedt, sep, set, dweights[cd.id]);
      }
    }
  }
  // B3a
  if (dweights)
    for (int s = 0; s < P.na; ++s)
      cudaMemsetAsync(dweights[P.c[s].id].dw, 0, (size_t)P.c[s].mc * P.c[s].k * P.c[s].k * sizeof(float), st);
  static const bool dw_tile = getenv("TFNAS_DW") && strcmp(getenv("TFNAS_DW"), "tile") == 0;
  if (!dw_tile && dws_supported(P)) {
    cudaStream_t dwst = st;
    if (side && P.stride == 2) {     // stride 2: dDW is a kernel of its own beside the transposed depthwise (DC, sD are final)
      cudaEventRecord(side->fork2, st);
      cudaStreamWaitEvent(wst, side->fork2, 0);
      dwst = wst;
    }
    launch_dws_bwd(P, S.DC, D, bn2, S.sD, UH, S.DA, dweights, st, dwst);
  } else if (P.stride == 1) {
    launch_dw_bwd<3, 1>(P, S.DC, D, bn2, S.sD, UH, S.DA, dweights, st);
    launch_dw_bwd<5, 1>(P, S.DC, D, bn2, S.sD, UH, S.DA, dweights, st);
  } else {
    launch_dw_bwd<3, 2>(P, S.DC, D, bn2, S.sD, UH, S.DA, dweights, st);
    launch_dw_bwd<5, 2>(P, S.DC, D, bn2, S.sD, UH, S.DA, dweights, st);
  }
  if (stop_at_da) return;      // S.DA = dL/d act(UH); the caller finishes (first-stem BN backward + conv weight gradient)
  if (side) {
    // Smat = sum_p du-hat x^T reads DA / UH / x, final since the transposed depthwise: forked alongside the dx GEMM
    cudaEventRecord(side->fork2, st);
    cudaStreamWaitEvent(wst, side->fork2, 0);
    for (int s = 0; s < P.na; ++s) {
      const Cand& cd = P.c[s];
      float* Sm = S.Smat + (size_t)cd.coff * ic;
      cudaMemsetAsync(Sm, 0, (size_t)cd.mc * ic * sizeof(float), wst);
      umma_wgrad(P, s, 1, S.DA, UH, x, nullptr, nullptr, nullptr, nullptr, nullptr, Sm, wst);
    }
  }
  // B3b
  OcTile Tx = oc_tile(ic, 24);
  if (umma_enabled()) {
    umma_dx(P, WX, CHX, S.DA, UH, bn1, dx, S.sU, st);
  } else {
    int tiles = cdiv(P.P, PW_TPX);
    int total_chunks = 0;
    for (int s = 0; s < P.na; ++s) total_chunks += cdiv(P.c[s].mc, PW_KC);
    int ksplit = max(1, min(total_chunks, cdiv(2 * sm_count(), tiles)));
    if (ksplit > 1) cudaMemsetAsync(dx, 0, (size_t)P.P * ic * sizeof(float), st);
    switch (Tx.TC) {
      case 4: launch_dx<4>(P, Tx, ksplit, S.DA, UH, bn1, dx, S.sU, st); break;
      case 8: launch_dx<8>(P, Tx, ksplit, S.DA, UH, bn1, dx, S.sU, st); break;
      case 12: launch_dx<12>(P, Tx, ksplit, S.DA, UH, bn1, dx, S.sU, st); break;
      case 16: launch_dx<16>(P, Tx, ksplit, S.DA, UH, bn1, dx, S.sU, st); break;
      default: launch_dx<24>(P, Tx, ksplit, S.DA, UH, bn1, dx, S.sU, st); break;
    }
  }
  if (dweights) {   // dW1 via Smat
    if (side) {       // Smat was accumulated on the side stream (forked before the dx GEMM); the finishing kernel needs the
      cudaEventRecord(side->fork2, st);           // BN1-backward sums of the dx GEMM too and then stays on the side stream:
      cudaStreamWaitEvent(wst, side->fork2, 0);   // nothing on the caller's stream reads dW1 (joined at the end of the call)
    }
    for (int s = 0; s < P.na; ++s) {
      const Cand& cd = P.c[s];
      float* Sm = S.Smat + (size_t)cd.coff * ic;
      if (!side) cudaMemsetAsync(Sm, 0, (size_t)cd.mc * ic * sizeof(float), st);
      const int smat_t = umma_enabled();     // tensor-core path writes Smat transposed [ic][mc]
      if (side) {
        // already accumulated on the side stream
      } else if (smat_t) {
        umma_wgrad(P, s, 1, S.DA, UH, x, nullptr, nullptr, nullptr, nullptr, nullptr, Sm, st);
      } else {
        int nsplit = max(1, min(cdiv(P.P, 2048), cdiv(6 * sm_count(), cdiv(cd.mc, WG_T) * cdiv(ic, WG_T))));
        dim3 grid(cdiv(cd.mc, WG_T), cdiv(ic, WG_T), nsplit);
        ProfScope ps("wgrad_w1", 4.0 * P.P * (2.0 * cd.mc + ic), 2.0 * P.P * (double)ic * cd.mc, st);
        if (relu) k_wgrad<1, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, s, S.DA, UH, x, nullptr, nullptr, nullptr, nullptr, Sm);
        else k_wgrad<1, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, s, S.DA, UH, x, nullptr, nullptr, nullptr, nullptr, Sm);
      }
      { ProfScope ps("w1fin", 12.0 * cd.mc * ic, 2.0 * cd.mc * ic * ic, wst);
        k_w1fin<<<cdiv(cd.mc, W1F_CH), NT, (size_t)W1F_CH * ic * 4, wst>>>(P, s, Sm, smat_t, bn1, S.sU, xmom, dweights[cd.id].w1); }
    }
  }
  // B4
  {
    { ProfScope ps("b4mm", 4.0 * P.MC * ic, 2.0 * P.MC * ic * ic, st);
      const int kt = cdiv(ic, B4_T);
      int nsplit = max(1, min(cdiv(P.MC, 2 * B4_KC), cdiv(2 * sm_count(), kt * kt)));
      k_b4mm<<<dim3(kt, kt, nsplit), NT, 0, st>>>(P, bn1, S.sU, 1.0 / (double)P.P, S.Mm, S.cvec2); }   // Mm and cvec (k' tiles of row 0)
    // few pixel tiles (14x14 / 7x7 planes): split the output channels over several CTAs per tile so the grid covers the SMs
    OcTile Tf = Tx;
    {
      const int tiles = cdiv(P.P, PW_TPX);
      const int tcs[4] = {16, 12, 8, 4};
      for (int t = 0; t < 4 && tiles * Tf.nchunk < 2 * sm_count(); ++t)
        if (tcs[t] < Tf.TC) Tf = oc_tile(ic, tcs[t]);
    }
    switch (Tf.TC) {
      case 4: launch_dxfin<4>(P, Tf, x, S.Mm, S.cvec2, xmom, dout, dx, st); break;
      case 8: launch_dxfin<8>(P, Tf, x, S.Mm, S.cvec2, xmom, dout, dx, st); break;
      case 12: launch_dxfin<12>(P, Tf, x, S.Mm, S.cvec2, xmom, dout, dx, st); break;
      case 16: launch_dxfin<16>(P, Tf, x, S.Mm, S.cvec2, xmom, dout, dx, st); break;
      default: launch_dxfin<24>(P, Tf, x, S.Mm, S.cvec2, xmom, dout, dx, st); break;
    }
  }
  if (alpha_mode && dlog_alphas) {
    ProfScope ps("alpha_grad", 128, 0, st);
    k_alpha_grad<<<1, 32, 0, st>>>(P.num_ops, mixw, latsave, S.dmix, dlat, T, dlog_alphas);
  }
  if (side) {         // every weight gradient is written and the workspace may be reused once the side stream has joined
    cudaEventRecord(side->join, wst);
    cudaStreamWaitEvent(st, side->join, 0);
  }
}
