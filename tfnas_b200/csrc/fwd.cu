// Forward phases F0..F4 of one MixedOP call (see DESIGN.md "Kernels").
//
//  F0  k_xsum, k_xcov, k_xfin   input moments mean_x[ic], cov_x[ic,ic]           (x read twice)
//      k_bn1                    analytic BN1 statistics  mu1 = W1 mu_x, v1 = diag(W1 cov W1^T)
//  F1a k_expand<TC>             expand 1x1 (K=ic) + BN1 normalisation -> UH
//  F1b k_dw_fwd<KS,S,ACT>       depthwise KSxKS/S over act(UH) -> D, BN2 sums
//  F2  k_se_pool, k_se_fc       SE squeeze (mean of act(BN2(d))) and the two FCs -> gate g
//  F3  k_project<TC,ACT>           c = act(BN2(d))*g ; z = W3 c -> Z, BN3 sums
//  F4  k_f4prep, k_f4           Gumbel-softmax mixing weights, latency dot, out = sum_i w_i BN3(z_i) (+x)
//
// Reference arithmetic: models/layers.py:539-561, models/model_search.py:86-91.
#include "kernels.h"
#include "pw.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// ----------------------------------------------------------------------------------------------
// F0: input moments
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_xsum(Plan P, const float* __restrict__ x, double* __restrict__ xsum) {
  const int k = blockIdx.x;
  const int nsplit = gridDim.y, sp = blockIdx.y;
  const int n0 = (int)((long long)P.N * sp / nsplit), n1 = (int)((long long)P.N * (sp + 1) / nsplit);
  double acc = 0.0;
  for (int n = n0; n < n1; ++n) {
    const float* p = x + ((size_t)n * P.ic + k) * P.HW;
    float s = 0.f;
    for (int i = threadIdx.x; i < P.HW; i += NT) s += p[i];
    acc += (double)s;
  }
  __shared__ double red[NT / 32];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < NT / 32; ++i) t += red[i];
    atomicAdd(&xsum[k], t);
  }
}

// centred second moment over flattened pixels p = (n, hw); persistent CTAs walk tiles of XC_TPX pixels and keep
// their partial ic x ic blocks in registers, so each CTA issues one double atomic per entry.
// smem: xs[icp][XC_LD] + mean[icp] (+ red[nblk*16] when several thread groups share a block)
#define XC_TPX 128
#define XC_LD (XC_TPX + 4)
__global__ void __launch_bounds__(NT) k_xcov(Plan P, const float* __restrict__ x,
                                              const double* __restrict__ xsum, double* __restrict__ xcov) {
  extern __shared__ float sm[];
  const int ic = P.ic, nb = (ic + 3) >> 2, icp = nb * 4, nblk = nb * nb;
  float* xs = sm;                      // [icp][XC_LD]
  float* mean = xs + icp * XC_LD;      // [icp]
  float* red = mean + icp;             // [nblk*16] when nsplit > 1
  const int tid = threadIdx.x;
  for (int k = tid; k < icp; k += NT) mean[k] = k < ic ? (float)(xsum[k] / (double)P.P) : 0.f;
  const int nsplit = nblk >= NT ? 1 : NT / nblk;
  const int myblk = nsplit == 1 ? tid : tid % nblk;
  const int mysp = nsplit == 1 ? 0 : tid / nblk;
  const bool active = nsplit == 1 ? true : (tid < nblk * nsplit);
  const int nloops = nsplit == 1 ? (nblk + NT - 1) / NT : 1;
  float acc[9][16];
#pragma unroll
  for (int l = 0; l < 9; ++l)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[l][e] = 0.f;
  const int ntiles = (P.P + XC_TPX - 1) / XC_TPX;
  const float inv_hw = 1.f / (float)P.HW;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * XC_TPX;
    __syncthreads();
    // thread -> pixel (tid & 127), channel rows (tid >> 7) + 2 i : coalesced along pixels
    {
      const int pp = tid & (XC_TPX - 1);
      const int p = p0 + pp;
      const bool pv = p < P.P;
      const int n = pv ? fast_div(p, P.HW, inv_hw) : 0, hw = p - n * P.HW;
      const float* src = x + (size_t)n * ic * P.HW + hw;
      for (int k = tid >> 7; k < icp; k += NT / XC_TPX)
        xs[k * XC_LD + pp] = (pv && k < ic) ? src[(size_t)k * P.HW] - mean[k] : 0.f;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int l = 0; l < 9; ++l) {
        if (l < nloops) {
          int blk = myblk + l * NT;
          if (blk < nblk) {
            int bi = blk / nb, bj = blk - bi * nb;
            if (bj >= bi) {
              const float* xi = xs + (bi * 4) * XC_LD;
              const float* xj = xs + (bj * 4) * XC_LD;
              for (int q = mysp; q < XC_TPX / 4; q += nsplit) {
                float4 a[4], b[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  a[r] = *(const float4*)(xi + r * XC_LD + q * 4);
                  b[r] = *(const float4*)(xj + r * XC_LD + q * 4);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    acc[l][r * 4 + c] += a[r].x * b[c].x + a[r].y * b[c].y + a[r].z * b[c].z + a[r].w * b[c].w;
              }
            }
          }
        }
      }
    }
  }
  // reduce across splits (if any) through shared memory, then one double atomic per entry per CTA
  __syncthreads();
  if (nsplit > 1) {
    for (int i = tid; i < nblk * 16; i += NT) red[i] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
      for (int e = 0; e < 16; ++e) atomicAdd(&red[myblk * 16 + e], acc[0][e]);
    }
    __syncthreads();
    for (int i = tid; i < nblk * 16; i += NT) {
      int blk = i >> 4, e = i & 15;
      int bi = blk / nb, bj = blk - bi * nb;
      if (bj < bi) continue;
      int r = bi * 4 + (e >> 2), c = bj * 4 + (e & 3);
      if (r < ic && c < ic) {
        double v = (double)red[i];
        atomicAdd(&xcov[r * ic + c], v);
        if (bj > bi) atomicAdd(&xcov[c * ic + r], v);
      }
    }
  } else {
#pragma unroll
    for (int l = 0; l < 9; ++l) {
      if (l < nloops) {
        int blk = myblk + l * NT;
        if (blk < nblk) {
          int bi = blk / nb, bj = blk - bi * nb;
          if (bj >= bi) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              int r = bi * 4 + (e >> 2), c = bj * 4 + (e & 3);
              if (r < ic && c < ic) {
                double v = (double)acc[l][e];
                atomicAdd(&xcov[r * ic + c], v);
                if (bj > bi) atomicAdd(&xcov[c * ic + r], v);
              }
            }
          }
        }
      }
    }
  }
}

// ONE pass over x for both input moments: S1[k] = sum_p x_k and S2[k][l] = sum_p x_k x_l (uncentred; centred in fp64 by
// k_xfin_raw, like the BN2 / BN3 statistics).  A CTA walks tiles of TP pixels: the tile is staged in shared memory with a row
// of ones appended (the Gram matrix's last column is then S1), threads own 4x4 blocks of the upper triangle (several pixel
// slices per block when there are fewer blocks than threads, several blocks per thread when there are more) and keep their
// partial sums in registers across tiles; one fp64 atomic per entry and CTA at the end.
// XM_MAXB = blocks per thread: ic <= 192 -> 49 x 50 / 2 = 1225 blocks / 256 threads = 5; the narrow instantiations (1 block
// for ic <= 87, 2 for ic <= 123) keep the register count low enough for 3-4 CTAs per SM
template <int XM_MAXB>
__global__ void __launch_bounds__(NT) k_xmom(Plan P, const float* __restrict__ x, int TP, int tp_shift,
                                              float* __restrict__ xpart) {
  extern __shared__ float xs[];        // [icp][LD], LD = TP + 2 (even: pixel pairs load as float2; rows 2 banks apart)
  const int ic = P.ic, nb = (ic + 1 + 3) >> 2, icp = nb * 4, nut = nb * (nb + 1) / 2, LD = TP + 2;
  const int tid = threadIdx.x;
  const int nsplit = nut >= NT ? 1 : NT / nut;
  const int sp = nut >= NT ? 0 : tid / nut;
  const bool active = nut >= NT ? true : tid < nut * nsplit;
  // gridDim.y > 1 (few pixel tiles, many blocks: the 7x7 stages): the CTAs of one tile slot share the upper triangle,
  // blocks [q_lo, q_hi) each, and write disjoint parts of the same xpart slice
  const int nper = (nut + gridDim.y - 1) / gridDim.y;
  const int q_lo = blockIdx.y * nper, q_hi = min(nut, q_lo + nper);
  int bi[XM_MAXB], bj[XM_MAXB];
#pragma unroll
  for (int l = 0; l < XM_MAXB; ++l) {
    int q = q_lo + (nut >= NT ? tid : tid % nut) + l * NT;
    bi[l] = -1; bj[l] = 0;
    if (active && q < q_hi && (l == 0 || nut > NT)) {
      int r = 0, rowlen = nb;
      while (q >= rowlen) { q -= rowlen; ++r; --rowlen; }
      bi[l] = r; bj[l] = r + q;
    }
  }
  float acc[XM_MAXB][16];
#pragma unroll
  for (int l = 0; l < XM_MAXB; ++l)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[l][e] = 0.f;
  const int ntiles = (P.P + TP - 1) >> tp_shift;
  const float inv_hw = 1.f / (float)P.HW;
  const bool vec = (P.HW & 3) == 0 && ((((uintptr_t)x) & 15) == 0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile << tp_shift;
    __syncthreads();
    if (vec) {
      // TP / 4 <= NT: a thread keeps its 4 pixels for the whole tile and walks the channel rows (no per-element division)
      const int q4s = tp_shift - 2, pp = (tid & ((1 << q4s) - 1)) << 2, p = p0 + pp;
      const bool pv = p < P.P;           // P.P % 4 == 0 here, so the four pixels are valid and in the same image
      const int n = pv ? fast_div(p, P.HW, inv_hw) : 0, hw = p - n * P.HW;
      const float* src = x + (size_t)n * ic * P.HW + hw;
      for (int k = tid >> q4s; k < icp; k += NT >> q4s) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pv) {
          if (k < ic) v = *(const float4*)(src + (size_t)k * P.HW);
          else if (k == ic) v = make_float4(1.f, 1.f, 1.f, 1.f);
        }
        float* d = xs + k * LD + pp;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      for (int i = tid; i < icp * TP; i += NT) {
        const int k = i >> tp_shift, pp = i & (TP - 1), p = p0 + pp;
        float v = 0.f;
        if (p < P.P) {
          if (k < ic) {
            const int n = fast_div(p, P.HW, inv_hw), hw = p - n * P.HW;
            v = x[((size_t)n * ic + k) * P.HW + hw];
          } else if (k == ic) v = 1.f;
        }
        xs[k * LD + pp] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < XM_MAXB; ++l) {
      if (bi[l] < 0) continue;
      const float* ra = xs + (bi[l] * 4) * LD;
      const float* rb = xs + (bj[l] * 4) * LD;
      for (int pp = 2 * sp; pp < TP; pp += 2 * nsplit) {       // two pixels per iteration: 8 vector loads for 32 FMAs
        float2 a[4], b[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { a[e] = *(const float2*)(ra + e * LD + pp); b[e] = *(const float2*)(rb + e * LD + pp); }
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[l][e] += a[e >> 2].x * b[e & 3].x + a[e >> 2].y * b[e & 3].y;
      }
    }
  }
  // the CTA's partial Gram blocks go to its slice of xpart ([cta][block][16], plain stores): k_xred sums the slices in
  // fp64.  (One fp64 atomic per entry and CTA measured 40-60 us per launch: 2-3 M atomics at ic >= 112.)
  float* mine = xpart + (size_t)blockIdx.x * nut * 16;
  if (nsplit > 1) {
    __syncthreads();
    float* red = xs;                 // [nsplit][nut][16] <= 256 * 16 floats
    if (active)
#pragma unroll
      for (int e = 0; e < 16; ++e) red[(sp * nut + (tid % nut)) * 16 + e] = acc[0][e];
    __syncthreads();
    for (int i = tid; i < nut * 16; i += NT) {
      float t = 0.f;
      for (int s2 = 0; s2 < nsplit; ++s2) t += red[s2 * nut * 16 + i];
      mine[i] = t;
    }
  } else {
#pragma unroll
    for (int l = 0; l < XM_MAXB; ++l) {
      if (bi[l] < 0) continue;
      const int q = q_lo + tid + l * NT;
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4)
        *(float4*)(mine + (size_t)q * 16 + e4 * 4) = make_float4(acc[l][e4 * 4], acc[l][e4 * 4 + 1], acc[l][e4 * 4 + 2], acc[l][e4 * 4 + 3]);
    }
  }
}

// sum the per-CTA partial blocks (fp64) and scatter them into S1[k] (the ones column) and the full symmetric S2[k][l].
// Eight threads per entry walk the CTA slices with stride 8 (a single thread per entry spent 27 us in ~300 dependent loads).
__global__ void __launch_bounds__(NT) k_xred(int ic, int nctas, const float* __restrict__ xpart, double* __restrict__ xsum,
                                              double* __restrict__ xx) {
  const int nb = (ic + 1 + 3) >> 2, nut = nb * (nb + 1) / 2;
  const int i = (blockIdx.x * NT + threadIdx.x) >> 3, sub = threadIdx.x & 7;
  const bool ok = i < nut * 16;
  double t = 0.0;
  if (ok)
    for (int c = sub; c < nctas; c += 8) t += (double)xpart[(size_t)c * nut * 16 + i];
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 4);
  if (!ok || sub != 0) return;
  int q = i >> 4, e = i & 15, rb = 0, rowlen = nb;
  while (q >= rowlen) { q -= rowlen; ++rb; --rowlen; }
  const int cb = rb + q;
  const int r = rb * 4 + (e >> 2), c = cb * 4 + (e & 3);
  if (cb == rb && (e >> 2) > (e & 3)) return;          // diagonal block: upper triangle only
  if (r < ic && c < ic) {
    xx[r * ic + c] = t;
    if (r != c) xx[c * ic + r] = t;
  } else if (r < ic && c == ic) {
    xsum[r] = t;
  }
}

// normalise the accumulators into mean / biased covariance (double, kept in `saved`)
__global__ void k_xfin(int ic, int Pn, const double* __restrict__ xsum, const double* __restrict__ xcov,
                       double* __restrict__ xmom) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double inv = 1.0 / (double)Pn;
  if (i < ic) xmom[i] = xsum[i] * inv;
  if (i < ic * ic) xmom[ic + i] = xcov[i] * inv;
}
// same from UNCENTRED sums (tensor-core one-pass moments): cov = E[x x^T] - mean mean^T, in double
__global__ void k_xfin_raw(int ic, int Pn, const double* __restrict__ xsum, const double* __restrict__ xx,
                           double* __restrict__ xmom) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double inv = 1.0 / (double)Pn;
  if (i < ic) xmom[i] = xsum[i] * inv;
  if (i < ic * ic) {
    const int r = i / ic, c = i - r * ic;
    xmom[ic + i] = xx[i] * inv - (xsum[r] * inv) * (xsum[c] * inv);
  }
}

// analytic BN1 statistics: mu1 = W1 mu_x, v1 = w^T cov w.  8 stacked mid channels per CTA; every thread walks its
// share of the ic x ic covariance once (coalesced, fp64) for all 8 channels.
#define BN1_CH 8
__global__ void __launch_bounds__(NT) k_bn1(Plan P, const double* __restrict__ xmom, float* __restrict__ bn1) {
  extern __shared__ float wsm[];       // [BN1_CH][ic]
  const int ic = P.ic, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.x * BN1_CH;
  for (int i = tid; i < BN1_CH * ic; i += NT) {
    const int cc = i / ic, k = i - cc * ic, cst = c0 + cc;
    float w = 0.f;
    if (cst < P.MC) {
      int s = 0;
      while (s + 1 < P.na && cst >= P.c[s + 1].coff) ++s;
      w = P.c[s].w1[(size_t)(cst - P.c[s].coff) * ic + k];
    }
    wsm[i] = w;
  }
  __syncthreads();
  const double* mean = xmom;
  const double* cov = xmom + ic;
  double acc[BN1_CH];
#pragma unroll
  for (int c = 0; c < BN1_CH; ++c) acc[c] = 0.0;
  const float inv_ic = 1.f / (float)ic;
  for (int idx = tid; idx < ic * ic; idx += NT) {
    const int k = fast_div(idx, ic, inv_ic), j = idx - k * ic;
    const double cv = cov[idx];
#pragma unroll
    for (int c = 0; c < BN1_CH; ++c) acc[c] += (double)(wsm[c * ic + k] * wsm[c * ic + j]) * cv;
  }
  __shared__ double red[NT / 32][BN1_CH];
#pragma unroll
  for (int c = 0; c < BN1_CH; ++c) {
    const double t = warp_sum_d(acc[c]);
    if (lane == 0) red[warp][c] = t;
  }
  __syncthreads();
  if (warp < BN1_CH && c0 + warp < P.MC) {      // warp c: finish channel c
    double mu = 0.0;
    for (int k = lane; k < ic; k += 32) mu += (double)wsm[warp * ic + k] * mean[k];
    mu = warp_sum_d(mu);
    if (lane == 0) {
      double v = 0.0;
      for (int w = 0; w < NT / 32; ++w) v += red[w][warp];
      bn1[c0 + warp] = (float)mu;
      bn1[P.MC + c0 + warp] = (float)(1.0 / sqrt(fmax(v, 0.0) + (double)BN_EPS));
    }
  }
}

// sums -> (mean, rstd)
__global__ void k_bnfin(int C, double invM, const double* __restrict__ st, float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = st[2 * c] * invM;
  double v = st[2 * c + 1] * invM - m * m;
  out[c] = (float)m;
  out[C + c] = (float)(1.0 / sqrt(fmax(v, 0.0) + (double)BN_EPS));
}

// ----------------------------------------------------------------------------------------------
// F1a: expand 1x1 (K = ic) + BN1 normalisation -> UH (pre-activation, normalised)
// ----------------------------------------------------------------------------------------------
template <int TC>
__global__ void __launch_bounds__(NT) k_expand(Plan P, const float* __restrict__ x, const float* __restrict__ bn1,
                                                float* __restrict__ UH) {
  __shared__ __align__(16) float ins[PW_KC * PW_LDP];
  __shared__ __align__(16) float ws[PW_KC * (8 * TC + 4)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Cand& cd = P.c[blockIdx.z];
  const int o0 = blockIdx.y * 8 * TC;
  if (o0 >= cd.mc) return;
  const int no = min(8 * TC, cd.mc - o0);
  Px4 px;
  px_decomp(px, blockIdx.x * PW_TPX + lane * 4, P.P, P.HW);
  float acc[TC][4];
#pragma unroll
  for (int j = 0; j < TC; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  for (int k0 = 0; k0 < P.ic; k0 += PW_KC) {
    const int nk = min(PW_KC, P.ic - k0);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PW_KC / 8; ++i) {
      const int kk = warp + i * 8;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      if (kk < nk) load4(d, x, px, P.ic, k0 + kk, P.HW);
      *(float4*)(ins + kk * PW_LDP + lane * 4) = make_float4(d[0], d[1], d[2], d[3]);
    }
    stage_w_t<TC>(ws, cd.w1, P.ic, o0, no, k0, nk);
    __syncthreads();
    if (warp * TC < no) pw_mma<TC>(acc, ins, ws, lane, warp);
  }
#pragma unroll
  for (int j = 0; j < TC; ++j) {
    const int c = o0 + warp * TC + j;
    if (c < cd.mc) {
      const float mu = bn1[cd.coff + c], r = bn1[P.MC + cd.coff + c];
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (acc[j][e] - mu) * r;
      store4(UH, o, px, P.MC, cd.coff + c, P.HW);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// F1b: depthwise KSxKS stride S over act(UH) -> D, BN2 sums
// ----------------------------------------------------------------------------------------------

static DwCfg dw_config(const Plan& P, int KS) {
  // tile over OUTPUT rows; CPB channels per CTA; the `a` tile is [CPB][IR][WP] floats (+ slack for over-reads)
  DwCfg c;
  const int S = P.stride, pad = KS / 2;
  int CPB = 1;
  while (CPB < 32 && CPB * P.HWo < 2048) CPB <<= 1;
  c.CPB = CPB;
  c.WP = dw_wp(P.W, pad);
  for (int tiles = 1;; ++tiles) {
    int R = cdiv(P.Ho, tiles);
    int IR = (R - 1) * S + KS;
    size_t smem = ((size_t)CPB * IR * c.WP + 16) * 4;
    if (smem <= 40 * 1024 || R == 1) {
      c.R = R; c.IR = IR; c.tiles = cdiv(P.Ho, R); c.smem = smem;
      break;
    }
  }
  return c;
}

static void dw_work(const Plan& P, int KS, int CPB, DwWork& w) {
  w.n = 0;
  w.gstart[0] = 0;
  for (int s = 0; s < P.na; ++s) {
    if (P.c[s].k != KS) continue;
    w.slot[w.n] = s;
    w.gstart[w.n + 1] = w.gstart[w.n] + cdiv(P.c[s].mc, CPB);
    ++w.n;
  }
}

template <int KS, int S, int ACT>
__global__ void __launch_bounds__(NT) k_dw_fwd(Plan P, DwWork Wk, DwCfg cfg, const float* __restrict__ UH,
                                                float* __restrict__ D, double* __restrict__ st2) {
  extern __shared__ __align__(16) float as[];   // [CPB][IR][WP] (+16 slack)
  constexpr int pad = KS / 2;
  const int tid = threadIdx.x, n = blockIdx.z;
  int e = 0;
  while (e + 1 < Wk.n && (int)blockIdx.y >= Wk.gstart[e + 1]) ++e;
  const Cand& cd = P.c[Wk.slot[e]];
  const int CPB = cfg.CPB, IR = cfg.IR, WP = cfg.WP;
  const int cbase = ((int)blockIdx.y - Wk.gstart[e]) * CPB;      // first local channel of this group
  const int nc = min(CPB, cd.mc - cbase);
  const int H = P.H, W = P.W, Ho = P.Ho, Wo = P.Wo;
  const int oy0 = blockIdx.x * cfg.R, oy1 = min(Ho, oy0 + cfg.R);
  const int r_lo = oy0 * S - pad;
  // zero the tile (halo + slack), then stage act(UH) for the rows this tile needs
  for (int i = tid; i < (CPB * IR * WP + 16) / 4; i += NT) ((float4*)as)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  {
    const int vr_lo = max(r_lo, 0), vr_hi = min((oy1 - 1) * S - pad + KS - 1, H - 1);
    const float* src = UH + (((size_t)n * P.MC + cd.coff + cbase) * H + vr_lo) * W;
    stage_planes(as, src, (size_t)H * W, nc, vr_hi - vr_lo + 1, W, IR, WP, vr_lo - r_lo, pad,
                 [](float v, int) { return act_f<ACT>(v); });
  }
  __syncthreads();
  const int TPC = NT / CPB;
  const int cl = tid / TPC, jl = tid - cl * TPC;
  float s1 = 0.f, s2 = 0.f;
  if (cl < nc) {
    float wr[KS * KS];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) wr[i] = cd.dw[(size_t)(cbase + cl) * KS * KS + i];
    const float* ab = as + (size_t)cl * IR * WP;
    float* dp = D + (((size_t)n * P.MC + cd.coff + cbase + cl) * Ho + oy0) * Wo;
    const int gpr = (Wo + 3) >> 2;                 // groups of 4 outputs per row
    const int ngroups = (oy1 - oy0) * gpr;
    const float inv_gpr = 1.f / (float)gpr;
    for (int g = jl; g < ngroups; g += TPC) {
      const int oyl = fast_div(g, gpr, inv_gpr), ox0 = (g - oyl * gpr) * 4;
      float o[4];
      dw_row4<KS, S>(o, ab + (size_t)(oyl * S) * WP + ox0 * S, WP, wr);
      float* q = dp + oyl * Wo + ox0;
      if ((Wo & 3) == 0) {
        *(float4*)q = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s1 += o[j]; s2 += o[j] * o[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ox0 + j < Wo) { q[j] = o[j]; s1 += o[j]; s2 += o[j] * o[j]; }
      }
    }
  }
  // TPC threads share a channel; reduce within the warp (or sub-warp group), one atomic per group
  const int GW = TPC < 32 ? TPC : 32;
  s1 = group_sum(s1, GW);
  s2 = group_sum(s2, GW);
  if ((jl & (GW - 1)) == 0 && cl < nc) {
    atomicAdd(&st2[2 * (cd.coff + cbase + cl)], (double)s1);
    atomicAdd(&st2[2 * (cd.coff + cbase + cl) + 1], (double)s2);
  }
}

// ----------------------------------------------------------------------------------------------
// F2: squeeze-excite
// ----------------------------------------------------------------------------------------------
// one warp per (n, SE-gated stacked channel)
template <int ACT>
__global__ void __launch_bounds__(NT) k_se_pool(Plan P, const float* __restrict__ D, const float* __restrict__ bn2,
                                                 float* __restrict__ sep) {
  const int widx = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  if (widx >= P.MCse) return;
  int s = -1;
  for (int i = 0; i < P.na; ++i)
    if (P.c[i].se > 0 && widx >= P.c[i].soff && widx < P.c[i].soff + P.c[i].mc) s = i;
  const int c = P.c[s].coff + (widx - P.c[s].soff);
  const float mu = bn2[c], r = bn2[P.MC + c];
  const float* d = D + ((size_t)n * P.MC + c) * P.HWo;
  float acc = 0.f;
  for (int i = lane; i < P.HWo; i += 32) acc += act_f<ACT>((d[i] - mu) * r);
  acc = warp_sum(acc);
  if (lane == 0) sep[(size_t)n * P.MCse + widx] = acc / (float)P.HWo;
}

// SE FC 1: t = Wr p + br  (hidden pre-activation, saved).  K = mc is long and the output narrow, so the K axis is
// split over blockIdx.z = slot * ksplit + part and combined with atomics into the zeroed `set` (bias added by part 0).
__global__ void __launch_bounds__(NT) k_se_fc1(Plan P, int ksplit, const float* __restrict__ sep, float* __restrict__ set) {
  const int slot = blockIdx.z / ksplit, part = blockIdx.z - slot * ksplit;
  const Cand& cd = P.c[slot];
  if (cd.se == 0 || (int)blockIdx.y * FC_TO >= cd.se) return;
  const int chunks = (cd.mc + FC_KC - 1) / FC_KC;
  const int ks = min(ksplit, chunks);               // this candidate's split: every part < ks owns at least one chunk
  if (part >= ks) return;
  const int k0 = (chunks * part / ks) * FC_KC, k1 = min(cd.mc, (chunks * (part + 1) / ks) * FC_KC);
  fc_tile<false>(P.N, cd.se, cd.mc, cd.rw,
                 [&](int n, int k) { return sep[(size_t)n * P.MCse + cd.soff + k]; },
                 [&](int n, int o, float a) { atomicAdd(&set[(size_t)n * P.SEH + cd.hoff + o], part == 0 ? a + cd.rb[o] : a); },
                 k0, k1);
}
// SE FC 2: g = sigmoid(We act(t) + be).  grid (N/32, mc/64, na)
template <int ACT>
__global__ void __launch_bounds__(NT) k_se_fc2(Plan P, const float* __restrict__ set, float* __restrict__ seg) {
  const Cand& cd = P.c[blockIdx.z];
  if (cd.se == 0 || (int)blockIdx.y * FC_TO >= cd.mc) return;
  fc_tile<false>(P.N, cd.mc, cd.se, cd.ew,
                 [&](int n, int k) { return act_f<ACT>(set[(size_t)n * P.SEH + cd.hoff + k]); },
                 [&](int n, int o, float a) { seg[(size_t)n * P.MCse + cd.soff + o] = sigmoid_f(a + cd.eb[o]); });
}

// ----------------------------------------------------------------------------------------------
// F3: project 1x1 (K = mc) with the BN2/act/SE-gate prologue fused -> Z, BN3 sums
// ----------------------------------------------------------------------------------------------
template <int TC, int ACT>
__global__ void __launch_bounds__(NT) k_project(Plan P, OcTile T, const float* __restrict__ D,
                                                 const float* __restrict__ bn2, const float* __restrict__ seg,
                                                 float* __restrict__ Zb, double* __restrict__ st3) {
  __shared__ __align__(16) float ins[PW_KC * PW_LDP];
  __shared__ __align__(16) float ws[PW_KC * (8 * TC + 4)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = blockIdx.z;
  const Cand& cd = P.c[slot];
  const int mc = cd.mc, oc = P.oc;
  const int o0 = blockIdx.y * T.occ, no = min(T.occ, oc - o0);
  Px4 px;
  px_decomp(px, blockIdx.x * PW_TPX + lane * 4, P.Q, P.HWo);
  float acc[TC][4];
#pragma unroll
  for (int j = 0; j < TC; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const bool gated = cd.se > 0;
  for (int k0 = 0; k0 < mc; k0 += PW_KC) {
    const int nk = min(PW_KC, mc - k0);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PW_KC / 8; ++i) {
      const int kk = warp + i * 8, k = k0 + kk;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (kk < nk) {
        const int cst = cd.coff + k;
        const float mu = bn2[cst], r = bn2[P.MC + cst];
        float d[4];
        load4(d, D, px, P.MC, cst, P.HWo);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float b = act_f<ACT>((d[e] - mu) * r);
          if (gated) b *= seg[(size_t)px.n[e] * P.MCse + cd.soff + k];
          o[e] = px.v[e] ? b : 0.f;
        }
      }
      *(float4*)(ins + kk * PW_LDP + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
    stage_w_t<TC>(ws, cd.w3, mc, o0, no, k0, nk);
    __syncthreads();
    if (warp < T.ng) pw_mma<TC>(acc, ins, ws, lane, warp);
  }
  if (warp < T.ng) {
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const int o = warp * TC + j;   // warp-uniform
      float s1 = 0.f, s2 = 0.f;
      if (o < no) {
        store4(Zb, acc[j], px, P.na * oc, slot * oc + o0 + o, P.HWo);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (px.v[e]) { s1 += acc[j][e]; s2 += acc[j][e] * acc[j][e]; }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0 && o < no) {
        atomicAdd(&st3[2 * (slot * oc + o0 + o)], (double)s1);
        atomicAdd(&st3[2 * (slot * oc + o0 + o) + 1], (double)s2);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// F4: mixing weights, latency, combine
// ----------------------------------------------------------------------------------------------
// one CTA: w = softmax((log_alpha + g)/T) over ALL num_ops (alpha mode) or 1.0 (sampled mode);
// coef[slot][c] = w * r3 ; bias[c] = -sum w r3 mu3 ; out_lat = sum w lat
// st3 != nullptr: the BN3 sums are finalised here first (mean, rstd -> bn3; one launch instead of k_bnfin + k_f4prep)
__global__ void k_f4prep(Plan P, int alpha_mode, const float* __restrict__ log_alphas,
                         const float* __restrict__ gumbel, const float* __restrict__ lat8, float T,
                         const double* __restrict__ st3, double invQ,
                         float* __restrict__ bn3, float* __restrict__ mixw, float* __restrict__ latsave,
                         float* __restrict__ coef, float* __restrict__ out_lat) {
  __shared__ float w[TFNAS_MAX_OPS];
  const int num_ops = P.num_ops;
  if (threadIdx.x == 0) {
    if (alpha_mode) {
      float l[TFNAS_MAX_OPS], m = -INFINITY;
      for (int i = 0; i < num_ops; ++i) { l[i] = (log_alphas[i] + gumbel[i]) / T; m = fmaxf(m, l[i]); }
      float s = 0.f;
      for (int i = 0; i < num_ops; ++i) { l[i] = expf(l[i] - m); s += l[i]; }
      float lat = 0.f;
      for (int i = 0; i < num_ops; ++i) {
        w[i] = l[i] / s;
        mixw[i] = w[i];
        latsave[i] = lat8[i];
        lat += w[i] * lat8[i];
      }
      *out_lat = lat;
    } else {
      for (int i = 0; i < TFNAS_MAX_OPS; ++i) { w[i] = 1.f; mixw[i] = 1.f; latsave[i] = 0.f; }
    }
  }
  __syncthreads();
  const int oc = P.oc, C3 = P.na * oc;
  for (int c = threadIdx.x; c < oc; c += blockDim.x) {
    float b = 0.f;
    for (int s = 0; s < P.na; ++s) {
      float wi = w[P.c[s].id];
      if (st3) {                       // same arithmetic as k_bnfin
        const double m = st3[2 * (s * oc + c)] * invQ;
        const double v = st3[2 * (s * oc + c) + 1] * invQ - m * m;
        bn3[s * oc + c] = (float)m;
        bn3[C3 + s * oc + c] = (float)(1.0 / sqrt(fmax(v, 0.0) + (double)BN_EPS));
      }
      float r3 = bn3[C3 + s * oc + c], mu3 = bn3[s * oc + c];
      coef[s * oc + c] = wi * r3;
      b -= wi * r3 * mu3;
    }
    coef[C3 + c] = b;
  }
}

// out[n,c,:] = sum_s coef[s,c] * Z[n,s,c,:] + bias[c] (+ x[n,c,:])
__global__ void __launch_bounds__(NT) k_f4(Plan P, const float* __restrict__ Zb, const float* __restrict__ coef,
                                            const float* __restrict__ x, float* __restrict__ out) {
  const int HWo = P.HWo, oc = P.oc, na = P.na;
  const size_t total = (size_t)P.N * oc * HWo;
  if ((HWo & 3) == 0) {
    const size_t nv = total >> 2;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < nv; i += (size_t)gridDim.x * NT) {
      size_t e = i << 2;
      int plane = (int)(e / HWo);
      int hw = (int)(e - (size_t)plane * HWo);
      int n = plane / oc, c = plane - n * oc;
      float b = coef[na * oc + c];
      float4 o = make_float4(b, b, b, b);
      for (int s = 0; s < na; ++s) {
        float cf = coef[s * oc + c];
        float4 z = *(const float4*)(Zb + ((size_t)(n * na + s) * oc + c) * HWo + hw);
        o.x += cf * z.x; o.y += cf * z.y; o.z += cf * z.z; o.w += cf * z.w;
      }
      if (P.residual) {
        float4 r = *(const float4*)(x + e);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *(float4*)(out + e) = o;
    }
  } else {
    for (size_t e = (size_t)blockIdx.x * NT + threadIdx.x; e < total; e += (size_t)gridDim.x * NT) {
      int plane = (int)(e / HWo);
      int hw = (int)(e - (size_t)plane * HWo);
      int n = plane / oc, c = plane - n * oc;
      float o = coef[na * oc + c];
      for (int s = 0; s < na; ++s) o += coef[s * oc + c] * Zb[((size_t)(n * na + s) * oc + c) * HWo + hw];
      if (P.residual) o += x[e];
      out[e] = o;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static int g_sm_count = 0;
int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

template <int KS, int S>
static void launch_dw_fwd(const Plan& P, const float* UH, float* D, double* st2, cudaStream_t st) {
  DwCfg cfg = dw_config(P, KS);
  DwWork w;
  dw_work(P, KS, cfg.CPB, w);
  if (!w.n) return;
  double mck = 0;
  for (int i = 0; i < w.n; ++i) mck += P.c[w.slot[i]].mc;
  dim3 grid(cfg.tiles, w.gstart[w.n], P.N);
  auto kern = P.act == TFNAS_ACT_RELU ? k_dw_fwd<KS, S, TFNAS_ACT_RELU> : k_dw_fwd<KS, S, TFNAS_ACT_SWISH>;
  ensure_smem(kern, (size_t)(cfg.smem));
  ProfScope ps(KS == 3 ? "dw_fwd_k3" : "dw_fwd_k5", 4.0 * mck * ((double)P.P + P.Q), 2.0 * KS * KS * mck * P.Q, st);
  kern<<<grid, NT, cfg.smem, st>>>(P, w, cfg, UH, D, st2);
}

template <int TC>
static void launch_project(const Plan& P, OcTile T, const float* D, const float* bn2, const float* seg, float* Zb,
                           double* st3, cudaStream_t st) {
  dim3 grid(cdiv(P.Q, PW_TPX), T.nchunk, P.na);
  ProfScope ps("project", 4.0 * P.Q * ((double)P.MC + (double)P.na * P.oc) + 4.0 * P.MC * P.oc,
               2.0 * P.Q * (double)P.MC * P.oc, st);
  if (P.act == TFNAS_ACT_RELU)
    k_project<TC, TFNAS_ACT_RELU><<<grid, NT, 0, st>>>(P, T, D, bn2, seg, Zb, st3);
  else
    k_project<TC, TFNAS_ACT_SWISH><<<grid, NT, 0, st>>>(P, T, D, bn2, seg, Zb, st3);
}

void launch_forward(const Plan& P, const float* x, const float* log_alphas, const float* gumbel,
                    const float* lat8, float T, int alpha_mode, float* out, float* out_lat,
                    char* saved, const SavedLayout& L, const FwdScratch& S, cudaStream_t st, const PreppedFwd* pre) {
  double* xmom = (double*)(saved + L.xmom);
  float* bn1 = (float*)(saved + L.bn1);
  float* bn2 = (float*)(saved + L.bn2);
  float* bn3 = (float*)(saved + L.bn3);
  float* mixw = (float*)(saved + L.mixw);
  float* latsave = (float*)(saved + L.lat);
  float* sep = (float*)(saved + L.sep);
  float* set = (float*)(saved + L.set);
  float* seg = (float*)(saved + L.seg);
  float* UH = (float*)(saved + L.UH);
  float* D = (float*)(saved + L.D);
  float* Zb = (float*)(saved + L.Z);
  const int ic = P.ic;
  const double xbytes = 4.0 * P.P * ic;
  // zero the accumulators (xsum, xcov, st2, st3 are contiguous in the workspace)
  size_t zbytes = (size_t)(ic + ic * ic + 2 * P.MC + 2 * P.na * P.oc) * sizeof(double);
  cudaMemsetAsync(S.xsum, 0, zbytes, st);
  // F0
  static const bool xmom_gemm = getenv("TFNAS_XMOM") && strcmp(getenv("TFNAS_XMOM"), "gemm") == 0;   // A/B: the two-pass tensor-core moments
  if (umma_enabled() && !xmom_gemm) {
    // one-pass moments: tile width by channel count so the staged tile stays under 52 KB (4 CTAs per SM)
    const int icp = ((ic + 1 + 3) >> 2) << 2;
    const int tp_shift = icp >= 96 ? 6 : icp >= 48 ? 7 : icp >= 24 ? 8 : 9, TP = 1 << tp_shift;
    const size_t smem = (size_t)icp * (TP + 2) * 4;
    const int tiles = cdiv(P.P, TP);
    const int nb4 = icp >> 2, nut = nb4 * (nb4 + 1) / 2, maxb = cdiv(nut, NT);
    { ProfScope ps("xmom", xbytes, 1.0 * P.P * ic * ic, st);
      const int grid = max(1, min(tiles, min(2 * sm_count(), XM_MAXCTA)));
      // fewer pixel tiles than SMs and several blocks per thread (ic = 192 on 7x7 planes: 98 tiles, 5 blocks per thread):
      // split the blocks of the triangle over gridDim.y CTAs per tile
      int nsb = 1;
      if (grid < sm_count() && maxb > 1) nsb = min(maxb, cdiv(2 * sm_count(), grid));
      const int mb = cdiv(cdiv(nut, nsb), NT);
      const dim3 g2(grid, nsb);
      if (mb <= 1) { ensure_smem(k_xmom<1>, smem); k_xmom<1><<<g2, NT, smem, st>>>(P, x, TP, tp_shift, S.xpart); }
      else if (mb <= 2) { ensure_smem(k_xmom<2>, smem); k_xmom<2><<<g2, NT, smem, st>>>(P, x, TP, tp_shift, S.xpart); }
      else { ensure_smem(k_xmom<5>, smem); k_xmom<5><<<g2, NT, smem, st>>>(P, x, TP, tp_shift, S.xpart); }
      count_launch(1);
      k_xred<<<cdiv(nut * 16 * 8, NT), NT, 0, st>>>(ic, grid, S.xpart, S.xsum, S.xcov); }
    { ProfScope ps("xfin", 16.0 * ic * ic, 0, st);
      k_xfin_raw<<<cdiv(ic * ic, 256), 256, 0, st>>>(ic, P.P, S.xsum, S.xcov, xmom); }
    { ProfScope ps("bn1", 4.0 * P.MC * ic + 8.0 * ic * ic, 2.0 * P.MC * ic * ic, st);
      k_bn1<<<cdiv(P.MC, BN1_CH), NT, (size_t)BN1_CH * ic * 4, st>>>(P, xmom, bn1); }
  } else if (umma_enabled()) {
    int split = max(1, min(P.N, 4 * sm_count() / max(ic, 1)));
    { ProfScope ps("xsum", xbytes, 1.0 * P.P * ic, st);
      k_xsum<<<dim3(ic, split), NT, 0, st>>>(P, x, S.xsum); }
    umma_covariance(P, x, S.xsum, st);
    { ProfScope ps("xfin", 16.0 * ic * ic, 0, st);
      k_xfin<<<cdiv(ic * ic, 256), 256, 0, st>>>(ic, P.P, S.xsum, S.xcov, xmom); }
    { ProfScope ps("bn1", 4.0 * P.MC * ic + 8.0 * ic * ic, 2.0 * P.MC * ic * ic, st);
      k_bn1<<<cdiv(P.MC, BN1_CH), NT, (size_t)BN1_CH * ic * 4, st>>>(P, xmom, bn1); }
  } else {
    int split = max(1, min(P.N, 4 * sm_count() / max(ic, 1)));
    { ProfScope ps("xsum", xbytes, 1.0 * P.P * ic, st);
      k_xsum<<<dim3(ic, split), NT, 0, st>>>(P, x, S.xsum); }
    int nb = (ic + 3) / 4, icp = nb * 4, nblk = nb * nb;
    int ctas = max(1, min(cdiv(P.P, XC_TPX), 4 * sm_count()));
    size_t smem = (size_t)(icp * XC_LD + icp + (nblk < NT ? nblk * 16 : 0)) * 4;
    ensure_smem(k_xcov, (size_t)(smem));
    { ProfScope ps("xcov", xbytes, 1.0 * P.P * ic * ic, st);
      k_xcov<<<ctas, NT, smem, st>>>(P, x, S.xsum, S.xcov); }
    { ProfScope ps("xfin", 16.0 * ic * ic, 0, st);
      k_xfin<<<cdiv(ic * ic, 256), 256, 0, st>>>(ic, P.P, S.xsum, S.xcov, xmom); }
    { ProfScope ps("bn1", 4.0 * P.MC * ic + 8.0 * ic * ic, 2.0 * P.MC * ic * ic, st);
      k_bn1<<<cdiv(P.MC, BN1_CH), NT, (size_t)BN1_CH * ic * 4, st>>>(P, xmom, bn1); }
  }
  // F1a
  UmWAll WE, WP;
  if (umma_enabled()) {
    if (pre) { WE = pre->WE; WP = pre->WP; }
    else umma_prep_fwd(P, S.umprep, WE, WP, st);
    umma_expand(P, WE, x, bn1, UH, st);
  } else {
    int maxmc = 0;
    for (int s = 0; s < P.na; ++s) maxmc = max(maxmc, P.c[s].mc);
    ProfScope ps("expand", xbytes + 4.0 * P.P * P.MC + 4.0 * P.MC * ic, 2.0 * P.P * (double)P.MC * ic, st);
    if (maxmc > 64) {
      dim3 grid(cdiv(P.P, PW_TPX), cdiv(maxmc, 128), P.na);
      k_expand<16><<<grid, NT, 0, st>>>(P, x, bn1, UH);
    } else {
      dim3 grid(cdiv(P.P, PW_TPX), cdiv(maxmc, 64), P.na);
      k_expand<8><<<grid, NT, 0, st>>>(P, x, bn1, UH);
    }
  }
  launch_forward_tail(P, umma_enabled() ? &WP : nullptr, x, log_alphas, gumbel, lat8, T, alpha_mode, out, out_lat, saved, L, S, st);
}

// F1b .. F4 from a given UH = BN1-normalised pre-activation input of the depthwise stage (saved + L.UH) and bn1 (saved + L.bn1).
// The second stem of the supernet (an MBConv without expand conv, models/model_search.py:220) enters here with UH = the
// normalised first-stem convolution.  WP: project weights already prepped (tcgen05 path), nullptr on the SIMT path.
// The BN2 / BN3 accumulators S.st2 / S.st3 must have been zeroed by the caller.
void launch_forward_tail(const Plan& P, const UmWAll* WPp, const float* x, const float* log_alphas, const float* gumbel,
                         const float* lat8, float T, int alpha_mode, float* out, float* out_lat, char* saved,
                         const SavedLayout& L, const FwdScratch& S, cudaStream_t st) {
  float* bn2 = (float*)(saved + L.bn2);
  float* bn3 = (float*)(saved + L.bn3);
  float* mixw = (float*)(saved + L.mixw);
  float* latsave = (float*)(saved + L.lat);
  float* sep = (float*)(saved + L.sep);
  float* set = (float*)(saved + L.set);
  float* seg = (float*)(saved + L.seg);
  float* UH = (float*)(saved + L.UH);
  float* D = (float*)(saved + L.D);
  float* Zb = (float*)(saved + L.Z);
  // F1b
  static const bool dw_tile = getenv("TFNAS_DW") && strcmp(getenv("TFNAS_DW"), "tile") == 0;
  if (!dw_tile && dws_supported(P)) {
    launch_dws_fwd(P, UH, D, S.st2, st);
  } else if (P.stride == 1) {
    launch_dw_fwd<3, 1>(P, UH, D, S.st2, st);
    launch_dw_fwd<5, 1>(P, UH, D, S.st2, st);
  } else {
    launch_dw_fwd<3, 2>(P, UH, D, S.st2, st);
    launch_dw_fwd<5, 2>(P, UH, D, S.st2, st);
  }
  { ProfScope ps("bnfin", 24.0 * P.MC, 0, st);
    k_bnfin<<<cdiv(P.MC, 256), 256, 0, st>>>(P.MC, 1.0 / (double)P.Q, S.st2, bn2); }
  // F2
  if (P.MCse > 0) {
    int maxmc = 0, maxse = 0;
    double fcw = 0;
    for (int s = 0; s < P.na; ++s)
      if (P.c[s].se > 0) { maxmc = max(maxmc, P.c[s].mc); maxse = max(maxse, P.c[s].se); fcw += 2.0 * P.c[s].mc * P.c[s].se; }
    const bool relu = P.act == TFNAS_ACT_RELU;
    { ProfScope ps("se_pool", 4.0 * P.Q * P.MCse, 4.0 * P.Q * P.MCse, st);
      dim3 g(cdiv(P.MCse * 32, NT), P.N);
      if (relu) k_se_pool<TFNAS_ACT_RELU><<<g, NT, 0, st>>>(P, D, bn2, sep);
      else k_se_pool<TFNAS_ACT_SWISH><<<g, NT, 0, st>>>(P, D, bn2, sep); }
    { ProfScope ps("se_fc", 4.0 * fcw + 8.0 * P.N * P.MCse, 2.0 * P.N * fcw, st);
      const int ksplit = max(1, min(8, cdiv(maxmc, 2 * FC_KC)));
      cudaMemsetAsync(set, 0, (size_t)P.N * P.SEH * sizeof(float), st);
      k_se_fc1<<<dim3(cdiv(P.N, FC_TN), cdiv(maxse, FC_TO), P.na * ksplit), NT, 0, st>>>(P, ksplit, sep, set);
      dim3 g2(cdiv(P.N, FC_TN), cdiv(maxmc, FC_TO), P.na);
      if (relu) k_se_fc2<TFNAS_ACT_RELU><<<g2, NT, 0, st>>>(P, set, seg);
      else k_se_fc2<TFNAS_ACT_SWISH><<<g2, NT, 0, st>>>(P, set, seg);
      count_launch(1); }
  }
  // F3
  if (umma_enabled()) {
    umma_project(P, *WPp, D, bn2, seg, Zb, S.st3, st);
  } else {
    OcTile T3 = oc_tile(P.oc, 16);
    switch (T3.TC) {
      case 4: launch_project<4>(P, T3, D, bn2, seg, Zb, S.st3, st); break;
      case 8: launch_project<8>(P, T3, D, bn2, seg, Zb, S.st3, st); break;
      case 12: launch_project<12>(P, T3, D, bn2, seg, Zb, S.st3, st); break;
      default: launch_project<16>(P, T3, D, bn2, seg, Zb, S.st3, st); break;
    }
  }
  // F4 (the BN3 sums are finalised inside k_f4prep)
  { ProfScope ps("f4prep", 36.0 * P.na * P.oc, 0, st);
    k_f4prep<<<1, 256, 0, st>>>(P, alpha_mode, log_alphas, gumbel, lat8, T, S.st3, 1.0 / (double)P.Q, bn3, mixw, latsave, S.coef,
                                out_lat); }
  size_t total = (size_t)P.N * P.oc * P.HWo;
  int blocks = (int)min((size_t)(8 * sm_count()), (total / 4 + NT - 1) / NT);
  { ProfScope ps("combine", 4.0 * total * (P.na + 1 + (P.residual ? 1 : 0)), 2.0 * total * P.na, st);
    k_f4<<<max(blocks, 1), NT, 0, st>>>(P, Zb, S.coef, x, out); }
}
