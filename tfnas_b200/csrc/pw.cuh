// Building blocks of the flattened-pixel pointwise (1x1 conv) GEMM kernels.
//
// Every 1x1 convolution on the path (expand, project, and their backward mirrors) is
//      Out[o][p] = sum_k W[o][k] * In[k][p]          p = flattened (n, h, w) pixel
// computed by a 256-thread CTA on a tile of PW_TPX = 128 pixels x (8 warps * TC) output channels,
// with K streamed through shared memory in chunks of PW_KC = 32 rows.  lane -> 4 consecutive pixels,
// warp -> TC consecutive output channels, so weight reads are warp-uniform broadcasts and input
// reads are conflict-free float4.  Each kernel supplies its own "stage one input row" prologue
// (BN/act/gate applied on load) and its own epilogue (BN statistics, stores).
#pragma once
#include "common.cuh"

#define PW_TPX 128
#define PW_KC 32
#define PW_LDP (PW_TPX + 4)

struct Px4 {
  int n[4];
  int hw[4];
  bool v[4];
  bool vec;   // 4 pixels valid, same image, 16B aligned
};

__device__ __forceinline__ void px_decomp(Px4& px, int p0, int total, int HW) {
  // one reciprocal division for the first pixel, the other three follow incrementally (p0 < 2^23)
  int n = 0, hw = 0;
  if (p0 < total) { n = fast_div(p0, HW, __frcp_rn((float)HW)); hw = p0 - n * HW; }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    px.v[e] = p0 + e < total;
    px.n[e] = px.v[e] ? n : 0;
    px.hw[e] = px.v[e] ? hw : 0;
    if (++hw == HW) { hw = 0; ++n; }
  }
  px.vec = ((HW & 3) == 0) && px.v[0];   // p0 % 4 == 0 and total % 4 == 0 then
}

// load 4 pixels of plane `ch` of a [N][C][HW] tensor
__device__ __forceinline__ void load4(float (&d)[4], const float* __restrict__ T, const Px4& px, int C, int ch, int HW) {
  if (px.vec) {
    float4 t = *(const float4*)(T + ((size_t)px.n[0] * C + ch) * HW + px.hw[0]);
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) d[e] = px.v[e] ? T[((size_t)px.n[e] * C + ch) * HW + px.hw[e]] : 0.f;
  }
}
__device__ __forceinline__ void store4(float* __restrict__ T, const float (&d)[4], const Px4& px, int C, int ch, int HW) {
  if (px.vec) {
    *(float4*)(T + ((size_t)px.n[0] * C + ch) * HW + px.hw[0]) = make_float4(d[0], d[1], d[2], d[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (px.v[e]) T[((size_t)px.n[e] * C + ch) * HW + px.hw[e]] = d[e];
  }
}
__device__ __forceinline__ void atomic_add4(float* __restrict__ T, const float (&d)[4], const Px4& px, int C, int ch, int HW) {
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (px.v[e]) atomicAdd(&T[((size_t)px.n[e] * C + ch) * HW + px.hw[e]], d[e]);
}

// weights stored [out][K] (row stride ldw): ws[kk][o] = Wg[(o0+o)*ldw + k0+kk]
template <int TC>
__device__ __forceinline__ void stage_w_t(float* ws, const float* __restrict__ Wg, int ldw, int o0, int no, int k0, int nk) {
  constexpr int WLD = 8 * TC + 4;
  for (int i = threadIdx.x; i < PW_KC * 8 * TC; i += NT) {
    int o = i / PW_KC, kk = i - o * PW_KC;
    float w = 0.f;
    if (o < no && kk < nk) w = Wg[(size_t)(o0 + o) * ldw + k0 + kk];
    ws[kk * WLD + o] = w;
  }
}
// weights stored [K][out] (row stride ldw): ws[kk][o] = Wg[(k0+kk)*ldw + o0+o]
template <int TC>
__device__ __forceinline__ void stage_w_n(float* ws, const float* __restrict__ Wg, int ldw, int o0, int no, int k0, int nk) {
  constexpr int WLD = 8 * TC + 4;
  for (int i = threadIdx.x; i < PW_KC * 8 * TC; i += NT) {
    int kk = i / (8 * TC), o = i - kk * (8 * TC);
    float w = 0.f;
    if (o < no && kk < nk) w = Wg[(size_t)(k0 + kk) * ldw + o0 + o];
    ws[kk * WLD + o] = w;
  }
}

// acc[j][e] += sum_kk ws[kk][warp*TC + j] * ins[kk][lane*4 + e]
template <int TC>
__device__ __forceinline__ void pw_mma(float (&acc)[TC][4], const float* ins, const float* ws, int lane, int warp) {
  constexpr int WLD = 8 * TC + 4;
  const float* cp = ins + lane * 4;
  const float* wp = ws + warp * TC;
#pragma unroll 4
  for (int kk = 0; kk < PW_KC; ++kk) {
    float4 xv = *(const float4*)(cp + kk * PW_LDP);
#pragma unroll
    for (int j = 0; j < TC; j += 4) {
      float4 wv = *(const float4*)(wp + kk * WLD + j);
      acc[j][0] += wv.x * xv.x; acc[j][1] += wv.x * xv.y; acc[j][2] += wv.x * xv.z; acc[j][3] += wv.x * xv.w;
      acc[j + 1][0] += wv.y * xv.x; acc[j + 1][1] += wv.y * xv.y; acc[j + 1][2] += wv.y * xv.z; acc[j + 1][3] += wv.y * xv.w;
      acc[j + 2][0] += wv.z * xv.x; acc[j + 2][1] += wv.z * xv.y; acc[j + 2][2] += wv.z * xv.z; acc[j + 2][3] += wv.z * xv.w;
      acc[j + 3][0] += wv.w * xv.x; acc[j + 3][1] += wv.w * xv.y; acc[j + 3][2] += wv.w * xv.z; acc[j + 3][3] += wv.w * xv.w;
    }
  }
}

// output-channel tiling of a pointwise GEMM: `nchunk` CTAs along the channel axis, each `occ`
// channels, handled by `ng` warps of TC channels.
struct OcTile { int occ, nchunk, ng, TC; };
static inline OcTile oc_tile(int oc, int maxTC) {
  OcTile best; best.TC = 0;
  double bestu = -1;
  const int tcs[5] = {4, 8, 12, 16, 24};
  for (int t = 0; t < 5; ++t) {
    int TC = tcs[t];
    if (TC > maxTC) break;
    int nchunk = cdiv(oc, 8 * TC);
    int occ = cdiv(cdiv(oc, nchunk), TC) * TC;
    int ng = occ / TC;
    // utilisation of the 8 warps' accumulators, mild preference for fewer chunks (less input re-read)
    double u = (double)oc / ((double)nchunk * 8 * TC) - 0.02 * nchunk;
    if (u > bestu + 1e-9) { bestu = u; best.occ = occ; best.nchunk = nchunk; best.ng = ng; best.TC = TC; }
  }
  return best;
}
