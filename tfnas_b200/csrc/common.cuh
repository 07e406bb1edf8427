// Shared device/host definitions for the tfnas_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tfnas_b200.h"

#define BN_EPS 1e-5f
#define NT 256  // threads per CTA of every tile kernel
#define SE_NB 8 // images per CTA in the squeeze-excite FC kernels

// One ACTIVE candidate in "slot" order (slot s = s-th set bit of cand_mask).
struct Cand {
  int id;     // original candidate index 0..7
  int mc;     // mid width
  int k;      // 3 / 5
  int se;     // SE hidden width (0 = none)
  int coff;   // offset in the stacked mid-channel space [0, MC)
  int soff;   // offset in the stacked SE-gated mid-channel space [0, MCse)  (valid if se>0)
  int hoff;   // offset in the stacked SE-hidden space [0, SEH)
  int pad_;
  const float *w1, *dw, *w3, *rw, *rb, *ew, *eb;
};

// Per-MixedOP-call plan, passed BY VALUE to kernels.
struct Plan {
  int N, ic, oc, H, W, Ho, Wo, stride, act, na;
  int MC, MCse, SEH;   // stacked widths over active candidates
  int HW, HWo;         // H*W, Ho*Wo
  int P, Q;            // N*H*W, N*Ho*Wo (assumed < 2^31)
  int residual;
  int num_ops;         // candidates of the MixedOP (softmax width in alpha mode)
  Cand c[TFNAS_MAX_OPS];
};

// Depthwise kernels: channel groups (CPB channels each) of the candidates sharing one kernel size.
struct DwWork {
  int n;                  // entries
  int slot[TFNAS_MAX_OPS];
  int gstart[TFNAS_MAX_OPS + 1];
};
struct DwCfg { int CPB, R, IR, WP, tiles; size_t smem; };
struct DwGrads { float* p[4]; };   // depthwise weight-grad pointers per DwWork entry

// MUFU-backed exp2 / reciprocal with flush-to-zero (one instruction each; no denormal-range fix-up code)
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / (1 + e^-x): e^-x overflows to +inf for x < -88 -> 0, underflows to 0 for large x -> 1
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.f + ex2_ftz(x * -1.4426950408889634f)); }

template <int ACT>
__device__ __forceinline__ float act_f(float x) {
  if (ACT == TFNAS_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == TFNAS_ACT_NONE) return x;          // the head's feature-mix conv reads the sink output as is
  return x * sigmoid_f(x);
}
// derivative of the activation at pre-activation x
template <int ACT>
__device__ __forceinline__ float act_df(float x) {
  if (ACT == TFNAS_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if (ACT == TFNAS_ACT_NONE) return 1.f;
  const float s = sigmoid_f(x);
  return s * (1.f + x * (1.f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum 16 per-lane values across the 32 lanes with 16 shuffles (instead of 80): afterwards lane l holds the
// total of v[l & 15] (lanes l and l+16 hold the same channel).
__device__ __forceinline__ float warp_sum16(float (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// sum over aligned groups of `width` lanes (width power of two <= 32)
__device__ __forceinline__ float group_sum(float v, int width) {
  for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// padded row width of a depthwise smem tile: interior [pad, pad+W), zero halo, multiple of 4
static inline int dw_wp(int W, int pad) { return (W + 2 * pad + 3) / 4 * 4; }

// Four consecutive outputs (ox0 % 4 == 0) of a KSxKS stride-S correlation from a zero-haloed smem tile.
// `ap` points at tile(row = oy*S, col = ox0*S) (16B aligned: WP % 4 == 0), wr = weights [KS*KS].
template <int KS, int S>
__device__ __forceinline__ void dw_row4(float (&o)[4], const float* ap, int WP, const float (&wr)[KS * KS]) {
  constexpr int NV = 3 * S + KS;            // input columns touched by 4 outputs
  constexpr int NV4 = (NV + 3) / 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < KS; ++ky) {
    float v[NV4 * 4];
#pragma unroll
    for (int q = 0; q < NV4; ++q) {
      float4 t = *(const float4*)(ap + ky * WP + q * 4);
      v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
    }
#pragma unroll
    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += wr[ky * KS + kx] * v[j * S + kx];
  }
}

// a / b for 0 <= a < 2^23 via a float reciprocal with a +-1 fix-up (a handful of instructions instead of ~20)
__device__ __forceinline__ int fast_div(int a, int b, float inv_b) {
  int q = __float2int_rz(((float)a + 0.5f) * inv_b);
  q -= (q * b > a);
  q += ((q + 1) * b <= a);
  return q;
}

// Stage `nch` channel planes (each `rows` x W floats, contiguous rows; consecutive channels `plane_stride`
// floats apart) into a zero-haloed smem tile [c][IR][WP] at (row_off + r, pad + col), applying f(value, c).
// When the planes are whole and contiguous (plane_stride == rows*W) the copy is one flat vectorised stream.
template <class F>
__device__ __forceinline__ void stage_planes(float* tile, const float* __restrict__ src, size_t plane_stride, int nch,
                                             int rows, int W, int IR, int WP, int row_off, int pad, F f) {
  const int tid = threadIdx.x;
  const int hw = rows * W;
  const float inv_w = 1.f / (float)W, inv_hw = 1.f / (float)hw;
  const bool flat = plane_stride == (size_t)hw;
  const bool vec = ((hw & 3) == 0) && ((plane_stride & 3) == 0) && ((((uintptr_t)src) & 15) == 0);
  if (vec) {
    const int nv = hw >> 2;
    const int total = flat ? nch * nv : nv;
    for (int c0 = 0; c0 < (flat ? 1 : nch); ++c0) {
      for (int i = tid; i < total; i += NT) {
        int c = c0, v = i;
        if (flat) { c = fast_div(i, nv, 1.f / (float)nv); v = i - c * nv; }
        const float4 t = *(const float4*)(src + (size_t)c * plane_stride + (size_t)v * 4);
        int r = fast_div(v * 4, W, inv_w), col = v * 4 - r * W;
        const float x[4] = {t.x, t.y, t.z, t.w};
        float* base = tile + ((size_t)c * IR + row_off) * WP + pad;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          base[r * WP + col] = f(x[e], c);
          if (++col == W) { col = 0; ++r; }
        }
      }
    }
  } else {
    const int total = nch * hw;
    for (int i = tid; i < total; i += NT) {
      const int c = fast_div(i, hw, inv_hw), p = i - c * hw;
      const int r = fast_div(p, W, inv_w), col = p - r * W;
      tile[((size_t)c * IR + row_off + r) * WP + pad + col] = f(src[(size_t)c * plane_stride + p], c);
    }
  }
}

// Same as stage_planes with a second tensor that shares the indexing (src2 = src + delta): f(v1, v2, c).
template <class F>
__device__ __forceinline__ void stage_planes2(float* tile, const float* __restrict__ src, ptrdiff_t delta, size_t plane_stride,
                                              int nch, int rows, int W, int IR, int WP, int row_off, int pad, F f) {
  const int tid = threadIdx.x;
  const int hw = rows * W;
  const float inv_w = 1.f / (float)W, inv_hw = 1.f / (float)hw;
  const bool flat = plane_stride == (size_t)hw;
  const bool vec = ((hw & 3) == 0) && ((plane_stride & 3) == 0) && ((((uintptr_t)src) & 15) == 0) &&
                   ((((uintptr_t)(src + delta)) & 15) == 0);
  if (vec) {
    const int nv = hw >> 2;
    const int total = flat ? nch * nv : nv;
    for (int c0 = 0; c0 < (flat ? 1 : nch); ++c0) {
      for (int i = tid; i < total; i += NT) {
        int c = c0, v = i;
        if (flat) { c = fast_div(i, nv, 1.f / (float)nv); v = i - c * nv; }
        const float* q = src + (size_t)c * plane_stride + (size_t)v * 4;
        const float4 t = *(const float4*)q, u = *(const float4*)(q + delta);
        int r = fast_div(v * 4, W, inv_w), col = v * 4 - r * W;
        const float x[4] = {t.x, t.y, t.z, t.w}, y[4] = {u.x, u.y, u.z, u.w};
        float* base = tile + ((size_t)c * IR + row_off) * WP + pad;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          base[r * WP + col] = f(x[e], y[e], c);
          if (++col == W) { col = 0; ++r; }
        }
      }
    }
  } else {
    const int total = nch * hw;
    for (int i = tid; i < total; i += NT) {
      const int c = fast_div(i, hw, inv_hw), p2 = i - c * hw;
      const int r = fast_div(p2, W, inv_w), col = p2 - r * W;
      const float* q = src + (size_t)c * plane_stride + p2;
      tile[((size_t)c * IR + row_off + r) * WP + pad + col] = f(q[0], q[delta], c);
    }
  }
}

// Small tiled GEMM used by the squeeze-excite path (FCs and their weight gradients):
//   Out[r][o] = sum_{k in [k_begin, k_end)} In(r, k) * W(o, k)        r < R rows, o < Nout outputs
// CTA tile 32 rows x 64 outputs (blockIdx.x, blockIdx.y), K streamed through smem in chunks of 64; thread (ty, tx)
// owns 2 rows x 4 outputs.  The caller supplies in(r, k), w(o, k) and store(r, o, acc).  IN_ROWFAST / W_OFAST say
// which index consecutive threads should walk when staging (pick the one that is contiguous in memory).
#define FC_TN 32
#define FC_TO 64
#define FC_KC 64
struct FcSmem {
  float ins[FC_KC][FC_TN + 1];
  __align__(16) float ws[FC_KC][FC_TO + 4];
};
template <bool IN_ROWFAST, bool W_OFAST, class InF, class WF, class StoreF>
__device__ __forceinline__ void fc_tile_g(FcSmem& sm, int R, int Nout, int k_begin, int k_end, InF in, WF w, StoreF store) {
  float (*ins)[FC_TN + 1] = sm.ins;
  float (*ws)[FC_TO + 4] = sm.ws;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * FC_TN, o0 = blockIdx.y * FC_TO;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = k_begin; k0 < k_end; k0 += FC_KC) {
    __syncthreads();
    for (int i = tid; i < FC_TN * FC_KC; i += NT) {
      int nn, kk;
      if (IN_ROWFAST) { kk = i / FC_TN; nn = i - kk * FC_TN; } else { nn = i / FC_KC; kk = i - nn * FC_KC; }
      ins[kk][nn] = (n0 + nn < R && k0 + kk < k_end) ? in(n0 + nn, k0 + kk) : 0.f;
    }
    for (int i = tid; i < FC_TO * FC_KC; i += NT) {
      int oo, kk;
      if (W_OFAST) { kk = i / FC_TO; oo = i - kk * FC_TO; } else { oo = i / FC_KC; kk = i - oo * FC_KC; }
      ws[kk][oo] = (o0 + oo < Nout && k0 + kk < k_end) ? w(o0 + oo, k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < FC_KC; ++kk) {
      const float a0 = ins[kk][ty * 2], a1 = ins[kk][ty * 2 + 1];
      const float4 wv = *(const float4*)&ws[kk][tx * 4];
      acc[0][0] += a0 * wv.x; acc[0][1] += a0 * wv.y; acc[0][2] += a0 * wv.z; acc[0][3] += a0 * wv.w;
      acc[1][0] += a1 * wv.x; acc[1][1] += a1 * wv.y; acc[1][2] += a1 * wv.z; acc[1][3] += a1 * wv.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty * 2 + i, o = o0 + tx * 4 + j;
      if (n < R && o < Nout) store(n, o, acc[i][j]);
    }
}
// weights from a dense matrix: WKN = stored [k][o] (o contiguous) instead of [o][k]
template <bool WKN, class LoadF, class StoreF>
__device__ __forceinline__ void fc_tile(int N, int Nout, int K, const float* __restrict__ Wg, LoadF load, StoreF store,
                                        int k_begin = 0, int k_end = -1) {
  if (k_end < 0) k_end = K;
  __shared__ __align__(16) FcSmem sm;
  fc_tile_g<false, WKN>(sm, N, Nout, k_begin, k_end, load,
                        [&](int o, int k) { return WKN ? Wg[(size_t)k * Nout + o] : Wg[(size_t)o * K + k]; }, store);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
