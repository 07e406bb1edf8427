// The two stems of Network.forward (reference models/model_search.py:219-220, :283-284):
//   first_stem  = ConvLayer(3, 32, kernel 3, stride 2) -> BN(batch statistics, no affine) -> ReLU     (models/layers.py:190-256)
//   second_stem = MBInvertedResBlock(32, 32, se 8, 16, kernel 3, stride 1, relu) WITHOUT an expand conv because
//                 mid_channels == in_channels (models/layers.py:479-482), i.e. depthwise -> BN -> ReLU -> SE -> 1x1 -> BN.
//
// The second stem is an MBConv whose "expand" output is the BN-normalised first-stem convolution, so it runs on the SAME
// phase kernels as a sampled MixedOP candidate (launch_forward_tail: F1b..F4; launch_backward(stop_at_da): B1..B3a) with
// UH := BN(conv(image)).  What is specific to the stems lives here:
//   k_stem_conv<PASS>   direct 3x3 / stride 2 convolution, 3 -> 32 channels; PASS 0 accumulates the BN sums only, PASS 1
//                       recomputes the convolution and writes the normalised output (the image is 3/32 of the output's
//                       size: recomputing is cheaper than writing, re-reading and re-writing the raw convolution)
//   k_stem_dustats      sums of du-hat = DA * relu'(UH) and du-hat * UH (BN backward of the first stem)
//   k_stem_conv_wgrad   dW[o][c][ky][kx] = sum_p dy[o][p] * image patch, dy = r (du-hat - m1 - UH m2) formed on load
#include <string.h>
#include "api_internal.h"

#define STEM_CI 3
#define STEM_CM 32
#define STEM_TAPS (STEM_CI * 9)

struct StemLayout {
  Plan P;
  SavedLayout L;
  size_t saved, ws, ws_bytes, total;
};

static int stem_layout(const TfnasStemDesc* d, const TfnasStemPtrs* w, int want_wgrad, StemLayout& B) {
  if (!d) return fail(TFNAS_E_INVALID, "stem: null descriptor");
  if (d->c_in != STEM_CI || d->c_mid != STEM_CM) return fail(TFNAS_E_UNSUPPORTED, "stem: only 3 -> 32 channels (got %d -> %d)", d->c_in, d->c_mid);
  if (d->N < 1 || d->H < 2 || d->W < 2 || d->se < 1 || d->c_out < 1 || d->c_out > 256) return fail(TFNAS_E_INVALID, "stem: bad shape");
  Plan& P = B.P;
  memset(&P, 0, sizeof(P));
  const int Ho = (d->H + 2 - 3) / 2 + 1, Wo = (d->W + 2 - 3) / 2 + 1;
  P.N = d->N; P.ic = STEM_CM; P.oc = d->c_out; P.H = Ho; P.W = Wo; P.Ho = Ho; P.Wo = Wo; P.stride = 1; P.act = TFNAS_ACT_RELU;
  P.na = 1; P.num_ops = 1; P.MC = STEM_CM; P.MCse = STEM_CM; P.SEH = d->se; P.HW = Ho * Wo; P.HWo = P.HW;
  const long long Pn = (long long)d->N * P.HW;
  if (Pn >= (1LL << 23)) return fail(TFNAS_E_UNSUPPORTED, "stem: N*H*W >= 2^23 pixels per call not supported");
  P.P = (int)Pn; P.Q = (int)Pn; P.residual = 0;
  Cand& c = P.c[0];
  c.id = 0; c.mc = STEM_CM; c.k = 3; c.se = d->se; c.coff = 0; c.soff = 0; c.hoff = 0;
  if (w) {
    if (!w->conv_w || !w->dw || !w->pw || !w->se_rw || !w->se_rb || !w->se_ew || !w->se_eb) return fail(TFNAS_E_INVALID, "stem: null weight pointer");
    c.w1 = nullptr; c.dw = w->dw; c.w3 = w->pw; c.rw = w->se_rw; c.rb = w->se_rb; c.ew = w->se_ew; c.eb = w->se_eb;
  }
  saved_layout(P, B.L);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  B.saved = take(B.L.total);
  FwdScratch F;
  BwdScratch Bs;
  B.ws_bytes = max(fwd_scratch(P, nullptr, F), bwd_scratch(P, want_wgrad, nullptr, Bs));
  B.ws = take(B.ws_bytes);
  B.total = o;
  return TFNAS_OK;
}

// ---- first-stem convolution ------------------------------------------------------------------------------------------
// thread = one output pixel (all 32 channels in registers); weights in shared memory as [tap][32] (float4 broadcasts).
// Persistent CTAs: PASS 0 keeps per-thread partial sums across its pixels and reduces once per CTA.
template <int PASS>
__global__ void __launch_bounds__(NT) k_stem_conv(int N, int H, int W, int Ho, int Wo, const float* __restrict__ img,
                                                   const float* __restrict__ wconv, const float* __restrict__ bn1,
                                                   double* __restrict__ sums, float* __restrict__ UH) {
  __shared__ __align__(16) float ws[STEM_TAPS][STEM_CM];
  __shared__ float mr[2 * STEM_CM];
  __shared__ double red[NT / 32][2 * STEM_CM];
  const int tid = threadIdx.x;
  for (int i = tid; i < STEM_TAPS * STEM_CM; i += NT) {
    const int o = i / STEM_TAPS, t = i - o * STEM_TAPS;          // wconv is [32][3][3][3] = [o][tap]
    ws[t][o] = wconv[i];
  }
  if (PASS == 1 && tid < 2 * STEM_CM) mr[tid] = bn1[tid];
  __syncthreads();
  const int HWo = Ho * Wo;
  const long long total = (long long)N * HWo;
  const float inv_hwo = 1.f / (float)HWo, inv_wo = 1.f / (float)Wo;
  float s1[STEM_CM], s2[STEM_CM];
  if (PASS == 0) {
#pragma unroll
    for (int c = 0; c < STEM_CM; ++c) s1[c] = s2[c] = 0.f;
  }
  for (long long q = (long long)blockIdx.x * NT + tid; q < total; q += (long long)gridDim.x * NT) {
    const int n = fast_div((int)q, HWo, inv_hwo), hw = (int)q - n * HWo;
    const int oy = fast_div(hw, Wo, inv_wo), ox = hw - oy * Wo;
    float acc[STEM_CM];
#pragma unroll
    for (int c = 0; c < STEM_CM; ++c) acc[c] = 0.f;
    const float* ip = img + (size_t)n * STEM_CI * H * W;
#pragma unroll
    for (int ci = 0; ci < STEM_CI; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = 2 * ox - 1 + kx;
          const float v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(ip + ((size_t)ci * H + iy) * W + ix) : 0.f;
          const float4* wr = (const float4*)ws[ci * 9 + ky * 3 + kx];
#pragma unroll
          for (int c4 = 0; c4 < STEM_CM / 4; ++c4) {
            const float4 w4 = wr[c4];
            acc[c4 * 4] += w4.x * v; acc[c4 * 4 + 1] += w4.y * v; acc[c4 * 4 + 2] += w4.z * v; acc[c4 * 4 + 3] += w4.w * v;
          }
        }
      }
    }
    if (PASS == 0) {
#pragma unroll
      for (int c = 0; c < STEM_CM; ++c) { s1[c] += acc[c]; s2[c] += acc[c] * acc[c]; }
    } else {
      float* op = UH + (size_t)n * STEM_CM * HWo + hw;
#pragma unroll
      for (int c = 0; c < STEM_CM; ++c) op[(size_t)c * HWo] = (acc[c] - mr[c]) * mr[STEM_CM + c];
    }
  }
  if (PASS == 0) {
    const int lane = tid & 31, warp = tid >> 5;
    // 16-value transpose-reduce: afterwards lane l holds the warp total of value l & 15
    float a0[16], a1[16], b0[16], b1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a0[i] = s1[i]; a1[i] = s1[16 + i]; b0[i] = s2[i]; b1[i] = s2[16 + i]; }
    const float ta0 = warp_sum16(a0), ta1 = warp_sum16(a1), tb0 = warp_sum16(b0), tb1 = warp_sum16(b1);
    if (lane < 16) {
      red[warp][2 * lane] = (double)ta0; red[warp][2 * lane + 1] = (double)tb0;
      red[warp][2 * (16 + lane)] = (double)ta1; red[warp][2 * (16 + lane) + 1] = (double)tb1;
    }
    __syncthreads();
    if (tid < 2 * STEM_CM) {
      double t = 0;
      for (int w = 0; w < NT / 32; ++w) t += red[w][tid];
      atomicAdd(&sums[tid], t);        // interleaved {sum, sum of squares} per channel, the layout k_stem_bnfin reads
    }
  }
}

__global__ void k_stem_bnfin(int C, double invM, const double* __restrict__ st, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = st[2 * c] * invM;
  const double v = st[2 * c + 1] * invM - m * m;
  out[c] = (float)m;
  out[C + c] = (float)(1.0 / sqrt(fmax(v, 0.0) + (double)BN_EPS));
}

// sums over (n, hw) of g = DA * relu'(UH) and g * UH per channel; grid (32, splits)
__global__ void __launch_bounds__(NT) k_stem_dustats(int N, int HW, const float* __restrict__ UH, const float* __restrict__ DA,
                                                      double* __restrict__ st) {
  const int c = blockIdx.x;
  const long long total = (long long)N * HW;
  const long long i0 = (total * blockIdx.y / gridDim.y) & ~3LL, i1 = (int)blockIdx.y + 1 == (int)gridDim.y ? total : (total * (blockIdx.y + 1) / gridDim.y) & ~3LL;
  const float inv_hw = 1.f / (float)HW;
  float s1 = 0.f, s2 = 0.f;
  if ((HW & 3) == 0) {
    for (long long i = (i0 >> 2) + threadIdx.x; i < (i1 >> 2); i += NT) {
      const int e = (int)(i << 2);
      const int n = fast_div(e, HW, inv_hw), hw = e - n * HW;
      const size_t a = ((size_t)n * STEM_CM + c) * HW + hw;
      const float4 u = *(const float4*)(UH + a), g = *(const float4*)(DA + a);
      const float g0 = u.x > 0.f ? g.x : 0.f, g1 = u.y > 0.f ? g.y : 0.f, g2 = u.z > 0.f ? g.z : 0.f, g3 = u.w > 0.f ? g.w : 0.f;
      s1 += g0 + g1 + g2 + g3;
      s2 += g0 * u.x + g1 * u.y + g2 * u.z + g3 * u.w;
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += NT) {
      const int n = fast_div((int)i, HW, inv_hw), hw = (int)i - n * HW;
      const size_t a = ((size_t)n * STEM_CM + c) * HW + hw;
      const float u = UH[a], g = u > 0.f ? DA[a] : 0.f;
      s1 += g;
      s2 += g * u;
    }
  }
  __shared__ double red[2][NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double a = warp_sum_d((double)s1), b = warp_sum_d((double)s2);
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0;
    for (int w = 0; w < NT / 32; ++w) t += red[threadIdx.x][w];
    atomicAdd(&st[2 * c + threadIdx.x], t);
  }
}

// Weight gradient of the first-stem convolution as a skinny GEMM dW[32][27] = sum_p dy[32][p] * patch[27][p].  Persistent
// CTAs walk tiles of WG_TP output pixels: dy = r (g - m1 - UH m2), g = DA * relu'(UH), is formed on load (float4 along the
// pixels) and staged [pixel][channel]; the 27 image taps of every pixel are gathered into [pixel][28].  thread = (output
// channel, group of 4 taps): per pixel one conflict-free scalar and one broadcast vector shared-memory load for 4 FMAs; the
// partial sums stay in registers across tiles, 4 float atomics per thread at the end.
#define WG_TP 128
__global__ void __launch_bounds__(NT) k_stem_conv_wgrad(int N, int H, int W, int Ho, int Wo, const float* __restrict__ img,
                                                         const float* __restrict__ UH, const float* __restrict__ DA,
                                                         const float* __restrict__ bn1, const double* __restrict__ sU,
                                                         float* __restrict__ dW) {
  __shared__ float dys[WG_TP][STEM_CM + 1];
  __shared__ __align__(16) float pts[WG_TP][28];
  __shared__ float cr[STEM_CM], cm1[STEM_CM], cm2[STEM_CM];
  const int tid = threadIdx.x, o = tid & 31, tg = tid >> 5;        // taps tg*4 .. tg*4+3 (tg == 7: nothing to do, 27 taps)
  const int HWo = Ho * Wo;
  const long long total = (long long)N * HWo;
  if (tid < STEM_CM) {
    const double invM = 1.0 / (double)total;
    cr[tid] = bn1[STEM_CM + tid];
    cm1[tid] = (float)(sU[2 * tid] * invM);
    cm2[tid] = (float)(sU[2 * tid + 1] * invM);
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float inv_hwo = 1.f / (float)HWo, inv_wo = 1.f / (float)Wo;
  const int ntiles = (int)((total + WG_TP - 1) / WG_TP);
  const bool vec = (HWo & 3) == 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int q0 = tile * WG_TP;
    __syncthreads();
    // dy tile: thread -> (channel, 4 consecutive pixels)
    for (int i = tid; i < STEM_CM * (WG_TP / 4); i += NT) {
      const int c = i / (WG_TP / 4), pp = (i - c * (WG_TP / 4)) * 4, q = q0 + pp;
      float u[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec && q + 3 < total) {
        const int n = fast_div(q, HWo, inv_hwo), hw = q - n * HWo;
        const size_t a = ((size_t)n * STEM_CM + c) * HWo + hw;
        const float4 uv = *(const float4*)(UH + a), gv = *(const float4*)(DA + a);
        u[0] = uv.x; u[1] = uv.y; u[2] = uv.z; u[3] = uv.w;
        g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (q + e < total) {
            const int n = fast_div(q + e, HWo, inv_hwo), hw = q + e - n * HWo;
            const size_t a = ((size_t)n * STEM_CM + c) * HWo + hw;
            u[e] = UH[a]; g[e] = DA[a];
          }
      }
      const float r = cr[c], m1 = cm1[c], m2 = cm2[c];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        dys[pp + e][c] = (q + e < total) ? r * ((u[e] > 0.f ? g[e] : 0.f) - m1 - u[e] * m2) : 0.f;
    }
    // patch tile: thread -> (tap, pixel), consecutive threads along the pixels
    for (int i = tid; i < 28 * WG_TP; i += NT) {
      const int t = i / WG_TP, pp = i - t * WG_TP, q = q0 + pp;
      float v = 0.f;
      if (t < STEM_TAPS && q < total) {
        const int n = fast_div(q, HWo, inv_hwo), hw = q - n * HWo;
        const int oy = fast_div(hw, Wo, inv_wo), ox = hw - oy * Wo;
        const int ci = t / 9, k9 = t - ci * 9, ky = k9 / 3, kx = k9 - ky * 3;
        const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((size_t)n * STEM_CI + ci) * H + iy) * W + ix);
      }
      pts[pp][t] = v;
    }
    __syncthreads();
    if (tg < 7) {
#pragma unroll 8
      for (int pp = 0; pp < WG_TP; ++pp) {
        const float d = dys[pp][o];
        const float4 pv = *(const float4*)&pts[pp][tg * 4];
        acc[0] += d * pv.x; acc[1] += d * pv.y; acc[2] += d * pv.z; acc[3] += d * pv.w;
      }
    }
  }
  if (tg < 7) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (tg * 4 + e < STEM_TAPS) atomicAdd(&dW[o * STEM_TAPS + tg * 4 + e], acc[e]);
  }
}

extern "C" {

size_t tfnas_stem_arena_bytes(const TfnasStemDesc* d, int want_wgrad) {
  static thread_local StemLayout B;
  if (stem_layout(d, nullptr, want_wgrad, B) != TFNAS_OK) return 0;
  return B.total;
}

int tfnas_stem_fwd(const TfnasStemDesc* d, const float* img, const TfnasStemPtrs* w, float* out, void* arena,
                   size_t arena_bytes, void* stream) {
  static thread_local StemLayout B;
  if (!w) return fail(TFNAS_E_INVALID, "stem: null weights");
  int rc = stem_layout(d, w, 0, B);
  if (rc != TFNAS_OK) return rc;
  if (!img || !out || !arena) return fail(TFNAS_E_INVALID, "stem: null tensor pointer");
  if (!aligned16(img) || !aligned16(out) || (((uintptr_t)arena) & 255)) return fail(TFNAS_E_INVALID, "stem: img / out must be 16-byte, the arena 256-byte aligned");
  if (arena_bytes < B.total) return fail(TFNAS_E_WORKSPACE, "stem arena %zu < %zu", arena_bytes, B.total);
  const Plan& P = B.P;
  char* A = (char*)arena;
  char* saved = A + B.saved;
  cudaStream_t st = (cudaStream_t)stream;
  FwdScratch S;
  fwd_scratch(P, A + B.ws, S);
  float* bn1 = (float*)(saved + B.L.bn1);
  float* UH = (float*)(saved + B.L.UH);
  cudaGetLastError();
  // the forward accumulators [xsum | xcov | st2 | st3] are contiguous: one memset; xsum/xcov hold the conv's BN sums here
  cudaMemsetAsync(S.xsum, 0, (size_t)(P.ic + P.ic * P.ic + 2 * P.MC + 2 * P.na * P.oc) * sizeof(double), st);
  const int grid = 2 * sm_count();
  const double cbytes = 4.0 * d->N * (3.0 * d->H * d->W), ubytes = 4.0 * (double)P.P * STEM_CM;
  { ProfScope ps("stem_conv_stats", cbytes, 2.0 * P.P * STEM_CM * STEM_TAPS, st);
    k_stem_conv<0><<<grid, NT, 0, st>>>(d->N, d->H, d->W, P.H, P.W, img, w->conv_w, nullptr, S.xsum, nullptr); }
  { ProfScope ps("stem_bnfin", 512, 0, st);
    k_stem_bnfin<<<1, 64, 0, st>>>(STEM_CM, 1.0 / (double)P.P, S.xsum, bn1); }
  { ProfScope ps("stem_conv", cbytes + ubytes, 2.0 * P.P * STEM_CM * STEM_TAPS, st);
    k_stem_conv<1><<<grid, NT, 0, st>>>(d->N, d->H, d->W, P.H, P.W, img, w->conv_w, bn1, nullptr, UH); }
  UmWAll WP;
  if (umma_enabled()) umma_prep_project(P, S.umprep, WP, st);
  launch_forward_tail(P, umma_enabled() ? &WP : nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 0, out, nullptr, saved, B.L, S, st);
  return check_cuda("tfnas_stem_fwd");
}

int tfnas_stem_bwd(const TfnasStemDesc* d, const float* img, const TfnasStemPtrs* w, const float* dout,
                   const TfnasStemPtrs* dw, void* arena, size_t arena_bytes, void* stream) {
  static thread_local StemLayout B;
  if (!w || !dw) return fail(TFNAS_E_INVALID, "stem: null weights / weight gradients");
  int rc = stem_layout(d, w, 1, B);
  if (rc != TFNAS_OK) return rc;
  if (!img || !dout || !arena) return fail(TFNAS_E_INVALID, "stem: null tensor pointer");
  if (!dw->conv_w || !dw->dw || !dw->pw || !dw->se_rw || !dw->se_rb || !dw->se_ew || !dw->se_eb) return fail(TFNAS_E_INVALID, "stem: null weight-grad pointer");
  if (!aligned16(img) || !aligned16(dout) || (((uintptr_t)arena) & 255)) return fail(TFNAS_E_INVALID, "stem: img / dout must be 16-byte, the arena 256-byte aligned");
  if (arena_bytes < B.total) return fail(TFNAS_E_WORKSPACE, "stem arena %zu < %zu", arena_bytes, B.total);
  const Plan& P = B.P;
  char* A = (char*)arena;
  const char* saved = A + B.saved;
  cudaStream_t st = (cudaStream_t)stream;
  BwdScratch S;
  bwd_scratch(P, 1, A + B.ws, S);
  TfnasCandPtrs g;
  memset(&g, 0, sizeof(g));
  g.dw = dw->dw; g.w3 = dw->pw; g.se_rw = dw->se_rw; g.se_rb = dw->se_rb; g.se_ew = dw->se_ew; g.se_eb = dw->se_eb;
  g.w1 = dw->pw;      // unused by the stop_at_da path (placeholder, never written)
  cudaGetLastError();
  // B1 .. B3a on the MixedOP phase kernels: S.DA = dL/d relu(UH); dDW, dW3 and the SE gradients are written on the way
  launch_backward(P, nullptr, dout, nullptr, 1.f, 0, saved, B.L, S, nullptr, nullptr, &g, st, 1);
  const float* bn1 = (const float*)(saved + B.L.bn1);
  const float* UH = (const float*)(saved + B.L.UH);
  // first-stem BN backward + convolution weight gradient (S.sU was zeroed with the other accumulators by launch_backward)
  const double ubytes = 4.0 * (double)P.P * STEM_CM;
  { ProfScope ps("stem_dustats", 2.0 * ubytes, 4.0 * P.P * STEM_CM, st);
    int splits = max(1, min((int)(((long long)P.P + 4095) / 4096), cdiv(4 * sm_count(), STEM_CM)));
    k_stem_dustats<<<dim3(STEM_CM, splits), NT, 0, st>>>(P.N, P.HW, UH, S.DA, S.sU); }
  cudaMemsetAsync(dw->conv_w, 0, (size_t)STEM_CM * STEM_TAPS * sizeof(float), st);
  { ProfScope ps("stem_conv_wgrad", 2.0 * ubytes + 4.0 * d->N * (3.0 * d->H * d->W), 2.0 * P.P * STEM_CM * STEM_TAPS, st);
    k_stem_conv_wgrad<<<4 * sm_count(), NT, 0, st>>>(d->N, d->H, d->W, P.H, P.W, img, UH, S.DA, bn1, S.sU, dw->conv_w); }
  return check_cuda("tfnas_stem_bwd");
}

}  // extern "C"
