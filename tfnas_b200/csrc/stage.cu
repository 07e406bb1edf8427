// MixedStage sink-connecting sum (reference models/model_search.py:202-204) and its backward.
#include "kernels.h"

struct SinkPtrs {
  const float* res[4];
  float* dres[4];
};

__device__ __forceinline__ void softmax_k(const float* betas, int K, float (&b)[4]) {
  float m = -INFINITY;
  for (int j = 0; j < K; ++j) m = fmaxf(m, betas[j]);
  float s = 0.f;
  for (int j = 0; j < K; ++j) { b[j] = expf(betas[j] - m); s += b[j]; }
  for (int j = 0; j < K; ++j) b[j] /= s;
  for (int j = K; j < 4; ++j) b[j] = 0.f;
}

__global__ void __launch_bounds__(NT) k_sink_fwd(int K, size_t numel, SinkPtrs p, const float* __restrict__ betas,
                                                  const float* __restrict__ cumlat, float* __restrict__ out,
                                                  float* __restrict__ out_lat) {
  float b[4];
  softmax_k(betas, K, b);
  if (blockIdx.x == 0 && threadIdx.x == 0 && cumlat && out_lat) {
    float l = 0.f;
    for (int j = 0; j < K; ++j) l += b[j] * cumlat[j];
    *out_lat = l;
  }
  const size_t stride = (size_t)gridDim.x * NT;
  if ((numel & 3) == 0) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < numel / 4; i += stride) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < K; ++j) {
        float4 r = ((const float4*)p.res[j])[i];
        o.x += b[j] * r.x; o.y += b[j] * r.y; o.z += b[j] * r.z; o.w += b[j] * r.w;
      }
      ((float4*)out)[i] = o;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < numel; i += stride) {
      float o = 0.f;
      for (int j = 0; j < K; ++j) o += b[j] * p.res[j][i];
      out[i] = o;
    }
  }
}

__global__ void __launch_bounds__(NT) k_sink_bwd(int K, size_t numel, SinkPtrs p, const float* __restrict__ betas,
                                                  const float* __restrict__ dout, double* __restrict__ dots) {
  float b[4];
  softmax_k(betas, K, b);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t stride = (size_t)gridDim.x * NT;
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < numel; i += stride) {
    float g = dout[i];
    for (int j = 0; j < K; ++j) {
      acc[j] += g * p.res[j][i];
      p.dres[j][i] = b[j] * g;
    }
  }
  __shared__ double red[NT / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = 0; j < 4; ++j) {
    double t = warp_sum_d((double)acc[j]);
    if (lane == 0) red[warp][j] = t;
  }
  __syncthreads();
  if ((int)threadIdx.x < K) {
    double t = 0;
    for (int w = 0; w < NT / 32; ++w) t += red[w][threadIdx.x];
    atomicAdd(&dots[threadIdx.x], t);
  }
}

__global__ void k_sink_fin(int K, const float* __restrict__ betas, const float* __restrict__ cumlat,
                           const float* __restrict__ dlat, const double* __restrict__ dots,
                           float* __restrict__ dbetas, float* __restrict__ dcumlat) {
  if (threadIdx.x != 0) return;
  float b[4];
  softmax_k(betas, K, b);
  const float dl = dlat ? *dlat : 0.f;
  float t[4], dot = 0.f;
  for (int j = 0; j < K; ++j) {
    t[j] = (float)dots[j] + (cumlat ? dl * cumlat[j] : 0.f);
    dot += t[j] * b[j];
  }
  for (int j = 0; j < K; ++j) {
    dbetas[j] = b[j] * (t[j] - dot);
    if (dcumlat) dcumlat[j] = b[j] * dl;
  }
}

void launch_sink_fwd(int K, size_t numel, const float* const* res, const float* betas, const float* cumlat,
                     float* out, float* out_lat, cudaStream_t st) {
  SinkPtrs p;
  for (int j = 0; j < 4; ++j) { p.res[j] = j < K ? res[j] : nullptr; p.dres[j] = nullptr; }
  int blocks = (int)min((size_t)(8 * sm_count()), (numel / 4 + NT - 1) / NT);
  ProfScope ps("sink_fwd", 4.0 * numel * (K + 1), 2.0 * numel * K, st);
  k_sink_fwd<<<max(blocks, 1), NT, 0, st>>>(K, numel, p, betas, cumlat, out, out_lat);
}

void launch_sink_bwd(int K, size_t numel, const float* const* res, const float* betas, const float* cumlat,
                     const float* dout, const float* dlat, float* const* dres, float* dbetas, float* dcumlat,
                     double* ws, cudaStream_t st) {
  SinkPtrs p;
  for (int j = 0; j < 4; ++j) { p.res[j] = j < K ? res[j] : nullptr; p.dres[j] = j < K ? dres[j] : nullptr; }
  cudaMemsetAsync(ws, 0, 4 * sizeof(double), st);
  int blocks = (int)min((size_t)(4 * sm_count()), (numel + NT - 1) / NT);
  { ProfScope ps("sink_bwd", 4.0 * numel * (2 * K + 1), 3.0 * numel * K, st);
    k_sink_bwd<<<max(blocks, 1), NT, 0, st>>>(K, numel, p, betas, dout, ws); }
  { ProfScope ps("sink_fin", 64, 0, st);
    k_sink_fin<<<1, 32, 0, st>>>(K, betas, cumlat, dlat, ws, dbetas, dcumlat); }
}

// -------------------------------------------------------------------------------------------------
// Batch-statistic BatchNorm (no affine, biased variance, eps 1e-5) fused with the activation, for the
// non-MixedOP layers of Network.forward (reference models/layers.py:90-110 BasicLayer with
// nn.BatchNorm2d(affine=False, track_running_stats=False) + ReLU / Swish; models/model_search.py:219-220,275).
// -------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float bn_act(float v) {
  if (ACT == TFNAS_ACT_RELU) return act_f<TFNAS_ACT_RELU>(v);
  if (ACT == TFNAS_ACT_SWISH) return act_f<TFNAS_ACT_SWISH>(v);
  return v;
}
template <int ACT>
__device__ __forceinline__ float bn_dact(float v) {
  if (ACT == TFNAS_ACT_RELU) return act_df<TFNAS_ACT_RELU>(v);
  if (ACT == TFNAS_ACT_SWISH) return act_df<TFNAS_ACT_SWISH>(v);
  return 1.f;
}

// grid (C, splits): per-channel sum and sum of squares over (n, hw)
__global__ void __launch_bounds__(NT) k_bn_stats(int N, int C, int HW, const float* __restrict__ x, double* __restrict__ st) {
  const int c = blockIdx.x;
  const long long total = (long long)N * HW;
  const long long i0 = total * blockIdx.y / gridDim.y, i1 = total * (blockIdx.y + 1) / gridDim.y;
  float s1 = 0.f, s2 = 0.f;
  const float inv_hw = 1.f / (float)HW;
  if ((HW & 3) == 0) {
    for (long long i = (i0 >> 2) + threadIdx.x; i < (i1 >> 2); i += NT) {
      const int e = (int)(i << 2);
      const int n = fast_div(e, HW, inv_hw), hw = e - n * HW;
      const float4 v = *(const float4*)(x + ((size_t)n * C + c) * HW + hw);
      s1 += v.x + v.y + v.z + v.w;
      s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += NT) {
      const int n = fast_div((int)i, HW, inv_hw), hw = (int)i - n * HW;
      const float v = x[((size_t)n * C + c) * HW + hw];
      s1 += v;
      s2 += v * v;
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

__global__ void k_bn_fin(int C, double invM, const double* __restrict__ st, float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = st[2 * c] * invM;
  double v = st[2 * c + 1] * invM - m * m;
  out[c] = (float)m;
  out[C + c] = (float)(1.0 / sqrt(fmax(v, 0.0) + (double)BN_EPS));
}

template <int ACT>
__global__ void __launch_bounds__(NT) k_bn_apply(int C, int HW, size_t total, const float* __restrict__ x,
                                                  const float* __restrict__ mr, float* __restrict__ y) {
  const float inv_hw = 1.f / (float)HW;
  const size_t stride = (size_t)gridDim.x * NT;
  if ((HW & 3) == 0) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < (total >> 2); i += stride) {
      const size_t e = i << 2;
      const int plane = (int)(e / HW);
      const int c = plane % C;
      const float mu = mr[c], r = mr[C + c];
      const float4 v = *(const float4*)(x + e);
      *(float4*)(y + e) = make_float4(bn_act<ACT>((v.x - mu) * r), bn_act<ACT>((v.y - mu) * r),
                                      bn_act<ACT>((v.z - mu) * r), bn_act<ACT>((v.w - mu) * r));
    }
  } else {
    for (size_t e = (size_t)blockIdx.x * NT + threadIdx.x; e < total; e += stride) {
      const int c = (int)((e / HW) % C);
      y[e] = bn_act<ACT>((x[e] - mr[c]) * mr[C + c]);
    }
  }
  (void)inv_hw;
}

// backward pass 1: g = dy * act'(xhat); sums of g and g*xhat per channel.  grid (C, splits)
template <int ACT>
__global__ void __launch_bounds__(NT) k_bn_bwd_stats(int N, int C, int HW, const float* __restrict__ x,
                                                      const float* __restrict__ mr, const float* __restrict__ dy,
                                                      double* __restrict__ st) {
  const int c = blockIdx.x;
  const float mu = mr[c], r = mr[C + c];
  const long long total = (long long)N * HW;
  const long long i0 = total * blockIdx.y / gridDim.y, i1 = total * (blockIdx.y + 1) / gridDim.y;
  const float inv_hw = 1.f / (float)HW;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += NT) {
    const int n = fast_div((int)i, HW, inv_hw), hw = (int)i - n * HW;
    const size_t a = ((size_t)n * C + c) * HW + hw;
    const float xh = (x[a] - mu) * r;
    const float g = dy[a] * bn_dact<ACT>(xh);
    s1 += g;
    s2 += g * xh;
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

// backward pass 2: dx = r (g - mean(g) - xhat mean(g xhat))
template <int ACT>
__global__ void __launch_bounds__(NT) k_bn_bwd_apply(int C, int HW, size_t total, double invM, const float* __restrict__ x,
                                                      const float* __restrict__ mr, const float* __restrict__ dy,
                                                      const double* __restrict__ st, float* __restrict__ dx) {
  const size_t stride = (size_t)gridDim.x * NT;
  for (size_t e = (size_t)blockIdx.x * NT + threadIdx.x; e < total; e += stride) {
    const int c = (int)((e / HW) % C);
    const float mu = mr[c], r = mr[C + c];
    const float m1 = (float)(st[2 * c] * invM), m2 = (float)(st[2 * c + 1] * invM);
    const float xh = (x[e] - mu) * r;
    const float g = dy[e] * bn_dact<ACT>(xh);
    dx[e] = r * (g - m1 - xh * m2);
  }
}

void launch_bn_act_fwd(int N, int C, int HW, int act, const float* x, float* y, float* mr, double* ws, cudaStream_t st) {
  const size_t total = (size_t)N * C * HW;
  cudaMemsetAsync(ws, 0, (size_t)2 * C * sizeof(double), st);
  int splits = max(1, min((int)(((long long)N * HW + 4095) / 4096), cdiv(4 * sm_count(), C)));
  { ProfScope ps("bn_stats", 4.0 * total, 3.0 * total, st);
    k_bn_stats<<<dim3(C, splits), NT, 0, st>>>(N, C, HW, x, ws); }
  { ProfScope ps("bn_fin", 24.0 * C, 0, st);
    k_bn_fin<<<cdiv(C, 256), 256, 0, st>>>(C, 1.0 / ((double)N * HW), ws, mr); }
  int blocks = (int)min((size_t)(8 * sm_count()), (total / 4 + NT - 1) / NT);
  ProfScope ps("bn_apply", 8.0 * total, 4.0 * total, st);
  if (act == TFNAS_ACT_RELU) k_bn_apply<TFNAS_ACT_RELU><<<max(blocks, 1), NT, 0, st>>>(C, HW, total, x, mr, y);
  else if (act == TFNAS_ACT_SWISH) k_bn_apply<TFNAS_ACT_SWISH><<<max(blocks, 1), NT, 0, st>>>(C, HW, total, x, mr, y);
  else k_bn_apply<2><<<max(blocks, 1), NT, 0, st>>>(C, HW, total, x, mr, y);
}

void launch_bn_act_bwd(int N, int C, int HW, int act, const float* x, const float* mr, const float* dy, float* dx,
                       double* ws, cudaStream_t st) {
  const size_t total = (size_t)N * C * HW;
  cudaMemsetAsync(ws, 0, (size_t)2 * C * sizeof(double), st);
  int splits = max(1, min((int)(((long long)N * HW + 4095) / 4096), cdiv(4 * sm_count(), C)));
  int blocks = (int)min((size_t)(8 * sm_count()), (total + NT - 1) / NT);
  const double invM = 1.0 / ((double)N * HW);
  { ProfScope ps("bn_bwd_stats", 8.0 * total, 6.0 * total, st);
    if (act == TFNAS_ACT_RELU) k_bn_bwd_stats<TFNAS_ACT_RELU><<<dim3(C, splits), NT, 0, st>>>(N, C, HW, x, mr, dy, ws);
    else if (act == TFNAS_ACT_SWISH) k_bn_bwd_stats<TFNAS_ACT_SWISH><<<dim3(C, splits), NT, 0, st>>>(N, C, HW, x, mr, dy, ws);
    else k_bn_bwd_stats<2><<<dim3(C, splits), NT, 0, st>>>(N, C, HW, x, mr, dy, ws); }
  ProfScope ps("bn_bwd_apply", 12.0 * total, 8.0 * total, st);
  if (act == TFNAS_ACT_RELU) k_bn_bwd_apply<TFNAS_ACT_RELU><<<max(blocks, 1), NT, 0, st>>>(C, HW, total, invM, x, mr, dy, ws, dx);
  else if (act == TFNAS_ACT_SWISH) k_bn_bwd_apply<TFNAS_ACT_SWISH><<<max(blocks, 1), NT, 0, st>>>(C, HW, total, invM, x, mr, dy, ws, dx);
  else k_bn_bwd_apply<2><<<max(blocks, 1), NT, 0, st>>>(C, HW, total, invM, x, mr, dy, ws, dx);
}

// ------------------------------------------------------------------------------------------------
// Plain depthwise KxK convolution, stride 1, padding K/2 (the second stem's depth_conv, models/layers.py:486-489
// with the widths of models/model_search.py:220): forward, input gradient and weight gradient.
// Direct kernels: a thread owns 4 consecutive outputs of one row, the KxK taps live in registers, inputs come
// through the read-only path (each input element is re-read from L1, never staged).
// ------------------------------------------------------------------------------------------------
// FLIP = false: y[r][c] = sum w[ky][kx] x[r+ky-p][c+kx-p]     (forward)
// FLIP = true : correlation with the flipped filter             (input gradient of the same convolution)
template <int KS, bool FLIP>
__global__ void __launch_bounds__(NT) k_dwc_s1(int C, int H, int W, const float* __restrict__ x, const float* __restrict__ w,
                                               float* __restrict__ y) {
  constexpr int pad = KS / 2;
  const int plane = blockIdx.y;                 // n * C + c
  const int c = plane % C;
  const int gpr = (W + 3) >> 2;                 // groups of 4 outputs per row
  const int g = blockIdx.x * NT + threadIdx.x;
  if (g >= H * gpr) return;
  const int r = g / gpr, c0 = (g - r * gpr) * 4;
  float wr[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) wr[i] = w[(size_t)c * KS * KS + (FLIP ? KS * KS - 1 - i : i)];
  const float* xp = x + (size_t)plane * H * W;
  float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < KS; ++ky) {
    const int rr = r + ky - pad;
    if (rr < 0 || rr >= H) continue;
    float v[4 + KS - 1];
#pragma unroll
    for (int j = 0; j < 4 + KS - 1; ++j) {
      const int cc = c0 + j - pad;
      v[j] = (cc >= 0 && cc < W) ? __ldg(xp + (size_t)rr * W + cc) : 0.f;
    }
#pragma unroll
    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += wr[ky * KS + kx] * v[j + kx];
  }
  float* q = y + (size_t)plane * H * W + (size_t)r * W + c0;
  if ((W & 3) == 0) {
    *(float4*)q = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c0 + j < W) q[j] = o[j];
  }
}

// dw[c][ky][kx] = sum_{n,r,col} dy[n,c,r,col] * x[n,c,r+ky-p,col+kx-p].  grid (C, splits over images)
template <int KS>
__global__ void __launch_bounds__(NT) k_dwc_wgrad(int N, int C, int H, int W, const float* __restrict__ x,
                                                  const float* __restrict__ dy, float* __restrict__ dw) {
  constexpr int pad = KS / 2;
  const int c = blockIdx.x;
  const int n0 = (int)((long long)N * blockIdx.y / gridDim.y), n1 = (int)((long long)N * (blockIdx.y + 1) / gridDim.y);
  float acc[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) acc[i] = 0.f;
  const int gpr = (W + 3) >> 2, groups = H * gpr;
  for (int n = n0; n < n1; ++n) {
    const float* xp = x + ((size_t)n * C + c) * H * W;
    const float* gp = dy + ((size_t)n * C + c) * H * W;
    for (int g = threadIdx.x; g < groups; g += NT) {
      const int r = g / gpr, c0 = (g - r * gpr) * 4;
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = c0 + j < W ? __ldg(gp + (size_t)r * W + c0 + j) : 0.f;
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
        const int rr = r + ky - pad;
        if (rr < 0 || rr >= H) continue;
        float v[4 + KS - 1];
#pragma unroll
        for (int j = 0; j < 4 + KS - 1; ++j) {
          const int cc = c0 + j - pad;
          v[j] = (cc >= 0 && cc < W) ? __ldg(xp + (size_t)rr * W + cc) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[ky * KS + kx] += d[j] * v[j + kx];
      }
    }
  }
  __shared__ float red[NT / 32][KS * KS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) {
    const float t = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < KS * KS) {
    float t = 0.f;
    for (int wq = 0; wq < NT / 32; ++wq) t += red[wq][threadIdx.x];
    atomicAdd(&dw[(size_t)c * KS * KS + threadIdx.x], t);
  }
}

void launch_dwconv_fwd(int N, int C, int H, int W, int KS, const float* x, const float* w, float* y, cudaStream_t st) {
  const double total = (double)N * C * H * W;
  dim3 grid(cdiv(H * ((W + 3) >> 2), NT), N * C);
  ProfScope ps("stem_dw_fwd", 8.0 * total, 2.0 * KS * KS * total, st);
  if (KS == 3) k_dwc_s1<3, false><<<grid, NT, 0, st>>>(C, H, W, x, w, y);
  else k_dwc_s1<5, false><<<grid, NT, 0, st>>>(C, H, W, x, w, y);
}

void launch_dwconv_bwd(int N, int C, int H, int W, int KS, const float* x, const float* w, const float* dy, float* dx,
                       float* dw, cudaStream_t st) {
  const double total = (double)N * C * H * W;
  if (dx) {
    dim3 grid(cdiv(H * ((W + 3) >> 2), NT), N * C);
    ProfScope ps("stem_dw_bwd", 8.0 * total, 2.0 * KS * KS * total, st);
    if (KS == 3) k_dwc_s1<3, true><<<grid, NT, 0, st>>>(C, H, W, dy, w, dx);
    else k_dwc_s1<5, true><<<grid, NT, 0, st>>>(C, H, W, dy, w, dx);
  }
  if (dw) {
    cudaMemsetAsync(dw, 0, (size_t)C * KS * KS * sizeof(float), st);
    const int splits = max(1, min(N, cdiv(4 * sm_count(), C)));
    ProfScope ps("stem_dw_wgrad", 8.0 * total, 2.0 * KS * KS * total, st);
    if (KS == 3) k_dwc_wgrad<3><<<dim3(C, splits), NT, 0, st>>>(N, C, H, W, x, dy, dw);
    else k_dwc_wgrad<5><<<dim3(C, splits), NT, 0, st>>>(N, C, H, W, x, dy, dw);
  }
}
