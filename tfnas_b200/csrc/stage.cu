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
