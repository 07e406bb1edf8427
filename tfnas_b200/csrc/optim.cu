// Step glue of the search loop (reference train_search.py:381-385, :414-422, :121): global-norm clipping fused with the
// SGD-momentum / Adam updates over a table of live tensors passed in kernel-parameter space, the log_softmax
// renormalisation of the architecture parameters, and softmax cross-entropy with its gradient.
#include <math.h>
#include "api_internal.h"

#define OPT_CHUNK 224          // tensors per launch: 224 * 28 B = 6.1 KB of kernel parameters
#define OPT_BX 8               // CTAs per tensor (grid-stride inside the tensor)
struct SgdChunk {
  float* p[OPT_CHUNK];
  float* g[OPT_CHUNK];
  float* b[OPT_CHUNK];
  int n[OPT_CHUNK];
};

// acc[0] += sum over the chunk's tensors of (grad_scale * g)^2
__global__ void __launch_bounds__(NT) k_sgd_norm(const __grid_constant__ SgdChunk c, float gscale, double* __restrict__ acc) {
  const int t = blockIdx.y;
  const float* __restrict__ g = c.g[t];
  const int n = c.n[t];
  float s = 0.f;
  for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += OPT_BX * NT) {
    const float v = g[i] * gscale;
    s += v * v;
  }
  __shared__ double red[NT / 32];
  const double w = warp_sum_d((double)s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0;
    for (int i = 0; i < NT / 32; ++i) tsum += red[i];
    if (tsum != 0.0) atomicAdd(acc, tsum);
  }
}

__global__ void __launch_bounds__(NT) k_sgd_update(const __grid_constant__ SgdChunk c, float gscale, float lr, float mom,
                                                    float wd, float max_norm, const double* __restrict__ acc,
                                                    float* __restrict__ norm_out) {
  const int t = blockIdx.y;
  float* __restrict__ p = c.p[t];
  float* __restrict__ g = c.g[t];
  float* __restrict__ b = c.b[t];
  const int n = c.n[t];
  const float total = (float)sqrt(*acc);
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.f);
  if (norm_out && t == 0 && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
  const float gs = gscale * coef;
  for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += OPT_BX * NT) {
    const float gi = g[i] * gs;
    const float pi = p[i];
    const float d = gi + wd * pi;
    const float bi = mom * b[i] + d;
    g[i] = gi;
    b[i] = bi;
    p[i] = pi - lr * bi;
  }
}

#define ADAM_MAXT 64
struct AdamTab {
  float* p[ADAM_MAXT];
  float* g[ADAM_MAXT];
  float* m[ADAM_MAXT];
  float* v[ADAM_MAXT];
  short n[ADAM_MAXT];
  short renorm[ADAM_MAXT];
};

// one CTA: thread (tensor, element) pairs are walked flat; every tensor has <= 64 elements
__global__ void __launch_bounds__(NT) k_adam(const __grid_constant__ AdamTab T, int nt, float gscale, float max_norm, float lr,
                                              float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  __shared__ double red[NT / 32];
  __shared__ float s_coef;
  float s = 0.f;
  for (int t = threadIdx.x >> 6; t < nt; t += NT >> 6) {
    const int e = threadIdx.x & 63;
    if (e < T.n[t]) { const float v = T.g[t][e] * gscale; s += v * v; }
  }
  const double w = warp_sum_d((double)s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int i = 0; i < NT / 32; ++i) tot += red[i];
    const float total = (float)sqrt(tot);
    s_coef = max_norm > 0.f ? fminf(max_norm / (total + 1e-6f), 1.f) : 1.f;
  }
  __syncthreads();
  const float gs = gscale * s_coef;
  const float step_size = lr / bc1;
  for (int t = threadIdx.x >> 6; t < nt; t += NT >> 6) {
    const int e = threadIdx.x & 63;
    if (e < T.n[t]) {
      float gi = T.g[t][e] * gs;
      T.g[t][e] = gi;
      const float pi = T.p[t][e];
      gi += wd * pi;
      const float m = T.m[t][e] + (gi - T.m[t][e]) * (1.f - b1);
      const float v = b2 * T.v[t][e] + (1.f - b2) * gi * gi;
      T.m[t][e] = m;
      T.v[t][e] = v;
      const float denom = sqrtf(v) / bc2_sqrt + eps;
      T.p[t][e] = pi - step_size * (m / denom);
    }
  }
  __syncthreads();
  // p = log_softmax(p) per tensor (train_search.py:421-422)
  for (int t = threadIdx.x; t < nt; t += NT) {
    if (!T.renorm[t]) continue;
    float* p = T.p[t];
    const int n = T.n[t];
    float mx = -INFINITY;
    for (int e = 0; e < n; ++e) mx = fmaxf(mx, p[e]);
    float se = 0.f;
    for (int e = 0; e < n; ++e) se += expf(p[e] - mx);
    const float lse = mx + logf(se);
    for (int e = 0; e < n; ++e) p[e] -= lse;
  }
}

// one CTA of 1024 threads: warp w handles rows w, w+32, ...; deterministic final sum
// eps = label smoothing: target distribution (1-eps) onehot + eps/C (train_eval.py:72-84); eps = 0 is nn.CrossEntropyLoss
__global__ void __launch_bounds__(1024) k_softmax_ce(int N, int C, const float* __restrict__ logits,
                                                      const long long* __restrict__ targets, float eps,
                                                      float* __restrict__ loss, float* __restrict__ dlogits) {
  __shared__ float part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invN = 1.f / (float)N;
  float acc = 0.f;
  for (int r = warp; r < N; r += 32) {
    const float* l = logits + (size_t)r * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, l[c]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f, sl = 0.f;
    for (int c = lane; c < C; c += 32) { se += expf(l[c] - mx); sl += l[c]; }
    se = warp_sum(se);
    sl = warp_sum(sl);
    const float lse = mx + logf(se);
    const int tg = (int)targets[r];
    const float toff = eps / (float)C, ton = 1.f - eps + toff;
    if (dlogits) {
      float* d = dlogits + (size_t)r * C;
      for (int c = lane; c < C; c += 32) d[c] = (expf(l[c] - lse) - (c == tg ? ton : toff)) * invN;
    }
    if (lane == 0) acc += eps == 0.f ? lse - l[tg] : lse - (1.f - eps) * l[tg] - toff * sl;
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += part[w];
    *loss = t * invN;
  }
}

extern "C" {

int tfnas_sgd_step(int n, const TfnasSgdTensor* t, float lr, float momentum, float weight_decay, float max_norm,
                   float grad_scale, float* total_norm_out, void* workspace, size_t ws_bytes, void* stream) {
  if (n < 0 || (n > 0 && !t)) return fail(TFNAS_E_INVALID, "sgd: bad tensor table");
  if (!workspace || ws_bytes < 16) return fail(TFNAS_E_WORKSPACE, "sgd workspace %zu < 16", ws_bytes);
  if (n == 0) return TFNAS_OK;
  for (int i = 0; i < n; ++i) {
    if (!t[i].p || !t[i].g || !t[i].buf) return fail(TFNAS_E_INVALID, "sgd: tensor %d has a null pointer", i);
    if (t[i].numel < 0 || t[i].numel >= (1LL << 31)) return fail(TFNAS_E_UNSUPPORTED, "sgd: tensor %d numel %lld", i, (long long)t[i].numel);
  }
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (double*)workspace;
  cudaGetLastError();
  cudaMemsetAsync(acc, 0, sizeof(double), st);
  static thread_local SgdChunk c;
  for (int pass = 0; pass < 2; ++pass) {
    for (int i0 = 0; i0 < n; i0 += OPT_CHUNK) {
      const int m = min(OPT_CHUNK, n - i0);
      double bytes = 0;
      for (int i = 0; i < m; ++i) {
        c.p[i] = t[i0 + i].p; c.g[i] = t[i0 + i].g; c.b[i] = t[i0 + i].buf; c.n[i] = (int)t[i0 + i].numel;
        bytes += 4.0 * (double)t[i0 + i].numel;
      }
      if (pass == 0) {
        ProfScope ps("sgd_norm", bytes, 2.0 * bytes / 4, st);
        k_sgd_norm<<<dim3(OPT_BX, m), NT, 0, st>>>(c, grad_scale, acc);
      } else {
        ProfScope ps("sgd_update", 6.0 * bytes, 6.0 * bytes / 4, st);
        k_sgd_update<<<dim3(OPT_BX, m), NT, 0, st>>>(c, grad_scale, lr, momentum, weight_decay, max_norm, acc,
                                                      i0 == 0 ? total_norm_out : nullptr);
      }
    }
  }
  return check_cuda("tfnas_sgd_step");
}

int tfnas_adam_step(int n, const TfnasAdamTensor* t, int step, float lr, float beta1, float beta2, float eps,
                    float weight_decay, float max_norm, float grad_scale, void* stream) {
  if (n < 1 || n > ADAM_MAXT || !t) return fail(TFNAS_E_INVALID, "adam: 1..%d tensors", ADAM_MAXT);
  if (step < 1) return fail(TFNAS_E_INVALID, "adam: step must be >= 1");
  static thread_local AdamTab T;
  for (int i = 0; i < n; ++i) {
    if (!t[i].p || !t[i].g || !t[i].m || !t[i].v) return fail(TFNAS_E_INVALID, "adam: tensor %d has a null pointer", i);
    if (t[i].numel < 1 || t[i].numel > 64) return fail(TFNAS_E_UNSUPPORTED, "adam: tensor %d numel %d not in 1..64", i, t[i].numel);
    T.p[i] = t[i].p; T.g[i] = t[i].g; T.m[i] = t[i].m; T.v[i] = t[i].v;
    T.n[i] = (short)t[i].numel; T.renorm[i] = (short)(t[i].renorm != 0);
  }
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  cudaStream_t st = (cudaStream_t)stream;
  cudaGetLastError();
  { ProfScope ps("adam", 1024, 0, st);
    k_adam<<<1, NT, 0, st>>>(T, n, grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2s); }
  return check_cuda("tfnas_adam_step");
}

int tfnas_softmax_ce_smooth(int N, int C, const float* logits, const int64_t* targets, float epsilon, float* loss,
                            float* dlogits, void* stream) {
  if (N < 1 || C < 1 || !logits || !targets || !loss) return fail(TFNAS_E_INVALID, "softmax_ce: bad arguments");
  if (!(epsilon >= 0.f && epsilon < 1.f)) return fail(TFNAS_E_INVALID, "softmax_ce: label smoothing %g not in [0, 1)", (double)epsilon);
  cudaStream_t st = (cudaStream_t)stream;
  cudaGetLastError();
  { ProfScope ps("softmax_ce", 8.0 * N * C, 4.0 * N * C, st);
    k_softmax_ce<<<1, 1024, 0, st>>>(N, C, logits, (const long long*)targets, epsilon, loss, dlogits); }
  return check_cuda("tfnas_softmax_ce");
}

int tfnas_softmax_ce(int N, int C, const float* logits, const int64_t* targets, float* loss, float* dlogits, void* stream) {
  return tfnas_softmax_ce_smooth(N, C, logits, targets, 0.f, loss, dlogits, stream);
}

}  // extern "C"
