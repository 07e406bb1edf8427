// The head of Network.forward (reference models/model_search.py:275-277, :299-302):
//   feature_mix_layer = ConvLayer(320, 1280, kernel 1) -> BN(batch statistics, no affine) -> Swish   (models/layers.py:190-256)
//   global_avg_pooling = AdaptiveAvgPool2d(1);  classifier = LinearLayer(1280, num_classes)           (models/layers.py:259-300)
//
// The 1x1 feature-mix convolution is the "project" GEMM of an MBConv whose input needs no BN / activation / gate, so it
// runs on the same tcgen05 kernels as the MixedOP phases with the identity activation: forward = F3 (GEMM + BN sums),
// backward = B1 (BN-backward sums), B2 (dz on load, dx = W^T dz) and the dW GEMM.  Specific to the head:
//   k_head_pool      p[n][c] = mean_hw swish(BN(z))                       (BN + Swish + pooling in one pass over z)
//   k_head_fc*       the classifier as three small tiled GEMMs (logits; dW, dp) + the bias gradient
//   k_head_dpool     G = dL/d BN(z) = dp / HW * swish'(BN(z))             (what B1 / B2 consume as "dout")
#include <string.h>
#include "api_internal.h"

struct HeadLayout {
  Plan P;
  SavedLayout L;
  size_t saved, ws, ws_bytes, pooled, dpooled, G, total;
};

static int head_layout(const TfnasHeadDesc* d, const TfnasHeadPtrs* w, HeadLayout& B) {
  if (!d) return fail(TFNAS_E_INVALID, "head: null descriptor");
  if (d->N < 1 || d->c_in < 1 || d->c_mid < 1 || d->H < 1 || d->W < 1 || d->num_classes < 1) return fail(TFNAS_E_INVALID, "head: bad shape");
  if (d->c_in > 32767) return fail(TFNAS_E_UNSUPPORTED, "head: c_in too large");
  Plan& P = B.P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.ic = d->c_in; P.oc = d->c_mid; P.H = d->H; P.W = d->W; P.Ho = d->H; P.Wo = d->W; P.stride = 1;
  P.act = TFNAS_ACT_NONE; P.na = 1; P.num_ops = 1; P.MC = d->c_in; P.MCse = 0; P.SEH = 0; P.HW = d->H * d->W; P.HWo = P.HW;
  const long long Pn = (long long)d->N * P.HW;
  if (Pn >= (1LL << 23)) return fail(TFNAS_E_UNSUPPORTED, "head: N*H*W >= 2^23 pixels per call not supported");
  P.P = (int)Pn; P.Q = (int)Pn; P.residual = 0;
  Cand& c = P.c[0];
  c.id = 0; c.mc = d->c_in; c.k = 3; c.se = 0;
  if (w) {
    if (!w->fm_w || !w->fc_w || !w->fc_b) return fail(TFNAS_E_INVALID, "head: null weight pointer");
    c.w3 = w->fm_w;
  }
  saved_layout(P, B.L);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  B.saved = take(B.L.total);
  FwdScratch F;
  BwdScratch Bs;
  B.ws_bytes = max(fwd_scratch(P, nullptr, F), bwd_scratch(P, 1, nullptr, Bs));
  B.ws = take(B.ws_bytes);
  B.pooled = take((size_t)d->N * d->c_mid * 4);
  B.dpooled = take((size_t)d->N * d->c_mid * 4);
  B.G = take((size_t)d->N * d->c_mid * P.HW * 4);
  B.total = o;
  return TFNAS_OK;
}

// bn2 = identity (mean 0, rstd 1) for the GEMM prologue, mixing weight 1
__global__ void k_head_init(int MC, float* __restrict__ bn2, float* __restrict__ mixw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < MC) { bn2[i] = 0.f; bn2[MC + i] = 1.f; }
  if (i < TFNAS_MAX_OPS) mixw[i] = 1.f;
}

__global__ void k_head_bnfin(int C, double invM, const double* __restrict__ st, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = st[2 * c] * invM;
  const double v = st[2 * c + 1] * invM - m * m;
  out[c] = (float)m;
  out[C + c] = (float)(1.0 / sqrt(fmax(v, 0.0) + (double)BN_EPS));
}

// one warp per (n, c): p = mean over hw of swish((z - mu) r)
__global__ void __launch_bounds__(NT) k_head_pool(int N, int C, int HW, const float* __restrict__ Z, const float* __restrict__ bn3,
                                                   float* __restrict__ pooled) {
  const long long w = ((long long)blockIdx.x * NT + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)N * C) return;
  const int c = (int)(w % C);
  const float mu = bn3[c], r = bn3[C + c];
  const float* z = Z + (size_t)w * HW;
  float s = 0.f;
  for (int i = lane; i < HW; i += 32) s += act_f<TFNAS_ACT_SWISH>((z[i] - mu) * r);
  s = warp_sum(s);
  if (lane == 0) pooled[w] = s / (float)HW;
}

// G[n][c][hw] = dp[n][c] / HW * swish'((z - mu) r)
__global__ void __launch_bounds__(NT) k_head_dpool(int N, int C, int HW, const float* __restrict__ Z, const float* __restrict__ bn3,
                                                    const float* __restrict__ dp, float* __restrict__ G) {
  const long long w = ((long long)blockIdx.x * NT + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)N * C) return;
  const int c = (int)(w % C);
  const float mu = bn3[c], r = bn3[C + c], g = dp[w] / (float)HW;
  const float* z = Z + (size_t)w * HW;
  float* o = G + (size_t)w * HW;
  for (int i = lane; i < HW; i += 32) o[i] = g * act_df<TFNAS_ACT_SWISH>((z[i] - mu) * r);
}

// logits[n][o] = sum_c p[n][c] W[o][c] + b[o]; split over c (blockIdx.z) into zeroed logits, the bias rides with split 0
#define HEAD_FC_SPLIT 16
__global__ void __launch_bounds__(NT) k_head_fc(int N, int C, int K, const float* __restrict__ p, const float* __restrict__ W,
                                                 const float* __restrict__ b, float* __restrict__ logits) {
  const int per = (C + gridDim.z - 1) / gridDim.z, k0 = blockIdx.z * per, k1 = min(C, k0 + per);
  const bool first = blockIdx.z == 0;
  fc_tile<false>(N, K, C, W, [&](int n, int c) { return p[(size_t)n * C + c]; },
                 [&](int n, int o, float v) { atomicAdd(&logits[(size_t)n * K + o], first ? v + b[o] : v); }, k0, k1);
}
// dp[n][c] = sum_o dl[n][o] W[o][c]
__global__ void __launch_bounds__(NT) k_head_fc_dp(int N, int C, int K, const float* __restrict__ dl, const float* __restrict__ W,
                                                    float* __restrict__ dp) {
  fc_tile<true>(N, C, K, W, [&](int n, int o) { return dl[(size_t)n * K + o]; },
                [&](int n, int c, float v) { dp[(size_t)n * C + c] = v; });
}
// dW[o][c] = sum_n dl[n][o] p[n][c]
__global__ void __launch_bounds__(NT) k_head_fc_dw(int N, int C, int K, const float* __restrict__ dl, const float* __restrict__ p,
                                                    float* __restrict__ dW) {
  fc_tile<true>(K, C, N, p, [&](int o, int n) { return dl[(size_t)n * K + o]; },
                [&](int o, int c, float v) { dW[(size_t)o * C + c] = v; });
}
__global__ void k_head_fc_db(int N, int K, const float* __restrict__ dl, float* __restrict__ db) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= K) return;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s += dl[(size_t)n * K + o];
  db[o] = s;
}

extern "C" {

size_t tfnas_head_arena_bytes(const TfnasHeadDesc* d) {
  static thread_local HeadLayout B;
  if (head_layout(d, nullptr, B) != TFNAS_OK) return 0;
  return B.total;
}

int tfnas_head_fwd(const TfnasHeadDesc* d, const float* x, const TfnasHeadPtrs* w, float* logits, void* arena,
                   size_t arena_bytes, void* stream) {
  static thread_local HeadLayout B;
  if (!w) return fail(TFNAS_E_INVALID, "head: null weights");
  int rc = head_layout(d, w, B);
  if (rc != TFNAS_OK) return rc;
  if (!x || !logits || !arena) return fail(TFNAS_E_INVALID, "head: null tensor pointer");
  if (!aligned16(x) || (((uintptr_t)arena) & 255)) return fail(TFNAS_E_INVALID, "head: x must be 16-byte, the arena 256-byte aligned");
  if (arena_bytes < B.total) return fail(TFNAS_E_WORKSPACE, "head arena %zu < %zu", arena_bytes, B.total);
  if (!umma_enabled()) return fail(TFNAS_E_UNSUPPORTED, "head: needs the tcgen05 GEMM path (TFNAS_GEMM=simt is a MixedOP debugging switch)");
  const Plan& P = B.P;
  char* A = (char*)arena;
  char* saved = A + B.saved;
  cudaStream_t st = (cudaStream_t)stream;
  FwdScratch S;
  fwd_scratch(P, A + B.ws, S);
  float* bn2 = (float*)(saved + B.L.bn2);
  float* bn3 = (float*)(saved + B.L.bn3);
  float* mixw = (float*)(saved + B.L.mixw);
  float* Zb = (float*)(saved + B.L.Z);
  float* pooled = (float*)(A + B.pooled);
  const int C = P.oc, K = d->num_classes;
  cudaGetLastError();
  cudaMemsetAsync(S.st3, 0, (size_t)2 * C * sizeof(double), st);
  { ProfScope ps("head_init", 8.0 * P.MC, 0, st);
    k_head_init<<<cdiv(max(P.MC, TFNAS_MAX_OPS), 256), 256, 0, st>>>(P.MC, bn2, mixw); }
  // the backward phases read the GEMM input from the saved buffer's D region (it is "the depthwise output" of an MBConv)
  float* D = (float*)(saved + B.L.D);
  cudaMemcpyAsync(D, x, (size_t)P.Q * P.MC * sizeof(float), cudaMemcpyDeviceToDevice, st);
  UmWAll WP;
  umma_prep_project(P, S.umprep, WP, st);
  umma_project(P, WP, D, bn2, nullptr, Zb, S.st3, st);
  { ProfScope ps("head_bnfin", 24.0 * C, 0, st);
    k_head_bnfin<<<cdiv(C, 256), 256, 0, st>>>(C, 1.0 / (double)P.Q, S.st3, bn3); }
  { ProfScope ps("head_pool", 4.0 * P.Q * C, 6.0 * P.Q * C, st);
    k_head_pool<<<cdiv((long long)P.N * C * 32, NT), NT, 0, st>>>(P.N, C, P.HW, Zb, bn3, pooled); }
  cudaMemsetAsync(logits, 0, (size_t)P.N * K * sizeof(float), st);
  { ProfScope ps("head_fc", 4.0 * (P.N * C + K * C), 2.0 * P.N * C * K, st);
    k_head_fc<<<dim3(cdiv(P.N, FC_TN), cdiv(K, FC_TO), HEAD_FC_SPLIT), NT, 0, st>>>(P.N, C, K, pooled, w->fc_w, w->fc_b, logits); }
  return check_cuda("tfnas_head_fwd");
}

int tfnas_head_bwd(const TfnasHeadDesc* d, const float* x, const TfnasHeadPtrs* w, const float* dlogits, float* dx,
                   const TfnasHeadPtrs* dw, void* arena, size_t arena_bytes, void* stream) {
  static thread_local HeadLayout B;
  if (!w) return fail(TFNAS_E_INVALID, "head: null weights");
  int rc = head_layout(d, w, B);
  if (rc != TFNAS_OK) return rc;
  if (!x || !dlogits || !dx || !arena) return fail(TFNAS_E_INVALID, "head: null tensor pointer");
  if (dw && (!dw->fm_w || !dw->fc_w || !dw->fc_b)) return fail(TFNAS_E_INVALID, "head: null weight-grad pointer");
  if (!aligned16(x) || !aligned16(dx) || (((uintptr_t)arena) & 255)) return fail(TFNAS_E_INVALID, "head: x / dx must be 16-byte, the arena 256-byte aligned");
  if (arena_bytes < B.total) return fail(TFNAS_E_WORKSPACE, "head arena %zu < %zu", arena_bytes, B.total);
  const Plan& P = B.P;
  char* A = (char*)arena;
  const char* saved = A + B.saved;
  cudaStream_t st = (cudaStream_t)stream;
  BwdScratch S;
  bwd_scratch(P, 1, A + B.ws, S);
  S.DC = dx;                      // the dc GEMM's output (dL/d input, identity BN2 / activation) IS the head's dx
  const float* bn3 = (const float*)(saved + B.L.bn3);
  const float* Zb = (const float*)(saved + B.L.Z);
  const float* pooled = (const float*)(A + B.pooled);
  float* dp = (float*)(A + B.dpooled);
  float* G = (float*)(A + B.G);
  const int C = P.oc, K = d->num_classes;
  cudaGetLastError();
  if (dw) {
    ProfScope ps("head_fc_bwd", 4.0 * (2.0 * P.N * C + 2.0 * K * C), 4.0 * P.N * C * K, st);
    k_head_fc_dw<<<dim3(cdiv(K, FC_TN), cdiv(C, FC_TO)), NT, 0, st>>>(P.N, C, K, dlogits, pooled, dw->fc_w);
    k_head_fc_db<<<cdiv(K, 128), 128, 0, st>>>(P.N, K, dlogits, dw->fc_b);
    count_launch(1);
  }
  { ProfScope ps("head_fc_dp", 4.0 * (P.N * C + K * C), 2.0 * P.N * C * K, st);
    k_head_fc_dp<<<dim3(cdiv(P.N, FC_TN), cdiv(C, FC_TO)), NT, 0, st>>>(P.N, C, K, dlogits, w->fc_w, dp); }
  { ProfScope ps("head_dpool", 8.0 * P.Q * C, 8.0 * P.Q * C, st);
    k_head_dpool<<<cdiv((long long)P.N * C * 32, NT), NT, 0, st>>>(P.N, C, P.HW, Zb, bn3, dp, G); }
  TfnasCandPtrs g;
  memset(&g, 0, sizeof(g));
  if (dw) { g.w3 = dw->fm_w; g.w1 = dw->fm_w; g.dw = dw->fm_w; }     // only w3 is written on the stop_at = 2 path
  launch_backward(P, nullptr, G, nullptr, 1.f, 0, saved, B.L, S, dx, nullptr, dw ? &g : nullptr, st, 2);
  return check_cuda("tfnas_head_bwd");
}

}  // extern "C"
