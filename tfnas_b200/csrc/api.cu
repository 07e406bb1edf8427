// C-ABI entry points (include/tfnas_b200.h): descriptor validation, buffer layout, launch sequencing.
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "kernels.h"
#include "api_internal.h"

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- optional per-launch profiling (CUDA events on the launching stream) -----------------------
struct ProfRec { const char* name; double bytes, flops; cudaEvent_t e0, e1; cudaStream_t st; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_prof_mu;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(const char* name, double bytes, double flops, cudaStream_t st) : rec_(nullptr), st_(st) {
  count_launch(1);
  if (!g_prof) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec* r = new ProfRec{name, bytes, flops, get_event(), get_event(), st};
  cudaEventRecord(r->e0, st);
  rec_ = r;
}
ProfScope::~ProfScope() {
  if (!rec_) return;
  ProfRec* r = (ProfRec*)rec_;
  cudaEventRecord(r->e1, st_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_recs.push_back(*r);
  delete r;
}

void ensure_smem_impl(const void* kern, size_t smem) {
  static std::map<const void*, size_t> limits;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = limits[kern];
  if (smem > cur) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cur = smem;
  }
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int build_plan(const TfnasMixedOpDesc* d, uint32_t mask, const TfnasCandPtrs* w, Plan& P) {
  if (!d) return fail(TFNAS_E_INVALID, "null descriptor");
  if (d->num_ops < 1 || d->num_ops > TFNAS_MAX_OPS) return fail(TFNAS_E_INVALID, "num_ops=%d out of range", d->num_ops);
  if (d->N < 1 || d->ic < 1 || d->oc < 1 || d->H < 1 || d->W < 1) return fail(TFNAS_E_INVALID, "non-positive shape");
  if (d->stride != 1 && d->stride != 2) return fail(TFNAS_E_INVALID, "stride must be 1 or 2 (got %d)", d->stride);
  if (d->act != TFNAS_ACT_RELU && d->act != TFNAS_ACT_SWISH) return fail(TFNAS_E_INVALID, "unknown act %d", d->act);
  mask &= (1u << d->num_ops) - 1u;
  if (!mask) return fail(TFNAS_E_INVALID, "empty candidate mask");
  if (d->ic > 192) return fail(TFNAS_E_UNSUPPORTED, "ic=%d > 192 not tiled by the dx / covariance kernels", d->ic);
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.ic = d->ic; P.oc = d->oc; P.H = d->H; P.W = d->W; P.stride = d->stride; P.act = d->act;
  P.num_ops = d->num_ops;
  P.HW = d->H * d->W;
  int coff = 0, soff = 0, hoff = 0, na = 0;
  for (int i = 0; i < d->num_ops; ++i) {
    if (!(mask >> i & 1)) continue;
    if (d->k[i] != 3 && d->k[i] != 5) return fail(TFNAS_E_INVALID, "candidate %d: kernel %d not in {3,5}", i, d->k[i]);
    if (d->mc[i] < 1 || d->mc[i] > 32767) return fail(TFNAS_E_INVALID, "candidate %d: mc=%d", i, d->mc[i]);
    if (d->se[i] < 0) return fail(TFNAS_E_INVALID, "candidate %d: se=%d", i, d->se[i]);
    Cand& c = P.c[na];
    c.id = i; c.mc = d->mc[i]; c.k = d->k[i]; c.se = d->se[i];
    c.coff = coff; c.soff = soff; c.hoff = hoff;
    coff += c.mc;
    if (c.se > 0) { soff += c.mc; hoff += c.se; }
    if (w) {
      c.w1 = w[i].w1; c.dw = w[i].dw; c.w3 = w[i].w3;
      c.rw = w[i].se_rw; c.rb = w[i].se_rb; c.ew = w[i].se_ew; c.eb = w[i].se_eb;
      if (!c.w1 || !c.dw || !c.w3) return fail(TFNAS_E_INVALID, "candidate %d: null weight pointer", i);
      if (c.se > 0 && (!c.rw || !c.rb || !c.ew || !c.eb)) return fail(TFNAS_E_INVALID, "candidate %d: null SE pointer", i);
    }
    ++na;
  }
  P.na = na; P.MC = coff; P.MCse = soff; P.SEH = hoff;
  // output size of the depthwise conv, padding k//2 (identical for k=3 and k=5)
  P.Ho = (d->H + 2 * 1 - 3) / d->stride + 1;
  P.Wo = (d->W + 2 * 1 - 3) / d->stride + 1;
  P.HWo = P.Ho * P.Wo;
  long long Pn = (long long)d->N * P.HW, Qn = (long long)d->N * P.HWo;
  if (Pn * (long long)P.MC >= (1LL << 40)) return fail(TFNAS_E_UNSUPPORTED, "tensor too large");
  if (Pn >= (1LL << 23) || Qn >= (1LL << 23)) return fail(TFNAS_E_UNSUPPORTED, "N*H*W >= 2^23 pixels per call not supported");
  P.P = (int)Pn; P.Q = (int)Qn;
  P.residual = (d->ic == d->oc && d->stride == 1);
  return TFNAS_OK;
}

void saved_layout(const Plan& P, SavedLayout& L) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.xmom = take((size_t)(P.ic + P.ic * P.ic) * sizeof(double));
  L.bn1 = take((size_t)2 * P.MC * 4);
  L.bn2 = take((size_t)2 * P.MC * 4);
  L.bn3 = take((size_t)2 * P.na * P.oc * 4);
  L.mixw = take(TFNAS_MAX_OPS * 4);
  L.lat = take(TFNAS_MAX_OPS * 4);
  L.sep = take((size_t)P.N * P.MCse * 4);
  L.set = take((size_t)P.N * P.SEH * 4);
  L.seg = take((size_t)P.N * P.MCse * 4);
  L.UH = take((size_t)P.N * P.MC * P.HW * 4);
  L.D = take((size_t)P.N * P.MC * P.HWo * 4);
  L.Z = take((size_t)P.N * P.na * P.oc * P.HWo * 4);
  L.total = o;
}

size_t fwd_scratch(const Plan& P, char* base, FwdScratch& S) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  // the four double accumulators must be CONTIGUOUS (one memset): take them as one block
  size_t nd = (size_t)P.ic + (size_t)P.ic * P.ic + 2 * (size_t)P.MC + 2 * (size_t)P.na * P.oc;
  size_t acc = take(nd * sizeof(double));
  size_t coef = take((size_t)(P.na * P.oc + P.oc) * 4);
  size_t umprep = take(umma_fwd_prep_bytes(P));
  const size_t nb4 = ((size_t)P.ic + 1 + 3) / 4;
  size_t xpart = take((size_t)XM_MAXCTA * (nb4 * (nb4 + 1) / 2) * 16 * sizeof(float));
  if (base) {
    S.umprep = (float*)(base + umprep);
    S.xpart = (float*)(base + xpart);
    S.xsum = (double*)(base + acc);
    S.xcov = S.xsum + P.ic;
    S.st2 = S.xcov + (size_t)P.ic * P.ic;
    S.st3 = S.st2 + 2 * (size_t)P.MC;
    S.coef = (float*)(base + coef);
  }
  return o;
}

size_t bwd_scratch(const Plan& P, int want_wgrad, char* base, BwdScratch& S) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  // everything the backward accumulates into with atomics sits in ONE zeroed region: [sG | sGY | sD | sU] doubles, then
  // cvec2, Mm, sedt, dg -- a single memset at the start of the call (S.zero_bytes) instead of four along the way
  size_t nd = (size_t)P.oc + (size_t)P.na * P.oc + 4 * (size_t)P.MC;
  size_t acc = take(nd * sizeof(double));
  size_t cvec2 = take((size_t)P.ic * 4);
  size_t Mm = take((size_t)P.ic * P.ic * 4);
  size_t sedt = take((size_t)P.N * P.SEH * 4 + 16);
  size_t dg = take((size_t)P.N * P.MCse * 4);
  const size_t zero_end = o;
  size_t dzc = take((size_t)P.na * P.oc * sizeof(float4));
  size_t dzc2 = take((size_t)P.na * P.oc * sizeof(float4));
  size_t a12 = take((size_t)2 * P.MC * 4);
  size_t dmix = take(TFNAS_MAX_OPS * 4);
  size_t sede = want_wgrad ? take((size_t)P.N * P.MCse * 4) : 0;
  size_t Smat = want_wgrad ? take((size_t)P.MC * P.ic * 4) : 0;
  size_t DC = take((size_t)P.N * P.MC * P.HWo * 4);
  size_t DA = take((size_t)P.N * P.MC * P.HW * 4);
  size_t umprep = take(umma_bwd_prep_bytes(P));
  if (base) {
    S.umprep = (float*)(base + umprep);
    S.sG = (double*)(base + acc);
    S.zero_bytes = zero_end - acc;
    S.sGY = S.sG + P.oc;
    S.sD = S.sGY + (size_t)P.na * P.oc;
    S.sU = S.sD + 2 * (size_t)P.MC;
    S.dzc = (float4*)(base + dzc);
    S.dzc2 = (float4*)(base + dzc2);
    S.cvec2 = (float*)(base + cvec2);
    S.Mm = (float*)(base + Mm);
    S.a12 = (float*)(base + a12);
    S.dmix = (float*)(base + dmix);
    S.dg = (float*)(base + dg);
    S.sede = want_wgrad ? (float*)(base + sede) : nullptr;
    S.sedt = (float*)(base + sedt);
    S.Smat = want_wgrad ? (float*)(base + Smat) : nullptr;
    S.DC = (float*)(base + DC);
    S.DA = (float*)(base + DA);
  }
  return o;
}

int check_cuda(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TFNAS_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return TFNAS_OK;
}

extern "C" {

int tfnas_version(void) { return TFNAS_ABI_VERSION; }
const char* tfnas_last_error(void) { return g_err; }
uint64_t tfnas_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t tfnas_mixedop_saved_bytes(const TfnasMixedOpDesc* d, uint32_t cand_mask) {
  Plan P;
  if (build_plan(d, cand_mask, nullptr, P) != TFNAS_OK) return 0;
  SavedLayout L;
  saved_layout(P, L);
  return L.total;
}

size_t tfnas_mixedop_workspace_bytes(const TfnasMixedOpDesc* d, uint32_t cand_mask, int want_wgrad) {
  Plan P;
  if (build_plan(d, cand_mask, nullptr, P) != TFNAS_OK) return 0;
  FwdScratch F;
  BwdScratch B;
  size_t f = fwd_scratch(P, nullptr, F), b = bwd_scratch(P, want_wgrad, nullptr, B);
  return f > b ? f : b;
}

int tfnas_mixedop_fwd(const TfnasMixedOpDesc* d, uint32_t cand_mask, const float* x, const TfnasCandPtrs* weights,
                      const float* log_alphas, const float* gumbel, const float* lat8, float T, float* out,
                      float* out_lat, void* saved, size_t saved_bytes, void* workspace, size_t ws_bytes,
                      void* stream) {
  Plan P;
  if (!weights) return fail(TFNAS_E_INVALID, "null weights");
  int rc = build_plan(d, cand_mask, weights, P);
  if (rc != TFNAS_OK) return rc;
  if (!x || !out || !saved || !workspace) return fail(TFNAS_E_INVALID, "null tensor pointer");
  if (!aligned16(x) || !aligned16(out) || !aligned16(saved) || !aligned16(workspace))
    return fail(TFNAS_E_INVALID, "x / out / saved / workspace must be 16-byte aligned");
  const uint32_t full = (1u << d->num_ops) - 1u;
  const int alpha_mode = (cand_mask & full) == full && d->num_ops > 1 ? 1 : 0;
  if (alpha_mode) {
    if (!log_alphas || !gumbel || !lat8 || !out_lat) return fail(TFNAS_E_INVALID, "alpha mode needs log_alphas/gumbel/lat8/out_lat");
    if (!(T > 0.f)) return fail(TFNAS_E_INVALID, "temperature must be > 0");
  } else if (P.na != 1) {
    return fail(TFNAS_E_INVALID, "cand_mask must be all candidates (alpha mode) or one-hot (sampled mode)");
  }
  SavedLayout L;
  saved_layout(P, L);
  if (saved_bytes < L.total) return fail(TFNAS_E_WORKSPACE, "saved buffer %zu < %zu", saved_bytes, L.total);
  FwdScratch S;
  size_t need = fwd_scratch(P, (char*)workspace, S);
  if (ws_bytes < need) return fail(TFNAS_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  cudaGetLastError();
  launch_forward(P, x, log_alphas, gumbel, lat8, T, alpha_mode, out, out_lat, (char*)saved, L, S, (cudaStream_t)stream);
  return check_cuda("tfnas_mixedop_fwd");
}

int tfnas_mixedop_bwd(const TfnasMixedOpDesc* d, uint32_t cand_mask, const float* x, const TfnasCandPtrs* weights,
                      const float* dout, const float* dlat, float T, const void* saved, size_t saved_bytes, float* dx,
                      float* dlog_alphas, const TfnasCandPtrs* dweights, void* workspace, size_t ws_bytes,
                      void* stream) {
  Plan P;
  if (!weights) return fail(TFNAS_E_INVALID, "null weights");
  int rc = build_plan(d, cand_mask, weights, P);
  if (rc != TFNAS_OK) return rc;
  if (!x || !dout || !saved || !workspace) return fail(TFNAS_E_INVALID, "null tensor pointer");
  if (!aligned16(x) || !aligned16(dout) || !aligned16(dx) || !aligned16(saved) || !aligned16(workspace))
    return fail(TFNAS_E_INVALID, "x / dout / dx / saved / workspace must be 16-byte aligned");
  if (!dx && dweights) return fail(TFNAS_E_INVALID, "weight gradients need dx");
  const uint32_t full = (1u << d->num_ops) - 1u;
  const int alpha_mode = (cand_mask & full) == full && d->num_ops > 1 ? 1 : 0;
  if (!alpha_mode && P.na != 1) return fail(TFNAS_E_INVALID, "cand_mask must be all candidates or one-hot");
  if (alpha_mode && !(T > 0.f)) return fail(TFNAS_E_INVALID, "temperature must be > 0");
  if (dweights) {
    for (int s = 0; s < P.na; ++s) {
      const TfnasCandPtrs& g = dweights[P.c[s].id];
      if (!g.w1 || !g.dw || !g.w3) return fail(TFNAS_E_INVALID, "candidate %d: null weight-grad pointer", P.c[s].id);
      if (P.c[s].se > 0 && (!g.se_rw || !g.se_rb || !g.se_ew || !g.se_eb))
        return fail(TFNAS_E_INVALID, "candidate %d: null SE weight-grad pointer", P.c[s].id);
    }
  }
  SavedLayout L;
  saved_layout(P, L);
  if (saved_bytes < L.total) return fail(TFNAS_E_WORKSPACE, "saved buffer %zu < %zu", saved_bytes, L.total);
  BwdScratch S;
  size_t need = bwd_scratch(P, dweights != nullptr, (char*)workspace, S);
  if (ws_bytes < need) return fail(TFNAS_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  cudaGetLastError();
  launch_backward(P, x, dout, dlat, T, alpha_mode, (const char*)saved, L, S, dx, dlog_alphas, dweights,
                  (cudaStream_t)stream);
  return check_cuda("tfnas_mixedop_bwd");
}

int tfnas_bn_act_fwd(int N, int C, int HW, int act, const float* x, float* y, float* mean_rstd, void* workspace,
                     size_t ws_bytes, void* stream) {
  if (N < 1 || C < 1 || HW < 1 || act < 0 || act > 2) return fail(TFNAS_E_INVALID, "bn_act: bad shape / act");
  if (!x || !y || !mean_rstd || !workspace) return fail(TFNAS_E_INVALID, "null pointer");
  if ((long long)N * HW >= (1LL << 23)) return fail(TFNAS_E_UNSUPPORTED, "bn_act: N*HW >= 2^23");
  if (ws_bytes < (size_t)2 * C * sizeof(double)) return fail(TFNAS_E_WORKSPACE, "bn_act workspace %zu < %zu", ws_bytes, (size_t)16 * C);
  cudaGetLastError();
  launch_bn_act_fwd(N, C, HW, act, x, y, mean_rstd, (double*)workspace, (cudaStream_t)stream);
  return check_cuda("tfnas_bn_act_fwd");
}

int tfnas_bn_act_bwd(int N, int C, int HW, int act, const float* x, const float* mean_rstd, const float* dy, float* dx,
                     void* workspace, size_t ws_bytes, void* stream) {
  if (N < 1 || C < 1 || HW < 1 || act < 0 || act > 2) return fail(TFNAS_E_INVALID, "bn_act: bad shape / act");
  if (!x || !dy || !dx || !mean_rstd || !workspace) return fail(TFNAS_E_INVALID, "null pointer");
  if ((long long)N * HW >= (1LL << 23)) return fail(TFNAS_E_UNSUPPORTED, "bn_act: N*HW >= 2^23");
  if (ws_bytes < (size_t)2 * C * sizeof(double)) return fail(TFNAS_E_WORKSPACE, "bn_act workspace %zu < %zu", ws_bytes, (size_t)16 * C);
  cudaGetLastError();
  launch_bn_act_bwd(N, C, HW, act, x, mean_rstd, dy, dx, (double*)workspace, (cudaStream_t)stream);
  return check_cuda("tfnas_bn_act_bwd");
}

int tfnas_dwconv_fwd(int N, int C, int H, int W, int K, int stride, const float* x, const float* w, float* y, void* stream) {
  if (N < 1 || C < 1 || H < 1 || W < 1 || (K != 3 && K != 5)) return fail(TFNAS_E_INVALID, "dwconv: bad shape / kernel size");
  if (stride != 1) return fail(TFNAS_E_UNSUPPORTED, "dwconv: only stride 1 (the second stem) is implemented");
  if ((long long)N * C > 65535) return fail(TFNAS_E_UNSUPPORTED, "dwconv: N*C > 65535 planes");
  if (!x || !w || !y) return fail(TFNAS_E_INVALID, "null pointer");
  cudaGetLastError();
  launch_dwconv_fwd(N, C, H, W, K, x, w, y, (cudaStream_t)stream);
  return check_cuda("tfnas_dwconv_fwd");
}

int tfnas_dwconv_bwd(int N, int C, int H, int W, int K, int stride, const float* x, const float* w, const float* dy,
                     float* dx, float* dw, void* stream) {
  if (N < 1 || C < 1 || H < 1 || W < 1 || (K != 3 && K != 5)) return fail(TFNAS_E_INVALID, "dwconv: bad shape / kernel size");
  if (stride != 1) return fail(TFNAS_E_UNSUPPORTED, "dwconv: only stride 1 (the second stem) is implemented");
  if ((long long)N * C > 65535) return fail(TFNAS_E_UNSUPPORTED, "dwconv: N*C > 65535 planes");
  if (!w || !dy || (dw && !x)) return fail(TFNAS_E_INVALID, "null pointer");
  cudaGetLastError();
  launch_dwconv_bwd(N, C, H, W, K, x, w, dy, dx, dw, (cudaStream_t)stream);
  return check_cuda("tfnas_dwconv_bwd");
}

int tfnas_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_recs) { g_pool.push_back(r.e0); g_pool.push_back(r.e1); }
  g_recs.clear();
  g_prof = on != 0;
  return TFNAS_OK;
}

int tfnas_prof_collect(TfnasProfEntry* out, int max_entries) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, TfnasProfEntry> agg;
  std::vector<std::string> order;
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return fail(TFNAS_E_CUDA, "prof: event sync failed");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto it = agg.find(r.name);
    if (it == agg.end()) {
      TfnasProfEntry e;
      memset(&e, 0, sizeof(e));
      strncpy(e.name, r.name, sizeof(e.name) - 1);
      it = agg.insert({r.name, e}).first;
      order.push_back(r.name);
    }
    it->second.ms += ms;
    it->second.bytes += r.bytes;
    it->second.flops += r.flops;
    it->second.launches += 1;
  }
  int n = 0;
  for (auto& k : order) {
    if (n >= max_entries) break;
    out[n++] = agg[k];
  }
  return n;
}

/* Raw launch timeline of the profiled region: per recorded launch its name, stream and start / end time in ms relative to
 * the first recorded launch's start (events on the launching streams; tools/timeline.py turns it into per-stream gaps). */
int tfnas_prof_timeline(TfnasProfLaunch* out, int max_entries) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_recs.empty()) return 0;
  int n = 0;
  cudaEvent_t base = g_recs[0].e0;
  for (auto& r : g_recs) {
    if (n >= max_entries) break;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return fail(TFNAS_E_CUDA, "prof: event sync failed");
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, base, r.e0);
    cudaEventElapsedTime(&b, base, r.e1);
    memset(&out[n], 0, sizeof(out[n]));
    strncpy(out[n].name, r.name, sizeof(out[n].name) - 1);
    out[n].stream = (uint64_t)(uintptr_t)r.st;
    out[n].start_ms = a;
    out[n].end_ms = b;
    ++n;
  }
  return n;
}

/* debug/test helper: byte offsets of the saved-buffer regions, in SavedLayout order (12 entries + total) */
int tfnas_debug_saved_layout(const TfnasMixedOpDesc* d, uint32_t cand_mask, size_t* out13) {
  Plan P;
  int rc = build_plan(d, cand_mask, nullptr, P);
  if (rc != TFNAS_OK) return rc;
  SavedLayout L;
  saved_layout(P, L);
  const size_t v[13] = {L.xmom, L.bn1, L.bn2, L.bn3, L.mixw, L.lat, L.sep, L.set, L.seg, L.UH, L.D, L.Z, L.total};
  for (int i = 0; i < 13; ++i) out13[i] = v[i];
  return TFNAS_OK;
}

/* debug/test helper: byte offsets of backward-workspace regions {sG, sGY, sD, sU, cvec2, Mm, dg, DC, DA, total} */
int tfnas_debug_bwd_layout(const TfnasMixedOpDesc* d, uint32_t cand_mask, int want_wgrad, size_t* out10) {
  Plan P;
  int rc = build_plan(d, cand_mask, nullptr, P);
  if (rc != TFNAS_OK) return rc;
  BwdScratch S;
  char* base = (char*)4096;
  size_t total = bwd_scratch(P, want_wgrad, base, S);
  const char* v[9] = {(char*)S.sG, (char*)S.sGY, (char*)S.sD, (char*)S.sU, (char*)S.cvec2, (char*)S.Mm, (char*)S.dg,
                      (char*)S.DC, (char*)S.DA};
  for (int i = 0; i < 9; ++i) out10[i] = (size_t)(v[i] - base);
  out10[9] = total;
  return TFNAS_OK;
}

int tfnas_stage_sink_fwd(int K, size_t numel, const float* const* res, const float* betas, const float* cumlat,
                         float* out, float* out_lat, void* stream) {
  if (K < 1 || K > 4) return fail(TFNAS_E_INVALID, "sink K=%d not in 1..4", K);
  if (!res || !betas || !out) return fail(TFNAS_E_INVALID, "null pointer");
  if (!aligned16(out)) return fail(TFNAS_E_INVALID, "sink: out must be 16-byte aligned");
  for (int j = 0; j < K; ++j)
    if (!res[j] || !aligned16(res[j])) return fail(TFNAS_E_INVALID, "sink: res[%d] null or not 16-byte aligned", j);
  cudaGetLastError();
  launch_sink_fwd(K, numel, res, betas, cumlat, out, out_lat, (cudaStream_t)stream);
  return check_cuda("tfnas_stage_sink_fwd");
}

int tfnas_stage_sink_bwd(int K, size_t numel, const float* const* res, const float* betas, const float* cumlat,
                         const float* dout, const float* dlat, float* const* dres, float* dbetas, float* dcumlat,
                         void* workspace, size_t ws_bytes, void* stream) {
  if (K < 1 || K > 4) return fail(TFNAS_E_INVALID, "sink K=%d not in 1..4", K);
  if (!res || !betas || !dout || !dres || !dbetas || !workspace) return fail(TFNAS_E_INVALID, "null pointer");
  if (ws_bytes < 4 * sizeof(double)) return fail(TFNAS_E_WORKSPACE, "sink workspace %zu < 32", ws_bytes);
  cudaGetLastError();
  launch_sink_bwd(K, numel, res, betas, cumlat, dout, dlat, dres, dbetas, dcumlat, (double*)workspace,
                  (cudaStream_t)stream);
  return check_cuda("tfnas_stage_sink_bwd");
}

}  // extern "C"
