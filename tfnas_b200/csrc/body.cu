// Supernet body executor: the six MixedStages (18 MixedOPs + 6 sink-connecting sums) of Network.forward and their
// backward, sequenced in C++ over ONE caller-provided arena (reference models/model_search.py:157-206 per stage,
// :291-296 the stage loop of Network.forward; backward = what autograd derives from them).
//
// One C-ABI call per direction replaces 24 autograd Functions, ~60 torch allocations and ~50 ctypes marshalling round trips
// per pass: the host cost of a pass is the kernel launches themselves.  Nothing here allocates, synchronises or touches a
// stream other than the caller's; the launch sequence depends only on (descriptor, candidate masks, pointers), so a call is
// CUDA-graph capturable.
//
// Arena layout (body_layout): per MixedOP its output tensor, its saved buffer and two latency scalars; per stage the sink
// output (the last stage writes the caller's `out`) and the small sink scratch; ONE workspace sized for the largest
// MixedOP; four gradient buffers (two stage-level, two block-level ping-pong).
#include <string.h>
#include "api_internal.h"

struct OpSlot {
  Plan P;
  SavedLayout L;
  size_t out, saved, lat, dlat;   // byte offsets in the arena (lat / dlat: one float each)
  size_t prep_f, prep_b;          // pre-split / pre-swizzled tcgen05 weights of the forward / backward GEMMs
  size_t out_numel, in_numel;
};
struct StageSlot {
  int first, K;                   // first MixedOP index, number of MixedOPs
  size_t out;                     // sink output (unused for the last stage)
  size_t cumlat, dcumlat, lat;    // float[4], float[4], float[1]
  size_t dots;                    // double[4]
  size_t numel;
};
struct BodyLayout {
  int nb, ns;
  OpSlot op[TFNAS_MAX_BLOCKS];
  StageSlot st[TFNAS_MAX_STAGES];
  size_t ws, ws_bytes;
  size_t gS[2], gB[2];
  size_t gS_numel, gB_numel;
  size_t total;
};

static int body_layout(const TfnasBodyDesc* d, const uint32_t* masks, const TfnasCandPtrs* w, int want_wgrad, BodyLayout& B) {
  if (!d || !masks) return fail(TFNAS_E_INVALID, "body: null descriptor / masks");
  if (d->num_stages < 1 || d->num_stages > TFNAS_MAX_STAGES) return fail(TFNAS_E_INVALID, "body: num_stages=%d", d->num_stages);
  int nb = 0;
  for (int s = 0; s < d->num_stages; ++s) {
    if (d->stage_blocks[s] < 1 || d->stage_blocks[s] > 4) return fail(TFNAS_E_UNSUPPORTED, "body: stage %d has %d blocks (1..4)", s, d->stage_blocks[s]);
    nb += d->stage_blocks[s];
  }
  if (nb != d->num_blocks || nb > TFNAS_MAX_BLOCKS) return fail(TFNAS_E_INVALID, "body: num_blocks=%d does not match the stages (%d)", d->num_blocks, nb);
  B.nb = nb;
  B.ns = d->num_stages;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  size_t wsb = 0, gS = 0, gB = 0;
  int bi = 0;
  int N = 0, C = 0, H = 0, W = 0;
  for (int s = 0; s < d->num_stages; ++s) {
    StageSlot& S = B.st[s];
    S.first = bi;
    S.K = d->stage_blocks[s];
    for (int j = 0; j < S.K; ++j, ++bi) {
      OpSlot& O = B.op[bi];
      const TfnasMixedOpDesc* od = &d->op[bi];
      int rc = build_plan(od, masks[bi], w ? w + (size_t)bi * TFNAS_MAX_OPS : nullptr, O.P);
      if (rc != TFNAS_OK) return rc;
      if (bi > 0 && (od->N != N || od->ic != C || od->H != H || od->W != W))
        return fail(TFNAS_E_INVALID, "body: MixedOP %d input [%d,%d,%d,%d] does not chain with [%d,%d,%d,%d]", bi, od->N, od->ic,
                    od->H, od->W, N, C, H, W);
      N = od->N; C = od->oc; H = O.P.Ho; W = O.P.Wo;
      if (j > 0 && (B.op[bi - 1].P.oc != O.P.oc || B.op[bi - 1].P.HWo != O.P.HWo))
        return fail(TFNAS_E_INVALID, "body: outputs of stage %d differ in shape (the sink sums them)", s);
      saved_layout(O.P, O.L);
      O.in_numel = (size_t)O.P.N * O.P.ic * O.P.HW;
      O.out_numel = (size_t)O.P.N * O.P.oc * O.P.HWo;
      O.out = take(O.out_numel * 4);
      O.saved = take(O.L.total);
      O.lat = take(4);
      O.dlat = take(4);
      O.prep_f = take(umma_fwd_prep_bytes(O.P));
      O.prep_b = take(umma_bwd_prep_bytes(O.P));
      FwdScratch F;
      BwdScratch Bs;
      wsb = max(wsb, max(fwd_scratch(O.P, nullptr, F), bwd_scratch(O.P, want_wgrad, nullptr, Bs)));
      gB = max(gB, max(O.in_numel, O.out_numel));
    }
    S.numel = B.op[bi - 1].out_numel;
    S.out = take(S.numel * 4);
    S.cumlat = take(16);
    S.dcumlat = take(16);
    S.lat = take(4);
    S.dots = take(32);
    gS = max(gS, max(S.numel, B.op[S.first].in_numel));
  }
  B.ws_bytes = wsb;
  B.ws = take(wsb);
  B.gS_numel = gS;
  B.gB_numel = gB;
  for (int i = 0; i < 2; ++i) B.gS[i] = take(gS * 4);
  for (int i = 0; i < 2; ++i) B.gB[i] = take(gB * 4);
  B.total = o;
  return TFNAS_OK;
}

// ---- small kernels of the executor ---------------------------------------------------------------------------------
// cumulative latencies of one stage (reference models/model_search.py:172-199: lat_list = [lat1, lat1+lat2, ...])
__global__ void k_cumlat(int K, const float* const __restrict__ l0, const float* const __restrict__ l1,
                         const float* const __restrict__ l2, const float* const __restrict__ l3, float* __restrict__ cum) {
  if (threadIdx.x != 0) return;
  const float* l[4] = {l0, l1, l2, l3};
  float c = 0.f;
  for (int j = 0; j < K; ++j) { c += *l[j]; cum[j] = c; }
}
// backward of k_cumlat: dlat_i = sum_{j >= i} dcum_j
__global__ void k_cumlat_bwd(int K, const float* __restrict__ dcum, float* d0, float* d1, float* d2, float* d3) {
  if (threadIdx.x != 0) return;
  float* d[4] = {d0, d1, d2, d3};
  float c = 0.f;
  for (int j = K - 1; j >= 0; --j) { c += dcum[j]; *d[j] = c; }
}
// out_lat = sum of the stage latencies (Network.forward :294-296 adds them to lut['base'] on the host side of the ABI)
struct LatPtrs { const float* p[TFNAS_MAX_STAGES]; };
__global__ void k_lat_total(int ns, LatPtrs L, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  float t = 0.f;
  for (int s = 0; s < ns; ++s) t += *L.p[s];
  *out = t;
}

__device__ __forceinline__ float softmax_at(const float* betas, int K, int j) {
  float m = -INFINITY;
  for (int i = 0; i < K; ++i) m = fmaxf(m, betas[i]);
  float s = 0.f;
  for (int i = 0; i < K; ++i) s += expf(betas[i] - m);
  return expf(betas[j] - m) / s;
}

// Backward of the sink without materialising K scaled copies of dout: per-block dots <dout, res_j> for d(beta), and only the
// LAST block's gradient beta_K * dout (the earlier blocks get theirs added onto the next block's dx by k_axpy_beta).
struct SinkRes { const float* res[4]; };
__global__ void __launch_bounds__(NT) k_sink_bwd_last(int K, size_t numel, SinkRes p, const float* __restrict__ betas,
                                                       const float* __restrict__ dout, float* __restrict__ dlast,
                                                       double* __restrict__ dots) {
  const float bl = softmax_at(betas, K, K - 1);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t stride = (size_t)gridDim.x * NT;
  if ((numel & 3) == 0) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < numel / 4; i += stride) {
      const float4 g = ((const float4*)dout)[i];
      for (int j = 0; j < K; ++j) {
        const float4 r = ((const float4*)p.res[j])[i];
        acc[j] += g.x * r.x + g.y * r.y + g.z * r.z + g.w * r.w;
      }
      ((float4*)dlast)[i] = make_float4(bl * g.x, bl * g.y, bl * g.z, bl * g.w);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < numel; i += stride) {
      const float g = dout[i];
      for (int j = 0; j < K; ++j) acc[j] += g * p.res[j][i];
      dlast[i] = bl * g;
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

// dst += softmax(betas)[j] * g   (gradient of block j's output = dx of block j+1 + its share of the sink)
__global__ void __launch_bounds__(NT) k_axpy_beta(size_t numel, const float* __restrict__ betas, int K, int j,
                                                   const float* __restrict__ g, float* __restrict__ dst) {
  const float b = softmax_at(betas, K, j);
  const size_t stride = (size_t)gridDim.x * NT;
  if ((numel & 3) == 0) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < numel / 4; i += stride) {
      const float4 a = ((const float4*)g)[i];
      float4 d = ((float4*)dst)[i];
      d.x += b * a.x; d.y += b * a.y; d.z += b * a.z; d.w += b * a.w;
      ((float4*)dst)[i] = d;
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < numel; i += stride) dst[i] += b * g[i];
  }
}

// d(beta)_k = b_k (t_k - sum_j t_j b_j), t_j = <dout, res_j> + dlat * cumlat_j ; dcumlat_j = b_j dlat
__global__ void k_sink_fin2(int K, const float* __restrict__ betas, const float* __restrict__ cumlat,
                            const float* __restrict__ dlat, const double* __restrict__ dots, float* __restrict__ dbetas,
                            float* __restrict__ dcumlat) {
  if (threadIdx.x != 0) return;
  float b[4], t[4], dot = 0.f;
  for (int j = 0; j < K; ++j) b[j] = softmax_at(betas, K, j);
  const float dl = dlat ? *dlat : 0.f;
  for (int j = 0; j < K; ++j) {
    t[j] = (float)dots[j] + (cumlat ? dl * cumlat[j] : 0.f);
    dot += t[j] * b[j];
  }
  for (int j = 0; j < K; ++j) {
    if (dbetas) dbetas[j] = b[j] * (t[j] - dot);
    if (dcumlat) dcumlat[j] = b[j] * dl;
  }
}

static int alpha_mode_of(const TfnasMixedOpDesc& od, uint32_t mask) {
  const uint32_t full = (1u << od.num_ops) - 1u;
  return ((mask & full) == full && od.num_ops > 1) ? 1 : 0;
}

extern "C" {

size_t tfnas_body_arena_bytes(const TfnasBodyDesc* d, const uint32_t* cand_masks, int want_wgrad) {
  static thread_local BodyLayout B;
  if (body_layout(d, cand_masks, nullptr, want_wgrad, B) != TFNAS_OK) return 0;
  return B.total;
}

int tfnas_body_fwd(const TfnasBodyDesc* d, const uint32_t* cand_masks, const float* x, const TfnasCandPtrs* weights,
                   const float* const* log_alphas, const float* const* betas, const float* gumbel, const float* lat,
                   float T, float* out, float* out_lat, void* arena, size_t arena_bytes, void* stream) {
  static thread_local BodyLayout B;
  if (!weights || !betas) return fail(TFNAS_E_INVALID, "body: null weights / betas");
  int rc = body_layout(d, cand_masks, weights, 0, B);
  if (rc != TFNAS_OK) return rc;
  if (!x || !out || !arena) return fail(TFNAS_E_INVALID, "body: null tensor pointer");
  if (!aligned16(x) || !aligned16(out) || (((uintptr_t)arena) & 255)) return fail(TFNAS_E_INVALID, "body: x / out must be 16-byte, the arena 256-byte aligned");
  // the arena of a forward must also hold the backward's scratch: size it with the same want_wgrad the backward will use
  if (arena_bytes < B.total) return fail(TFNAS_E_WORKSPACE, "body arena %zu < %zu", arena_bytes, B.total);
  int any_alpha = 0;
  for (int i = 0; i < B.nb; ++i) {
    const int am = alpha_mode_of(d->op[i], cand_masks[i]);
    any_alpha |= am;
    if (am) {
      if (!log_alphas || !log_alphas[i] || !gumbel || !lat) return fail(TFNAS_E_INVALID, "body: MixedOP %d in alpha mode needs log_alphas / gumbel / lat", i);
      if (!(T > 0.f)) return fail(TFNAS_E_INVALID, "temperature must be > 0");
    } else if (B.op[i].P.na != 1) {
      return fail(TFNAS_E_INVALID, "body: MixedOP %d: cand_mask must be all candidates or one-hot", i);
    }
  }
  if (any_alpha && !out_lat) return fail(TFNAS_E_INVALID, "body: alpha mode needs out_lat");
  for (int s = 0; s < B.ns; ++s)
    if (!betas[s]) return fail(TFNAS_E_INVALID, "body: null betas of stage %d", s);
  char* A = (char*)arena;
  cudaStream_t st = (cudaStream_t)stream;
  cudaGetLastError();
  // every weight of the pass is known up front: convert them all for the tcgen05 GEMMs in a few launches
  static thread_local PreppedFwd PF[TFNAS_MAX_BLOCKS];
  const bool pre = umma_enabled() != 0;
  if (pre) {
    umma_prep_batch_begin();
    for (int i = 0; i < B.nb; ++i) umma_prep_fwd(B.op[i].P, (float*)(A + B.op[i].prep_f), PF[i].WE, PF[i].WP, st);
    umma_prep_batch_flush(st);
  }
  const float* cur = x;
  LatPtrs LP;
  memset(&LP, 0, sizeof(LP));
  for (int s = 0; s < B.ns; ++s) {
    const StageSlot& S = B.st[s];
    const float* res[4] = {nullptr, nullptr, nullptr, nullptr};
    const float* lats[4] = {nullptr, nullptr, nullptr, nullptr};
    int stage_alpha = 0;
    for (int j = 0; j < S.K; ++j) {
      const int i = S.first + j;
      const OpSlot& O = B.op[i];
      const int am = alpha_mode_of(d->op[i], cand_masks[i]);
      stage_alpha |= am;
      FwdScratch F;
      fwd_scratch(O.P, A + B.ws, F);
      float* o = (float*)(A + O.out);
      float* l = (float*)(A + O.lat);
      if (!am) cudaMemsetAsync(l, 0, 4, st);
      launch_forward(O.P, cur, am ? log_alphas[i] : nullptr, am ? gumbel + (size_t)i * TFNAS_MAX_OPS : nullptr,
                     am ? lat + (size_t)i * TFNAS_MAX_OPS : nullptr, T, am, o, l, A + O.saved, O.L, F, st, pre ? &PF[i] : nullptr);
      res[j] = o;
      lats[j] = l;
      cur = o;
    }
    float* so = (s == B.ns - 1) ? out : (float*)(A + S.out);
    float* cum = (float*)(A + S.cumlat);
    float* sl = (float*)(A + S.lat);
    if (stage_alpha) {
      count_launch(1);
      k_cumlat<<<1, 32, 0, st>>>(S.K, lats[0], lats[1] ? lats[1] : lats[0], lats[2] ? lats[2] : lats[0],
                                 lats[3] ? lats[3] : lats[0], cum);
    }
    launch_sink_fwd(S.K, S.numel, res, betas[s], stage_alpha ? cum : nullptr, so, stage_alpha ? sl : nullptr, st);
    LP.p[s] = sl;
    if (!stage_alpha) cudaMemsetAsync(sl, 0, 4, st);
    cur = so;
  }
  if (out_lat) {
    count_launch(1);
    k_lat_total<<<1, 32, 0, st>>>(B.ns, LP, out_lat);
  }
  return check_cuda("tfnas_body_fwd");
}

int tfnas_body_bwd(const TfnasBodyDesc* d, const uint32_t* cand_masks, const float* x, const TfnasCandPtrs* weights,
                   const float* const* betas, const float* dout, const float* dlat, float T, float* dx,
                   float* const* dlog_alphas, float* const* dbetas, const TfnasCandPtrs* dweights, void* arena,
                   size_t arena_bytes, void* stream) {
  static thread_local BodyLayout B;
  if (!weights || !betas) return fail(TFNAS_E_INVALID, "body: null weights / betas");
  int rc = body_layout(d, cand_masks, weights, dweights != nullptr, B);
  if (rc != TFNAS_OK) return rc;
  if (!x || !dout || !arena) return fail(TFNAS_E_INVALID, "body: null tensor pointer");
  if (!aligned16(x) || !aligned16(dout) || !aligned16(dx) || (((uintptr_t)arena) & 255))
    return fail(TFNAS_E_INVALID, "body: x / dout / dx must be 16-byte, the arena 256-byte aligned");
  if (arena_bytes < B.total) return fail(TFNAS_E_WORKSPACE, "body arena %zu < %zu", arena_bytes, B.total);
  if (dweights && !dx) return fail(TFNAS_E_INVALID, "body: weight gradients need dx");
  for (int i = 0; i < B.nb; ++i) {
    const int am = alpha_mode_of(d->op[i], cand_masks[i]);
    if (am && !(T > 0.f)) return fail(TFNAS_E_INVALID, "temperature must be > 0");
    if (!am && B.op[i].P.na != 1) return fail(TFNAS_E_INVALID, "body: MixedOP %d: cand_mask must be all candidates or one-hot", i);
    if (dweights) {
      for (int s = 0; s < B.op[i].P.na; ++s) {
        const int id = B.op[i].P.c[s].id;
        const TfnasCandPtrs& g = dweights[(size_t)i * TFNAS_MAX_OPS + id];
        if (!g.w1 || !g.dw || !g.w3) return fail(TFNAS_E_INVALID, "body: MixedOP %d candidate %d: null weight-grad pointer", i, id);
        if (B.op[i].P.c[s].se > 0 && (!g.se_rw || !g.se_rb || !g.se_ew || !g.se_eb))
          return fail(TFNAS_E_INVALID, "body: MixedOP %d candidate %d: null SE weight-grad pointer", i, id);
      }
    }
  }
  char* A = (char*)arena;
  cudaStream_t st = (cudaStream_t)stream;
  cudaGetLastError();
  static thread_local PreppedBwd PB[TFNAS_MAX_BLOCKS];
  const bool pre = umma_enabled() != 0;
  if (pre) {      // the backward GEMMs' weights (dc: W3^T; dx: W1^T scaled by the forward's BN1 rstd) of all MixedOPs at once
    umma_prep_batch_begin();
    for (int i = (dx ? 0 : 1); i < B.nb; ++i)
      umma_prep_bwd(B.op[i].P, (const float*)(A + B.op[i].saved + B.op[i].L.bn1), (float*)(A + B.op[i].prep_b), PB[i].WD, PB[i].WX,
                    PB[i].CH, st);
    umma_prep_batch_flush(st);
  }
  const float* gstage = dout;          // gradient w.r.t. the current stage's (sink) output
  int gs_idx = 0;
  for (int s = B.ns - 1; s >= 0; --s) {
    const StageSlot& S = B.st[s];
    int stage_alpha = 0;
    for (int j = 0; j < S.K; ++j) stage_alpha |= alpha_mode_of(d->op[S.first + j], cand_masks[S.first + j]);
    SinkRes R;
    for (int j = 0; j < 4; ++j) R.res[j] = j < S.K ? (const float*)(A + B.op[S.first + j].out) : nullptr;
    const float* cum = stage_alpha ? (const float*)(A + S.cumlat) : nullptr;
    float* dcum = (float*)(A + S.dcumlat);
    double* dots = (double*)(A + S.dots);
    float* gcur = (float*)(A + B.gB[0]);
    int gb_idx = 0;
    cudaMemsetAsync(dots, 0, 4 * sizeof(double), st);
    {
      int blocks = (int)min((size_t)(4 * sm_count()), (S.numel / 4 + NT - 1) / NT);
      ProfScope ps("sink_bwd", 4.0 * S.numel * (S.K + 2), 2.0 * S.numel * S.K, st);
      k_sink_bwd_last<<<max(blocks, 1), NT, 0, st>>>(S.K, S.numel, R, betas[s], gstage, gcur, dots);
    }
    {
      ProfScope ps("sink_fin", 64, 0, st);
      k_sink_fin2<<<1, 32, 0, st>>>(S.K, betas[s], cum, stage_alpha ? dlat : nullptr, dots, dbetas ? dbetas[s] : nullptr,
                                    stage_alpha ? dcum : nullptr);
    }
    if (stage_alpha) {
      float* dl[4];
      for (int j = 0; j < 4; ++j) dl[j] = (float*)(A + B.op[S.first + (j < S.K ? j : 0)].dlat);
      count_launch(1);
      k_cumlat_bwd<<<1, 32, 0, st>>>(S.K, dcum, dl[0], dl[1], dl[2], dl[3]);
    }
    for (int j = S.K - 1; j >= 0; --j) {
      const int i = S.first + j;
      const OpSlot& O = B.op[i];
      const int am = alpha_mode_of(d->op[i], cand_masks[i]);
      const float* xin = (j > 0) ? (const float*)(A + B.op[i - 1].out) : (s > 0 ? (const float*)(A + B.st[s - 1].out) : x);
      float* dxo;
      if (j > 0) dxo = (float*)(A + B.gB[gb_idx ^ 1]);
      else if (s > 0) dxo = (float*)(A + B.gS[gs_idx]);
      else dxo = dx;
      BwdScratch Bs;
      bwd_scratch(O.P, dweights != nullptr, A + B.ws, Bs);
      launch_backward(O.P, xin, gcur, am ? (const float*)(A + O.dlat) : nullptr, T, am, A + O.saved, O.L, Bs, dxo,
                      (am && dlog_alphas) ? dlog_alphas[i] : nullptr, dweights ? dweights + (size_t)i * TFNAS_MAX_OPS : nullptr, st, 0,
                      (pre && (dxo || i > 0)) ? &PB[i] : nullptr);
      if (j > 0) {
        // gradient of block j-1's output: dx of block j plus its share beta_{j-1} * gstage of the sink
        int blocks = (int)min((size_t)(4 * sm_count()), (S.numel / 4 + NT - 1) / NT);
        ProfScope ps("sink_axpy", 12.0 * S.numel, 2.0 * S.numel, st);
        k_axpy_beta<<<max(blocks, 1), NT, 0, st>>>(S.numel, betas[s], S.K, j - 1, gstage, dxo);
        gb_idx ^= 1;
        gcur = dxo;
      }
    }
    if (s > 0) {
      gstage = (const float*)(A + B.gS[gs_idx]);
      gs_idx ^= 1;
    }
  }
  return check_cuda("tfnas_body_bwd");
}

}  // extern "C"
