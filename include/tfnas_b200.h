/*
 * tfnas_b200 — C ABI of the B200-native TF-NAS supernet search hot path.
 *
 * The reference (AberHu/TF-NAS) has no FFI/plugin seam: its hot path is the Python
 * class API of models/model_search.py, and all arithmetic is delegated to PyTorch.
 * This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md): plain pointers and sizes, no torch types.  Every entry point
 *   - returns 0 on success, a negative TFNAS_E_* code otherwise (message via
 *     tfnas_last_error()), never throws across the ABI;
 *   - never allocates or frees caller memory, never synchronises the device;
 *   - orders all its work on the cudaStream_t it is given (passed as void*): the weight-gradient GEMMs of a backward
 *     call may run on a library-owned side stream that is forked from and joined back into the caller's stream by
 *     events INSIDE the call (TFNAS_SIDE_STREAM=0 keeps everything on the caller's stream).
 * All tensors are fp32, contiguous NCHW, device memory, 16-byte aligned.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   tfnas_mixedop_fwd / _bwd      MixedOP.forward        models/model_search.py:58-91
 *                                 MBInvertedResBlock.forward  models/layers.py:539-561
 *                                 F.gumbel_softmax call  models/model_search.py:87
 *                                 get_lookup_latency dot models/model_search.py:88,90
 *                                 (+ the autograd mirror of all of the above)
 *   tfnas_stage_sink_fwd / _bwd   MixedStage.forward sink sum  models/model_search.py:202-204
 */
#ifndef TFNAS_B200_H
#define TFNAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFNAS_MAX_OPS 8
#define TFNAS_ABI_VERSION 1

enum {
  TFNAS_OK = 0,
  TFNAS_E_INVALID = -1,     /* bad descriptor / argument */
  TFNAS_E_UNSUPPORTED = -2, /* shape outside what the kernels tile */
  TFNAS_E_WORKSPACE = -3,   /* workspace / saved buffer too small */
  TFNAS_E_CUDA = -4         /* a CUDA runtime call failed */
};

enum { TFNAS_ACT_RELU = 0, TFNAS_ACT_SWISH = 1 };

/* Shape of one MixedOP (reference ctor models/model_search.py:33-47). */
typedef struct TfnasMixedOpDesc {
  int32_t N, ic, oc, H, W;
  int32_t stride;             /* 1 or 2 */
  int32_t act;                /* TFNAS_ACT_* */
  int32_t num_ops;            /* <= TFNAS_MAX_OPS */
  int32_t mc[TFNAS_MAX_OPS];  /* mid (expanded) width per candidate: arbitrary >= 1 */
  int32_t k[TFNAS_MAX_OPS];   /* depthwise kernel: 3 or 5 */
  int32_t se[TFNAS_MAX_OPS];  /* squeeze-excite hidden width, 0 = no SE */
} TfnasMixedOpDesc;

/* Device pointers of one candidate's parameters (reference state_dict names in comments). */
typedef struct TfnasCandPtrs {
  float* w1;    /* inverted_bottleneck.conv.weight        [mc, ic]   */
  float* dw;    /* depth_conv.conv.weight                 [mc, k*k]  */
  float* w3;    /* point_linear.conv.weight               [oc, mc]   */
  float* se_rw; /* squeeze_excite.conv_reduce.weight      [se, mc]   (NULL if se==0) */
  float* se_rb; /* squeeze_excite.conv_reduce.bias        [se]       */
  float* se_ew; /* squeeze_excite.conv_expand.weight      [mc, se]   */
  float* se_eb; /* squeeze_excite.conv_expand.bias        [mc]       */
} TfnasCandPtrs;

int tfnas_version(void);
const char* tfnas_last_error(void);

/* Bytes of the buffer kept from forward to backward ("saved": D, Z, statistics, SE state). */
size_t tfnas_mixedop_saved_bytes(const TfnasMixedOpDesc* d, uint32_t cand_mask);
/* Bytes of scratch for one forward or backward call (max of both). */
size_t tfnas_mixedop_workspace_bytes(const TfnasMixedOpDesc* d, uint32_t cand_mask, int want_wgrad);

/*
 * Forward.  cand_mask = 0xFF: alpha mode (MixedOP.forward(sampling=False)):
 *   w = softmax((log_alphas + gumbel)/T), out = sum_i w_i * op_i(x), out_lat = sum_i w_i*lat8[i].
 * cand_mask one-hot: sampled mode (sampling=True): out = op_idx(x); log_alphas/gumbel/lat8/out_lat
 *   may be NULL.
 * weights: host array of num_ops TfnasCandPtrs (entries of inactive candidates are ignored).
 * log_alphas, gumbel, lat8: device arrays of num_ops floats.  out_lat: device scalar.
 */
int tfnas_mixedop_fwd(const TfnasMixedOpDesc* d, uint32_t cand_mask,
                      const float* x, const TfnasCandPtrs* weights,
                      const float* log_alphas, const float* gumbel, const float* lat8, float T,
                      float* out, float* out_lat,
                      void* saved, size_t saved_bytes,
                      void* workspace, size_t ws_bytes, void* stream);

/*
 * Backward.  dout: dL/dout; dlat: device scalar dL/dout_lat (NULL = 0).
 * dx: dL/dx (written); NULL = input needs no gradient (then only dlog_alphas is produced).
 * dlog_alphas: device [num_ops] (written; alpha mode only, else may be NULL).
 * dweights: host array of num_ops TfnasCandPtrs receiving dL/dW of the ACTIVE candidates
 *   (written, not accumulated), or NULL when weight gradients are not wanted (alpha step).
 */
int tfnas_mixedop_bwd(const TfnasMixedOpDesc* d, uint32_t cand_mask,
                      const float* x, const TfnasCandPtrs* weights,
                      const float* dout, const float* dlat, float T,
                      const void* saved, size_t saved_bytes,
                      float* dx, float* dlog_alphas, const TfnasCandPtrs* dweights,
                      void* workspace, size_t ws_bytes, void* stream);

/*
 * MixedStage sink (models/model_search.py:202-204): out = sum_j softmax(betas)_j * res[j],
 * out_lat = sum_j softmax(betas)_j * cumlat[j].  res: host array of K device pointers, each
 * numel floats.  cumlat: device [K] or NULL (then out_lat untouched).
 */
int tfnas_stage_sink_fwd(int K, size_t numel, const float* const* res, const float* betas,
                         const float* cumlat, float* out, float* out_lat, void* stream);
/*
 * Backward of the sink: dres[j] = beta_j * dout (written), dbetas[k] (written),
 * dcumlat[j] = beta_j * dlat (written if non-NULL).  workspace: >= 32 bytes.
 */
int tfnas_stage_sink_bwd(int K, size_t numel, const float* const* res, const float* betas,
                         const float* cumlat, const float* dout, const float* dlat,
                         float* const* dres, float* dbetas, float* dcumlat,
                         void* workspace, size_t ws_bytes, void* stream);

/* Test helper: byte offsets of the regions of the saved buffer, in the order
 * {xmom, bn1, bn2, bn3, mixw, lat, se_p, se_t, se_g, UH, D, Z, total} (13 entries). */
int tfnas_debug_saved_layout(const TfnasMixedOpDesc* d, uint32_t cand_mask, size_t* out13);

/*
 * Batch-statistic BatchNorm (no affine, biased variance, eps 1e-5, no running stats) fused with the activation:
 * the BN + act of the reference's ConvLayer / second stem (models/layers.py:90-110, 469-477 with affine=False as in
 * models/model_search.py:219-220,275).  x, y, dy, dx: [N, C, HW] fp32.  act: TFNAS_ACT_RELU, TFNAS_ACT_SWISH or
 * TFNAS_ACT_NONE.  mean_rstd: device [2*C] (written by fwd, read by bwd).  workspace: >= 16*C bytes.
 */
#define TFNAS_ACT_NONE 2
int tfnas_bn_act_fwd(int N, int C, int HW, int act, const float* x, float* y, float* mean_rstd, void* workspace,
                     size_t ws_bytes, void* stream);
int tfnas_bn_act_bwd(int N, int C, int HW, int act, const float* x, const float* mean_rstd, const float* dy, float* dx,
                     void* workspace, size_t ws_bytes, void* stream);

/*
 * Depthwise KxK convolution (groups = C, padding K/2, no bias) of the second stem: the reference's
 * MBInvertedResBlock.depth_conv when mid_channels == in_channels (models/layers.py:479-489, built by
 * models/model_search.py:220).  K in {3, 5}; stride 1 only.  x, y, dy, dx: [N, C, H, W]; w, dw: [C, K*K].
 * bwd: dx (input gradient) and dw (weight gradient, written not accumulated) may each be NULL.
 */
int tfnas_dwconv_fwd(int N, int C, int H, int W, int K, int stride, const float* x, const float* w, float* y, void* stream);
int tfnas_dwconv_bwd(int N, int C, int H, int W, int K, int stride, const float* x, const float* w, const float* dy,
                     float* dx, float* dw, void* stream);

/* Test helper (host only): pipeline configuration of the persistent tcgen05 GEMM `which` (0 expand, 1 project, 2 dc,
 * 3 dx; tfnas_b200/csrc/umma_ws.cu) for an N chunk of max_nc columns:
 * out5 = {operand stages, weight slots, dynamic shared-memory bytes, producer groups, threads per CTA}.
 * TFNAS_E_UNSUPPORTED: nothing fits, the call would use the per-tile kernel. */
int tfnas_debug_ws_config(int which, int max_nc, uint32_t* out5);

/* Test helper: byte offsets of the backward workspace regions
 * {sG, sGY, sD, sU, cvec2, Mm, dg, DC, DA, total} (10 entries). */
int tfnas_debug_bwd_layout(const TfnasMixedOpDesc* d, uint32_t cand_mask, int want_wgrad, size_t* out10);

/*
 * Per-kernel timing for bench.py's roofline: when enabled, every kernel launch is bracketed by
 * CUDA events on its stream.  tfnas_prof_collect synchronises those events and aggregates by
 * kernel name: total milliseconds, algorithmic bytes and flops (as modelled in DESIGN.md) and
 * launch count.  Returns the number of entries written (<= max_entries) or a negative error.
 * Enabling (or disabling) clears previously recorded launches.
 */
typedef struct TfnasProfEntry {
  char name[32];
  double ms, bytes, flops;
  int64_t launches;
} TfnasProfEntry;
int tfnas_prof_enable(int on);
int tfnas_prof_collect(TfnasProfEntry* out, int max_entries);
/* The same records un-aggregated, in launch order: start / end in ms relative to the first recorded launch. */
typedef struct TfnasProfLaunch {
  char name[32];
  uint64_t stream;
  double start_ms, end_ms;
} TfnasProfLaunch;
int tfnas_prof_timeline(TfnasProfLaunch* out, int max_entries);

/*
 * Test helper for the tcgen05 building blocks: C[n][p] = sum_k B[n][k] * A[k][p] (A: [K][M], B: [N][K],
 * C: [N][M], all device fp32) computed with the 3-term tf32 split on the tensor cores.  N <= 256.
 * wp_scratch: >= ceil(K/32) * 2 * roundup(N,16) * 128 bytes.
 */
int tfnas_umma_selftest(int M, int N, int K, const float* A, const float* B, float* C, float* wp_scratch,
                        size_t wp_bytes, int variant, void* stream);


/* =====================================================================================================================
 * Supernet body: the six MixedStages of Network.forward (reference models/model_search.py:291-296; per stage
 * MixedStage.forward :157-206) executed in ONE call per direction over one caller-provided arena -- the MixedOPs in
 * forward order, each stage closed by its sink-connecting sum (start_res == 1: the sink sums the outputs of all blocks
 * of the stage, as for every stage of the reference Network).  Replaces the per-MixedOP / per-sink calls above when the
 * whole body runs (same kernels, same arithmetic); the launch sequence depends only on (descriptor, masks, pointers), so a
 * call can be captured in a CUDA graph.
 * ===================================================================================================================== */
#define TFNAS_MAX_BLOCKS 32
#define TFNAS_MAX_STAGES 8
typedef struct TfnasBodyDesc {
  int32_t num_stages;                      /* <= TFNAS_MAX_STAGES */
  int32_t num_blocks;                      /* total MixedOPs = sum(stage_blocks) <= TFNAS_MAX_BLOCKS */
  int32_t stage_blocks[TFNAS_MAX_STAGES];  /* MixedOPs per stage, 1..4 */
  TfnasMixedOpDesc op[TFNAS_MAX_BLOCKS];   /* in forward order; op[i+1] input shape == op[i] output shape */
} TfnasBodyDesc;

/* Bytes of the arena for one forward(+backward) pass with these candidate masks (0 on an invalid descriptor).  The
 * arena keeps every MixedOP's output and saved buffer from forward to backward; size it with the want_wgrad the backward
 * will use.  256-byte aligned device memory. */
size_t tfnas_body_arena_bytes(const TfnasBodyDesc* d, const uint32_t* cand_masks, int want_wgrad);

/*
 * Forward.  cand_masks[i]: all candidates (alpha mode) or one-hot (sampled) per MixedOP, as for tfnas_mixedop_fwd.
 * weights: host array [num_blocks][TFNAS_MAX_OPS]; log_alphas: host array of num_blocks device pointers (alpha mode);
 * betas: host array of num_stages device pointers; gumbel, lat: device [num_blocks][TFNAS_MAX_OPS] (alpha mode).
 * out: [N, oc_last, Ho, Wo]; out_lat: device scalar = sum over stages of sum_j softmax(betas)_j * cumlat_j (the caller
 * adds lut['base'], models/model_search.py:282), may be NULL when no MixedOP is in alpha mode.
 */
int tfnas_body_fwd(const TfnasBodyDesc* d, const uint32_t* cand_masks, const float* x, const TfnasCandPtrs* weights,
                   const float* const* log_alphas, const float* const* betas, const float* gumbel, const float* lat,
                   float T, float* out, float* out_lat, void* arena, size_t arena_bytes, void* stream);
/*
 * Backward of the same pass (same descriptor, masks, x, weights, betas, arena).
 * dout: dL/dout; dlat: device scalar dL/dout_lat or NULL.  dx: dL/dx or NULL
 * (then the first MixedOP only produces its d log_alpha).  dlog_alphas: host array of num_blocks device pointers [num_ops]
 * or NULL; dbetas: host array of num_stages device pointers or NULL; dweights: host array [num_blocks][TFNAS_MAX_OPS] of
 * the active candidates' gradient tensors (written, not accumulated) or NULL.
 */
int tfnas_body_bwd(const TfnasBodyDesc* d, const uint32_t* cand_masks, const float* x, const TfnasCandPtrs* weights,
                   const float* const* betas, const float* dout, const float* dlat, float T, float* dx,
                   float* const* dlog_alphas, float* const* dbetas, const TfnasCandPtrs* dweights, void* arena,
                   size_t arena_bytes, void* stream);


/* =====================================================================================================================
 * The two stems of Network.forward (reference models/model_search.py:219-220, :283-284): first_stem = ConvLayer(3, 32,
 * k3, s2, BN, ReLU) (models/layers.py:190-256), second_stem = MBInvertedResBlock(32, 32, se 8, 16, k3, s1, relu) without an
 * expand conv (models/layers.py:479-482).  One call per direction over a caller-provided arena; BN = batch statistics,
 * biased variance, eps 1e-5, no affine.  The image needs no gradient.
 * ===================================================================================================================== */
typedef struct TfnasStemDesc {
  int32_t N, H, W;    /* image batch [N, c_in, H, W] */
  int32_t c_in;       /* 3 */
  int32_t c_mid;      /* 32: first-stem output = second-stem mid width */
  int32_t se;         /* 8: squeeze-excite hidden width of the second stem */
  int32_t c_out;      /* 16: second-stem output channels */
} TfnasStemDesc;
typedef struct TfnasStemPtrs {
  float* conv_w;      /* first_stem.conv.weight                         [c_mid, c_in, 3, 3] */
  float* dw;          /* second_stem.depth_conv.conv.weight             [c_mid, 1, 3, 3]    */
  float* se_rw;       /* second_stem.squeeze_excite.conv_reduce.weight  [se, c_mid]          */
  float* se_rb;       /*                            conv_reduce.bias    [se]                 */
  float* se_ew;       /*                            conv_expand.weight  [c_mid, se]          */
  float* se_eb;       /*                            conv_expand.bias    [c_mid]              */
  float* pw;          /* second_stem.point_linear.conv.weight           [c_out, c_mid]       */
} TfnasStemPtrs;
size_t tfnas_stem_arena_bytes(const TfnasStemDesc* d, int want_wgrad);
/* out: [N, c_out, H/2, W/2] (the input of the first MixedOP).  The arena keeps what the backward needs. */
int tfnas_stem_fwd(const TfnasStemDesc* d, const float* img, const TfnasStemPtrs* w, float* out, void* arena,
                   size_t arena_bytes, void* stream);
/* dw: gradient tensors of the seven parameters (written, not accumulated). */
int tfnas_stem_bwd(const TfnasStemDesc* d, const float* img, const TfnasStemPtrs* w, const float* dout,
                   const TfnasStemPtrs* dw, void* arena, size_t arena_bytes, void* stream);


/* =====================================================================================================================
 * The head of Network.forward (reference models/model_search.py:275-277, :299-302): feature_mix_layer = ConvLayer(320,
 * 1280, k1, BN, Swish), AdaptiveAvgPool2d(1), LinearLayer(1280, num_classes).  BN = batch statistics, no affine.
 * ===================================================================================================================== */
typedef struct TfnasHeadDesc {
  int32_t N, H, W;        /* input [N, c_in, H, W] (the last stage's output) */
  int32_t c_in;           /* 320 */
  int32_t c_mid;          /* 1280 */
  int32_t num_classes;
} TfnasHeadDesc;
typedef struct TfnasHeadPtrs {
  float* fm_w;            /* feature_mix_layer.conv.weight  [c_mid, c_in]        */
  float* fc_w;            /* classifier.linear.weight       [num_classes, c_mid] */
  float* fc_b;            /* classifier.linear.bias         [num_classes]        */
} TfnasHeadPtrs;
size_t tfnas_head_arena_bytes(const TfnasHeadDesc* d);
int tfnas_head_fwd(const TfnasHeadDesc* d, const float* x, const TfnasHeadPtrs* w, float* logits, void* arena,
                   size_t arena_bytes, void* stream);
/* dx: dL/dx (written).  dw: gradients of the three parameters (written) or NULL (alpha step: weights frozen). */
int tfnas_head_bwd(const TfnasHeadDesc* d, const float* x, const TfnasHeadPtrs* w, const float* dlogits, float* dx,
                   const TfnasHeadPtrs* dw, void* arena, size_t arena_bytes, void* stream);

/* =====================================================================================================================
 * Step glue of train_w_arch (reference train_search.py:381-385 and :414-422): global-norm gradient clipping fused with
 * the optimiser update over a table of the LIVE tensors (parameters whose gradient exists this step -- the reference's
 * torch.optim skips tensors with grad None), and the loss of train_search.py:121.
 * ===================================================================================================================== */
typedef struct TfnasSgdTensor {
  float* p;        /* parameter (updated in place) */
  float* g;        /* gradient (scaled in place by grad_scale * clip coefficient, like clip_grad_norm_) */
  float* buf;      /* momentum buffer (zero-initialised by the caller the first time the tensor is live) */
  int64_t numel;
} TfnasSgdTensor;
/*
 * nn.utils.clip_grad_norm_(max_norm) + torch.optim.SGD(momentum, weight_decay, dampening 0, no nesterov):
 *   g *= grad_scale (1/world after a SUM all-reduce); coef = min(1, max_norm / (||g||_2 + 1e-6)) (max_norm <= 0: no clip);
 *   d = coef*g + wd*p; buf = momentum*buf + d; p -= lr*buf.
 * t: host array of n entries.  workspace: >= 16 bytes of device memory.  total_norm_out: device float or NULL.
 */
int tfnas_sgd_step(int n, const TfnasSgdTensor* t, float lr, float momentum, float weight_decay, float max_norm,
                   float grad_scale, float* total_norm_out, void* workspace, size_t ws_bytes, void* stream);

typedef struct TfnasAdamTensor {
  float* p; float* g; float* m; float* v;   /* parameter, gradient, exp_avg, exp_avg_sq (zero-initialised) */
  int32_t numel;                            /* <= 64 (architecture parameters: 8 log_alphas / <= 4 betas per tensor) */
  int32_t renorm;                           /* 1: p = log_softmax(p) after the update (train_search.py:421-422) */
} TfnasAdamTensor;
/*
 * clip_grad_norm_ + torch.optim.Adam(betas, eps, weight_decay; not amsgrad) + the log_softmax renormalisation of every
 * architecture parameter, in one launch.  n <= 64 tensors.  step: 1-based update count (bias correction).
 */
int tfnas_adam_step(int n, const TfnasAdamTensor* t, int step, float lr, float beta1, float beta2, float eps,
                    float weight_decay, float max_norm, float grad_scale, void* stream);

/*
 * nn.CrossEntropyLoss (mean reduction) forward + gradient in one launch: loss = mean_i (logsumexp(logits_i) -
 * logits_i[target_i]); dlogits = (softmax(logits) - onehot(target)) / N.  logits, dlogits: [N, C] fp32; targets: int64.
 */
int tfnas_softmax_ce(int N, int C, const float* logits, const int64_t* targets, float* loss, float* dlogits, void* stream);
/*
 * The same with label smoothing — the derived-network criterion CrossEntropyLabelSmooth (train_eval.py:72-84): the target
 * distribution is (1 - epsilon) onehot + epsilon / C.  epsilon in [0, 1); 0 is tfnas_softmax_ce.
 */
int tfnas_softmax_ce_smooth(int N, int C, const float* logits, const int64_t* targets, float epsilon, float* loss,
                            float* dlogits, void* stream);

/* Weight-gradient GEMMs of the backward calls on a library-owned side stream (forked from / joined into the caller's stream
 * inside the call): 1 = on (default; also TFNAS_SIDE_STREAM=0/1 at first use), 0 = everything on the caller's stream (e.g.
 * for per-kernel timing without a concurrent neighbour).  Returns the previous setting (-1: was still undecided). */
int tfnas_config_side_stream(int on);

/* Number of kernel launches issued through this library since load (bench "gpu_launches"). */
uint64_t tfnas_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TFNAS_B200_H */
