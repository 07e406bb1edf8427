// Host-side launch wrappers (implemented in fwd.cu / bwd.cu / stage.cu), called from api.cu.
#pragma once
#include "common.cuh"

// Layout of the buffer kept from forward to backward.  All offsets in BYTES from the base.
struct SavedLayout {
  size_t xmom;   // double [ic + ic*ic] : mean_x, cov_x (biased)
  size_t bn1;    // float  [2*MC]       : mu1, r1
  size_t bn2;    // float  [2*MC]       : mu2, r2
  size_t bn3;    // float  [2*na*oc]    : mu3, r3
  size_t mixw;   // float  [8]          : mixing weights by ORIGINAL candidate id (1.0 in sampled mode)
  size_t lat;    // float  [8]          : LUT latencies by original id
  size_t sep;    // float  [N*MCse]     : SE pooled input p
  size_t set;    // float  [N*SEH]      : SE hidden pre-activation t
  size_t seg;    // float  [N*MCse]     : SE gate g
  size_t UH;     // float  [N*MC*HW]    : BN1-normalised expand output (pre-activation)
  size_t D;      // float  [N*MC*HWo]   : depthwise output (pre-BN2)
  size_t Z;      // float  [N*na*oc*HWo]: project output (pre-BN3)
  size_t total;
};

struct FwdScratch {   // all device pointers into the workspace
  double* xsum;   // [ic]
  double* xcov;   // [ic*ic]
  double* st2;    // [2*MC]
  double* st3;    // [2*na*oc]
  float* coef;    // [na*oc + oc]
  float* umprep;  // pre-split / pre-swizzled weights for the tcgen05 GEMMs (expand, project)
  float* xpart;   // [XM_MAXCTA][upper-triangular 4x4 blocks of the (ic+1)^2 Gram matrix][16]: per-CTA partial input moments
};
#define XM_MAXCTA 320

struct BwdScratch {
  float* DC;      // [N*MC*HWo]  dL/dc, then dL/dd-hat in place
  float* DA;      // [N*MC*HW]   dL/da (output of the transposed depthwise)
  double* sG;     // [oc]
  double* sGY;    // [na*oc]
  float* dg;      // [N*MCse]    SE: sum_hw dc*b, then dp in place
  double* sD;     // [2*MC]
  double* sU;     // [2*MC]
  float4* dzc;    // [na*oc]  per (slot, out channel) BN3-backward coefficients {A, B, C}: dz = A*(g - B - yhat*C)
  float4* dzc2;   // [na*oc]  the same folded into dz = a*g + b*z + c (tcgen05 dc prologue: two FMAs per element)
  float* cvec2;   // [ic]        (cvec2 and Mm adjacent: zeroed by one memset)
  float* Mm;      // [ic*ic]
  float* a12;     // [2*MC]  (unused since b4mm forms the BN1-backward coefficients itself)
  size_t zero_bytes;   // bytes from sG to the end of dg: the atomically accumulated region, zeroed once per call
  float* dmix;    // [8]  dL/dw_i (data term)
  float* sede;    // [N*MCse]  SE: dL/de (pre-sigmoid)     (weight-grad mode)
  float* sedt;    // [N*SEH]   SE: dL/dt (pre-act hidden)  (weight-grad mode)
  float* Smat;    // [MC*ic]   sum_p du-hat x^T            (weight-grad mode)
  float* umprep;  // pre-split / pre-swizzled weights for the tcgen05 GEMMs (dc, dx)
};

int sm_count();

// tcgen05 GEMM path (umma_pw.cu); TFNAS_GEMM=simt selects the CUDA-core kernels instead (A/B debugging only)
int umma_enabled();
size_t umma_fwd_prep_bytes(const Plan& P);
size_t umma_bwd_prep_bytes(const Plan& P);
// One GEMM's weight-side geometry for one candidate slot
struct UmW {
  const float* wp;   // prepped weights: [nN][nK][2][Nc*128 B]
  int Nout;          // true output channels
  int Nc;            // channels per N chunk (multiple of 16, <= 256)
  int nN;            // number of N chunks
  int nK;            // number of K chunks (of 32)
};
struct UmWAll { UmW s[TFNAS_MAX_OPS]; };

struct DxChunks {            // chunk c of the stacked K axis -> (slot, first local channel)
  int total;
  int first[TFNAS_MAX_OPS + 1];   // first chunk index of each slot
};

void umma_prep_fwd(const Plan& P, float* prep_buf, UmWAll& WE, UmWAll& WP, cudaStream_t st);
// deferred prep (see umma_pw.cu): queue the jobs of several MixedOPs, convert them in a few launches
void umma_prep_batch_begin();
void umma_prep_batch_flush(cudaStream_t st);
// geometry of weights prepped ahead of the call (body executor); nullptr = the call preps its own into its workspace
struct PreppedFwd { UmWAll WE, WP; };
struct PreppedBwd { UmWAll WD; UmW WX; DxChunks CH; };
void umma_prep_bwd(const Plan& P, const float* bn1, float* prep_buf, UmWAll& WD, UmW& WX, DxChunks& CH, cudaStream_t st);
void umma_expand(const Plan& P, const UmWAll& WA, const float* x, const float* bn1, float* UH, cudaStream_t st);
void umma_project(const Plan& P, const UmWAll& WA, const float* D, const float* bn2, const float* seg, float* Zb,
                  double* st3, cudaStream_t st);
void umma_dc(const Plan& P, const UmWAll& WA, const float* G, const float* Zb, const float4* dzc2,
             const float* D, const float* bn2, float* DC, float* dg, double* sD, cudaStream_t st);
// MODE 0: dW3[o][c] += sum_p dz c~ ; MODE 1: SmatT[k][c] += sum_p du-hat x   (out must be zeroed by the caller)
void umma_wgrad(const Plan& P, int slot, int mode, const float* A0, const float* A1, const float* B0, const float* B1,
                const float* bn2, const float* seg, const float* bn3, const float4* dzc, float* out, cudaStream_t st);
// input covariance on the tensor cores: acc = [sum x (ic, input) | centred second moments (ic*ic, accumulated)] doubles
void umma_covariance(const Plan& P, const float* x, double* acc, cudaStream_t st);
void umma_dx(const Plan& P, const UmW& W, const DxChunks& CH, const float* DA, const float* UH, const float* bn1,
             float* dx, double* sU, cudaStream_t st);

// persistent warp-specialised versions (umma_ws.cu); return false when the shape does not fit (caller falls back).
// which: 0 expand, 1 project, 2 dc, 3 dx  (TFNAS_WS=comma list selects a subset, "none" disables)
int ws_enabled(int which);
bool ws_expand(const Plan& P, const UmWAll& WA, const float* x, const float* bn1, float* UH, cudaStream_t st);
bool ws_project(const Plan& P, const UmWAll& WA, const float* D, const float* bn2, const float* seg, float* Zb, double* st3,
                cudaStream_t st);
bool ws_dc(const Plan& P, const UmWAll& WA, const float* G, const float* Zb, const float4* dzc2, const float* D,
           const float* bn2, float* DC, float* dg, double* sD, cudaStream_t st);
bool ws_dx(const Plan& P, const UmW& W, const DxChunks& CH, const float* DA, const float* UH, float* dx, double* sU,
           cudaStream_t st);

// register sliding-window depthwise kernels (dws.cu): stride-1 MixedOPs; TFNAS_DW=tile forces the smem-tile kernels
bool dws_supported(const Plan& P);
void launch_dws_fwd(const Plan& P, const float* UH, float* D, double* st2, cudaStream_t st);
void launch_dws_bwd(const Plan& P, const float* DC, const float* D, const float* bn2, const double* sD, const float* UH,
                    float* DA, const TfnasCandPtrs* dweights, cudaStream_t st, cudaStream_t wg_st);

void launch_forward(const Plan& P, const float* x, const float* log_alphas, const float* gumbel,
                    const float* lat8, float T, int alpha_mode, float* out, float* out_lat,
                    char* saved, const SavedLayout& L, const FwdScratch& S, cudaStream_t st, const PreppedFwd* pre = nullptr);

void launch_forward_tail(const Plan& P, const UmWAll* WPp, const float* x, const float* log_alphas, const float* gumbel,
                         const float* lat8, float T, int alpha_mode, float* out, float* out_lat, char* saved,
                         const SavedLayout& L, const FwdScratch& S, cudaStream_t st);
// project weights only (the second stem has no expand conv)
void umma_prep_project(const Plan& P, float* prep_buf, UmWAll& WP, cudaStream_t st);
// dc weights only
void umma_prep_dc(const Plan& P, float* prep_buf, UmWAll& WD, cudaStream_t st);

void launch_backward(const Plan& P, const float* x, const float* dout, const float* dlat, float T,
                     int alpha_mode, const char* saved, const SavedLayout& L, const BwdScratch& S,
                     float* dx, float* dlog_alphas, const TfnasCandPtrs* dweights, cudaStream_t st, int stop_at_da = 0,
                     const PreppedBwd* pre = nullptr);

void launch_sink_fwd(int K, size_t numel, const float* const* res, const float* betas,
                     const float* cumlat, float* out, float* out_lat, cudaStream_t st);
void launch_sink_bwd(int K, size_t numel, const float* const* res, const float* betas,
                     const float* cumlat, const float* dout, const float* dlat, float* const* dres,
                     float* dbetas, float* dcumlat, double* ws, cudaStream_t st);

void launch_bn_act_fwd(int N, int C, int HW, int act, const float* x, float* y, float* mr, double* ws, cudaStream_t st);
void launch_bn_act_bwd(int N, int C, int HW, int act, const float* x, const float* mr, const float* dy, float* dx,
                       double* ws, cudaStream_t st);

void launch_dwconv_fwd(int N, int C, int H, int W, int KS, const float* x, const float* w, float* y, cudaStream_t st);
void launch_dwconv_bwd(int N, int C, int H, int W, int KS, const float* x, const float* w, const float* dy, float* dx,
                       float* dw, cudaStream_t st);

void count_launch(int n);

// Raise a kernel's dynamic shared-memory limit only when a launch needs more than any launch before it did
// (cudaFuncSetAttribute costs about a microsecond of host time per call).  Keyed by the kernel's address.
void ensure_smem_impl(const void* kern, size_t smem);
template <class K>
static inline void ensure_smem(K kern, size_t smem) { ensure_smem_impl((const void*)kern, smem); }

// RAII launch marker: counts the launch and, when profiling is enabled (tfnas_prof_enable), brackets
// it with CUDA events on the launching stream and records its algorithmic bytes / flops.
struct ProfScope {
  ProfScope(const char* name, double bytes, double flops, cudaStream_t st);
  ~ProfScope();
  void* rec_;
  cudaStream_t st_;
};
