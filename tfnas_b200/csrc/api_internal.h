// Host-side helpers shared by the C-ABI translation units (api.cu: single MixedOP / sink; body.cu: the supernet executor;
// optim.cu: step glue).
#pragma once
#include "kernels.h"

// record an error message for tfnas_last_error() and return `code`
int fail(int code, const char* fmt, ...);
// cudaGetLastError() -> TFNAS_E_CUDA with the message, else TFNAS_OK
int check_cuda(const char* what);
// the float4 / bulk-copy paths of the kernels assume what the header states: 16-byte aligned boundary tensors
static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// descriptor validation + per-call plan (slot order, stacked widths, output geometry); w may be NULL (sizes only)
int build_plan(const TfnasMixedOpDesc* d, uint32_t mask, const TfnasCandPtrs* w, Plan& P);
void saved_layout(const Plan& P, SavedLayout& L);
size_t fwd_scratch(const Plan& P, char* base, FwdScratch& S);
size_t bwd_scratch(const Plan& P, int want_wgrad, char* base, BwdScratch& S);
