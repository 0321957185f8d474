// Epilogue flags and side-argument structs shared by the GEMM translation units (gemm_tcgen05.cu) and
// the C ABI (af3_abi.cu); the flag values are the AF3_EPI_* constants of include/af3b200.h.
#pragma once
#include "common.h"

namespace af3 {

enum : int { EPI_BIAS = 1, EPI_GELU = 2, EPI_RESID = 4, EPI_SWIGLU = 8, EPI_F32OUT = 16, EPI_ROPE = 32,
              EPI_SWIGLU_CONCAT = 64 /* host-side only: weight rows are [gate; up], not interleaved (stripped before dispatch) */ };

// rotary embedding + KV-cache append in the epilogue of the few-token fused q/k/v projection
struct RopeEpilogue {
    const float* cs;   // [n_tok][64] (cos, sin) pairs, bf16-rounded
    bf16* k_cache;
    bf16* v_cache;
    const int* pos;    // device int: cache slot of this step
    int H, Hkv, Tmax;
};

// RMSNorm fusion across few-token GEMMs (see GemmArgs in gemm_tcgen05.cu): consumer side (norm_w ...) and / or producer side (sumsq_out ...)
struct NormFusion {
    const bf16* norm_w;
    const float* norm_part;
    int norm_parts, norm_ld;
    float norm_eps;
    float* sumsq_out;
    int sumsq_ld;
};

int gemm_bf16(cudaStream_t stream, const bf16* x, int ldx, const bf16* w, int ldw, void* out, int ldo, int n_tok,
              int n_feat, int K, int flags, const bf16* bias, const bf16* resid, int ld_res, int res_period,
              void* workspace, size_t workspace_bytes, const RopeEpilogue* rope, const NormFusion* nf);
size_t gemm_workspace_bytes();

}  // namespace af3
