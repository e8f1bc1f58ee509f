// Host entry points of the tensor-core dense-layer path (defined in linear_tc.cu), called by the
// C-ABI functions in linear_simt.cu when math_mode == HGIN_MATH_TF32 and the shapes qualify.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hgin {
namespace tcgemm {

struct TnDebug {   // overrides of the MN-major descriptor fields (hgin_debug_gemm_tn)
    int tma_swizzle;
    int lbo;
    int sbo;
    int layout_type;
    int k_step_bytes;
};

struct PostArgs {  // activation derivative of the layer BELOW, applied to dx on its way out
    const void *z;   // rows of the same storage type as dx
    int64_t ldz;
    int act;
    const float *alpha;
    float *dalpha;  // [1] out: sum dx * min(z, 0), or NULL
    // GIN self branch on the same epilogue (hgin_linear_bwd_post_self): dx leaves as (1 + eps) * (dz W) * act'(z)
    // and ddot = sum (dz W) * act(z); both NULL for the plain post-activation
    const float *self_eps;
    float *ddot;
};

bool fused_bwd_enabled();
void set_fused_bwd(int on);
bool fused_dw_enabled();
void set_fused_dw(int on);
bool fwd_eligible(int64_t rows, const float *x1, int64_t ld1, int k1, int k2, int n, const float *z, int64_t ldz,
                  const float *out, int64_t ldo);
bool bwd_eligible(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act, const float *x1,
                  int64_t ld1, int k1, int k2, int n, int c0, int c1, const float *dx, int64_t lddx,
                  const float *dot_x, int64_t ld_dot);
int64_t fwd_workspace_bytes(int k1, int n);
int64_t bwd_workspace_bytes(int64_t rows, int k1, int k2, int n);
int32_t linear_fwd(int64_t rows, const float *x1, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                   const float *W, const float *bias, int n, int act, const float *alpha, float *z, int64_t ldz,
                   float *out, int64_t ldo, int accumulate_out, void *workspace, cudaStream_t s);
int32_t linear_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act,
                   const float *alpha, const float *x1, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                   const float *W, int n, int c0, int c1, float *dx, int64_t lddx, const float *dot_x,
                   int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace,
                   const TnDebug *dbg, const PostArgs *post, cudaStream_t s);


// ---- bf16 rows (HGIN_DTYPE_BF16; linear_tc_bf16.cu): x1 / z / out / g / dx / dot_x / post->z are bf16, x2, W, bias and
// all reductions fp32 --------------------------------------------------------------------------------------------
bool fwd_eligible_bf16(int64_t rows, const void *x1, int64_t ld1, int k1, int k2, int n, const void *z, int64_t ldz,
                       const void *out, int64_t ldo);
bool bwd_eligible_bf16(int64_t rows, const void *g, int64_t ldg, const void *z, int64_t ldz, int act, const void *x1,
                       int64_t ld1, int k1, int k2, int n, int c0, int c1, const void *dx, int64_t lddx,
                       const void *dot_x, int64_t ld_dot);
int64_t fwd_workspace_bytes_bf16(int k1, int n);
int64_t bwd_workspace_bytes_bf16(int64_t rows, int k1, int k2, int n);
int32_t linear_fwd_bf16(int64_t rows, const void *x1, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                        const float *W, const float *bias, int n, int act, const float *alpha, void *z, int64_t ldz,
                        void *out, int64_t ldo, int accumulate_out, void *workspace, cudaStream_t s);
int32_t linear_bwd_bf16(int64_t rows, const void *g, int64_t ldg, const void *z, int64_t ldz, int act,
                        const float *alpha, const void *x1, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                        const float *W, int n, int c0, int c1, void *dx, int64_t lddx, const void *dot_x,
                        int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace,
                        const TnDebug *dbg, const PostArgs *post, cudaStream_t s);

}  // namespace tcgemm
}  // namespace hgin
