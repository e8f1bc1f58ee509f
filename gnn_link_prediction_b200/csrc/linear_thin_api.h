// Host entry points of the thin-contraction (K <= 8) dense-layer kernels (linear_thin.cu).
// `dtype` (HGIN_DTYPE_*) is the storage type of the WIDE side: z / out / g of the K <= 8 layers, x / dx / post_z of
// the head; the narrow side (x of the thin layers, the single head column) is always fp32.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hgin {
namespace thin {

bool fwd_eligible(const float *x1, int k1, int k2, int n, const void *z, int64_t ldz, const void *out, int64_t ldo, int dtype);
bool bwd_eligible(const void *g, int64_t ldg, const void *z, int64_t ldz, int act, int k1, int k2, int n, int c0,
                  int c1, const void *dx, const float *dot_x, int dtype);
int64_t bwd_workspace_bytes(int n, int k);
int32_t linear_fwd(int64_t rows, const float *x, int64_t ldx, int k, const float *W, const float *bias, int n, int act,
                   const float *alpha, void *z, int64_t ldz, void *out, int64_t ldo, int accumulate_out, int dtype,
                   cudaStream_t s);
int32_t linear_bwd(int64_t rows, const void *g, int64_t ldg, const void *z, int64_t ldz, int act, const float *alpha,
                   const float *x, int64_t ldx, int k, const float *W, int n, int c0, int c1, const float *dot_x,
                   int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace, int dtype,
                   cudaStream_t s);

// n = 1 head (Linear(k, 1), models.py:328)
bool head_fwd_eligible(const void *x1, int64_t ld1, int k1, int k2, int n, int dtype);
bool head_bwd_eligible(const void *x1, int64_t ld1, int k1, int k2, int n, int c0, int c1, const void *dx,
                       int64_t lddx, const void *dot_x, const float *W, int dtype);
int32_t head_fwd(int64_t rows, const void *x, int64_t ldx, int k, const float *W, const float *bias, int act,
                 const float *alpha, float *z, int64_t ldz, float *out, int64_t ldo, int accumulate_out, int dtype,
                 cudaStream_t s);
int32_t head_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act, const float *alpha,
                 const void *x, int64_t ldx, int k, const float *W, void *dx, int64_t lddx, float *dW, float *db,
                 float *dalpha, void *workspace, const void *post_z, int64_t ld_post, int post_act,
                 const float *post_alpha, float *post_dalpha, int dtype, cudaStream_t s);

}  // namespace thin
}  // namespace hgin
