// Host entry points of the thin-contraction (K <= 8) dense-layer kernels (linear_thin.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hgin {
namespace thin {

bool fwd_eligible(const float *x1, int k1, int k2, int n, const float *z, int64_t ldz, const float *out, int64_t ldo);
bool bwd_eligible(const float *g, int64_t ldg, const float *z, int64_t ldz, int act, int k1, int k2, int n, int c0,
                  int c1, const float *dx, const float *dot_x);
int64_t bwd_workspace_bytes(int n, int k);
int32_t linear_fwd(int64_t rows, const float *x, int64_t ldx, int k, const float *W, const float *bias, int n, int act,
                   const float *alpha, float *z, int64_t ldz, float *out, int64_t ldo, int accumulate_out,
                   cudaStream_t s);
int32_t linear_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act, const float *alpha,
                   const float *x, int64_t ldx, int k, const float *W, int n, int c0, int c1, const float *dot_x,
                   int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace, cudaStream_t s);

// n = 1 head (Linear(k, 1), models.py:328)
bool head_fwd_eligible(const float *x1, int64_t ld1, int k1, int k2, int n);
bool head_bwd_eligible(const float *x1, int64_t ld1, int k1, int k2, int n, int c0, int c1, const float *dx,
                       int64_t lddx, const float *dot_x, const float *W);
int32_t head_fwd(int64_t rows, const float *x, int64_t ldx, int k, const float *W, const float *bias, int act,
                 const float *alpha, float *z, int64_t ldz, float *out, int64_t ldo, int accumulate_out, cudaStream_t s);
int32_t head_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act, const float *alpha,
                 const float *x, int64_t ldx, int k, const float *W, float *dx, int64_t lddx, float *dW, float *db,
                 float *dalpha, void *workspace, const float *post_z, int64_t ld_post, int post_act,
                 const float *post_alpha, float *post_dalpha, cudaStream_t s);

}  // namespace thin
}  // namespace hgin
