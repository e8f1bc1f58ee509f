// K2 / K3 — dense layer forward and backward, fp32 SIMT path (HGIN_MATH_FP32).
//
// Forward  (models.py:217, 236-239; readout models.py:366-374):   z = [x1|x2] W^T + b, out (+)= act(z)
// Backward (autograd, train.py:43):   dz = g * act'(z);  dx = dz W;  dW = dz^T [x1|x2];  db = sum dz;
//                                     dalpha = sum g*min(z,0);  d(eps) = sum dx_self * x_dst.
// All reductions over rows are two-stage and deterministic: every CTA writes its partial to the
// caller's workspace and a second kernel adds the partials in CTA order.
#include <string.h>

#include "gemm_simt.cuh"
#include "linear_tc_api.h"
#include "linear_thin_api.h"

namespace hgin {
namespace {

using simt::BK;
using simt::THREADS;

// ---- operand accessors ------------------------------------------------------------------------
struct ConcatRows {  // element (m, k) of [x1 | x2]
    const float *x1; int64_t ld1; int k1;
    const float *x2; int64_t ld2; int k2;
    int64_t rows;
    __device__ __forceinline__ float operator()(int64_t m, int64_t k) const {
        if (m >= rows) return 0.0f;
        if (k < k1) return __ldg(x1 + m * ld1 + k);
        if (k < k1 + k2) return __ldg(x2 + m * ld2 + (k - k1));
        return 0.0f;
    }
};

struct WeightRows {  // element (n, k) of W[n][k]
    const float *W; int n, k;
    __device__ __forceinline__ float operator()(int64_t j, int64_t kk) const {
        return (j < n && kk < k) ? __ldg(W + j * k + kk) : 0.0f;
    }
};

struct WeightCols {  // element (c, nn) = W[nn][c0 + c]: output index contiguous, contraction over rows of W
    const float *W; int n, k, c0, c1;
    __device__ __forceinline__ float operator()(int64_t c, int64_t nn) const {
        return (c0 + c < c1 && nn < n) ? __ldg(W + nn * k + c0 + c) : 0.0f;
    }
};

struct GradZ {  // element (m, nn) of dz = g * act'(z)
    const float *g; int64_t ldg;
    const float *z; int64_t ldz;
    int64_t rows; int n; int act; float alpha;
    __device__ __forceinline__ float operator()(int64_t m, int64_t nn) const {
        if (m >= rows || nn >= n) return 0.0f;
        const float gv = __ldg(g + m * ldg + nn);
        if (act == HGIN_ACT_NONE) return gv;
        return act_backward(gv, __ldg(z + m * ldz + nn), act, alpha);
    }
};

// element (nn, m) of dz^T.  Side sums ride on the loads: each thread fetches a FIXED nn (the tile
// extent divides the thread count), so it can accumulate db[nn] = sum_m dz locally; dalpha likewise.
struct GradZT {
    GradZ dz;
    float *dalpha_acc;  // thread-local
    float *db_acc;      // thread-local
    bool want_alpha;
    bool want_db;
    __device__ __forceinline__ float operator()(int64_t nn, int64_t m) const {
        if (m >= dz.rows || nn >= dz.n) return 0.0f;
        float v = __ldg(dz.g + m * dz.ldg + nn);
        if (dz.act != HGIN_ACT_NONE) {
            const float zv = __ldg(dz.z + m * dz.ldz + nn);
            if (want_alpha && !(zv > 0.0f)) *dalpha_acc += v * zv;   // at::prelu_backward: x > 0 ? 0 : x*g
            v = act_backward(v, zv, dz.act, dz.alpha);
        }
        if (want_db) *db_acc += v;
        return v;
    }
};

struct ConcatColsT {  // element (k, m) of [x1 | x2]^T
    ConcatRows x;
    __device__ __forceinline__ float operator()(int64_t k, int64_t m) const { return x(m, k); }
};

// ---- forward ----------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(THREADS)
linear_fwd_kernel(ConcatRows fa, WeightRows fb, const float *__restrict__ bias, int act,
                  const float *__restrict__ alpha_ptr, float *__restrict__ z, int64_t ldz,
                  float *__restrict__ out, int64_t ldo, int accumulate_out, int vec_ok) {
    __shared__ __align__(16) float smem[T::SMEM_FLOATS];
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * T::BM;
    const int n0 = blockIdx.y * T::BN;
    float acc[T::TM][T::TN];
#pragma unroll
    for (int i = 0; i < T::TM; ++i)
#pragma unroll
        for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.0f;
    simt::mainloop<T, true, true>(acc, fa, fb, m0, n0, 0, fb.k, smem);

    const float alpha = (act == HGIN_ACT_PRELU) ? __ldg(alpha_ptr) : 0.0f;
    const int tx = threadIdx.x % T::TX, ty = threadIdx.x / T::TX;
#pragma unroll
    for (int i = 0; i < T::TM; ++i) {
        const int64_t m = m0 + T::row_of(ty, i);
        if (m >= fa.rows) continue;
#pragma unroll
        for (int gj = 0; gj < T::GN; ++gj) {
            const int nb = n0 + T::col_of(tx, gj * T::VN);
            float zv[T::VN], ov[T::VN];
#pragma unroll
            for (int j = 0; j < T::VN; ++j) {
                const int n = nb + j;
                zv[j] = acc[i][gj * T::VN + j] + ((bias && n < fb.n) ? __ldg(bias + n) : 0.0f);
                ov[j] = act_forward(zv[j], act, alpha);
            }
            if (T::VN == 4 && vec_ok && nb + 3 < fb.n) {
                if (z) *reinterpret_cast<float4 *>(z + m * ldz + nb) = make_float4(zv[0], zv[1], zv[2], zv[3]);
                if (out) {
                    float4 *po = reinterpret_cast<float4 *>(out + m * ldo + nb);
                    if (accumulate_out) {
                        const float4 old = *po;
                        ov[0] += old.x; ov[1] += old.y; ov[2] += old.z; ov[3] += old.w;
                    }
                    *po = make_float4(ov[0], ov[1], ov[2], ov[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < T::VN; ++j) {
                    const int n = nb + j;
                    if (n >= fb.n) continue;
                    if (z) z[m * ldz + n] = zv[j];
                    if (out) out[m * ldo + n] = accumulate_out ? out[m * ldo + n] + ov[j] : ov[j];
                }
            }
        }
    }
}

// ---- input gradient: dx[:, c0:c1] = (dz W)[:, c0:c1], optional dot with dot_x ------------------
template <class T>
__global__ void __launch_bounds__(THREADS)
linear_bwd_dx_kernel(GradZ fa, const float *__restrict__ alpha_ptr, WeightCols fb, float *__restrict__ dx,
                     int64_t lddx, const float *__restrict__ dot_x, int64_t ld_dot,
                     float *__restrict__ dot_partials, int vec_ok) {
    __shared__ __align__(16) float smem[T::SMEM_FLOATS];
    __shared__ float red[32];
    fa.alpha = (fa.act == HGIN_ACT_PRELU) ? __ldg(alpha_ptr) : 0.0f;  // slope lives in device memory
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * T::BM;
    const int n0 = blockIdx.y * T::BN;  // offset inside [c0, c1)
    const int width = fb.c1 - fb.c0;
    float acc[T::TM][T::TN];
#pragma unroll
    for (int i = 0; i < T::TM; ++i)
#pragma unroll
        for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.0f;
    simt::mainloop<T, true, false>(acc, fa, fb, m0, n0, 0, fb.n, smem);

    const int tx = threadIdx.x % T::TX, ty = threadIdx.x / T::TX;
    float dot = 0.0f;
#pragma unroll
    for (int i = 0; i < T::TM; ++i) {
        const int64_t m = m0 + T::row_of(ty, i);
        if (m >= fa.rows) continue;
#pragma unroll
        for (int gj = 0; gj < T::GN; ++gj) {
            const int cb = n0 + T::col_of(tx, gj * T::VN);
            if (dot_x) {
#pragma unroll
                for (int j = 0; j < T::VN; ++j)
                    if (cb + j < width) dot += acc[i][gj * T::VN + j] * __ldg(dot_x + m * ld_dot + cb + j);
            }
            if (!dx) continue;
            if (T::VN == 4 && vec_ok && cb + 3 < width) {
                *reinterpret_cast<float4 *>(dx + m * lddx + cb) =
                    make_float4(acc[i][gj * 4 + 0], acc[i][gj * 4 + 1], acc[i][gj * 4 + 2], acc[i][gj * 4 + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < T::VN; ++j)
                    if (cb + j < width) dx[m * lddx + cb + j] = acc[i][gj * T::VN + j];
            }
        }
    }
    if (dot_partials) {
        dot = block_sum(dot, red);
        if (threadIdx.x == 0) dot_partials[blockIdx.y * gridDim.x + blockIdx.x] = dot;
    }
}

// ---- weight gradient: partial[cta][n][k] over this CTA's row range -------------------------------
// SWAP = false: tile rows index n (dz columns), tile columns index k; SWAP = true: the roles are
// exchanged so that a tiny n sits on the narrow BN side of the tile.  db and dalpha are summed on
// the side by the dz accessor (first tile along the x side only, so every dz element counts once).
template <class T, bool SWAP>
__global__ void __launch_bounds__(THREADS)
linear_bwd_dw_kernel(GradZ dzf, const float *__restrict__ alpha_ptr, ConcatRows xf, int want_alpha, int want_db,
                     int64_t rows_per_cta, float *__restrict__ partials /* [gridDim.x][n][k] */,
                     float *__restrict__ db_partials /* [gridDim.x][n] */, float *__restrict__ alpha_partials) {
    __shared__ __align__(16) float smem[T::SMEM_FLOATS];
    __shared__ float red[32];
    __shared__ float dbs[THREADS];
    dzf.alpha = (dzf.act == HGIN_ACT_PRELU) ? __ldg(alpha_ptr) : 0.0f;
    const int k = xf.k1 + xf.k2;
    const int r0 = blockIdx.y * T::BM;   // tile-row offset (n if !SWAP, k if SWAP)
    const int q0 = blockIdx.z * T::BN;   // tile-col offset (k if !SWAP, n if SWAP)
    const int64_t mbeg = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
    const int64_t mend = min(mbeg + rows_per_cta, dzf.rows);
    float dalpha = 0.0f, db_local = 0.0f;
    const bool first_x_tile = SWAP ? blockIdx.y == 0 : blockIdx.z == 0;
    GradZT fz{dzf, &dalpha, &db_local, want_alpha && first_x_tile, want_db && first_x_tile};
    ConcatColsT fx{xf};
    float acc[T::TM][T::TN];
#pragma unroll
    for (int i = 0; i < T::TM; ++i)
#pragma unroll
        for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.0f;
    fz.dz.rows = mend;   // accessors bound rows by .rows: restrict to this CTA's range
    fx.x.rows = mend;
    if (mbeg < mend) {
        if (SWAP) simt::mainloop<T, false, false>(acc, fx, fz, r0, q0, mbeg, mend, smem);
        else simt::mainloop<T, false, false>(acc, fz, fx, r0, q0, mbeg, mend, smem);
    }
    const int tx = threadIdx.x % T::TX, ty = threadIdx.x / T::TX;
    float *dst = partials + static_cast<int64_t>(blockIdx.x) * dzf.n * k;
#pragma unroll
    for (int i = 0; i < T::TM; ++i) {
#pragma unroll
        for (int j = 0; j < T::TN; ++j) {
            const int rr = r0 + T::row_of(ty, i), qq = q0 + T::col_of(tx, j);
            const int n = SWAP ? qq : rr, kk = SWAP ? rr : qq;
            if (n < dzf.n && kk < k) dst[static_cast<int64_t>(n) * k + kk] = acc[i][j];
        }
    }
    if (want_db && first_x_tile) {
        // threads t, t + EXT, t + 2 EXT, ... fetched the same dz column: add them in a fixed order
        constexpr int EXT = SWAP ? T::BN : T::BM;
        dbs[threadIdx.x] = db_local;
        __syncthreads();
        if (threadIdx.x < EXT) {
            float s = 0.0f;
            for (int t = threadIdx.x; t < THREADS; t += EXT) s += dbs[t];
            const int n = (SWAP ? q0 : r0) + threadIdx.x;
            if (n < dzf.n) db_partials[static_cast<int64_t>(blockIdx.x) * dzf.n + n] = s;
        }
    }
    if (alpha_partials && want_alpha && first_x_tile) {
        dalpha = block_sum(dalpha, red);
        if (threadIdx.x == 0) alpha_partials[(SWAP ? blockIdx.z : blockIdx.y) * gridDim.x + blockIdx.x] = dalpha;
    }
}

// out[i] = sum_p partials[p][i] for i < count, fixed association (four chains).
__global__ void __launch_bounds__(256)
reduce_partials_simt_kernel(const float *__restrict__ partials, int num_partials, int64_t count,
                            float *__restrict__ out) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        int p = 0;
        for (; p + 3 < num_partials; p += 4) {
            s0 += partials[static_cast<int64_t>(p) * count + i];
            s1 += partials[static_cast<int64_t>(p + 1) * count + i];
            s2 += partials[static_cast<int64_t>(p + 2) * count + i];
            s3 += partials[static_cast<int64_t>(p + 3) * count + i];
        }
        for (; p < num_partials; ++p) s0 += partials[static_cast<int64_t>(p) * count + i];
        out[i] = (s0 + s1) + (s2 + s3);
    }
}

// Deterministic sum of `count` floats into out[0] (single CTA, fixed tree).
__global__ void __launch_bounds__(1024) reduce_scalar_kernel(const float *__restrict__ v, int64_t count,
                                                             float *__restrict__ out) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) s += v[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

using T128 = simt::Tile<128, 128, 8, 8>;
using T64 = simt::Tile<128, 64, 8, 4>;
using T32 = simt::Tile<128, 32, 4, 4>;
using T16 = simt::Tile<128, 16, 4, 2>;
using T8 = simt::Tile<128, 8, 4, 1>;

constexpr int kDwSplit = kNumSMs * 2;  // row ranges of the weight-gradient pass

inline int64_t dw_splits(int64_t rows) {
    int64_t s = ceil_div(rows, 128);
    if (s < 1) s = 1;
    return s < kDwSplit ? s : kDwSplit;
}

}  // namespace
}  // namespace hgin

extern "C" int64_t hgin_linear_fwd_workspace_bytes(int64_t rows, int32_t k, int32_t n, int32_t math_mode) {
    if (rows < 0 || k <= 0 || n <= 0) return -1;
    // (fp32 rows under HGIN_MATH_BF16 run the tf32 kernels; bf16 rows need half of that)
    return math_mode != HGIN_MATH_FP32 ? hgin::tcgemm::fwd_workspace_bytes(k, n) : 0;
}

extern "C" int32_t hgin_linear_fwd(int64_t rows, const float *x1, int64_t ld1, int32_t k1, const float *x2,
                                   int64_t ld2, int32_t k2, const float *W, const float *bias, int32_t n,
                                   int32_t act, const float *alpha, float *z, int64_t ldz, float *out, int64_t ldo,
                                   int32_t accumulate_out, void *workspace, int64_t workspace_bytes,
                                   int32_t math_mode, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(rows >= 0 && k1 > 0 && k2 >= 0 && n > 0, "hgin_linear_fwd: bad sizes rows=%lld k1=%d k2=%d n=%d",
                   (long long)rows, k1, k2, n);
    HGIN_CHECK_ARG(act >= HGIN_ACT_NONE && act <= HGIN_ACT_RELU, "hgin_linear_fwd: bad act %d", act);
    HGIN_CHECK_ARG(act != HGIN_ACT_PRELU || alpha, "hgin_linear_fwd: PReLU needs alpha");
    HGIN_CHECK_ARG(math_mode >= HGIN_MATH_FP32 && math_mode <= HGIN_MATH_BF16, "hgin_linear_fwd: bad math_mode %d", math_mode);
    if (rows == 0) return HGIN_OK;
    HGIN_CHECK_ARG(x1 && W && (k2 == 0 || x2) && (z || out), "hgin_linear_fwd: null pointer");
    HGIN_CHECK_ARG(ld1 >= k1 && (k2 == 0 || ld2 >= k2) && (!z || ldz >= n) && (!out || ldo >= n),
                   "hgin_linear_fwd: leading dimension too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (thin::head_fwd_eligible(x1, ld1, k1, k2, n, HGIN_DTYPE_F32) && aligned16(W))
        return thin::head_fwd(rows, x1, ld1, k1, W, bias, act, alpha, z, ldz, out, ldo, accumulate_out, HGIN_DTYPE_F32, s);
    if (thin::fwd_eligible(x1, k1, k2, n, z, ldz, out, ldo, HGIN_DTYPE_F32))
        return thin::linear_fwd(rows, x1, ld1, k1, W, bias, n, act, alpha, z, ldz, out, ldo, accumulate_out, HGIN_DTYPE_F32, s);
    if (math_mode != HGIN_MATH_FP32 && tcgemm::fwd_eligible(rows, x1, ld1, k1, k2, n, z, ldz, out, ldo)) {
        if (!workspace || workspace_bytes < tcgemm::fwd_workspace_bytes(k1 + k2, n))
            return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_linear_fwd: workspace too small for the tf32 path");
        return tcgemm::linear_fwd(rows, x1, ld1, k1, x2, ld2, k2, W, bias, n, act, alpha, z, ldz, out, ldo,
                                  accumulate_out, workspace, s);
    }
    ConcatRows fa{x1, ld1, k1, x2, ld2, k2, rows};
    WeightRows fb{W, n, k1 + k2};
    const int vec_ok = (!z || (ldz % 4 == 0 && aligned16(z))) && (!out || (ldo % 4 == 0 && aligned16(out)));
    const unsigned gx = static_cast<unsigned>(ceil_div(rows, 128));
#define HGIN_FWD(T)                                                                                             \
    linear_fwd_kernel<T><<<dim3(gx, static_cast<unsigned>(ceil_div(n, T::BN))), THREADS, 0, s>>>(               \
        fa, fb, bias, act, alpha, z, ldz, out, ldo, accumulate_out, vec_ok)
    if (n > 64) HGIN_FWD(T128);
    else if (n > 32) HGIN_FWD(T64);
    else if (n > 16) HGIN_FWD(T32);
    else if (n > 8) HGIN_FWD(T16);
    else HGIN_FWD(T8);
#undef HGIN_FWD
    HGIN_CHECK_LAUNCH("hgin_linear_fwd");
    return HGIN_OK;
}

static int64_t simt_bwd_workspace_bytes(int64_t rows, int32_t k, int32_t n) {
    using namespace hgin;
    const int64_t dw = dw_splits(rows) * n * (k + 1) * 4;
    const int64_t scal = (ceil_div(rows, 128) * ceil_div(k > n ? k : n, 8) + dw_splits(rows) * (ceil_div(n, 8) + 1)) * 4;
    const int64_t simt = align_up(dw, 256) + align_up(scal, 256) + 512;
    const int64_t th = thin::bwd_workspace_bytes(n, k);
    return simt > th ? simt : th;
}

extern "C" int64_t hgin_linear_bwd_workspace_bytes(int64_t rows, int32_t k, int32_t n, int32_t math_mode) {
    if (rows < 0 || k <= 0 || n <= 0) return -1;
    const int64_t simt = simt_bwd_workspace_bytes(rows, k, n);
    if (math_mode == HGIN_MATH_FP32) return simt;
    const int64_t tc = hgin::tcgemm::bwd_workspace_bytes(rows, k, k < 4 ? k : 4, n);   // >= the bf16 kernels' need
    return simt > tc ? simt : tc;
}

// dx *= act'(post_z) in place, post_dalpha = sum dx * min(post_z, 0): the generic (non-fused) form of
// the post-activation used when the tensor-core kernel does not take the shape.
namespace hgin {
namespace {
__global__ void __launch_bounds__(256)
post_apply_kernel(int64_t rows, int width, float *__restrict__ dx, int64_t lddx, const float *__restrict__ z, int64_t ldz,
                  int act, const float *__restrict__ alpha_ptr, float *__restrict__ partials) {
    __shared__ float red[32];
    const float alpha = (act == HGIN_ACT_PRELU) ? __ldg(alpha_ptr) : 0.0f;
    float dal = 0.0f;
    const int64_t total = rows * width;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = i / width;
        const int c = static_cast<int>(i % width);
        const float zv = __ldg(z + m * ldz + c);
        const float v = dx[m * lddx + c];
        if (act == HGIN_ACT_PRELU && !(zv > 0.f)) dal = fmaf(v, zv, dal);
        dx[m * lddx + c] = act_backward(v, zv, act, alpha);
    }
    dal = block_sum(dal, red);
    if (threadIdx.x == 0 && partials) partials[blockIdx.x] = dal;
}
constexpr int kPostCtas = kNumSMs * 8;
}  // namespace
}  // namespace hgin

static int32_t linear_bwd_impl(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz,
                               int32_t act, const float *alpha, const float *x1, int64_t ld1, int32_t k1,
                               const float *x2, int64_t ld2, int32_t k2, const float *W, int32_t n, int32_t c0,
                               int32_t c1, float *dx, int64_t lddx, const float *dot_x, int64_t ld_dot,
                               float *ddot, float *dW, float *db, float *dalpha, void *workspace,
                               int64_t workspace_bytes, int32_t math_mode, void *stream,
                               const hgin::tcgemm::PostArgs *post) {
    using namespace hgin;
    const int k = k1 + k2;
    HGIN_CHECK_ARG(rows >= 0 && k1 > 0 && k2 >= 0 && n > 0, "hgin_linear_bwd: bad sizes");
    HGIN_CHECK_ARG(act >= HGIN_ACT_NONE && act <= HGIN_ACT_RELU, "hgin_linear_bwd: bad act %d", act);
    HGIN_CHECK_ARG(act == HGIN_ACT_NONE || z, "hgin_linear_bwd: activation backward needs z");
    HGIN_CHECK_ARG(act != HGIN_ACT_PRELU || alpha, "hgin_linear_bwd: PReLU needs alpha");
    HGIN_CHECK_ARG(0 <= c0 && c0 <= c1 && c1 <= k, "hgin_linear_bwd: bad column range [%d,%d) of %d", c0, c1, k);
    HGIN_CHECK_ARG(math_mode >= HGIN_MATH_FP32 && math_mode <= HGIN_MATH_BF16, "hgin_linear_bwd: bad math_mode %d", math_mode);
    HGIN_CHECK_ARG(!ddot || dot_x, "hgin_linear_bwd: ddot needs dot_x");
    HGIN_CHECK_ARG(g && W && x1 && (k2 == 0 || x2), "hgin_linear_bwd: null pointer");
    const int64_t need = hgin_linear_bwd_workspace_bytes(rows, k, n, math_mode);
    if (workspace_bytes < need || !workspace)
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_linear_bwd: workspace %lld < %lld bytes",
                    (long long)workspace_bytes, (long long)need);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t splits = dw_splits(rows);
    float *dw_partials = static_cast<float *>(workspace);
    float *scal = reinterpret_cast<float *>(static_cast<char *>(workspace) + align_up(splits * n * (k + 1) * 4, 256));

    const bool tc_ok = math_mode != HGIN_MATH_FP32 && rows > 0 &&
        tcgemm::bwd_eligible(rows, g, ldg, z, ldz, act, x1, ld1, k1, k2, n, c0, c1, dx, lddx, dot_x, ld_dot) &&
        (!post || (c1 - c0 >= 16 && post->ldz % 4 == 0 && aligned16(post->z)));
    if (rows > 0 && thin::head_bwd_eligible(x1, ld1, k1, k2, n, c0, c1, dx, lddx, dot_x, W, HGIN_DTYPE_F32) && !ddot &&
        !(post && post->self_eps) && (!post || (post->ldz % 4 == 0 && aligned16(post->z))))
        return thin::head_bwd(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, W, dx, lddx, dW, db, dalpha, workspace,
                              post ? post->z : nullptr, post ? post->ldz : 0, post ? post->act : HGIN_ACT_NONE,
                              post ? post->alpha : nullptr, post ? post->dalpha : nullptr, HGIN_DTYPE_F32,
                              static_cast<cudaStream_t>(stream));
    if (post && !tc_ok) return HGIN_ERR_UNSUPPORTED;   // hgin_linear_bwd_post then runs the generic form
    if (!post && rows > 0 && thin::bwd_eligible(g, ldg, z, ldz, act, k1, k2, n, c0, c1, dx, dot_x, HGIN_DTYPE_F32))
        return thin::linear_bwd(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, W, n, c0, c1, dot_x, ld_dot, ddot, dW, db,
                                dalpha, workspace, HGIN_DTYPE_F32, static_cast<cudaStream_t>(stream));
    if (tc_ok) {
        return tcgemm::linear_bwd(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, x2, ld2, k2, W, n, c0, c1, dx, lddx,
                                  dot_x, ld_dot, ddot, dW, db, dalpha, workspace, nullptr, post,
                                  static_cast<cudaStream_t>(stream));
    }
    if (rows == 0) {  // empty batch: all reductions are zero
        if (dW) cudaMemsetAsync(dW, 0, sizeof(float) * n * k, s);
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * n, s);
        if (dalpha) cudaMemsetAsync(dalpha, 0, sizeof(float), s);
        if (ddot) cudaMemsetAsync(ddot, 0, sizeof(float), s);
        return HGIN_OK;
    }

    // The PReLU slope stays in device memory (kernels read it through `alpha`), so a captured
    // graph replays with the current value.
    GradZ dzf{g, ldg, z, ldz, rows, n, act, 0.0f};

    // ---- input gradient ----
    const int width = c1 - c0;
    if (width > 0 && (dx || ddot)) {
        WeightCols fb{W, n, k, c0, c1};
        const unsigned gx = static_cast<unsigned>(ceil_div(rows, 128));
        const int vec_ok = !dx || (lddx % 4 == 0 && aligned16(dx));
        float *dot_partials = ddot ? scal : nullptr;
        unsigned gy = 1;
#define HGIN_DX(T)                                                                                              \
    gy = static_cast<unsigned>(ceil_div(width, T::BN));                                                         \
    linear_bwd_dx_kernel<T><<<dim3(gx, gy), THREADS, 0, s>>>(dzf, alpha, fb, dx, lddx, dot_x, ld_dot,           \
                                                             dot_partials, vec_ok)
        if (width > 64) { HGIN_DX(T128); }
        else if (width > 32) { HGIN_DX(T64); }
        else if (width > 16) { HGIN_DX(T32); }
        else if (width > 8) { HGIN_DX(T16); }
        else { HGIN_DX(T8); }
#undef HGIN_DX
        if (ddot) reduce_scalar_kernel<<<1, 1024, 0, s>>>(dot_partials, static_cast<int64_t>(gx) * gy, ddot);
    } else if (ddot) {
        cudaMemsetAsync(ddot, 0, sizeof(float), s);
    }

    // ---- weight / bias / slope gradients ----
    if (dW || db || dalpha) {
        ConcatRows xf{x1, ld1, k1, x2, ld2, k2, rows};
        const int64_t rows_per_cta = align_up(ceil_div(rows, splits), BK);
        const unsigned gx = static_cast<unsigned>(ceil_div(rows, rows_per_cta));
        float *alpha_partials = (dalpha && act == HGIN_ACT_PRELU) ? scal : nullptr;
        float *db_partials = dw_partials + static_cast<int64_t>(gx) * n * k;
        const bool swap = n <= 16 && k > n;   // tiny n: put it on the narrow side of the tile
        const int wide = swap ? k : n, narrow = swap ? n : k;
        const unsigned gy = static_cast<unsigned>(ceil_div(wide, 128));
        unsigned gz = 1;
#define HGIN_DW(T)                                                                                              \
    gz = static_cast<unsigned>(ceil_div(narrow, T::BN));                                                        \
    if (swap)                                                                                                   \
        linear_bwd_dw_kernel<T, true><<<dim3(gx, gy, gz), THREADS, 0, s>>>(                                     \
            dzf, alpha, xf, alpha_partials != nullptr, db != nullptr, rows_per_cta, dw_partials, db_partials,   \
            alpha_partials);                                                                                    \
    else                                                                                                        \
        linear_bwd_dw_kernel<T, false><<<dim3(gx, gy, gz), THREADS, 0, s>>>(                                    \
            dzf, alpha, xf, alpha_partials != nullptr, db != nullptr, rows_per_cta, dw_partials, db_partials,   \
            alpha_partials)
        if (narrow > 64) { HGIN_DW(T128); }
        else if (narrow > 32) { HGIN_DW(T64); }
        else if (narrow > 16) { HGIN_DW(T32); }
        else if (narrow > 8) { HGIN_DW(T16); }
        else { HGIN_DW(T8); }
#undef HGIN_DW
        const unsigned alpha_count = gx * (swap ? gz : gy);
        if (dW)
            reduce_partials_simt_kernel<<<grid_for(static_cast<int64_t>(n) * k, 256, 4), 256, 0, s>>>(
                dw_partials, static_cast<int>(gx), static_cast<int64_t>(n) * k, dW);
        if (db) reduce_partials_simt_kernel<<<1, 256, 0, s>>>(db_partials, static_cast<int>(gx), n, db);
        if (dalpha) {
            if (alpha_partials) reduce_scalar_kernel<<<1, 1024, 0, s>>>(alpha_partials, alpha_count, dalpha);
            else cudaMemsetAsync(dalpha, 0, sizeof(float), s);
        }
    }
    HGIN_CHECK_LAUNCH("hgin_linear_bwd");
    return HGIN_OK;
}

extern "C" int32_t hgin_linear_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz,
                                   int32_t act, const float *alpha, const float *x1, int64_t ld1, int32_t k1,
                                   const float *x2, int64_t ld2, int32_t k2, const float *W, int32_t n, int32_t c0,
                                   int32_t c1, float *dx, int64_t lddx, const float *dot_x, int64_t ld_dot,
                                   float *ddot, float *dW, float *db, float *dalpha, void *workspace,
                                   int64_t workspace_bytes, int32_t math_mode, void *stream) {
    return linear_bwd_impl(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, x2, ld2, k2, W, n, c0, c1, dx, lddx, dot_x,
                           ld_dot, ddot, dW, db, dalpha, workspace, workspace_bytes, math_mode, stream, nullptr);
}

extern "C" int32_t hgin_linear_bwd_post(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz,
                                        int32_t act, const float *alpha, const float *x1, int64_t ld1, int32_t k1,
                                        const float *x2, int64_t ld2, int32_t k2, const float *W, int32_t n, int32_t c0,
                                        int32_t c1, float *dx, int64_t lddx, float *dW, float *db, float *dalpha,
                                        const float *post_z, int64_t ld_post, int32_t post_act, const float *post_alpha,
                                        float *post_dalpha, void *workspace, int64_t workspace_bytes, int32_t math_mode,
                                        void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(post_act >= HGIN_ACT_NONE && post_act <= HGIN_ACT_RELU, "hgin_linear_bwd_post: bad post_act %d", post_act);
    HGIN_CHECK_ARG(post_act == HGIN_ACT_NONE || (post_z && dx && c1 > c0), "hgin_linear_bwd_post: post-activation needs post_z and dx");
    HGIN_CHECK_ARG(post_act != HGIN_ACT_PRELU || post_alpha, "hgin_linear_bwd_post: PReLU needs post_alpha");
    HGIN_CHECK_ARG(post_act == HGIN_ACT_NONE || ld_post >= c1 - c0, "hgin_linear_bwd_post: ld_post too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (post_act == HGIN_ACT_NONE) {
        if (post_dalpha) cudaMemsetAsync(post_dalpha, 0, sizeof(float), s);
        return linear_bwd_impl(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, x2, ld2, k2, W, n, c0, c1, dx, lddx, nullptr,
                               0, nullptr, dW, db, dalpha, workspace, workspace_bytes, math_mode, stream, nullptr);
    }
    tcgemm::PostArgs post{post_z, ld_post, post_act, post_alpha, post_dalpha, nullptr, nullptr};
    int32_t rc = linear_bwd_impl(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, x2, ld2, k2, W, n, c0, c1, dx, lddx,
                                 nullptr, 0, nullptr, dW, db, dalpha, workspace, workspace_bytes, math_mode, stream, &post);
    if (rc != HGIN_ERR_UNSUPPORTED) return rc;
    // generic form: plain backward, then one elementwise pass over dx
    rc = linear_bwd_impl(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, x2, ld2, k2, W, n, c0, c1, dx, lddx, nullptr, 0,
                         nullptr, dW, db, dalpha, workspace, workspace_bytes, math_mode, stream, nullptr);
    if (rc != HGIN_OK) return rc;
    if (workspace_bytes < static_cast<int64_t>(kPostCtas) * 4)
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_linear_bwd_post: workspace too small");
    float *partials = static_cast<float *>(workspace);   // the backward's partials are consumed by now (stream order)
    const bool want = post_dalpha && post_act == HGIN_ACT_PRELU;
    if (rows > 0) {
        const int grid = grid_for(rows * (c1 - c0), 256 * 4, 8);
        post_apply_kernel<<<grid, 256, 0, s>>>(rows, c1 - c0, dx, lddx, post_z, ld_post, post_act, post_alpha,
                                               want ? partials : nullptr);
        if (want) reduce_scalar_kernel<<<1, 1024, 0, s>>>(partials, grid, post_dalpha);
    }
    if (post_dalpha && (!want || rows == 0)) cudaMemsetAsync(post_dalpha, 0, sizeof(float), s);
    HGIN_CHECK_LAUNCH("hgin_linear_bwd_post");
    return HGIN_OK;
}

extern "C" int32_t hgin_linear_bwd_post_self(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz,
                                             int32_t act, const float *alpha, const float *x1, int64_t ld1, int32_t k1,
                                             const float *W, int32_t n, float *dx, int64_t lddx, float *dW, float *db,
                                             float *dalpha, const float *post_z, int64_t ld_post, int32_t post_act,
                                             const float *post_alpha, float *post_dalpha, const float *self_eps,
                                             float *post_ddot, void *workspace, int64_t workspace_bytes,
                                             int32_t math_mode, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(post_act == HGIN_ACT_PRELU || post_act == HGIN_ACT_RELU, "hgin_linear_bwd_post_self: bad post_act %d", post_act);
    HGIN_CHECK_ARG(post_z && dx && ld_post >= k1, "hgin_linear_bwd_post_self: needs post_z [rows, k1] and dx");
    HGIN_CHECK_ARG(post_act != HGIN_ACT_PRELU || post_alpha, "hgin_linear_bwd_post_self: PReLU needs post_alpha");
    tcgemm::PostArgs post{post_z, ld_post, post_act, post_alpha, post_dalpha, self_eps, post_ddot};
    // only the tensor-core input-gradient kernel carries this epilogue: other shapes / math modes return
    // HGIN_ERR_UNSUPPORTED and the caller runs hgin_linear_bwd + hgin_gin_combine_post instead
    const int32_t rc = linear_bwd_impl(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, nullptr, 0, 0, W, n, 0, k1, dx, lddx,
                                       nullptr, 0, nullptr, dW, db, dalpha, workspace, workspace_bytes, math_mode, stream,
                                       &post);
    if (rc == HGIN_ERR_UNSUPPORTED)
        return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_bwd_post_self: shapes / math mode outside the tensor-core path");
    return rc;
}

// ---- typed entry points: rows stored as float or bf16 -----------------------------------------------------
extern "C" int32_t hgin_linear_fwd_t(int32_t in_dtype, int32_t out_dtype, int64_t rows, const void *x1, int64_t ld1,
                                     int32_t k1, const float *x2, int64_t ld2, int32_t k2, const float *W,
                                     const float *bias, int32_t n, int32_t act, const float *alpha, void *z, int64_t ldz,
                                     void *out, int64_t ldo, int32_t accumulate_out, void *workspace,
                                     int64_t workspace_bytes, int32_t math_mode, void *stream) {
    using namespace hgin;
    if (in_dtype == HGIN_DTYPE_F32 && out_dtype == HGIN_DTYPE_F32)
        return hgin_linear_fwd(rows, static_cast<const float *>(x1), ld1, k1, x2, ld2, k2, W, bias, n, act, alpha,
                               static_cast<float *>(z), ldz, static_cast<float *>(out), ldo, accumulate_out, workspace,
                               workspace_bytes, math_mode, stream);
    HGIN_CHECK_ARG((in_dtype == HGIN_DTYPE_F32 || in_dtype == HGIN_DTYPE_BF16) && (out_dtype == HGIN_DTYPE_F32 || out_dtype == HGIN_DTYPE_BF16),
                   "hgin_linear_fwd_t: bad dtype %d / %d", in_dtype, out_dtype);
    HGIN_CHECK_ARG(rows >= 0 && k1 > 0 && k2 >= 0 && n > 0, "hgin_linear_fwd_t: bad sizes rows=%lld k1=%d k2=%d n=%d",
                   (long long)rows, k1, k2, n);
    HGIN_CHECK_ARG(act >= HGIN_ACT_NONE && act <= HGIN_ACT_RELU, "hgin_linear_fwd_t: bad act %d", act);
    HGIN_CHECK_ARG(act != HGIN_ACT_PRELU || alpha, "hgin_linear_fwd_t: PReLU needs alpha");
    if (rows == 0) return HGIN_OK;
    HGIN_CHECK_ARG(x1 && W && (k2 == 0 || x2) && (z || out), "hgin_linear_fwd_t: null pointer");
    HGIN_CHECK_ARG(ld1 >= k1 && (k2 == 0 || ld2 >= k2) && (!z || ldz >= n) && (!out || ldo >= n),
                   "hgin_linear_fwd_t: leading dimension too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // head (n = 1): wide bf16 input, fp32 output column
    if (in_dtype == HGIN_DTYPE_BF16 && out_dtype == HGIN_DTYPE_F32) {
        if (thin::head_fwd_eligible(x1, ld1, k1, k2, n, HGIN_DTYPE_BF16) && aligned16(W))
            return thin::head_fwd(rows, x1, ld1, k1, W, bias, act, alpha, static_cast<float *>(z), ldz, static_cast<float *>(out),
                                  ldo, accumulate_out, HGIN_DTYPE_BF16, s);
        return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_fwd_t: bf16 rows -> fp32 output is the n = 1 head only");
    }
    // thin (K <= 8): narrow fp32 input, wide bf16 output
    if (in_dtype == HGIN_DTYPE_F32 && out_dtype == HGIN_DTYPE_BF16) {
        if (thin::fwd_eligible(static_cast<const float *>(x1), k1, k2, n, z, ldz, out, ldo, HGIN_DTYPE_BF16))
            return thin::linear_fwd(rows, static_cast<const float *>(x1), ld1, k1, W, bias, n, act, alpha, z, ldz, out, ldo,
                                    accumulate_out, HGIN_DTYPE_BF16, s);
        return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_fwd_t: fp32 rows -> bf16 output is the K <= 8 layer only");
    }
    if (tcgemm::fwd_eligible_bf16(rows, x1, ld1, k1, k2, n, z, ldz, out, ldo)) {
        if (!workspace || workspace_bytes < tcgemm::fwd_workspace_bytes_bf16(k1 + k2, n))
            return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_linear_fwd_t: workspace too small for the bf16 path");
        return tcgemm::linear_fwd_bf16(rows, x1, ld1, k1, x2, ld2, k2, W, bias, n, act, alpha, z, ldz, out, ldo,
                                       accumulate_out, workspace, s);
    }
    return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_fwd_t: shape outside the bf16 kernels (rows=%lld k1=%d k2=%d n=%d)",
                (long long)rows, k1, k2, n);
}

extern "C" int32_t hgin_linear_bwd_t(int32_t g_dtype, int32_t x_dtype, int64_t rows, const void *g, int64_t ldg,
                                     const void *z, int64_t ldz, int32_t act, const float *alpha, const void *x1,
                                     int64_t ld1, int32_t k1, const float *x2, int64_t ld2, int32_t k2, const float *W,
                                     int32_t n, int32_t c0, int32_t c1, void *dx, int64_t lddx, const void *dot_x,
                                     int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha,
                                     const void *post_z, int64_t ld_post, int32_t post_act, const float *post_alpha,
                                     float *post_dalpha, const float *self_eps, float *post_ddot, void *workspace,
                                     int64_t workspace_bytes, int32_t math_mode, void *stream) {
    using namespace hgin;
    const bool post_on = post_z != nullptr && post_act != HGIN_ACT_NONE;
    if (g_dtype == HGIN_DTYPE_F32 && x_dtype == HGIN_DTYPE_F32) {
        if (post_on && (self_eps || post_ddot))
            return hgin_linear_bwd_post_self(rows, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, act,
                                             alpha, static_cast<const float *>(x1), ld1, k1, W, n, static_cast<float *>(dx), lddx,
                                             dW, db, dalpha, static_cast<const float *>(post_z), ld_post, post_act, post_alpha,
                                             post_dalpha, self_eps, post_ddot, workspace, workspace_bytes, math_mode, stream);
        if (post_on)
            return hgin_linear_bwd_post(rows, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, act, alpha,
                                        static_cast<const float *>(x1), ld1, k1, x2, ld2, k2, W, n, c0, c1,
                                        static_cast<float *>(dx), lddx, dW, db, dalpha, static_cast<const float *>(post_z),
                                        ld_post, post_act, post_alpha, post_dalpha, workspace, workspace_bytes, math_mode,
                                        stream);
        return hgin_linear_bwd(rows, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, act, alpha,
                               static_cast<const float *>(x1), ld1, k1, x2, ld2, k2, W, n, c0, c1, static_cast<float *>(dx),
                               lddx, static_cast<const float *>(dot_x), ld_dot, ddot, dW, db, dalpha, workspace,
                               workspace_bytes, math_mode, stream);
    }
    const int k = k1 + k2;
    HGIN_CHECK_ARG((g_dtype == HGIN_DTYPE_F32 || g_dtype == HGIN_DTYPE_BF16) && (x_dtype == HGIN_DTYPE_F32 || x_dtype == HGIN_DTYPE_BF16),
                   "hgin_linear_bwd_t: bad dtype %d / %d", g_dtype, x_dtype);
    HGIN_CHECK_ARG(rows >= 0 && k1 > 0 && k2 >= 0 && n > 0, "hgin_linear_bwd_t: bad sizes");
    HGIN_CHECK_ARG(act >= HGIN_ACT_NONE && act <= HGIN_ACT_RELU, "hgin_linear_bwd_t: bad act %d", act);
    HGIN_CHECK_ARG(act == HGIN_ACT_NONE || z, "hgin_linear_bwd_t: activation backward needs z");
    HGIN_CHECK_ARG(act != HGIN_ACT_PRELU || alpha, "hgin_linear_bwd_t: PReLU needs alpha");
    HGIN_CHECK_ARG(0 <= c0 && c0 <= c1 && c1 <= k, "hgin_linear_bwd_t: bad column range [%d,%d) of %d", c0, c1, k);
    HGIN_CHECK_ARG(!ddot || dot_x, "hgin_linear_bwd_t: ddot needs dot_x");
    HGIN_CHECK_ARG(g && W && x1 && (k2 == 0 || x2), "hgin_linear_bwd_t: null pointer");
    HGIN_CHECK_ARG(post_act >= HGIN_ACT_NONE && post_act <= HGIN_ACT_RELU, "hgin_linear_bwd_t: bad post_act %d", post_act);
    HGIN_CHECK_ARG(!post_on || (dx && c1 > c0 && ld_post >= c1 - c0), "hgin_linear_bwd_t: post-activation needs dx and post_z rows");
    HGIN_CHECK_ARG(post_act != HGIN_ACT_PRELU || !post_on || post_alpha, "hgin_linear_bwd_t: PReLU needs post_alpha");
    HGIN_CHECK_ARG(!(self_eps || post_ddot) || (post_on && k2 == 0 && c0 == 0 && c1 == k1),
                   "hgin_linear_bwd_t: the self-branch epilogue needs a post-activation over all columns of x1");
    const int64_t need = hgin_linear_bwd_workspace_bytes(rows, k, n, HGIN_MATH_BF16);
    if (workspace_bytes < need || !workspace)
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_linear_bwd_t: workspace %lld < %lld bytes", (long long)workspace_bytes,
                    (long long)need);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (rows == 0) {
        if (dW) cudaMemsetAsync(dW, 0, sizeof(float) * n * k, s);
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * n, s);
        if (dalpha) cudaMemsetAsync(dalpha, 0, sizeof(float), s);
        if (ddot) cudaMemsetAsync(ddot, 0, sizeof(float), s);
        if (post_dalpha) cudaMemsetAsync(post_dalpha, 0, sizeof(float), s);
        if (post_ddot) cudaMemsetAsync(post_ddot, 0, sizeof(float), s);
        return HGIN_OK;
    }
    // head (n = 1): g / z fp32 columns, x / dx / post_z bf16
    if (g_dtype == HGIN_DTYPE_F32 && x_dtype == HGIN_DTYPE_BF16) {
        if (thin::head_bwd_eligible(x1, ld1, k1, k2, n, c0, c1, dx, lddx, dot_x, W, HGIN_DTYPE_BF16) && !ddot && !self_eps &&
            !post_ddot && (!post_on || (ld_post % 4 == 0 && (reinterpret_cast<uintptr_t>(post_z) & 7u) == 0)))
            return thin::head_bwd(rows, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, act, alpha, x1,
                                  ld1, k1, W, dx, lddx, dW, db, dalpha, workspace, post_on ? post_z : nullptr, ld_post,
                                  post_on ? post_act : HGIN_ACT_NONE, post_alpha, post_dalpha, HGIN_DTYPE_BF16, s);
        return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_bwd_t: fp32 gradient with bf16 rows is the n = 1 head only");
    }
    // thin (K <= 8): g / z bf16, x / dot_x fp32, no dx
    if (g_dtype == HGIN_DTYPE_BF16 && x_dtype == HGIN_DTYPE_F32) {
        if (!post_on && thin::bwd_eligible(g, ldg, z, ldz, act, k1, k2, n, c0, c1, dx, static_cast<const float *>(dot_x),
                                           HGIN_DTYPE_BF16))
            return thin::linear_bwd(rows, g, ldg, z, ldz, act, alpha, static_cast<const float *>(x1), ld1, k1, W, n, c0, c1,
                                    static_cast<const float *>(dot_x), ld_dot, ddot, dW, db, dalpha, workspace,
                                    HGIN_DTYPE_BF16, s);
        return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_bwd_t: bf16 gradient with fp32 rows is the K <= 8 layer only");
    }
    if (tcgemm::bwd_eligible_bf16(rows, g, ldg, z, ldz, act, x1, ld1, k1, k2, n, c0, c1, dx, lddx, dot_x, ld_dot) &&
        (!post_on || (c1 - c0 >= 16 && ld_post % 8 == 0 && aligned16(post_z)))) {
        tcgemm::PostArgs post{post_z, ld_post, post_act, post_alpha, post_dalpha, self_eps, post_ddot};
        return tcgemm::linear_bwd_bf16(rows, g, ldg, z, ldz, act, alpha, x1, ld1, k1, x2, ld2, k2, W, n, c0, c1, dx, lddx,
                                       dot_x, ld_dot, ddot, dW, db, dalpha, workspace, nullptr, post_on ? &post : nullptr, s);
    }
    return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_bwd_t: shape outside the bf16 kernels (rows=%lld k1=%d k2=%d n=%d)",
                (long long)rows, k1, k2, n);
}

extern "C" int32_t hgin_debug_gemm_tn_bf16(int64_t rows, const void *a, int32_t n, const void *b, int32_t k, float *out,
                                           void *workspace, int64_t workspace_bytes, int32_t lbo, int32_t sbo,
                                           int32_t layout_type, int32_t k_step_bytes, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(rows > 0 && a && b && out && n >= 16 && n <= 128 && n % 16 == 0 && k >= 16 && k <= 128 && k % 16 == 0,
                   "hgin_debug_gemm_tn_bf16: bad arguments");
    if (!workspace || workspace_bytes < tcgemm::bwd_workspace_bytes_bf16(rows, k, 0, n))
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_debug_gemm_tn_bf16: workspace too small");
    tcgemm::TnDebug d{0, lbo < 0 ? 8192 : lbo, sbo < 0 ? 1024 : sbo, layout_type < 0 ? 2 : layout_type,
                      k_step_bytes < 0 ? 2048 : k_step_bytes};
    // dz = a (no activation), dW = a^T b
    return tcgemm::linear_bwd_bf16(rows, a, n, nullptr, 0, HGIN_ACT_NONE, nullptr, b, k, k, nullptr, 0, 0, out /*W unused*/, n,
                                   0, 0, nullptr, 0, nullptr, 0, nullptr, out, nullptr, nullptr, workspace, &d, nullptr,
                                   static_cast<cudaStream_t>(stream));
}

extern "C" int32_t hgin_debug_gemm_tn(int64_t rows, const float *a, int32_t n, const float *b, int32_t k, float *out,
                                      void *workspace, int64_t workspace_bytes, int32_t tma_swizzle, int32_t lbo,
                                      int32_t sbo, int32_t layout_type, int32_t k_step_bytes, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(rows > 0 && a && b && out && n >= 16 && n <= 128 && n % 16 == 0 && k >= 16 && k <= 128 && k % 16 == 0,
                   "hgin_debug_gemm_tn: bad arguments");
    if (!workspace || workspace_bytes < tcgemm::bwd_workspace_bytes(rows, k, 0, n))
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_debug_gemm_tn: workspace too small");
    tcgemm::TnDebug d{tma_swizzle < 0 ? 4 : tma_swizzle, lbo < 0 ? 4096 : lbo, sbo < 0 ? 512 : sbo,
                      layout_type < 0 ? 1 : layout_type, k_step_bytes < 0 ? 1024 : k_step_bytes};
    // dz = a (no activation), dW = a^T b
    return tcgemm::linear_bwd(rows, a, n, nullptr, 0, HGIN_ACT_NONE, nullptr, b, k, k, nullptr, 0, 0, out /*W unused*/, n,
                              0, 0, nullptr, 0, nullptr, 0, nullptr, out, nullptr, nullptr, workspace, &d, nullptr,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int32_t hgin_set_option(const char *name, int32_t value) {
    using namespace hgin;
    HGIN_CHECK_ARG(name != nullptr, "hgin_set_option: null name");
    if (strcmp(name, "fused_bwd") == 0) {
        tcgemm::set_fused_bwd(value);
        return HGIN_OK;
    }
    if (strcmp(name, "fused_dw") == 0) {
        tcgemm::set_fused_dw(value);
        return HGIN_OK;
    }
    return fail(HGIN_ERR_INVALID_ARGUMENT, "hgin_set_option: unknown option '%s'", name);
}
