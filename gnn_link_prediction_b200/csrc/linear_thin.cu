// Dense layers with a tiny contraction (K <= 8): layer 0 of the GIN stack, where the MLP input
// is [agg(3) | (1+eps) x_dst(3)] (models.py:287-290), and every GIN layer at config.json's
// NODE_EMBEDDING_SIZE = 8.  These are pure streaming kernels (24-32 B in, up to 1 KB out per row),
// so the tiled GEMM engine only wastes issue slots on them: here each lane owns 4 output columns,
// keeps its W slice in registers and walks rows, reading g/z and writing z/out as coalesced
// 128-bit vectors.  fp32 FMA throughout (this path serves both math modes).
//
// Backward computes, in ONE pass over g and z:
//   dW[n][k] = sum_m dz[m][n] h[m][k],  db[n] = sum_m dz[m][n],  dalpha = sum g * min(z, 0),
//   T[n][c]  = sum_m dz[m][n] dot_x[m][c]  ->  ddot = sum_{n,c} W[n][c0 + c] T[n][c]
// (the last equals sum_m sum_c (dz W)[m][c0+c] * dot_x[m][c] = d(eps) without materialising dz W).
// Reductions are deterministic: lanes -> CTA (shared memory, fixed order) -> partials -> final pass.
#include "hgin_common.cuh"
#include "linear_thin_api.h"

namespace hgin {
namespace thin {
namespace {

constexpr int THREADS = 256;
constexpr int KMAX = 8;
constexpr int DMAX = 4;                      // max dot_x columns
constexpr int NACC = 4 * KMAX + 4 * DMAX + 4 + 1;  // per-thread accumulators: dW, T, db, dalpha = 53

__global__ void __launch_bounds__(THREADS)
thin_fwd_kernel(int64_t rows, const float *__restrict__ x, int64_t ldx, int k, const float *__restrict__ W,
                const float *__restrict__ bias, int n, int act, const float *__restrict__ alpha_ptr,
                float *__restrict__ z, int64_t ldz, float *__restrict__ out, int64_t ldo, int accumulate_out) {
    const int lpr = n >> 2;                        // lanes per row
    const int cg = threadIdx.x % lpr;              // column group: columns 4cg .. 4cg+3
    const int64_t slot = (static_cast<int64_t>(blockIdx.x) * THREADS + threadIdx.x) / lpr;
    const int64_t num_slots = (static_cast<int64_t>(gridDim.x) * THREADS) / lpr;
    float w[4][KMAX], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        b[i] = bias ? __ldg(bias + cg * 4 + i) : 0.0f;
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) w[i][kk] = kk < k ? __ldg(W + static_cast<int64_t>(cg * 4 + i) * k + kk) : 0.0f;
    }
    const float alpha = act == HGIN_ACT_PRELU ? __ldg(alpha_ptr) : 0.0f;
    for (int64_t m = slot; m < rows; m += num_slots) {
        float xv[KMAX];
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) xv[kk] = kk < k ? __ldg(x + m * ldx + kk) : 0.0f;
        float zz[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float s = 0.0f;
#pragma unroll
            for (int kk = 0; kk < KMAX; ++kk) s = fmaf(xv[kk], w[i][kk], s);
            zz[i] = s + b[i];
        }
        if (z) *reinterpret_cast<float4 *>(z + m * ldz + cg * 4) = make_float4(zz[0], zz[1], zz[2], zz[3]);
        if (out) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = act_forward(zz[i], act, alpha);
            float4 *po = reinterpret_cast<float4 *>(out + m * ldo + cg * 4);
            if (accumulate_out) {
                const float4 old = *po;
                o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
            }
            *po = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// part[cta][n][kp], kp = k + d + 1: columns [0,k) = dW, [k, k+d) = T, k+d = db;  alpha_part[cta].
__global__ void __launch_bounds__(THREADS)
thin_bwd_kernel(int64_t rows, const float *__restrict__ g, int64_t ldg, const float *__restrict__ z, int64_t ldz,
                int act, const float *__restrict__ alpha_ptr, const float *__restrict__ x, int64_t ldx, int k,
                const float *__restrict__ dot_x, int64_t ld_dot, int d, int n, float *__restrict__ part,
                float *__restrict__ alpha_part) {
    __shared__ float sm[THREADS * 8];
    __shared__ float red[32];
    const int lpr = n >> 2;
    const int cg = threadIdx.x % lpr;
    const int grp = threadIdx.x / lpr;             // row slot inside the CTA
    const int slots_per_cta = THREADS / lpr;
    const int64_t slot = static_cast<int64_t>(blockIdx.x) * slots_per_cta + grp;
    const int64_t num_slots = static_cast<int64_t>(gridDim.x) * slots_per_cta;
    const float alpha = act == HGIN_ACT_PRELU ? __ldg(alpha_ptr) : 0.0f;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
    // acc layout: [i*KMAX + kk] dW, [32 + i*DMAX + c] T, [48 + i] db, [52] dalpha
    for (int64_t m = slot; m < rows; m += num_slots) {
        const float4 gv = __ldg(reinterpret_cast<const float4 *>(g + m * ldg) + cg);
        float dz[4] = {gv.x, gv.y, gv.z, gv.w};
        if (act != HGIN_ACT_NONE) {
            const float4 zv4 = __ldg(reinterpret_cast<const float4 *>(z + m * ldz) + cg);
            const float zv[4] = {zv4.x, zv4.y, zv4.z, zv4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (act == HGIN_ACT_PRELU && !(zv[i] > 0.0f)) acc[52] += dz[i] * zv[i];
                dz[i] = act_backward(dz[i], zv[i], act, alpha);
            }
        }
        float xv[KMAX], dv[DMAX];
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk) xv[kk] = kk < k ? __ldg(x + m * ldx + kk) : 0.0f;
#pragma unroll
        for (int c = 0; c < DMAX; ++c) dv[c] = c < d ? __ldg(dot_x + m * ld_dot + c) : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int kk = 0; kk < KMAX; ++kk) acc[i * KMAX + kk] = fmaf(dz[i], xv[kk], acc[i * KMAX + kk]);
#pragma unroll
            for (int c = 0; c < DMAX; ++c) acc[32 + i * DMAX + c] = fmaf(dz[i], dv[c], acc[32 + i * DMAX + c]);
            acc[48 + i] += dz[i];
        }
    }
    // CTA combine, 8 accumulators at a time: sm[thread][8] -> sum over row slots in slot order
    const int kp = k + d + 1;
    float *dst = part + static_cast<int64_t>(blockIdx.x) * n * kp;
#pragma unroll
    for (int round = 0; round < 7; ++round) {   // rounds 0-3: dW of column i; 4-5: T; 6: db (+ dalpha)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int a = round * 8 + j;
            sm[threadIdx.x * 8 + j] = a < NACC - 1 ? acc[a] : 0.0f;
        }
        __syncthreads();
        // one thread per (column group, j): sums the row slots
        for (int idx = threadIdx.x; idx < lpr * 8; idx += THREADS) {
            const int c_g = idx / 8, j = idx % 8;
            float s = 0.0f;
            for (int sl = 0; sl < slots_per_cta; ++sl) s += sm[(sl * lpr + c_g) * 8 + j];
            const int a = round * 8 + j;
            if (a < 32) {                       // dW: i = a / 8, kk = a % 8
                const int i = a >> 3, kk = a & 7;
                if (kk < k) dst[static_cast<int64_t>(c_g * 4 + i) * kp + kk] = s;
            } else if (a < 48) {                // T: i = (a-32)/4, c = (a-32)%4
                const int i = (a - 32) >> 2, c = (a - 32) & 3;
                if (c < d) dst[static_cast<int64_t>(c_g * 4 + i) * kp + k + c] = s;
            } else if (a < 52) {                // db
                dst[static_cast<int64_t>(c_g * 4 + (a - 48)) * kp + k + d] = s;
            }
        }
    }
    __syncthreads();
    const float da = block_sum(acc[52], red);
    if (threadIdx.x == 0 && alpha_part) alpha_part[blockIdx.x] = da;
}

// dW / db from the partials; ddot = sum_{n,c} W[n][c0+c] * T[n][c]; dalpha = sum of alpha partials.
__global__ void __launch_bounds__(1024)
thin_finalize_kernel(const float *__restrict__ part, int num_part, int n, int k, int d, const float *__restrict__ W,
                     int c0, float *__restrict__ dW, float *__restrict__ db, float *__restrict__ ddot,
                     const float *__restrict__ alpha_part, float *__restrict__ dalpha) {
    __shared__ float red[32];
    const int kp = k + d + 1;
    const int total = n * kp;
    float dot = 0.0f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;   // four independent chains, fixed association
        int p = 0;
        for (; p + 3 < num_part; p += 4) {
            s0 += part[static_cast<int64_t>(p) * total + i];
            s1 += part[static_cast<int64_t>(p + 1) * total + i];
            s2 += part[static_cast<int64_t>(p + 2) * total + i];
            s3 += part[static_cast<int64_t>(p + 3) * total + i];
        }
        for (; p < num_part; ++p) s0 += part[static_cast<int64_t>(p) * total + i];
        const float s = (s0 + s1) + (s2 + s3);
        const int nn = i / kp, c = i % kp;
        if (c < k) {
            if (dW) dW[nn * k + c] = s;
        } else if (c < k + d) {
            dot = fmaf(__ldg(W + nn * k + c0 + (c - k)), s, dot);
        } else if (db) {
            db[nn] = s;
        }
    }
    dot = block_sum(dot, red);
    if (threadIdx.x == 0 && ddot) ddot[0] = dot;
    if (dalpha) {
        float s = 0.0f;
        if (alpha_part)
            for (int i = threadIdx.x; i < num_part; i += blockDim.x) s += alpha_part[i];
        s = block_sum(s, red);
        if (threadIdx.x == 0) dalpha[0] = s;
    }
}

inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int thin_ctas(int64_t rows, int n) {
    const int slots = THREADS / (n >> 2);
    const int64_t want = ceil_div(rows, static_cast<int64_t>(slots) * 4);
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 4;
    return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

bool fwd_eligible(const float *x1, int k1, int k2, int n, const float *z, int64_t ldz, const float *out, int64_t ldo) {
    return x1 && k2 == 0 && k1 <= KMAX && n >= 4 && n <= 128 && pow2(n) && (!z || (ldz % 4 == 0 && aligned16(z))) &&
           (!out || (ldo % 4 == 0 && aligned16(out)));
}

bool bwd_eligible(const float *g, int64_t ldg, const float *z, int64_t ldz, int act, int k1, int k2, int n, int c0,
                  int c1, const float *dx, const float *dot_x) {
    const int d = dot_x ? c1 - c0 : 0;
    return k2 == 0 && k1 <= KMAX && n >= 4 && n <= 128 && pow2(n) && dx == nullptr && d <= DMAX &&
           (c1 == c0 || dot_x != nullptr) && ldg % 4 == 0 && aligned16(g) &&
           (act == HGIN_ACT_NONE || (z && ldz % 4 == 0 && aligned16(z)));
}

int64_t bwd_workspace_bytes(int n, int k) {
    return align_up(static_cast<int64_t>(kNumSMs) * 4 * (static_cast<int64_t>(n) * (k + DMAX + 1) + 1) * 4, 256) + 256;
}

int32_t linear_fwd(int64_t rows, const float *x, int64_t ldx, int k, const float *W, const float *bias, int n, int act,
                   const float *alpha, float *z, int64_t ldz, float *out, int64_t ldo, int accumulate_out,
                   cudaStream_t s) {
    thin_fwd_kernel<<<thin_ctas(rows, n), THREADS, 0, s>>>(rows, x, ldx, k, W, bias, n, act, alpha, z, ldz, out, ldo,
                                                          accumulate_out);
    HGIN_CHECK_LAUNCH("hgin_linear_fwd(thin)");
    return HGIN_OK;
}

int32_t linear_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act, const float *alpha,
                   const float *x, int64_t ldx, int k, const float *W, int n, int c0, int c1, const float *dot_x,
                   int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace, cudaStream_t s) {
    const int d = dot_x ? c1 - c0 : 0;
    const int ctas = thin_ctas(rows, n);
    float *part = static_cast<float *>(workspace);
    float *alpha_part = part + static_cast<int64_t>(ctas) * n * (k + d + 1);
    const bool want_alpha = dalpha && act == HGIN_ACT_PRELU;
    thin_bwd_kernel<<<ctas, THREADS, 0, s>>>(rows, g, ldg, z, ldz, act, alpha, x, ldx, k, dot_x, ld_dot, d, n, part,
                                             want_alpha ? alpha_part : nullptr);
    thin_finalize_kernel<<<1, 1024, 0, s>>>(part, ctas, n, k, d, W, c0, dW, db, ddot, want_alpha ? alpha_part : nullptr,
                                            dalpha);
    HGIN_CHECK_LAUNCH("hgin_linear_bwd(thin)");
    return HGIN_OK;
}

}  // namespace thin
}  // namespace hgin
