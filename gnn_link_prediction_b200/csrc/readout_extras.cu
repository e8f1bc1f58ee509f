// The non-default branches of HetroGIN's readout and layer loop (models.py:301-330, 347-371):
//   * elementwise activations other than PReLU / ReLU (`act = eval(act)`, models.py:301, 330),
//   * BatchNorm1d between Linear and the activation (`mlp_bn`, models.py:303-313),
//   * dropout on every layer output (models.py:358-359),
//   * per-graph mean / max pool of the raw path features, gathered back per path (`global_feats`, models.py:347-352).
// All of them are row-streaming, HBM-bound passes (or tiny per-graph reductions); reductions are two-stage with a
// fixed association (run-to-run deterministic), column statistics accumulate in fp64.
#include <cuda_bf16.h>
#include <math.h>

#include "hgin_common.cuh"

namespace hgin {
namespace {

// ---- activation table ------------------------------------------------------------------------
// forward / derivative of every HGIN_ACT_* at pre-activation x; `a` = learnable PReLU slope (HGIN_ACT_PRELU),
// p0 / p1 = the constructor constants of the torch module (LeakyReLU.negative_slope, ELU.alpha, Softplus.beta/threshold).
__device__ __forceinline__ float actx_fwd(float x, int act, float a, float p0, float p1) {
    switch (act) {
        case HGIN_ACT_PRELU: return x > 0.f ? x : a * x;
        case HGIN_ACT_RELU: return x > 0.f ? x : 0.f;
        case HGIN_ACT_LEAKY_RELU: return x > 0.f ? x : p0 * x;
        case HGIN_ACT_ELU: return x > 0.f ? x : p0 * expm1f(x);
        case HGIN_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        case HGIN_ACT_TANH: return tanhf(x);
        case HGIN_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
        case HGIN_ACT_SILU: return x / (1.f + expf(-x));
        case HGIN_ACT_SOFTPLUS: return x * p0 > p1 ? x : log1pf(expf(x * p0)) / p0;
        default: return x;
    }
}
__device__ __forceinline__ float actx_grad(float x, int act, float a, float p0, float p1) {
    switch (act) {
        case HGIN_ACT_PRELU: return x > 0.f ? 1.f : a;
        case HGIN_ACT_RELU: return x > 0.f ? 1.f : 0.f;
        case HGIN_ACT_LEAKY_RELU: return x > 0.f ? 1.f : p0;
        case HGIN_ACT_ELU: return x > 0.f ? 1.f : p0 * expf(x);
        case HGIN_ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }
        case HGIN_ACT_TANH: { const float t = tanhf(x); return 1.f - t * t; }
        case HGIN_ACT_GELU: {
            const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
            return cdf + x * pdf;
        }
        case HGIN_ACT_SILU: { const float s = 1.f / (1.f + expf(-x)); return s * (1.f + x * (1.f - s)); }
        case HGIN_ACT_SOFTPLUS: { if (x * p0 > p1) return 1.f; const float z = expf(x * p0); return z / (z + 1.f); }
        default: return 1.f;
    }
}

// ---- 4-wide row pieces in either storage type --------------------------------------------------
template <typename T> struct Pack4;
template <> struct Pack4<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Pack4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[4]) {
        const uint2 q = *reinterpret_cast<const uint2 *>(p);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&q.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&q.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        const __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 q;
        q.x = *reinterpret_cast<const uint32_t *>(&a);
        q.y = *reinterpret_cast<const uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(p) = q;
    }
};
template <typename T> __device__ __forceinline__ float ld1(const T *p);
template <> __device__ __forceinline__ float ld1<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T *p, float v);
template <> __device__ __forceinline__ void st1<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// One "piece" = V consecutive elements of a row (V = 4 when widths, leading dimensions and base pointers allow 4-wide
// accesses, else 1).  Pieces are dealt to threads grid-stride; piece -> (row, first column).
template <typename T, int V>
__device__ __forceinline__ void load_piece(const T *base, int64_t ld, int64_t r, int c, float (&v)[4]) {
    if constexpr (V == 4) Pack4<T>::load(base + r * ld + c, v);
    else v[0] = ld1<T>(base + r * ld + c);
}
template <typename T, int V>
__device__ __forceinline__ void store_piece(T *base, int64_t ld, int64_t r, int c, const float (&v)[4]) {
    if constexpr (V == 4) Pack4<T>::store(base + r * ld + c, v);
    else st1<T>(base + r * ld + c, v[0]);
}

constexpr int kThreads = 256;
constexpr int kMaxCtas = kNumSMs * 8;

inline int piece_ctas(int64_t pieces) {
    const int64_t want = ceil_div(pieces, static_cast<int64_t>(kThreads) * 4);
    return static_cast<int>(want < 1 ? 1 : (want > kMaxCtas ? kMaxCtas : want));
}

template <typename T>
inline bool vec4_ok(int n, std::initializer_list<const void *> ptrs, std::initializer_list<int64_t> lds) {
    if (n % 4 != 0) return false;
    for (int64_t ld : lds) if (ld % 4 != 0) return false;
    for (const void *p : ptrs) if (p != nullptr && (reinterpret_cast<uintptr_t>(p) % (4 * sizeof(T))) != 0) return false;
    return true;
}

// ---- standalone activation ----------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
act_fwd_kernel(int64_t rows, int n, const T *__restrict__ z, int64_t ldz, int act, const float *__restrict__ alpha,
               float p0, float p1, T *__restrict__ out, int64_t ldo) {
    const float a = alpha ? __ldg(alpha) : 0.f;
    const int ppr = n / V;
    const int64_t pieces = rows * ppr;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < pieces;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t r = i / ppr;
        const int c = static_cast<int>(i - r * ppr) * V;
        float v[4];
        load_piece<T, V>(z, ldz, r, c, v);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = actx_fwd(v[j], act, a, p0, p1);
        store_piece<T, V>(out, ldo, r, c, v);
    }
}

// dz = g * act'(z);  partial[cta] = sum g * min(z, 0) (the PReLU slope gradient), or untouched when partial == NULL
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
act_bwd_kernel(int64_t rows, int n, const T *__restrict__ g, int64_t ldg, const T *__restrict__ z, int64_t ldz, int act,
               const float *__restrict__ alpha, float p0, float p1, T *__restrict__ dz, int64_t lddz,
               float *__restrict__ partial) {
    __shared__ float red[32];
    const float a = alpha ? __ldg(alpha) : 0.f;
    const int ppr = n / V;
    const int64_t pieces = rows * ppr;
    float da = 0.f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < pieces;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t r = i / ppr;
        const int c = static_cast<int>(i - r * ppr) * V;
        float gv[4], zv[4];
        load_piece<T, V>(g, ldg, r, c, gv);
        load_piece<T, V>(z, ldz, r, c, zv);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            da += gv[j] * fminf(zv[j], 0.f);
            gv[j] *= actx_grad(zv[j], act, a, p0, p1);
        }
        store_piece<T, V>(dz, lddz, r, c, gv);
    }
    if (partial != nullptr) {
        da = block_sum(da, red);
        if (threadIdx.x == 0) partial[blockIdx.x] = da;
    }
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(const float *__restrict__ partial, int count, float *out) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += partial[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

// ---- dropout (Philox4x32-10, counter = piece index of the [rows, n] matrix) ----------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
dropout_kernel(int64_t rows, int n, const T *__restrict__ x, int64_t ldx, float p, float scale, uint64_t seed,
               uint64_t offset, T *__restrict__ out, int64_t ldo) {
    const int n4 = (n + 3) / 4;                 // the mask is a function of (row, column) only: the same for V = 1 and 4
    const int ppr = n / V;
    const int64_t pieces = rows * ppr;
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < pieces;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t r = i / ppr;
        const int c = static_cast<int>(i - r * ppr) * V;
        const uint64_t group = static_cast<uint64_t>(r) * n4 + (c >> 2);
        const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(group), static_cast<uint32_t>(group >> 32),
                                                   static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)), key);
        const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
        float v[4];
        load_piece<T, V>(x, ldx, r, c, v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float u = static_cast<float>(w[(c + j) & 3] >> 8) * (1.0f / 16777216.0f);   // [0, 1)
            v[j] = u >= p ? v[j] * scale : 0.f;
        }
        store_piece<T, V>(out, ldo, r, c, v);
    }
}

// ---- BatchNorm1d column statistics ----------------------------------------------------------------
// blockDim = (32, 8): x over columns (a warp reads 32 consecutive elements of a row), y over rows.  Each CTA owns a
// contiguous row range and writes partial[cta][k][c] (k = 0: sum, 1: sum of squares) in fp64.
constexpr int kBnCols = 8;                    // columns per thread -> n <= 256 per launch (host loops over chunks)
constexpr int kBnRowsY = 8;

template <typename T>
__global__ void __launch_bounds__(32 * kBnRowsY)
bn_stats_kernel(int64_t rows, int n, const T *__restrict__ z, int64_t ldz, double *__restrict__ partial) {
    __shared__ double red[kBnRowsY][32];
    double s[kBnCols], ss[kBnCols];
#pragma unroll
    for (int j = 0; j < kBnCols; ++j) s[j] = ss[j] = 0.0;
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rows) ? r0 + per : rows;
    // four rows per iteration with all loads issued first: one 4-byte load per thread in flight ran at 2.2 TB/s
    constexpr int RU = 4;
    int64_t r = r0 + threadIdx.y;
    for (; r + (RU - 1) * kBnRowsY < r1; r += RU * kBnRowsY) {
        float v[RU][kBnCols];
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
            for (int j = 0; j < kBnCols; ++j) {
                const int c = threadIdx.x + 32 * j;
                v[u][j] = c < n ? ld1<T>(z + (r + u * kBnRowsY) * ldz + c) : 0.f;
            }
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
            for (int j = 0; j < kBnCols; ++j) {
                const double d = static_cast<double>(v[u][j]);
                s[j] += d;
                ss[j] = fma(d, d, ss[j]);
            }
    }
    for (; r < r1; r += kBnRowsY) {
#pragma unroll
        for (int j = 0; j < kBnCols; ++j) {
            const int c = threadIdx.x + 32 * j;
            if (c < n) {
                const double d = static_cast<double>(ld1<T>(z + r * ldz + c));
                s[j] += d;
                ss[j] = fma(d, d, ss[j]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < kBnCols; ++j) {
            const int c = threadIdx.x + 32 * j;
            red[threadIdx.y][threadIdx.x] = k == 0 ? s[j] : ss[j];
            __syncthreads();
            if (threadIdx.y == 0 && c < n) {
                double t = 0.0;
                for (int y = 0; y < kBnRowsY; ++y) t += red[y][threadIdx.x];
                partial[(static_cast<int64_t>(blockIdx.x) * 2 + k) * n + c] = t;
            }
            __syncthreads();
        }
    }
}

// sums[k * n_total + c0 + c] = sum over CTAs (fixed order) of partial[cta][k][c]
__global__ void bn_reduce_kernel(const double *__restrict__ partial, int ctas, int n, int kinds, double *__restrict__ sums,
                                 int n_total, int c0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kinds * n) return;
    const int k = i / n, c = i - k * n;
    double t = 0.0;
    for (int b = 0; b < ctas; ++b) t += partial[(static_cast<int64_t>(b) * kinds + k) * n + c];
    sums[static_cast<int64_t>(k) * n_total + c0 + c] = t;
}

__global__ void bn_set_count_kernel(double *sums, int n, double count) { sums[2 * n] = count; }

// mean / invstd from the (all-reduced) sums; running statistics updated as torch.nn.BatchNorm1d does in training mode
// (momentum form, unbiased variance).  use_running: eval mode, statistics come from the running buffers.
__global__ void bn_finalize_kernel(int n, const double *__restrict__ sums, double eps, double momentum, int use_running,
                                   float *__restrict__ mean, float *__restrict__ invstd, float *running_mean,
                                   float *running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (use_running) {
        mean[c] = running_mean[c];
        invstd[c] = static_cast<float>(1.0 / sqrt(static_cast<double>(running_var[c]) + eps));
        return;
    }
    const double cnt = sums[2 * n];
    const double m = sums[c] / cnt;
    double var = sums[n + c] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = static_cast<float>(m);
    invstd[c] = static_cast<float>(1.0 / sqrt(var + eps));
    if (running_mean) running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * m);
    if (running_var) {
        const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// out = act(gamma * (z - mean) * invstd + beta)
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_act_fwd_kernel(int64_t rows, int n, const T *__restrict__ z, int64_t ldz, const float *__restrict__ mean,
                  const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                  int act, const float *__restrict__ alpha, float p0, float p1, T *__restrict__ out, int64_t ldo) {
    const float a = alpha ? __ldg(alpha) : 0.f;
    const int ppr = n / V;
    const int64_t pieces = rows * ppr;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < pieces;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t r = i / ppr;
        const int c = static_cast<int>(i - r * ppr) * V;
        float v[4];
        load_piece<T, V>(z, ldz, r, c, v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float sc = __ldg(invstd + c + j) * (gamma ? __ldg(gamma + c + j) : 1.f);
            const float y = (v[j] - __ldg(mean + c + j)) * sc + (beta ? __ldg(beta + c + j) : 0.f);
            v[j] = actx_fwd(y, act, a, p0, p1);
        }
        store_piece<T, V>(out, ldo, r, c, v);
    }
}

// Backward reductions: per column  sum dy  and  sum dy * xhat  with  dy = g * act'(y),  y recomputed from z; and the
// scalar  sum g * min(y, 0)  (PReLU slope).  Same thread layout as bn_stats_kernel; partial[cta][k][c], k < 2, followed by
// one double per CTA for the slope at partial[ctas * 2 * n + cta].
template <typename T>
__global__ void __launch_bounds__(32 * kBnRowsY)
bn_bwd_reduce_kernel(int64_t rows, int n, const T *__restrict__ g, int64_t ldg, const T *__restrict__ z, int64_t ldz,
                     const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                     const float *__restrict__ beta, int act, const float *__restrict__ alpha, float p0, float p1,
                     double *__restrict__ partial, double *__restrict__ partial_alpha) {
    __shared__ double red[kBnRowsY][32];
    const float a = alpha ? __ldg(alpha) : 0.f;
    double s[kBnCols], ss[kBnCols], da = 0.0;
    float mu[kBnCols], is[kBnCols], ga[kBnCols], be[kBnCols];
#pragma unroll
    for (int j = 0; j < kBnCols; ++j) {
        const int c = threadIdx.x + 32 * j;
        s[j] = ss[j] = 0.0;
        mu[j] = c < n ? __ldg(mean + c) : 0.f;
        is[j] = c < n ? __ldg(invstd + c) : 0.f;
        ga[j] = (c < n && gamma) ? __ldg(gamma + c) : 1.f;
        be[j] = (c < n && beta) ? __ldg(beta + c) : 0.f;
    }
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rows) ? r0 + per : rows;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += kBnRowsY) {
#pragma unroll
        for (int j = 0; j < kBnCols; ++j) {
            const int c = threadIdx.x + 32 * j;
            if (c < n) {
                const float xh = (ld1<T>(z + r * ldz + c) - mu[j]) * is[j];
                const float y = xh * ga[j] + be[j];
                const float gv = ld1<T>(g + r * ldg + c);
                const float dy = gv * actx_grad(y, act, a, p0, p1);
                da += static_cast<double>(gv * fminf(y, 0.f));
                s[j] += static_cast<double>(dy);
                ss[j] += static_cast<double>(dy * xh);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < kBnCols; ++j) {
            const int c = threadIdx.x + 32 * j;
            red[threadIdx.y][threadIdx.x] = k == 0 ? s[j] : ss[j];
            __syncthreads();
            if (threadIdx.y == 0 && c < n) {
                double t = 0.0;
                for (int y = 0; y < kBnRowsY; ++y) t += red[y][threadIdx.x];
                partial[(static_cast<int64_t>(blockIdx.x) * 2 + k) * n + c] = t;
            }
            __syncthreads();
        }
    }
    red[threadIdx.y][threadIdx.x] = da;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && partial_alpha != nullptr) {
        double t = 0.0;
        for (int y = 0; y < kBnRowsY; ++y)
            for (int x = 0; x < 32; ++x) t += red[y][x];
        partial_alpha[blockIdx.x] = t;
    }
}

__global__ void bn_reduce_alpha_kernel(const double *__restrict__ partial_alpha, int ctas, double *sums, int n, int accumulate) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double t = accumulate ? sums[2 * n] : 0.0;
        for (int b = 0; b < ctas; ++b) t += partial_alpha[b];
        sums[2 * n] = t;
    }
}

// dz = gamma * invstd * (dy - sum_dy / N - xhat * sum_dy_xhat / N)   (training);   dz = gamma * invstd * dy   (eval)
// CTA 0 also writes dgamma = sum dy * xhat, dbeta = sum dy, dalpha = sums[2n] as fp32.
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(int64_t rows, int n, const T *__restrict__ g, int64_t ldg, const T *__restrict__ z, int64_t ldz,
                    const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                    const float *__restrict__ beta, int act, const float *__restrict__ alpha, float p0, float p1,
                    const double *__restrict__ sums, double count, int training, T *__restrict__ dz, int64_t lddz,
                    float *dgamma, float *dbeta, float *dalpha) {
    const float a = alpha ? __ldg(alpha) : 0.f;
    const float inv_n = training ? static_cast<float>(1.0 / count) : 0.f;
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < n; c += kThreads) {
            if (dbeta) dbeta[c] = static_cast<float>(sums[c]);
            if (dgamma) dgamma[c] = static_cast<float>(sums[n + c]);
        }
        if (threadIdx.x == 0 && dalpha) dalpha[0] = static_cast<float>(sums[2 * n]);
    }
    const int ppr = n / V;
    const int64_t pieces = rows * ppr;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < pieces;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
        const int64_t r = i / ppr;
        const int c = static_cast<int>(i - r * ppr) * V;
        float gv[4], zv[4];
        load_piece<T, V>(g, ldg, r, c, gv);
        load_piece<T, V>(z, ldz, r, c, zv);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float is = __ldg(invstd + c + j), ga = gamma ? __ldg(gamma + c + j) : 1.f;
            const float xh = (zv[j] - __ldg(mean + c + j)) * is;
            const float y = xh * ga + (beta ? __ldg(beta + c + j) : 0.f);
            const float dy = gv[j] * actx_grad(y, act, a, p0, p1);
            const float m1 = static_cast<float>(sums[c + j]) * inv_n, m2 = static_cast<float>(sums[n + c + j]) * inv_n;
            gv[j] = ga * is * (dy - m1 - xh * m2);
        }
        store_piece<T, V>(dz, lddz, r, c, gv);
    }
}

// ---- per-graph pools of the raw path features (global_feats) -----------------------------------------
// thread per (graph, column): rows of the segment are visited in CSR order and added left to right
// (= zeros().scatter_add_ on the CPU), mean = sum / max(count, 1); max over the same rows, 0 for an empty segment.
__global__ void segment_pool_kernel(int64_t segments, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                    const float *__restrict__ x, int64_t ldx, int f, float *__restrict__ mean_out,
                                    float *__restrict__ max_out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= segments * f) return;
    const int64_t s = i / f;
    const int c = static_cast<int>(i - s * f);
    const int e0 = __ldg(rowptr + s), e1 = __ldg(rowptr + s + 1);
    float sum = 0.f, mx = -INFINITY;
#pragma unroll 8
    for (int e = e0; e < e1; ++e) {
        const int64_t r = col ? __ldg(col + e) : e;
        const float v = __ldg(x + r * ldx + c);
        sum = __fadd_rn(sum, v);
        mx = fmaxf(mx, v);
    }
    const int cnt = e1 - e0;
    mean_out[s * f + c] = __fdiv_rn(sum, static_cast<float>(cnt > 1 ? cnt : 1));
    max_out[s * f + c] = cnt > 0 ? mx : 0.f;
}

// tail[p] = [ origin[p][0:f_origin] | mean[seg[p]] | max[seg[p]] ]  — the constant columns of the readout input
// (models.py:364-371); seg ids are int64 (PyG's batch vector) or int32.
template <typename I>
__global__ void readout_tail_kernel(int64_t rows, const I *__restrict__ seg, int64_t segments, const float *__restrict__ origin,
                                    int64_t ld_origin, int f_origin, const float *__restrict__ mean_in,
                                    const float *__restrict__ max_in, int f, float *__restrict__ tail, int64_t ld_tail) {
    const int w = f_origin + 2 * f;
    const int64_t total = rows * w;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / w;
        const int c = static_cast<int>(i - r * w);
        float v;
        if (c < f_origin) {
            v = __ldg(origin + r * ld_origin + c);
        } else {
            int64_t s = static_cast<int64_t>(seg[r]);
            s = s < 0 ? 0 : (s >= segments ? segments - 1 : s);
            v = c < f_origin + f ? __ldg(mean_in + s * f + (c - f_origin)) : __ldg(max_in + s * f + (c - f_origin - f));
        }
        tail[r * ld_tail + c] = v;
    }
}

inline bool act_known(int act) { return act >= HGIN_ACT_NONE && act <= HGIN_ACT_SOFTPLUS; }

template <typename T>
int32_t act_fwd_impl(int64_t rows, int n, const T *z, int64_t ldz, int act, const float *alpha, float p0, float p1, T *out,
                     int64_t ldo, cudaStream_t s) {
    if (rows == 0) return HGIN_OK;
    if (vec4_ok<T>(n, {z, out}, {ldz, ldo}))
        act_fwd_kernel<T, 4><<<piece_ctas(rows * (n / 4)), kThreads, 0, s>>>(rows, n, z, ldz, act, alpha, p0, p1, out, ldo);
    else
        act_fwd_kernel<T, 1><<<piece_ctas(rows * n), kThreads, 0, s>>>(rows, n, z, ldz, act, alpha, p0, p1, out, ldo);
    HGIN_CHECK_LAUNCH("hgin_act_fwd");
    return HGIN_OK;
}

template <typename T>
int32_t act_bwd_impl(int64_t rows, int n, const T *g, int64_t ldg, const T *z, int64_t ldz, int act, const float *alpha,
                     float p0, float p1, T *dz, int64_t lddz, float *dalpha, float *ws, cudaStream_t s) {
    if (rows == 0) {
        if (dalpha) cudaMemsetAsync(dalpha, 0, sizeof(float), s);
        return HGIN_OK;
    }
    float *partial = dalpha ? ws : nullptr;
    int ctas;
    if (vec4_ok<T>(n, {g, z, dz}, {ldg, ldz, lddz})) {
        ctas = piece_ctas(rows * (n / 4));
        act_bwd_kernel<T, 4><<<ctas, kThreads, 0, s>>>(rows, n, g, ldg, z, ldz, act, alpha, p0, p1, dz, lddz, partial);
    } else {
        ctas = piece_ctas(rows * n);
        act_bwd_kernel<T, 1><<<ctas, kThreads, 0, s>>>(rows, n, g, ldg, z, ldz, act, alpha, p0, p1, dz, lddz, partial);
    }
    if (dalpha) sum_partials_kernel<<<1, 1024, 0, s>>>(partial, ctas, dalpha);
    HGIN_CHECK_LAUNCH("hgin_act_bwd");
    return HGIN_OK;
}

inline int bn_ctas(int64_t rows) {
    const int64_t want = ceil_div(rows, 64);
    const int64_t cap = kNumSMs * 4;
    return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace
}  // namespace hgin

using namespace hgin;

#define HGIN_DISPATCH_DTYPE(dtype, CALL_F32, CALL_BF16)                                        \
    do {                                                                                       \
        if ((dtype) == HGIN_DTYPE_F32) return CALL_F32;                                        \
        if ((dtype) == HGIN_DTYPE_BF16) return CALL_BF16;                                      \
        return fail(HGIN_ERR_INVALID_ARGUMENT, "unknown dtype %d", static_cast<int>(dtype));   \
    } while (0)

extern "C" int64_t hgin_elementwise_workspace_bytes(void) { return static_cast<int64_t>(kMaxCtas) * sizeof(float); }

extern "C" int32_t hgin_act_fwd(int32_t dtype, int64_t rows, int32_t n, const void *z, int64_t ldz, int32_t act,
                                const float *alpha, float p0, float p1, void *out, int64_t ldo, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && n > 0 && ldz >= n && ldo >= n, "hgin_act_fwd: bad shape rows=%lld n=%d", (long long)rows, n);
    HGIN_CHECK_ARG(rows == 0 || (z && out), "hgin_act_fwd: null matrix");
    HGIN_CHECK_ARG(act_known(act) && (act != HGIN_ACT_PRELU || alpha), "hgin_act_fwd: activation %d (PReLU needs its slope)", act);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HGIN_DISPATCH_DTYPE(dtype,
        act_fwd_impl<float>(rows, n, static_cast<const float *>(z), ldz, act, alpha, p0, p1, static_cast<float *>(out), ldo, s),
        act_fwd_impl<__nv_bfloat16>(rows, n, static_cast<const __nv_bfloat16 *>(z), ldz, act, alpha, p0, p1,
                                    static_cast<__nv_bfloat16 *>(out), ldo, s));
}

extern "C" int32_t hgin_act_bwd(int32_t dtype, int64_t rows, int32_t n, const void *g, int64_t ldg, const void *z, int64_t ldz,
                                int32_t act, const float *alpha, float p0, float p1, void *dz, int64_t lddz, float *dalpha,
                                void *workspace, int64_t workspace_bytes, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && n > 0 && ldg >= n && ldz >= n && lddz >= n, "hgin_act_bwd: bad shape");
    HGIN_CHECK_ARG(rows == 0 || (g && z && dz), "hgin_act_bwd: null matrix");
    HGIN_CHECK_ARG(act_known(act) && (act != HGIN_ACT_PRELU || alpha), "hgin_act_bwd: activation %d (PReLU needs its slope)", act);
    if (dalpha && workspace_bytes < hgin_elementwise_workspace_bytes())
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_act_bwd: workspace %lld < %lld", (long long)workspace_bytes,
                    (long long)hgin_elementwise_workspace_bytes());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *ws = static_cast<float *>(workspace);
    HGIN_DISPATCH_DTYPE(dtype,
        act_bwd_impl<float>(rows, n, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, act, alpha, p0, p1,
                            static_cast<float *>(dz), lddz, dalpha, ws, s),
        act_bwd_impl<__nv_bfloat16>(rows, n, static_cast<const __nv_bfloat16 *>(g), ldg, static_cast<const __nv_bfloat16 *>(z),
                                    ldz, act, alpha, p0, p1, static_cast<__nv_bfloat16 *>(dz), lddz, dalpha, ws, s));
}

namespace {
template <typename T>
int32_t dropout_impl(int64_t rows, int n, const T *x, int64_t ldx, float p, uint64_t seed, uint64_t offset, T *out, int64_t ldo,
                     cudaStream_t s) {
    if (rows == 0) return HGIN_OK;
    const float scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
    if (vec4_ok<T>(n, {x, out}, {ldx, ldo}))
        dropout_kernel<T, 4><<<piece_ctas(rows * (n / 4)), kThreads, 0, s>>>(rows, n, x, ldx, p, scale, seed, offset, out, ldo);
    else
        dropout_kernel<T, 1><<<piece_ctas(rows * n), kThreads, 0, s>>>(rows, n, x, ldx, p, scale, seed, offset, out, ldo);
    HGIN_CHECK_LAUNCH("hgin_dropout");
    return HGIN_OK;
}
}  // namespace

extern "C" int32_t hgin_dropout(int32_t dtype, int64_t rows, int32_t n, const void *x, int64_t ldx, float p, uint64_t seed,
                                uint64_t offset, void *out, int64_t ldo, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && n > 0 && ldx >= n && ldo >= n, "hgin_dropout: bad shape");
    HGIN_CHECK_ARG(p >= 0.f && p <= 1.f, "hgin_dropout: p = %f outside [0, 1]", static_cast<double>(p));
    HGIN_CHECK_ARG(rows == 0 || (x && out), "hgin_dropout: null matrix");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HGIN_DISPATCH_DTYPE(dtype,
        dropout_impl<float>(rows, n, static_cast<const float *>(x), ldx, p, seed, offset, static_cast<float *>(out), ldo, s),
        dropout_impl<__nv_bfloat16>(rows, n, static_cast<const __nv_bfloat16 *>(x), ldx, p, seed, offset,
                                    static_cast<__nv_bfloat16 *>(out), ldo, s));
}

extern "C" int64_t hgin_bn_workspace_bytes(int64_t rows, int32_t n) {
    const int64_t ctas = bn_ctas(rows);
    const int nc = n < 32 * kBnCols ? n : 32 * kBnCols;
    return align_up(ctas * 2 * nc * static_cast<int64_t>(sizeof(double)), 256) + align_up(ctas * sizeof(double), 256);
}

namespace {
template <typename T>
int32_t bn_stats_impl(int64_t rows, int n, const T *z, int64_t ldz, double *sums, double *ws, cudaStream_t s) {
    const int ctas = bn_ctas(rows);
    for (int c0 = 0; c0 < n; c0 += 32 * kBnCols) {
        const int nc = (n - c0) < 32 * kBnCols ? (n - c0) : 32 * kBnCols;
        bn_stats_kernel<T><<<ctas, dim3(32, kBnRowsY), 0, s>>>(rows, nc, z + c0, ldz, ws);
        bn_reduce_kernel<<<static_cast<int>(ceil_div(2 * nc, 128)), 128, 0, s>>>(ws, ctas, nc, 2, sums, n, c0);
    }
    bn_set_count_kernel<<<1, 1, 0, s>>>(sums, n, static_cast<double>(rows));
    HGIN_CHECK_LAUNCH("hgin_bn_stats");
    return HGIN_OK;
}
}  // namespace

extern "C" int32_t hgin_bn_stats(int32_t dtype, int64_t rows, int32_t n, const void *z, int64_t ldz, double *sums,
                                 void *workspace, int64_t workspace_bytes, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && n > 0 && ldz >= n && sums, "hgin_bn_stats: bad arguments");
    HGIN_CHECK_ARG(rows == 0 || z, "hgin_bn_stats: null matrix");
    if (workspace_bytes < hgin_bn_workspace_bytes(rows, n))
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_bn_stats: workspace %lld < %lld", (long long)workspace_bytes,
                    (long long)hgin_bn_workspace_bytes(rows, n));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double *ws = static_cast<double *>(workspace);
    HGIN_DISPATCH_DTYPE(dtype, bn_stats_impl<float>(rows, n, static_cast<const float *>(z), ldz, sums, ws, s),
                        bn_stats_impl<__nv_bfloat16>(rows, n, static_cast<const __nv_bfloat16 *>(z), ldz, sums, ws, s));
}

extern "C" int32_t hgin_bn_finalize(int32_t n, const double *sums, double eps, double momentum, int32_t use_running,
                                    float *mean, float *invstd, float *running_mean, float *running_var, void *stream) {
    HGIN_CHECK_ARG(n > 0 && mean && invstd, "hgin_bn_finalize: bad arguments");
    HGIN_CHECK_ARG(use_running ? (running_mean && running_var) : sums != nullptr,
                   "hgin_bn_finalize: %s", use_running ? "eval mode needs the running statistics" : "null sums");
    bn_finalize_kernel<<<static_cast<int>(ceil_div(n, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        n, sums, eps, momentum, use_running, mean, invstd, running_mean, running_var);
    HGIN_CHECK_LAUNCH("hgin_bn_finalize");
    return HGIN_OK;
}

namespace {
template <typename T>
int32_t bn_act_fwd_impl(int64_t rows, int n, const T *z, int64_t ldz, const float *mean, const float *invstd, const float *gamma,
                        const float *beta, int act, const float *alpha, float p0, float p1, T *out, int64_t ldo, cudaStream_t s) {
    if (rows == 0) return HGIN_OK;
    if (vec4_ok<T>(n, {z, out}, {ldz, ldo}))
        bn_act_fwd_kernel<T, 4><<<piece_ctas(rows * (n / 4)), kThreads, 0, s>>>(rows, n, z, ldz, mean, invstd, gamma, beta, act,
                                                                                 alpha, p0, p1, out, ldo);
    else
        bn_act_fwd_kernel<T, 1><<<piece_ctas(rows * n), kThreads, 0, s>>>(rows, n, z, ldz, mean, invstd, gamma, beta, act, alpha,
                                                                           p0, p1, out, ldo);
    HGIN_CHECK_LAUNCH("hgin_bn_act_fwd");
    return HGIN_OK;
}

template <typename T>
int32_t bn_bwd_reduce_impl(int64_t rows, int n, const T *g, int64_t ldg, const T *z, int64_t ldz, const float *mean,
                           const float *invstd, const float *gamma, const float *beta, int act, const float *alpha, float p0,
                           float p1, double *sums, double *ws, cudaStream_t s) {
    const int ctas = bn_ctas(rows);
    const int ncap = n < 32 * kBnCols ? n : 32 * kBnCols;
    double *pa = ws + align_up(static_cast<int64_t>(ctas) * 2 * ncap * sizeof(double), 256) / sizeof(double);
    for (int c0 = 0; c0 < n; c0 += 32 * kBnCols) {
        const int nc = (n - c0) < 32 * kBnCols ? (n - c0) : 32 * kBnCols;
        bn_bwd_reduce_kernel<T><<<ctas, dim3(32, kBnRowsY), 0, s>>>(rows, nc, g + c0, ldg, z + c0, ldz, mean + c0, invstd + c0,
                                                                     gamma ? gamma + c0 : nullptr, beta ? beta + c0 : nullptr,
                                                                     act, alpha, p0, p1, ws, pa);
        bn_reduce_kernel<<<static_cast<int>(ceil_div(2 * nc, 128)), 128, 0, s>>>(ws, ctas, nc, 2, sums, n, c0);
        bn_reduce_alpha_kernel<<<1, 32, 0, s>>>(pa, ctas, sums, n, c0 > 0 ? 1 : 0);
    }
    HGIN_CHECK_LAUNCH("hgin_bn_act_bwd_reduce");
    return HGIN_OK;
}

template <typename T>
int32_t bn_bwd_apply_impl(int64_t rows, int n, const T *g, int64_t ldg, const T *z, int64_t ldz, const float *mean,
                          const float *invstd, const float *gamma, const float *beta, int act, const float *alpha, float p0,
                          float p1, const double *sums, double count, int training, T *dz, int64_t lddz, float *dgamma,
                          float *dbeta, float *dalpha, cudaStream_t s) {
    const bool v4 = vec4_ok<T>(n, {g, z, dz}, {ldg, ldz, lddz});
    const int ctas = piece_ctas(rows > 0 ? rows * (v4 ? n / 4 : n) : 1);
    if (v4)
        bn_bwd_apply_kernel<T, 4><<<ctas, kThreads, 0, s>>>(rows, n, g, ldg, z, ldz, mean, invstd, gamma, beta, act, alpha, p0, p1,
                                                           sums, count, training, dz, lddz, dgamma, dbeta, dalpha);
    else
        bn_bwd_apply_kernel<T, 1><<<ctas, kThreads, 0, s>>>(rows, n, g, ldg, z, ldz, mean, invstd, gamma, beta, act, alpha, p0, p1,
                                                           sums, count, training, dz, lddz, dgamma, dbeta, dalpha);
    HGIN_CHECK_LAUNCH("hgin_bn_act_bwd_apply");
    return HGIN_OK;
}
}  // namespace

extern "C" int32_t hgin_bn_act_fwd(int32_t dtype, int64_t rows, int32_t n, const void *z, int64_t ldz, const float *mean,
                                   const float *invstd, const float *gamma, const float *beta, int32_t act, const float *alpha,
                                   float p0, float p1, void *out, int64_t ldo, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && n > 0 && ldz >= n && ldo >= n && mean && invstd, "hgin_bn_act_fwd: bad arguments");
    HGIN_CHECK_ARG(rows == 0 || (z && out), "hgin_bn_act_fwd: null matrix");
    HGIN_CHECK_ARG(act_known(act) && (act != HGIN_ACT_PRELU || alpha), "hgin_bn_act_fwd: activation %d", act);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HGIN_DISPATCH_DTYPE(dtype,
        bn_act_fwd_impl<float>(rows, n, static_cast<const float *>(z), ldz, mean, invstd, gamma, beta, act, alpha, p0, p1,
                               static_cast<float *>(out), ldo, s),
        bn_act_fwd_impl<__nv_bfloat16>(rows, n, static_cast<const __nv_bfloat16 *>(z), ldz, mean, invstd, gamma, beta, act, alpha,
                                       p0, p1, static_cast<__nv_bfloat16 *>(out), ldo, s));
}

extern "C" int32_t hgin_bn_act_bwd_reduce(int32_t dtype, int64_t rows, int32_t n, const void *g, int64_t ldg, const void *z,
                                          int64_t ldz, const float *mean, const float *invstd, const float *gamma,
                                          const float *beta, int32_t act, const float *alpha, float p0, float p1, double *sums,
                                          void *workspace, int64_t workspace_bytes, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && n > 0 && ldg >= n && ldz >= n && mean && invstd && sums, "hgin_bn_act_bwd_reduce: bad arguments");
    HGIN_CHECK_ARG(rows == 0 || (g && z), "hgin_bn_act_bwd_reduce: null matrix");
    HGIN_CHECK_ARG(act_known(act) && (act != HGIN_ACT_PRELU || alpha), "hgin_bn_act_bwd_reduce: activation %d", act);
    if (workspace_bytes < hgin_bn_workspace_bytes(rows, n))
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_bn_act_bwd_reduce: workspace %lld < %lld", (long long)workspace_bytes,
                    (long long)hgin_bn_workspace_bytes(rows, n));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double *ws = static_cast<double *>(workspace);
    HGIN_DISPATCH_DTYPE(dtype,
        bn_bwd_reduce_impl<float>(rows, n, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, mean, invstd,
                                  gamma, beta, act, alpha, p0, p1, sums, ws, s),
        bn_bwd_reduce_impl<__nv_bfloat16>(rows, n, static_cast<const __nv_bfloat16 *>(g), ldg,
                                          static_cast<const __nv_bfloat16 *>(z), ldz, mean, invstd, gamma, beta, act, alpha, p0,
                                          p1, sums, ws, s));
}

extern "C" int32_t hgin_bn_act_bwd_apply(int32_t dtype, int64_t rows, int32_t n, const void *g, int64_t ldg, const void *z,
                                         int64_t ldz, const float *mean, const float *invstd, const float *gamma,
                                         const float *beta, int32_t act, const float *alpha, float p0, float p1,
                                         const double *sums, double count, int32_t training, void *dz, int64_t lddz,
                                         float *dgamma, float *dbeta, float *dalpha, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && n > 0 && ldg >= n && ldz >= n && lddz >= n && mean && invstd && sums,
                   "hgin_bn_act_bwd_apply: bad arguments");
    HGIN_CHECK_ARG(rows == 0 || (g && z && dz), "hgin_bn_act_bwd_apply: null matrix");
    HGIN_CHECK_ARG(!training || count >= 1.0, "hgin_bn_act_bwd_apply: count %f", count);
    HGIN_CHECK_ARG(act_known(act) && (act != HGIN_ACT_PRELU || alpha), "hgin_bn_act_bwd_apply: activation %d", act);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HGIN_DISPATCH_DTYPE(dtype,
        bn_bwd_apply_impl<float>(rows, n, static_cast<const float *>(g), ldg, static_cast<const float *>(z), ldz, mean, invstd,
                                 gamma, beta, act, alpha, p0, p1, sums, count, training, static_cast<float *>(dz), lddz, dgamma,
                                 dbeta, dalpha, s),
        bn_bwd_apply_impl<__nv_bfloat16>(rows, n, static_cast<const __nv_bfloat16 *>(g), ldg,
                                         static_cast<const __nv_bfloat16 *>(z), ldz, mean, invstd, gamma, beta, act, alpha, p0, p1,
                                         sums, count, training, static_cast<__nv_bfloat16 *>(dz), lddz, dgamma, dbeta, dalpha, s));
}

extern "C" int32_t hgin_segment_pool(int64_t segments, const int32_t *rowptr, const int32_t *col, const float *x, int64_t ldx,
                                     int32_t f, float *mean_out, float *max_out, void *stream) {
    HGIN_CHECK_ARG(segments >= 0 && f > 0 && ldx >= f, "hgin_segment_pool: bad shape");
    HGIN_CHECK_ARG(segments == 0 || (rowptr && x && mean_out && max_out), "hgin_segment_pool: null pointer");
    if (segments == 0) return HGIN_OK;
    segment_pool_kernel<<<static_cast<int>(ceil_div(segments * f, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        segments, rowptr, col, x, ldx, f, mean_out, max_out);
    HGIN_CHECK_LAUNCH("hgin_segment_pool");
    return HGIN_OK;
}

extern "C" int32_t hgin_readout_tail(int64_t rows, const void *segment_ids, int32_t index_bytes, int64_t segments,
                                     const float *origin, int64_t ld_origin, int32_t f_origin, const float *mean_in,
                                     const float *max_in, int32_t f, float *tail, int64_t ld_tail, void *stream) {
    HGIN_CHECK_ARG(rows >= 0 && f > 0 && f_origin >= 0 && ld_tail >= f_origin + 2 * f && segments >= 1,
                   "hgin_readout_tail: bad shape");
    HGIN_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "hgin_readout_tail: index_bytes must be 4 or 8");
    HGIN_CHECK_ARG(f_origin == 0 || (origin && ld_origin >= f_origin), "hgin_readout_tail: origin");
    HGIN_CHECK_ARG(rows == 0 || (segment_ids && mean_in && max_in && tail), "hgin_readout_tail: null pointer");
    if (rows == 0) return HGIN_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ctas = grid_for(rows * (f_origin + 2 * f), 256 * 4, 8);
    if (index_bytes == 8)
        readout_tail_kernel<int64_t><<<ctas, 256, 0, s>>>(rows, static_cast<const int64_t *>(segment_ids), segments, origin,
                                                          ld_origin, f_origin, mean_in, max_in, f, tail, ld_tail);
    else
        readout_tail_kernel<int32_t><<<ctas, 256, 0, s>>>(rows, static_cast<const int32_t *>(segment_ids), segments, origin,
                                                          ld_origin, f_origin, mean_in, max_in, f, tail, ld_tail);
    HGIN_CHECK_LAUNCH("hgin_readout_tail");
    return HGIN_OK;
}
