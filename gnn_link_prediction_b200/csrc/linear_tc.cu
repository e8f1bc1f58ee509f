// Host side of the tensor-core dense-layer path (HGIN_MATH_TF32): TMA tensor maps, the small
// helper kernels (dz = g * act'(z) with fused db / dalpha / tail-dW reductions, weight repacking)
// and the dispatch used by hgin_linear_fwd / hgin_linear_bwd.  Kernels: linear_tc.cuh.
#include <stdlib.h>

#include "linear_tc.cuh"
#include "linear_tc_fused.cuh"
#include "linear_tc_dw.cuh"
#include "tail_sums.cuh"

namespace hgin {
namespace tcgemm {

// ---- TMA tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// fp32 matrix [outer x inner], row pitch ld elements, box [box_outer x box_inner].
static bool make_map(CUtensorMap *m, const float *base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                     int box_outer, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool tma_ok(const float *p, int64_t ld) { return p && aligned16(p) && ld % 4 == 0; }

// ---- helper kernels -------------------------------------------------------------------------------
// dst[r][c] = src[r * ld + c0 + c]   (repack W[:, c0:c0+cols] to a dense, 16B-aligned matrix)
__global__ void __launch_bounds__(256) pack_cols_kernel(const float *__restrict__ src, int rows, int ld, int c0,
                                                        int cols, float *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x)
        dst[i] = __ldg(src + static_cast<int64_t>(i / cols) * ld + c0 + i % cols);
}
// dst[c][r] = src[r * ld + c0 + c]   (W^T restricted to columns [c0, c0+cols): K-major B of the dx GEMM)
__global__ void __launch_bounds__(256) transpose_cols_kernel(const float *__restrict__ src, int rows, int ld, int c0,
                                                             int cols, float *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
        const int c = i / rows, r = i % rows;
        dst[i] = __ldg(src + static_cast<int64_t>(r) * ld + c0 + c);
    }
}

// dz = g * act'(z) written densely [rows x n]; per-CTA partials of
//   db[nn] = sum_m dz,  tail[nn][t] = sum_m dz * x2[m][t] (t < k2 <= 4),  dalpha = sum g * min(z, 0).
// Partial layout: part[cta][nn][k2 + 1] with the LAST column = db (same convention as the SIMT
// weight-gradient partials), alpha_part[cta].
constexpr int DZ_THREADS = 256;
__global__ void __launch_bounds__(DZ_THREADS, 2)
dz_prepare_kernel(int64_t rows, int n, const float *__restrict__ g, int64_t ldg, const float *__restrict__ z,
                  int64_t ldz, int act, const float *__restrict__ alpha_ptr, const float *__restrict__ x2,
                  int64_t ld2, int k2, float *__restrict__ dz, float *__restrict__ part,
                  float *__restrict__ alpha_part, int want_sums, int write_dz) {
    extern __shared__ float sm[];  // [slots][n][5] for the cross-slot combine
    __shared__ float red[32];
    const int tpr = n / 4;                       // threads per row (float4 each)
    const int slots = DZ_THREADS / tpr;          // rows processed per iteration
    const int slot = threadIdx.x / tpr;
    const int cg = threadIdx.x % tpr;            // column group -> columns 4cg .. 4cg+3
    const bool active = slot < slots;
    const float alpha = (act == HGIN_ACT_PRELU) ? __ldg(alpha_ptr) : 0.0f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    float tail[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t) tail[i][t] = 0.f;
    float dalpha = 0.f;
    if (active) {
        // RIF rows per thread and iteration with every g / z vector requested before the first use
        // (one row per thread kept ~16 KB of reads in flight per SM; same per-thread row order)
        constexpr int RIF = 4;
        const int64_t stride = static_cast<int64_t>(gridDim.x) * slots;
        for (int64_t m0 = static_cast<int64_t>(blockIdx.x) * slots + slot; m0 < rows; m0 += stride * RIF) {
            float4 gq[RIF], zq[RIF];
            float xq[RIF][4];   // the rank-k2 tail's input columns, requested with the rows (not after them)
#pragma unroll
            for (int u = 0; u < RIF; ++u) {
                const int64_t m = m0 + u * stride;
                gq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                zq[u] = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
                for (int t = 0; t < 4; ++t) xq[u][t] = 0.f;
                if (m < rows) {
                    gq[u] = __ldg(reinterpret_cast<const float4 *>(g + m * ldg) + cg);
                    if (act != HGIN_ACT_NONE) zq[u] = __ldg(reinterpret_cast<const float4 *>(z + m * ldz) + cg);
                    if (want_sums) {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (t < k2) xq[u][t] = __ldg(x2 + m * ld2 + t);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RIF; ++u) {
                const int64_t m = m0 + u * stride;
                if (m >= rows) break;
                float d[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
                if (act != HGIN_ACT_NONE) {
                    const float zv[4] = {zq[u].x, zq[u].y, zq[u].z, zq[u].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (act == HGIN_ACT_PRELU && !(zv[i] > 0.f)) dalpha += d[i] * zv[i];
                        d[i] = act_backward(d[i], zv[i], act, alpha);
                    }
                }
                if (write_dz) reinterpret_cast<float4 *>(dz + m * n)[cg] = make_float4(d[0], d[1], d[2], d[3]);
                if (want_sums) {
                    const float xv[4] = {xq[u][0], xq[u][1], xq[u][2], xq[u][3]};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        db[i] += d[i];
#pragma unroll
                        for (int t = 0; t < 4; ++t) tail[i][t] = fmaf(d[i], xv[t], tail[i][t]);
                    }
                }
            }
        }
    }
    if (!want_sums) return;
    // combine the row slots of this CTA in a fixed order
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float *dst = sm + (static_cast<int64_t>(slot) * n + cg * 4 + i) * 5;
#pragma unroll
            for (int t = 0; t < 4; ++t) dst[t] = tail[i][t];
            dst[4] = db[i];
        }
    }
    __syncthreads();
    const int kp = k2 + 1;
    for (int i = threadIdx.x; i < n * kp; i += DZ_THREADS) {
        const int nn = i / kp, t = i % kp;
        const int src_t = (t == k2) ? 4 : t;
        float s = 0.f;
        for (int sl = 0; sl < slots; ++sl) s += sm[(static_cast<int64_t>(sl) * n + nn) * 5 + src_t];
        part[(static_cast<int64_t>(blockIdx.x) * n + nn) * kp + t] = s;
    }
    dalpha = block_sum(dalpha, red);
    if (threadIdx.x == 0 && alpha_part) alpha_part[blockIdx.x] = dalpha;
}

// out[nn][col0 + k] (ld = ldw) = sum_p part[p][nn][k] for k < kcols; the optional extra column
// (index kcols) goes to db.
static inline int reduce_grid(int64_t total) {
    const int64_t want = (total + 31) / 32;
    return static_cast<int>(want < 1 ? 1 : (want > kNumSMs * 8 ? kNumSMs * 8 : want));
}

// One CTA per 32 consecutive elements (grid-stride over such groups): warp w adds the partials w, w + 8, ... with the lanes on
// consecutive elements (coalesced, the loads of a warp independent), the eight sums are combined in a fixed order.  (A thread
// per element walking all ~148 partials was a chain of ~37 dependent steps: 14.5 us per call, 22 calls per Cfg-C step.)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float *__restrict__ part, int num_part, int n, int kcols, int has_db_col,
                       float *__restrict__ dW, int ldw, int col0, float *__restrict__ db) {
    __shared__ float red[8][32];
    const int kp = kcols + (has_db_col ? 1 : 0);
    const int64_t total = static_cast<int64_t>(n) * kp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t g0 = static_cast<int64_t>(blockIdx.x) * 32; g0 < total; g0 += static_cast<int64_t>(gridDim.x) * 32) {
        const int64_t i = g0 + lane;
        float s0 = 0.f, s1 = 0.f;
        if (i < total) {
            int pp = warp;
            for (; pp + 8 < num_part; pp += 16) {      // two independent chains per warp
                s0 += part[static_cast<int64_t>(pp) * total + i];
                s1 += part[static_cast<int64_t>(pp + 8) * total + i];
            }
            if (pp < num_part) s0 += part[static_cast<int64_t>(pp) * total + i];
        }
        red[warp][lane] = s0 + s1;
        __syncthreads();
        if (warp == 0 && i < total) {
            const float s = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) +
                            ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
            const int nn = static_cast<int>(i / kp), k = static_cast<int>(i % kp);
            if (has_db_col && k == kcols) {
                if (db) db[nn] = s;
            } else if (dW) {
                dW[static_cast<int64_t>(nn) * ldw + col0 + k] = s;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) reduce_scalar_tc_kernel(const float *__restrict__ v, int count,
                                                                float *__restrict__ out) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += v[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

// ---- eligibility ----------------------------------------------------------------------------------
bool fwd_eligible(int64_t rows, const float *x1, int64_t ld1, int k1, int k2, int n, const float *z, int64_t ldz,
                  const float *out, int64_t ldo) {
    return rows >= BM && k1 >= 16 && k1 <= 128 && k1 % 4 == 0 && k2 <= 4 && n >= 16 && n <= 128 && n % 16 == 0 &&
           tma_ok(x1, ld1) && (!z || tma_ok(z, ldz)) && (!out || tma_ok(out, ldo)) && encode_fn() != nullptr;
}

bool bwd_eligible(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act, const float *x1,
                  int64_t ld1, int k1, int k2, int n, int c0, int c1, const float *dx, int64_t lddx,
                  const float *dot_x, int64_t ld_dot) {
    const int width = c1 - c0;
    const bool dx_ok = width == 0 || (width >= 16 && width <= 128 && width % 16 == 0 && (!dx || tma_ok(dx, lddx)) &&
                                      (!dot_x || tma_ok(dot_x, ld_dot)));
    return rows >= BM && k1 >= 16 && k1 <= 128 && k1 % 16 == 0 && k2 <= 4 && n >= 16 && n <= 128 && n % 16 == 0 &&
           tma_ok(g, ldg) && (act == HGIN_ACT_NONE || tma_ok(z, ldz)) && tma_ok(x1, ld1) && dx_ok &&
           encode_fn() != nullptr;
}

static int dz_ctas() { return kNumSMs * 2; }   // two resident CTAs per SM (the tail accumulators need > 80 registers)

int64_t fwd_workspace_bytes(int k1, int n) { return align_up(static_cast<int64_t>(n) * k1 * 4, 1024) + 1024; }

int64_t bwd_workspace_bytes(int64_t rows, int k1, int k2, int n) {
    int64_t b = 0;
    b += align_up(rows * n * 4, 1024);                                     // dz
    b += align_up(static_cast<int64_t>(128) * n * 4, 1024);                // W^T slice
    b += align_up(static_cast<int64_t>(dz_ctas()) * n * (k2 + 1) * 4, 1024);  // db / tail partials
    b += align_up(static_cast<int64_t>(kNumSMs) * 4 * 4, 1024);            // dalpha partials (<= 4 per CTA of the fused kernels)
    b += align_up(static_cast<int64_t>(kNumSMs) * n * k1 * 4, 1024);       // dW partials
    b += align_up(static_cast<int64_t>(kNumSMs) * 4 * 2, 1024);            // dot / dot2 partials
    b += align_up(static_cast<int64_t>(kNumSMs) * n * 4, 1024);            // db partials of the weight-gradient kernel
    return b + 1024;
}

static int g_fused_bwd = -1;
bool fused_bwd_enabled() {
    if (g_fused_bwd < 0) g_fused_bwd = (getenv("HGIN_FUSED_BWD") && atoi(getenv("HGIN_FUSED_BWD")) != 0) ? 1 : 0;
    return g_fused_bwd == 1;
}
void set_fused_bwd(int on) { g_fused_bwd = on ? 1 : 0; }

static int g_fused_dw = -1;
bool fused_dw_enabled() {
    // opt-in: measured 1.17 ms against 0.94 ms for dz_prepare + gemm_tn on a 2.5 M x 128 x 128 layer —
    // ring slots stay occupied through transform and store, halving the bytes actually in flight
    if (g_fused_dw < 0) g_fused_dw = (getenv("HGIN_FUSED_DW") && atoi(getenv("HGIN_FUSED_DW")) != 0) ? 1 : 0;
    return g_fused_dw == 1;
}
void set_fused_dw(int on) { g_fused_dw = on ? 1 : 0; }

static bool g_attr_set = false;
static int32_t set_attrs() {
    if (g_attr_set) return HGIN_OK;
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel<EPI_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, NtSmem::total);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_kernel<EPI_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, NtSmem::total);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TnSmem::total);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem::total);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dw_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DwSmem::total);
    if (e != cudaSuccess) return fail(HGIN_ERR_CUDA, "tensor-core kernels: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    g_attr_set = true;
    return HGIN_OK;
}

static char *carve(char *&p, int64_t bytes) {
    char *r = p;
    p += align_up(bytes, 1024);
    return r;
}

// ---- forward --------------------------------------------------------------------------------------
int32_t linear_fwd(int64_t rows, const float *x1, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                   const float *W, const float *bias, int n, int act, const float *alpha, float *z, int64_t ldz,
                   float *out, int64_t ldo, int accumulate_out, void *workspace, cudaStream_t s) {
    if (int32_t rc = set_attrs()) return rc;
    const int k = k1 + k2;
    char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
    float *Wp = reinterpret_cast<float *>(carve(ws, static_cast<int64_t>(n) * k1 * 4));
    pack_cols_kernel<<<grid_for(n * k1, 256, 1), 256, 0, s>>>(W, n, k, 0, k1, Wp);

    CUtensorMap tm_a, tm_b, tm_o, tm_z, tm_e;
    bool ok = make_map(&tm_a, x1, k1, rows, ld1, KB, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_map(&tm_b, Wp, k1, n, k1, KB, n, CU_TENSOR_MAP_SWIZZLE_128B);
    float *o_base = out ? out : z;
    const int64_t o_ld = out ? ldo : ldz;
    ok = ok && make_map(&tm_o, o_base, n, rows, o_ld, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    ok = ok && make_map(&tm_z, z ? z : o_base, n, rows, z ? ldz : o_ld, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    ok = ok && make_map(&tm_e, o_base, n, rows, o_ld, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_fwd(tf32): cuTensorMapEncodeTiled failed");

    NtParams p{};
    p.rows = rows;
    p.num_tiles = static_cast<int>(ceil_div(rows, BM));
    p.num_kb = static_cast<int>(ceil_div(k1, KB));
    p.n = n;
    p.bias = bias;
    p.alpha = alpha;
    p.act = act;
    p.x2 = x2;
    p.ld2 = ld2;
    p.k2 = k2;
    p.w_tail = W + k1;
    p.ldw = k;
    p.want_z = z != nullptr;
    p.want_out = out != nullptr;
    p.use_e = (out != nullptr && accumulate_out) ? 1 : 0;
    p.dot_partials = nullptr;
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    gemm_nt_kernel<EPI_FWD><<<grid, NT_THREADS, NtSmem::total, s>>>(tm_a, tm_b, tm_o, tm_z, tm_e, p);
    HGIN_CHECK_LAUNCH("hgin_linear_fwd(tf32)");
    return HGIN_OK;
}

// ---- backward -------------------------------------------------------------------------------------
int32_t linear_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz, int act,
                   const float *alpha, const float *x1, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                   const float *W, int n, int c0, int c1, float *dx, int64_t lddx, const float *dot_x,
                   int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace,
                   const TnDebug *dbg, const PostArgs *post, cudaStream_t s) {
    if (int32_t rc = set_attrs()) return rc;
    const int k = k1 + k2;
    const bool post_on = post && post->z && post->act != HGIN_ACT_NONE;
    if (post_on && ddot) return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_bwd(tf32): post-activation and ddot together");
    char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
    float *dz = reinterpret_cast<float *>(carve(ws, rows * n * 4));
    float *Wt = reinterpret_cast<float *>(carve(ws, static_cast<int64_t>(128) * n * 4));
    float *sum_part = reinterpret_cast<float *>(carve(ws, static_cast<int64_t>(dz_ctas()) * n * (k2 + 1) * 4));
    float *alpha_part = reinterpret_cast<float *>(carve(ws, static_cast<int64_t>(kNumSMs) * 4 * 4));
    float *dw_part = reinterpret_cast<float *>(carve(ws, static_cast<int64_t>(kNumSMs) * n * k1 * 4));
    float *dot_part = reinterpret_cast<float *>(carve(ws, static_cast<int64_t>(kNumSMs) * 4 * 2));
    float *dot2_part = dot_part + kNumSMs;
    float *db_part = reinterpret_cast<float *>(carve(ws, static_cast<int64_t>(kNumSMs) * n * 4));

    // 0. fused single-pass backward when the shapes allow it (linear_tc_fused.cuh)
    // Opt-in (HGIN_FUSED_BWD=1 or hgin_set_option("fused_bwd", 1)): on B200 the fused kernel is bound
    // by bytes in flight — W (64 KB) + the x tile (64 KB) leave shared memory for a 2-stage ring only,
    // ~64 KB of DRAM reads outstanding per SM against the ~130 KB that 43 GB/s/SM x ~3 us loaded
    // latency require — and measures 2.3 ms against 1.5 ms for the three-pass path.
    if (fused_bwd_enabled() && !dbg && !post_on && act != HGIN_ACT_NONE && dW && n % 32 == 0 && k1 % 16 == 0 && k1 >= 32 && c0 == 0 && c1 == k1 && (dx || ddot)) {
        const int width = k1;
        transpose_cols_kernel<<<grid_for(n * width, 256, 1), 256, 0, s>>>(W, n, k, 0, width, Wt);
        CUtensorMap tm_g, tm_z, tm_wt, tm_h, tm_dx, tm_e;
        bool ok = make_map(&tm_g, g, n, rows, ldg, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
                  make_map(&tm_z, z ? z : g, n, rows, z ? ldz : ldg, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
                  make_map(&tm_wt, Wt, n, width, n, 32, width, CU_TENSOR_MAP_SWIZZLE_128B) &&
                  make_map(&tm_h, x1, k1, rows, ld1, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
                  make_map(&tm_dx, dx ? dx : dz, dx ? width : n, rows, dx ? lddx : n, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
                  make_map(&tm_e, dot_x ? dot_x : dz, dot_x ? width : n, rows, dot_x ? ld_dot : n, 32, BM,
                           CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_bwd(tf32 fused): cuTensorMapEncodeTiled failed");
        FusedParams p{};
        p.rows = rows;
        p.num_tiles = static_cast<int>(ceil_div(rows, BM));
        p.n = n;
        p.k1 = k1;
        p.act = act;
        p.alpha = alpha;
        p.x2 = x2;
        p.ld2 = ld2;
        p.k2 = k2;
        p.want_dx = dx != nullptr;
        p.use_e = ddot != nullptr;
        p.want_sums = 1;
        const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
        p.dot_partials = ddot ? dot_part : nullptr;
        p.dw_partials = dw_part;
        p.sum_partials = sum_part;
        p.alpha_partials = alpha_part;
        bwd_fused_kernel<<<grid, FUSED_THREADS, FusedSmem::total, s>>>(tm_g, tm_z, tm_wt, tm_h, tm_dx, tm_e, p);
        reduce_partials_kernel<<<reduce_grid(n * k1), 256, 0, s>>>(dw_part, grid, n, k1, 0, dW, k, 0, nullptr);
        if (db || k2 > 0)
            reduce_partials_kernel<<<reduce_grid(n * (k2 + 1)), 256, 0, s>>>(sum_part, grid, n, k2, 1, dW, k, k1, db);
        if (dalpha) {
            if (act == HGIN_ACT_PRELU) reduce_scalar_tc_kernel<<<1, 1024, 0, s>>>(alpha_part, grid * 4, dalpha);
            else cudaMemsetAsync(dalpha, 0, sizeof(float), s);
        }
        if (ddot) reduce_scalar_tc_kernel<<<1, 1024, 0, s>>>(dot_part, grid, ddot);
        HGIN_CHECK_LAUNCH("hgin_linear_bwd(tf32 fused)");
        return HGIN_OK;
    }

    // 1. dz, the weight gradient and the cheap reductions.  Preferred: ONE kernel that reads g, z, x
    //    once (dz through TMEM into the dW MMA, and written out for step 2); otherwise dz_prepare
    //    here and gemm_tn in step 3.
    const float *dz_src = dz;     // what step 2 contracts with W
    int64_t dz_ld = n;
    bool dw_done = false;
    if (fused_dw_enabled() && !dbg && act != HGIN_ACT_NONE && dW && (n == 32 || n == 64 || n == 128) && k1 % 16 == 0) {
        CUtensorMap tm_g, tm_z, tm_h, tm_dz;
        const bool act_on = act != HGIN_ACT_NONE;
        bool ok = make_map(&tm_g, g, n, rows, ldg, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
                  make_map(&tm_z, act_on ? z : g, n, rows, act_on ? ldz : ldg, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
                  make_map(&tm_h, x1, k1, rows, ld1, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
                  make_map(&tm_dz, dz, n, rows, n, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_bwd(tf32): cuTensorMapEncodeTiled failed (dW fused)");
        DwParams p{};
        p.rows = rows;
        p.rows_per_cta = align_up(ceil_div(rows, kNumSMs), BM);
        p.n = n;
        p.k1 = k1;
        p.act = act;
        p.alpha = alpha;
        p.x2 = x2;
        p.ld2 = ld2;
        p.k2 = k2;
        p.store_dz = act_on ? 1 : 0;
        p.dw_partials = dw_part;
        p.sum_partials = sum_part;
        p.alpha_partials = alpha_part;
        const int grid = static_cast<int>(ceil_div(rows, p.rows_per_cta));
        dw_fused_kernel<<<grid, DW_THREADS, DwSmem::total, s>>>(tm_g, tm_z, tm_h, tm_dz, p);
        reduce_partials_kernel<<<reduce_grid(n * k1), 256, 0, s>>>(dw_part, grid, n, k1, 0, dW, k, 0, nullptr);
        if (db || k2 > 0)
            reduce_partials_kernel<<<reduce_grid(n * (k2 + 1)), 256, 0, s>>>(sum_part, grid, n, k2, 1, dW, k, k1, db);
        if (dalpha) {
            if (act == HGIN_ACT_PRELU) reduce_scalar_tc_kernel<<<1, 1024, 0, s>>>(alpha_part, grid * 4, dalpha);
            else cudaMemsetAsync(dalpha, 0, sizeof(float), s);
        }
        if (!act_on) {   // dz == g: step 2 reads g in place
            dz_src = g;
            dz_ld = ldg;
        }
        dw_done = true;
    }
    // act == NONE: g already IS dz (the producer of g applied this layer's activation derivative,
    // hgin_linear_bwd_post / hgin_gin_combine_post) and is consumed in place by both GEMMs; db comes
    // out of the weight-gradient MMA (ones column), so no pass over g remains unless the rank-k2
    // tail of dW is wanted.
    bool db_from_mma = false;
    if (!dw_done) {
        const bool inplace = act == HGIN_ACT_NONE && !dbg;
        db_from_mma = inplace && dW && db && k1 % 32 == 0;
        const bool tail = dW && k2 > 0;
        const int want_sums = inplace ? ((tail || (db && !db_from_mma)) ? 1 : 0) : ((dW || db || dalpha) ? 1 : 0);
        const bool want_alpha = dalpha && act == HGIN_ACT_PRELU;
        if (inplace) {
            dz_src = g;
            dz_ld = ldg;
        }
        if (inplace && want_sums) {     // dz is read in place by the GEMMs: only the column sums are missing
            static bool attr = false;
            if (!attr) {
                cudaFuncSetAttribute(tail_sums_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                attr = true;
            }
            const int slots = TAIL_THREADS / (n / 4);
            const int ctas = static_cast<int>(ceil_div(rows, slots) < tail_ctas() ? ceil_div(rows, slots) : tail_ctas());
            tail_sums_kernel<float><<<ctas, TAIL_THREADS, tail_smem<float>(n), s>>>(rows, n, g, ldg, x2, ld2, k2, sum_part);
            reduce_partials_kernel<<<reduce_grid(n * (k2 + 1)), 256, 0, s>>>(sum_part, ctas, n, k2, 1, dW, k, k1,
                                                                                 db_from_mma ? nullptr : db);
        } else if (!inplace) {
            const int tpr = n / 4, slots = DZ_THREADS / tpr;
            const int ctas = static_cast<int>(ceil_div(rows, slots) < dz_ctas() ? ceil_div(rows, slots) : dz_ctas());
            dz_prepare_kernel<<<ctas, DZ_THREADS, static_cast<size_t>(slots) * n * 5 * 4, s>>>(
                rows, n, g, ldg, z, ldz, act, alpha, x2, ld2, k2, dz, sum_part, want_alpha ? alpha_part : nullptr, want_sums,
                inplace ? 0 : 1);
            if (want_sums && ((db && !db_from_mma) || tail))
                reduce_partials_kernel<<<reduce_grid(n * (k2 + 1)), 256, 0, s>>>(sum_part, ctas, n, k2, 1, dW, k, k1,
                                                                                     db_from_mma ? nullptr : db);
            if (want_alpha) reduce_scalar_tc_kernel<<<1, 1024, 0, s>>>(alpha_part, ctas, dalpha);
        }
        if (dalpha && !want_alpha) cudaMemsetAsync(dalpha, 0, sizeof(float), s);
    }

    // 2. input gradient: dx[:, c0:c1] = dz * W[:, c0:c1]
    const int width = c1 - c0;
    if (width > 0 && (dx || ddot)) {
        transpose_cols_kernel<<<grid_for(n * width, 256, 1), 256, 0, s>>>(W, n, k, c0, width, Wt);
        CUtensorMap tm_a, tm_b, tm_o, tm_e;
        bool ok = make_map(&tm_a, dz_src, n, rows, dz_ld, KB, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
                  make_map(&tm_b, Wt, n, width, n, KB, width, CU_TENSOR_MAP_SWIZZLE_128B);
        // without a dx destination the store map still needs a valid (never written) target
        ok = ok && make_map(&tm_o, dx ? dx : dz, dx ? width : n, rows, dx ? lddx : n, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        const float *e_src = post_on ? static_cast<const float *>(post->z) : (dot_x ? dot_x : dz);
        const int64_t e_ld = post_on ? post->ldz : (dot_x ? ld_dot : n);
        ok = ok && make_map(&tm_e, e_src, (post_on || dot_x) ? width : n, rows, e_ld, 32, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_bwd(tf32): cuTensorMapEncodeTiled failed (dx)");
        NtParams p{};
        p.rows = rows;
        p.num_tiles = static_cast<int>(ceil_div(rows, BM));
        p.num_kb = static_cast<int>(ceil_div(n, KB));
        p.n = width;
        p.act = post_on ? post->act : HGIN_ACT_NONE;
        p.alpha = post_on ? post->alpha : nullptr;
        p.want_out = dx != nullptr;
        p.use_e = post_on ? 2 : (ddot != nullptr ? 1 : 0);
        const bool post_alpha = post_on && post->dalpha && post->act == HGIN_ACT_PRELU;
        p.dot_partials = (ddot || post_alpha) ? dot_part : nullptr;
        p.self_eps = post_on ? post->self_eps : nullptr;
        p.dot2_partials = (post_on && post->ddot) ? dot2_part : nullptr;
        const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
        gemm_nt_kernel<EPI_DX><<<grid, NT_THREADS, NtSmem::total, s>>>(tm_a, tm_b, tm_o, tm_o, tm_e, p);
        if (ddot) reduce_scalar_tc_kernel<<<1, 1024, 0, s>>>(dot_part, grid, ddot);
        if (p.dot2_partials) reduce_scalar_tc_kernel<<<1, 1024, 0, s>>>(dot2_part, grid, post->ddot);
        if (post_alpha) reduce_scalar_tc_kernel<<<1, 1024, 0, s>>>(dot_part, grid, post->dalpha);
        else if (post && post->dalpha) cudaMemsetAsync(post->dalpha, 0, sizeof(float), s);
    } else {
        if (ddot) cudaMemsetAsync(ddot, 0, sizeof(float), s);
        if (post && post->dalpha) cudaMemsetAsync(post->dalpha, 0, sizeof(float), s);
        if (post && post->ddot) cudaMemsetAsync(post->ddot, 0, sizeof(float), s);
    }

    // 3. weight gradient: dW[:, :k1] = dz^T x1 (unless step 1 already produced it)
    if (dW && !dw_done) {
        const CUtensorMapSwizzle swz = dbg ? static_cast<CUtensorMapSwizzle>(dbg->tma_swizzle) : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
        CUtensorMap tm_a, tm_b;
        bool ok = make_map(&tm_a, dz_src, n, rows, dz_ld, 32, TN_ROWS, swz) &&
                  make_map(&tm_b, x1, k1, rows, ld1, 32, TN_ROWS, swz);
        if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_bwd(tf32): cuTensorMapEncodeTiled failed (dW)");
        TnParams p{};
        p.rows = rows;
        p.rows_per_cta = align_up(ceil_div(rows, kNumSMs), TN_ROWS);
        p.n = n;
        p.k = k1;
        p.partials = dw_part;
        p.ones_col = db_from_mma ? 1 : 0;
        p.db_partials = db_part;
        p.lbo = dbg ? dbg->lbo : TN_BOX_BYTES;
        p.sbo = dbg ? dbg->sbo : 512;
        p.layout_type = dbg ? dbg->layout_type : 1;
        p.k_step_bytes = dbg ? dbg->k_step_bytes : 1024;
        const int grid = static_cast<int>(ceil_div(rows, p.rows_per_cta));
        gemm_tn_kernel<<<grid, THREADS, TnSmem::total, s>>>(tm_a, tm_b, p);
        reduce_partials_kernel<<<reduce_grid(n * k1), 256, 0, s>>>(dw_part, grid, n, k1, 0, dW, k, 0, nullptr);
        if (db_from_mma) reduce_partials_kernel<<<reduce_grid(n), 256, 0, s>>>(db_part, grid, n, 0, 1, nullptr, 0, 0, db);
    }
    HGIN_CHECK_LAUNCH("hgin_linear_bwd(tf32)");
    return HGIN_OK;
}

}  // namespace tcgemm
}  // namespace hgin
