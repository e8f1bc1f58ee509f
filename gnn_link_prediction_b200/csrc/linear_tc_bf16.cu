// Host side of the tensor-core dense-layer path on bf16 rows (HGIN_DTYPE_BF16): TMA tensor maps, weight
// repacking to bf16, dz = g * act'(z) with the fused db / dalpha / tail-dW reductions, and the dispatch used by
// hgin_linear_fwd_t / hgin_linear_bwd_t.  Kernels: linear_tc_bf16.cuh.
#include "linear_tc_bf16.cuh"
#include "tail_sums.cuh"

namespace hgin {
namespace tcgemm {
namespace {

typedef CUresult (*EncodeTiledFn16)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn16 encode_fn16() {
    static EncodeTiledFn16 fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn16>(p);
    }();
    return fn;
}

// bf16 matrix [outer x inner], row pitch ld elements, box [box_outer x box_inner], SWIZZLE_128B.
bool make_map16(CUtensorMap *m, const bf16 *base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
    EncodeTiledFn16 fn = encode_fn16();
    if (!fn) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16 *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tma_ok16(const void *p, int64_t ld) { return p && aligned16(p) && ld % 8 == 0; }

// dst[r][c] = bf16(src[r * ld + c0 + c])   (W[:, c0:c0+cols] as a dense bf16 matrix: K-major B of the forward GEMM)
__global__ void __launch_bounds__(256) pack_cols16_kernel(const float *__restrict__ src, int rows, int ld, int c0, int cols,
                                                          bf16 *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(__ldg(src + static_cast<int64_t>(i / cols) * ld + c0 + i % cols));
}
// dst[c][r] = bf16(src[r * ld + c0 + c])   (W^T restricted to columns [c0, c0+cols): K-major B of the dx GEMM)
__global__ void __launch_bounds__(256) transpose_cols16_kernel(const float *__restrict__ src, int rows, int ld, int c0,
                                                               int cols, bf16 *__restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
        const int c = i / rows, r = i % rows;
        dst[i] = __float2bfloat16_rn(__ldg(src + static_cast<int64_t>(r) * ld + c0 + c));
    }
}

__device__ __forceinline__ void ld4(const bf16 *p, float (&o)[4]) {
    const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
    o[0] = __uint_as_float(q.x << 16); o[1] = __uint_as_float(q.x & 0xffff0000u);
    o[2] = __uint_as_float(q.y << 16); o[3] = __uint_as_float(q.y & 0xffff0000u);
}
__device__ __forceinline__ void st4(bf16 *p, const float (&o)[4]) {
    *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
}

// dz = g * act'(z) written densely [rows x n] (bf16); per-CTA partials of
//   db[nn] = sum_m dz,  tail[nn][t] = sum_m dz * x2[m][t] (t < k2 <= 4),  dalpha = sum g * min(z, 0).
// Same structure and partial layout as dz_prepare_kernel of linear_tc.cu; g / z / dz are bf16, x2 fp32, the sums
// are taken over the UNROUNDED fp32 products.
constexpr int DZ16_THREADS = 256;
__global__ void __launch_bounds__(DZ16_THREADS, 2)
dz_prepare16_kernel(int64_t rows, int n, const bf16 *__restrict__ g, int64_t ldg, const bf16 *__restrict__ z, int64_t ldz,
                    int act, const float *__restrict__ alpha_ptr, const float *__restrict__ x2, int64_t ld2, int k2,
                    bf16 *__restrict__ dz, float *__restrict__ part, float *__restrict__ alpha_part, int want_sums,
                    int write_dz) {
    extern __shared__ float sm[];  // [slots][n][5] for the cross-slot combine
    __shared__ float red[32];
    const int tpr = n / 4;                       // threads per row (4 columns each)
    const int slots = DZ16_THREADS / tpr;        // rows processed per iteration
    const int slot = threadIdx.x / tpr;
    const int cg = threadIdx.x % tpr;
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
        constexpr int RIF = 4;
        const int64_t stride = static_cast<int64_t>(gridDim.x) * slots;
        for (int64_t m0 = static_cast<int64_t>(blockIdx.x) * slots + slot; m0 < rows; m0 += stride * RIF) {
            float gq[RIF][4], zq[RIF][4], xq[RIF][4];
#pragma unroll
            for (int u = 0; u < RIF; ++u) {
                const int64_t m = m0 + u * stride;
#pragma unroll
                for (int t = 0; t < 4; ++t) { gq[u][t] = 0.f; zq[u][t] = 1.f; xq[u][t] = 0.f; }
                if (m < rows) {
                    ld4(g + m * ldg + cg * 4, gq[u]);
                    if (act != HGIN_ACT_NONE) ld4(z + m * ldz + cg * 4, zq[u]);
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
                float d[4] = {gq[u][0], gq[u][1], gq[u][2], gq[u][3]};
                if (act != HGIN_ACT_NONE) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (act == HGIN_ACT_PRELU && !(zq[u][i] > 0.f)) dalpha += d[i] * zq[u][i];
                        d[i] = act_backward(d[i], zq[u][i], act, alpha);
                    }
                }
                if (write_dz) st4(dz + m * n + cg * 4, d);
                if (want_sums) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        db[i] += d[i];
#pragma unroll
                        for (int t = 0; t < 4; ++t) tail[i][t] = fmaf(d[i], xq[u][t], tail[i][t]);
                    }
                }
            }
        }
    }
    if (!want_sums) {
        if (alpha_part) {
            dalpha = block_sum(dalpha, red);
            if (threadIdx.x == 0) alpha_part[blockIdx.x] = dalpha;
        }
        return;
    }
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
    for (int i = threadIdx.x; i < n * kp; i += DZ16_THREADS) {
        const int nn = i / kp, t = i % kp;
        const int src_t = (t == k2) ? 4 : t;
        float s = 0.f;
        for (int sl = 0; sl < slots; ++sl) s += sm[(static_cast<int64_t>(sl) * n + nn) * 5 + src_t];
        part[(static_cast<int64_t>(blockIdx.x) * n + nn) * kp + t] = s;
    }
    dalpha = block_sum(dalpha, red);
    if (threadIdx.x == 0 && alpha_part) alpha_part[blockIdx.x] = dalpha;
}

static inline int reduce_grid(int64_t total) {
    const int64_t want = (total + 31) / 32;
    return static_cast<int>(want < 1 ? 1 : (want > kNumSMs * 8 ? kNumSMs * 8 : want));
}

// One CTA per 32 consecutive elements (grid-stride over such groups): warp w adds the partials w, w + 8, ... with the lanes on
// consecutive elements (coalesced, the loads of a warp independent), the eight sums are combined in a fixed order.  (A thread
// per element walking all ~148 partials was a chain of ~37 dependent steps: 14.5 us per call, 22 calls per Cfg-C step.)
__global__ void __launch_bounds__(256)
reduce_partials16_kernel(const float *__restrict__ part, int num_part, int n, int kcols, int has_db_col,
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

__global__ void __launch_bounds__(1024) reduce_scalar16_kernel(const float *__restrict__ v, int count, float *__restrict__ out) {
    __shared__ float red[32];
    float s = 0.0f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += v[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

int dz16_ctas() { return kNumSMs * 2; }

bool g_attr16_set = false;
int32_t set_attrs16() {
    if (g_attr16_set) return HGIN_OK;
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_bf16_kernel<EPI_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Nt16Smem::total);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_nt_bf16_kernel<EPI_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, Nt16Smem::total);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tn_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Tn16Smem::total);
    if (e != cudaSuccess) return fail(HGIN_ERR_CUDA, "bf16 tensor-core kernels: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    g_attr16_set = true;
    return HGIN_OK;
}

char *carve16(char *&p, int64_t bytes) {
    char *r = p;
    p += align_up(bytes, 1024);
    return r;
}

}  // namespace

bool fwd_eligible_bf16(int64_t rows, const void *x1, int64_t ld1, int k1, int k2, int n, const void *z, int64_t ldz,
                       const void *out, int64_t ldo) {
    return rows >= BM && k1 >= 16 && k1 <= 128 && k1 % 16 == 0 && k2 <= 4 && n >= 16 && n <= 128 && n % 16 == 0 &&
           tma_ok16(x1, ld1) && (!z || tma_ok16(z, ldz)) && (!out || tma_ok16(out, ldo)) && encode_fn16() != nullptr;
}

bool bwd_eligible_bf16(int64_t rows, const void *g, int64_t ldg, const void *z, int64_t ldz, int act, const void *x1,
                       int64_t ld1, int k1, int k2, int n, int c0, int c1, const void *dx, int64_t lddx,
                       const void *dot_x, int64_t ld_dot) {
    const int width = c1 - c0;
    const bool dx_ok = width == 0 || (width >= 16 && width <= 128 && width % 16 == 0 && (!dx || tma_ok16(dx, lddx)) &&
                                      (!dot_x || tma_ok16(dot_x, ld_dot)));
    return rows >= BM && k1 >= 16 && k1 <= 128 && k1 % 16 == 0 && k2 <= 4 && n >= 16 && n <= 128 && n % 16 == 0 &&
           tma_ok16(g, ldg) && (act == HGIN_ACT_NONE || tma_ok16(z, ldz)) && tma_ok16(x1, ld1) && dx_ok &&
           encode_fn16() != nullptr;
}

int64_t fwd_workspace_bytes_bf16(int k1, int n) { return align_up(static_cast<int64_t>(n) * k1 * 2, 1024) + 1024; }

int64_t bwd_workspace_bytes_bf16(int64_t rows, int k1, int k2, int n) {
    int64_t b = 0;
    b += align_up(rows * n * 2, 1024);                                       // dz
    b += align_up(static_cast<int64_t>(128) * n * 2, 1024);                  // W^T slice
    b += align_up(static_cast<int64_t>(dz16_ctas()) * n * (k2 + 1) * 4, 1024);  // db / tail partials
    b += align_up(static_cast<int64_t>(dz16_ctas()) * 4, 1024);              // dalpha partials
    b += align_up(static_cast<int64_t>(kNumSMs) * n * k1 * 4, 1024);         // dW partials
    b += align_up(static_cast<int64_t>(kNumSMs) * 4 * 2, 1024);              // dot / dot2 partials
    b += align_up(static_cast<int64_t>(kNumSMs) * n * 4, 1024);              // db partials of the weight-gradient kernel
    return b + 1024;
}

int32_t linear_fwd_bf16(int64_t rows, const void *x1v, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                        const float *W, const float *bias, int n, int act, const float *alpha, void *zv, int64_t ldz,
                        void *outv, int64_t ldo, int accumulate_out, void *workspace, cudaStream_t s) {
    if (int32_t rc = set_attrs16()) return rc;
    const bf16 *x1 = static_cast<const bf16 *>(x1v);
    bf16 *z = static_cast<bf16 *>(zv), *out = static_cast<bf16 *>(outv);
    const int k = k1 + k2;
    char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
    bf16 *Wp = reinterpret_cast<bf16 *>(carve16(ws, static_cast<int64_t>(n) * k1 * 2));
    pack_cols16_kernel<<<grid_for(n * k1, 256, 1), 256, 0, s>>>(W, n, k, 0, k1, Wp);

    CUtensorMap tm_a, tm_b, tm_o, tm_z, tm_e;
    bool ok = make_map16(&tm_a, x1, k1, rows, ld1, KB16, BM) && make_map16(&tm_b, Wp, k1, n, k1, KB16, n);
    bf16 *o_base = out ? out : z;
    const int64_t o_ld = out ? ldo : ldz;
    ok = ok && make_map16(&tm_o, o_base, n, rows, o_ld, CW16, BM);
    ok = ok && make_map16(&tm_z, z ? z : o_base, n, rows, z ? ldz : o_ld, CW16, BM);
    ok = ok && make_map16(&tm_e, o_base, n, rows, o_ld, CW16, BM);
    if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_fwd_t(bf16): cuTensorMapEncodeTiled failed");

    NtParams p{};
    p.rows = rows;
    p.num_tiles = static_cast<int>(ceil_div(rows, BM));
    p.num_kb = static_cast<int>(ceil_div(k1, KB16));
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
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    gemm_nt_bf16_kernel<EPI_FWD><<<grid, NT_THREADS, Nt16Smem::total, s>>>(tm_a, tm_b, tm_o, tm_z, tm_e, p);
    HGIN_CHECK_LAUNCH("hgin_linear_fwd_t(bf16)");
    return HGIN_OK;
}

int32_t linear_bwd_bf16(int64_t rows, const void *gv, int64_t ldg, const void *zv, int64_t ldz, int act,
                        const float *alpha, const void *x1v, int64_t ld1, int k1, const float *x2, int64_t ld2, int k2,
                        const float *W, int n, int c0, int c1, void *dxv, int64_t lddx, const void *dot_xv,
                        int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha, void *workspace,
                        const TnDebug *dbg, const PostArgs *post, cudaStream_t s) {
    if (int32_t rc = set_attrs16()) return rc;
    const bf16 *g = static_cast<const bf16 *>(gv), *z = static_cast<const bf16 *>(zv);
    const bf16 *x1 = static_cast<const bf16 *>(x1v), *dot_x = static_cast<const bf16 *>(dot_xv);
    bf16 *dx = static_cast<bf16 *>(dxv);
    const int k = k1 + k2;
    const bool post_on = post && post->z && post->act != HGIN_ACT_NONE;
    if (post_on && ddot) return fail(HGIN_ERR_UNSUPPORTED, "hgin_linear_bwd_t(bf16): post-activation and ddot together");
    char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
    bf16 *dz = reinterpret_cast<bf16 *>(carve16(ws, rows * n * 2));
    bf16 *Wt = reinterpret_cast<bf16 *>(carve16(ws, static_cast<int64_t>(128) * n * 2));
    float *sum_part = reinterpret_cast<float *>(carve16(ws, static_cast<int64_t>(dz16_ctas()) * n * (k2 + 1) * 4));
    float *alpha_part = reinterpret_cast<float *>(carve16(ws, static_cast<int64_t>(dz16_ctas()) * 4));
    float *dw_part = reinterpret_cast<float *>(carve16(ws, static_cast<int64_t>(kNumSMs) * n * k1 * 4));
    float *dot_part = reinterpret_cast<float *>(carve16(ws, static_cast<int64_t>(kNumSMs) * 4 * 2));
    float *dot2_part = dot_part + kNumSMs;
    float *db_part = reinterpret_cast<float *>(carve16(ws, static_cast<int64_t>(kNumSMs) * n * 4));

    // 1. dz.  act == NONE: g already IS dz and both GEMMs read it in place; db comes out of the weight-gradient
    //    MMA (ones column), so no pass over g remains unless the rank-k2 tail of dW is wanted.
    const bf16 *dz_src = dz;
    int64_t dz_ld = n;
    const bool inplace = act == HGIN_ACT_NONE && !dbg;
    const bool db_from_mma = inplace && dW && db && k1 % 64 == 0;
    {
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
                cudaFuncSetAttribute(tail_sums_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                attr = true;
            }
            const int slots = TAIL_THREADS / (n / 8);
            const int ctas = static_cast<int>(ceil_div(rows, slots) < tail_ctas() ? ceil_div(rows, slots) : tail_ctas());
            tail_sums_kernel<bf16><<<ctas, TAIL_THREADS, tail_smem<bf16>(n), s>>>(rows, n, g, ldg, x2, ld2, k2, sum_part);
            reduce_partials16_kernel<<<reduce_grid(n * (k2 + 1)), 256, 0, s>>>(sum_part, ctas, n, k2, 1, dW, k, k1,
                                                                                   db_from_mma ? nullptr : db);
        } else if (!inplace) {
            const int tpr = n / 4, slots = DZ16_THREADS / tpr;
            const int ctas = static_cast<int>(ceil_div(rows, slots) < dz16_ctas() ? ceil_div(rows, slots) : dz16_ctas());
            dz_prepare16_kernel<<<ctas, DZ16_THREADS, static_cast<size_t>(slots) * n * 5 * 4, s>>>(
                rows, n, g, ldg, z, ldz, act, alpha, x2, ld2, k2, dz, sum_part, want_alpha ? alpha_part : nullptr, want_sums,
                inplace ? 0 : 1);
            if (want_sums && ((db && !db_from_mma) || tail))
                reduce_partials16_kernel<<<reduce_grid(n * (k2 + 1)), 256, 0, s>>>(sum_part, ctas, n, k2, 1, dW, k, k1,
                                                                                       db_from_mma ? nullptr : db);
            if (want_alpha) reduce_scalar16_kernel<<<1, 1024, 0, s>>>(alpha_part, ctas, dalpha);
        }
        if (dalpha && !want_alpha) cudaMemsetAsync(dalpha, 0, sizeof(float), s);
    }

    // 2. input gradient: dx[:, c0:c1] = dz * W[:, c0:c1]
    const int width = c1 - c0;
    if (width > 0 && (dx || ddot)) {
        transpose_cols16_kernel<<<grid_for(n * width, 256, 1), 256, 0, s>>>(W, n, k, c0, width, Wt);
        CUtensorMap tm_a, tm_b, tm_o, tm_e;
        bool ok = make_map16(&tm_a, dz_src, n, rows, dz_ld, KB16, BM) && make_map16(&tm_b, Wt, n, width, n, KB16, width);
        // without a dx destination the store map still needs a valid (never written) target
        ok = ok && make_map16(&tm_o, dx ? dx : dz, dx ? width : n, rows, dx ? lddx : n, CW16, BM);
        const bf16 *e_src = post_on ? static_cast<const bf16 *>(post->z) : (dot_x ? dot_x : dz);
        const int64_t e_ld = post_on ? post->ldz : (dot_x ? ld_dot : n);
        ok = ok && make_map16(&tm_e, e_src, (post_on || dot_x) ? width : n, rows, e_ld, CW16, BM);
        if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_bwd_t(bf16): cuTensorMapEncodeTiled failed (dx)");
        NtParams p{};
        p.rows = rows;
        p.num_tiles = static_cast<int>(ceil_div(rows, BM));
        p.num_kb = static_cast<int>(ceil_div(n, KB16));
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
        gemm_nt_bf16_kernel<EPI_DX><<<grid, NT_THREADS, Nt16Smem::total, s>>>(tm_a, tm_b, tm_o, tm_o, tm_e, p);
        if (ddot) reduce_scalar16_kernel<<<1, 1024, 0, s>>>(dot_part, grid, ddot);
        if (p.dot2_partials) reduce_scalar16_kernel<<<1, 1024, 0, s>>>(dot2_part, grid, post->ddot);
        if (post_alpha) reduce_scalar16_kernel<<<1, 1024, 0, s>>>(dot_part, grid, post->dalpha);
        else if (post && post->dalpha) cudaMemsetAsync(post->dalpha, 0, sizeof(float), s);
    } else {
        if (ddot) cudaMemsetAsync(ddot, 0, sizeof(float), s);
        if (post && post->dalpha) cudaMemsetAsync(post->dalpha, 0, sizeof(float), s);
        if (post && post->ddot) cudaMemsetAsync(post->ddot, 0, sizeof(float), s);
    }

    // 3. weight gradient: dW[:, :k1] = dz^T x1
    if (dW) {
        CUtensorMap tm_a, tm_b;
        bool ok = make_map16(&tm_a, dz_src, n, rows, dz_ld, 64, TN16_ROWS) && make_map16(&tm_b, x1, k1, rows, ld1, 64, TN16_ROWS);
        if (!ok) return fail(HGIN_ERR_CUDA, "hgin_linear_bwd_t(bf16): cuTensorMapEncodeTiled failed (dW)");
        TnParams p{};
        p.rows = rows;
        p.rows_per_cta = align_up(ceil_div(rows, kNumSMs), TN16_ROWS);
        p.n = n;
        p.k = k1;
        p.partials = dw_part;
        p.ones_col = db_from_mma ? 1 : 0;
        p.db_partials = db_part;
        p.lbo = dbg ? dbg->lbo : TN16_BOX_BYTES;
        p.sbo = dbg ? dbg->sbo : 1024;
        p.layout_type = dbg ? dbg->layout_type : static_cast<int>(kLayoutSwizzle128B);
        p.k_step_bytes = dbg ? dbg->k_step_bytes : 2048;
        const int grid = static_cast<int>(ceil_div(rows, p.rows_per_cta));
        gemm_tn_bf16_kernel<<<grid, THREADS, Tn16Smem::total, s>>>(tm_a, tm_b, p);
        reduce_partials16_kernel<<<reduce_grid(n * k1), 256, 0, s>>>(dw_part, grid, n, k1, 0, dW, k, 0, nullptr);
        if (db_from_mma) reduce_partials16_kernel<<<reduce_grid(n), 256, 0, s>>>(db_part, grid, n, 0, 1, nullptr, 0, 0, db);
    }
    HGIN_CHECK_LAUNCH("hgin_linear_bwd_t(bf16)");
    return HGIN_OK;
}

}  // namespace tcgemm
}  // namespace hgin
