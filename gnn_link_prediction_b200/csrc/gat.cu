// Graph-attention aggregation for HetroGAT (models.py:380-506 -> PyG 2.0.2 GATConv.forward / message on bipartite
// inputs with add_self_loops=True): softmax over each destination's incoming edges of
//     e_ij = leaky_relu(a_src[j][h] + a_dst[i][h]),   out[i][h][:] = sum_j softmax_j(e_ij) * xs[j][h][:] + bias,
// on the same destination-sorted / source-sorted CSRs as the GIN kernels.  The edge set PyG builds on the fly —
// edges with (source id == destination id) removed, one loop (i, i) appended for every i < min(N_src, N_dst) — is applied
// inside the kernels (skip col == row, one virtual neighbour at the end), so no second adjacency is materialised.
//
// One warp per row, lanes across the H*C feature columns (4 consecutive columns per lane and chunk, up to 4 chunks), a
// head spans C/4 consecutive lanes (C a power of two in [4, 128]).  No atomics: the forward pass and the two backward
// passes (per destination: d a_dst and the softmax dot D; per source over the transposed CSR: d xs and d a_src) each own
// their output rows, neighbours are visited in CSR order -> run-to-run deterministic.  Nothing per-edge is stored: the
// backward passes recompute alpha_ij from the saved row maxima and denominators.
#include <math.h>

#include "hgin_common.cuh"

namespace hgin {
namespace {

constexpr int kMaxChunks = 4;   // H*C <= 512

struct GatShape {
    int heads, c, hc, chunks, group;   // group = lanes per head (C / 4)
    float slope;
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : slope * v; }

__device__ __forceinline__ float group_sum(float v, int group) {
    for (int o = group >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ---- forward ----------------------------------------------------------------------------------------------
// row_max / row_sum: [num_rows, H] saved for the backward passes (sum BEFORE the +1e-16 of torch_geometric.utils.softmax)
__global__ void __launch_bounds__(256)
gat_fwd_kernel(int num_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int n_loop,
               const float *__restrict__ xs, int64_t ld_xs, const float *__restrict__ a_src, const float *__restrict__ a_dst,
               const float *__restrict__ bias, GatShape sh, int accumulate, float *__restrict__ out, int64_t ld_out,
               float *__restrict__ row_max, float *__restrict__ row_sum) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < num_rows; row += warps) {
        const int beg = rowptr ? __ldg(rowptr + row) : 0, end = rowptr ? __ldg(rowptr + row + 1) : 0;
        const bool loop = row < n_loop;
        int hd[kMaxChunks];
        float ad[kMaxChunks], mx[kMaxChunks], sum[kMaxChunks];
        float4 acc[kMaxChunks];
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int f = (c * 32 + lane) * 4;
            hd[c] = f < sh.hc ? f / sh.c : 0;
            ad[c] = (c < sh.chunks && f < sh.hc) ? __ldg(a_dst + static_cast<int64_t>(row) * sh.heads + hd[c]) : 0.f;
            mx[c] = -INFINITY;
            sum[c] = 0.f;
            acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // pass 1: row maximum of e_ij per head
        for (int e = beg; e <= end; ++e) {
            int j;
            if (e < end) {
                j = __ldg(col + e);
                if (j == row) continue;       // remove_self_loops: ids compared across the two node types
            } else {
                if (!loop) break;
                j = row;                      // add_self_loops: (i, i) last
            }
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c)
                if (c < sh.chunks)
                    mx[c] = fmaxf(mx[c], lrelu(__ldg(a_src + static_cast<int64_t>(j) * sh.heads + hd[c]) + ad[c], sh.slope));
        }
        // pass 2: weights, denominator, weighted sum of the source rows
        for (int e = beg; e <= end; ++e) {
            int j;
            if (e < end) {
                j = __ldg(col + e);
                if (j == row) continue;
            } else {
                if (!loop) break;
                j = row;
            }
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c) {
                const int f = (c * 32 + lane) * 4;
                if (c < sh.chunks && f < sh.hc) {
                    const float w = expf(lrelu(__ldg(a_src + static_cast<int64_t>(j) * sh.heads + hd[c]) + ad[c], sh.slope) - mx[c]);
                    const float4 x = ldg4(xs + static_cast<int64_t>(j) * ld_xs + f);
                    sum[c] += w;
                    acc[c].x = fmaf(w, x.x, acc[c].x);
                    acc[c].y = fmaf(w, x.y, acc[c].y);
                    acc[c].z = fmaf(w, x.z, acc[c].z);
                    acc[c].w = fmaf(w, x.w, acc[c].w);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int f = (c * 32 + lane) * 4;
            if (c < sh.chunks && f < sh.hc) {
                const float inv = 1.0f / (sum[c] + 1e-16f);
                const float4 b = bias ? ldg4(bias + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 r = make_float4(fmaf(acc[c].x, inv, b.x), fmaf(acc[c].y, inv, b.y), fmaf(acc[c].z, inv, b.z),
                                       fmaf(acc[c].w, inv, b.w));
                float *o = out + static_cast<int64_t>(row) * ld_out + f;
                if (accumulate) {
                    const float4 old = *reinterpret_cast<const float4 *>(o);
                    r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
                }
                *reinterpret_cast<float4 *>(o) = r;
                if ((f % sh.c) == 0) {     // first lane of the head
                    row_max[static_cast<int64_t>(row) * sh.heads + hd[c]] = mx[c];
                    row_sum[static_cast<int64_t>(row) * sh.heads + hd[c]] = sum[c];
                }
            }
        }
    }
}

// ---- backward, destination side -------------------------------------------------------------------------------
// per destination row i and head h:   dalpha_ij = <g[i][h], xs[j][h]>,   D_i = sum_j alpha_ij dalpha_ij,
//   d a_dst[i][h] = sum_j l'_ij alpha_ij (dalpha_ij - D_i) = S1 - D_i * S2   with l' the leaky-relu derivative.
__global__ void __launch_bounds__(256)
gat_bwd_dst_kernel(int num_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int n_loop,
                   const float *__restrict__ xs, int64_t ld_xs, const float *__restrict__ a_src, const float *__restrict__ a_dst,
                   const float *__restrict__ row_max, const float *__restrict__ row_sum, const float *__restrict__ g,
                   int64_t ld_g, GatShape sh, float *__restrict__ d_a_dst, float *__restrict__ dot_d) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < num_rows; row += warps) {
        const int beg = rowptr ? __ldg(rowptr + row) : 0, end = rowptr ? __ldg(rowptr + row + 1) : 0;
        const bool loop = row < n_loop;
        int hd[kMaxChunks];
        float ad[kMaxChunks], mx[kMaxChunks], inv[kMaxChunks], s0[kMaxChunks], s1[kMaxChunks], s2[kMaxChunks];
        float4 gr[kMaxChunks];
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int f = (c * 32 + lane) * 4;
            const bool on = c < sh.chunks && f < sh.hc;
            hd[c] = on ? f / sh.c : 0;
            const int64_t rh = static_cast<int64_t>(row) * sh.heads + hd[c];
            ad[c] = on ? __ldg(a_dst + rh) : 0.f;
            mx[c] = on ? __ldg(row_max + rh) : 0.f;
            inv[c] = on ? 1.0f / (__ldg(row_sum + rh) + 1e-16f) : 0.f;
            gr[c] = on ? ldg4(g + static_cast<int64_t>(row) * ld_g + f) : make_float4(0.f, 0.f, 0.f, 0.f);
            s0[c] = s1[c] = s2[c] = 0.f;
        }
        for (int e = beg; e <= end; ++e) {
            int j;
            if (e < end) {
                j = __ldg(col + e);
                if (j == row) continue;
            } else {
                if (!loop) break;
                j = row;
            }
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c) {
                if (c < sh.chunks) {      // (warp-uniform: the shuffles below are convergent)
                    const int f = (c * 32 + lane) * 4;
                    const bool on = f < sh.hc;
                    const float raw = on ? __ldg(a_src + static_cast<int64_t>(j) * sh.heads + hd[c]) + ad[c] : 0.f;
                    const float alpha = on ? expf(lrelu(raw, sh.slope) - mx[c]) * inv[c] : 0.f;
                    const float4 x = on ? ldg4(xs + static_cast<int64_t>(j) * ld_xs + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float da = group_sum(dot4(gr[c], x), sh.group);
                    const float lp = raw > 0.f ? 1.f : sh.slope;
                    s0[c] = fmaf(alpha, da, s0[c]);
                    s1[c] = fmaf(lp * alpha, da, s1[c]);
                    s2[c] = fmaf(lp, alpha, s2[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int f = (c * 32 + lane) * 4;
            if (c < sh.chunks && f < sh.hc && (f % sh.c) == 0) {
                const int64_t rh = static_cast<int64_t>(row) * sh.heads + hd[c];
                d_a_dst[rh] = s1[c] - s0[c] * s2[c];
                dot_d[rh] = s0[c];
            }
        }
    }
}

// ---- backward, source side (transposed CSR: rows = sources, cols = destinations) -----------------------------------
//   d xs[j][h] = sum_i alpha_ij g[i][h],    d a_src[j][h] = sum_i l'_ij alpha_ij (dalpha_ij - D_i)
__global__ void __launch_bounds__(256)
gat_bwd_src_kernel(int num_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int n_loop,
                   const float *__restrict__ xs, int64_t ld_xs, const float *__restrict__ a_src, const float *__restrict__ a_dst,
                   const float *__restrict__ row_max, const float *__restrict__ row_sum, const float *__restrict__ dot_d,
                   const float *__restrict__ g, int64_t ld_g, GatShape sh, float *__restrict__ d_xs, int64_t ld_dxs,
                   float *__restrict__ d_a_src) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < num_rows; row += warps) {
        const int beg = rowptr ? __ldg(rowptr + row) : 0, end = rowptr ? __ldg(rowptr + row + 1) : 0;
        const bool loop = row < n_loop;
        int hd[kMaxChunks];
        float as[kMaxChunks], das[kMaxChunks];
        float4 xr[kMaxChunks], acc[kMaxChunks];
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int f = (c * 32 + lane) * 4;
            const bool on = c < sh.chunks && f < sh.hc;
            hd[c] = on ? f / sh.c : 0;
            as[c] = on ? __ldg(a_src + static_cast<int64_t>(row) * sh.heads + hd[c]) : 0.f;
            xr[c] = on ? ldg4(xs + static_cast<int64_t>(row) * ld_xs + f) : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            das[c] = 0.f;
        }
        for (int e = beg; e <= end; ++e) {
            int i;
            if (e < end) {
                i = __ldg(col + e);
                if (i == row) continue;
            } else {
                if (!loop) break;
                i = row;
            }
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c) {
                if (c < sh.chunks) {
                    const int f = (c * 32 + lane) * 4;
                    const bool on = f < sh.hc;
                    const int64_t ih = static_cast<int64_t>(i) * sh.heads + hd[c];
                    const float raw = on ? as[c] + __ldg(a_dst + ih) : 0.f;
                    const float alpha = on ? expf(lrelu(raw, sh.slope) - __ldg(row_max + ih)) / (__ldg(row_sum + ih) + 1e-16f) : 0.f;
                    const float4 gv = on ? ldg4(g + static_cast<int64_t>(i) * ld_g + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float da = group_sum(dot4(gv, xr[c]), sh.group);
                    const float lp = raw > 0.f ? 1.f : sh.slope;
                    if (on) das[c] = fmaf(lp * alpha, da - __ldg(dot_d + ih), das[c]);
                    acc[c].x = fmaf(alpha, gv.x, acc[c].x);
                    acc[c].y = fmaf(alpha, gv.y, acc[c].y);
                    acc[c].z = fmaf(alpha, gv.z, acc[c].z);
                    acc[c].w = fmaf(alpha, gv.w, acc[c].w);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kMaxChunks; ++c) {
            const int f = (c * 32 + lane) * 4;
            if (c < sh.chunks && f < sh.hc) {
                *reinterpret_cast<float4 *>(d_xs + static_cast<int64_t>(row) * ld_dxs + f) = acc[c];
                if ((f % sh.c) == 0) d_a_src[static_cast<int64_t>(row) * sh.heads + hd[c]] = das[c];
            }
        }
    }
}

int32_t make_shape(int heads, int c, float slope, GatShape *sh, const char *who) {
    const int hc = heads * c;
    if (!(heads >= 1 && c >= 4 && c <= 128 && (c & (c - 1)) == 0 && hc <= 128 * kMaxChunks))
        return fail(HGIN_ERR_UNSUPPORTED, "%s: heads=%d, channels per head=%d: channels must be a power of two in [4, 128] and "
                    "heads * channels <= %d", who, heads, c, 128 * kMaxChunks);
    sh->heads = heads;
    sh->c = c;
    sh->hc = hc;
    sh->chunks = (hc + 127) / 128;
    sh->group = c / 4;
    sh->slope = slope;
    return HGIN_OK;
}

inline int row_grid(int64_t rows) { return grid_for(rows, 8, 8); }   // 8 warps (rows) per CTA

}  // namespace
}  // namespace hgin

using namespace hgin;

extern "C" int32_t hgin_gat_fwd(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_src,
                                const float *xs, int64_t ld_xs, const float *a_src, const float *a_dst, const float *bias,
                                int32_t heads, int32_t channels, float negative_slope, int32_t add_self_loops,
                                int32_t accumulate, float *out, int64_t ld_out, float *row_max, float *row_sum, void *stream) {
    GatShape sh;
    const int32_t st = make_shape(heads, channels, negative_slope, &sh, "hgin_gat_fwd");
    if (st != HGIN_OK) return st;
    HGIN_CHECK_ARG(num_rows >= 0 && num_rows < (1ll << 31) && num_src >= 0 && num_src < (1ll << 31), "hgin_gat_fwd: bad row counts");
    HGIN_CHECK_ARG(ld_xs >= sh.hc && ld_out >= sh.hc && ld_xs % 4 == 0 && ld_out % 4 == 0, "hgin_gat_fwd: leading dimensions");
    HGIN_CHECK_ARG(num_rows == 0 || (xs && a_src && a_dst && out && row_max && row_sum), "hgin_gat_fwd: null pointer");
    HGIN_CHECK_ARG(aligned16(xs) && aligned16(out) && aligned16(bias), "hgin_gat_fwd: 16-byte alignment of xs / out / bias");
    // (col may be NULL when the relation has no edges at all: only the appended loops contribute then)
    if (num_rows == 0) return HGIN_OK;
    const int n_loop = add_self_loops ? static_cast<int>(num_rows < num_src ? num_rows : num_src) : 0;
    gat_fwd_kernel<<<row_grid(num_rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<int>(num_rows), rowptr, col, n_loop, xs, ld_xs, a_src, a_dst, bias, sh, accumulate, out, ld_out, row_max,
        row_sum);
    HGIN_CHECK_LAUNCH("hgin_gat_fwd");
    return HGIN_OK;
}

extern "C" int32_t hgin_gat_bwd(int64_t num_dst, const int32_t *rowptr_dst, const int32_t *col_dst, int64_t num_src,
                                const int32_t *rowptr_src, const int32_t *col_src, const float *xs, int64_t ld_xs,
                                const float *a_src, const float *a_dst, const float *row_max, const float *row_sum,
                                const float *g, int64_t ld_g, int32_t heads, int32_t channels, float negative_slope,
                                int32_t add_self_loops, float *d_xs, int64_t ld_dxs, float *d_a_src, float *d_a_dst,
                                float *dot_ws, void *stream) {
    GatShape sh;
    const int32_t st = make_shape(heads, channels, negative_slope, &sh, "hgin_gat_bwd");
    if (st != HGIN_OK) return st;
    HGIN_CHECK_ARG(num_dst >= 0 && num_dst < (1ll << 31) && num_src >= 0 && num_src < (1ll << 31), "hgin_gat_bwd: bad row counts");
    HGIN_CHECK_ARG(ld_xs >= sh.hc && ld_g >= sh.hc && ld_dxs >= sh.hc && ld_xs % 4 == 0 && ld_g % 4 == 0 && ld_dxs % 4 == 0,
                   "hgin_gat_bwd: leading dimensions");
    HGIN_CHECK_ARG(aligned16(xs) && aligned16(g) && aligned16(d_xs), "hgin_gat_bwd: 16-byte alignment of xs / g / d_xs");
    HGIN_CHECK_ARG((num_dst == 0 && num_src == 0) || (xs && a_src && a_dst && row_max && row_sum && g && d_xs && d_a_src &&
                                                       d_a_dst && dot_ws),
                   "hgin_gat_bwd: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_loop = add_self_loops ? static_cast<int>(num_dst < num_src ? num_dst : num_src) : 0;
    if (num_dst > 0)
        gat_bwd_dst_kernel<<<row_grid(num_dst), 256, 0, s>>>(static_cast<int>(num_dst), rowptr_dst, col_dst, n_loop, xs, ld_xs,
                                                             a_src, a_dst, row_max, row_sum, g, ld_g, sh, d_a_dst, dot_ws);
    if (num_src > 0)
        gat_bwd_src_kernel<<<row_grid(num_src), 256, 0, s>>>(static_cast<int>(num_src), rowptr_src, col_src, n_loop, xs, ld_xs,
                                                             a_src, a_dst, row_max, row_sum, dot_ws, g, ld_g, sh, d_xs, ld_dxs,
                                                             d_a_src);
    HGIN_CHECK_LAUNCH("hgin_gat_bwd");
    return HGIN_OK;
}
