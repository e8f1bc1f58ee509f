// Graph-attention aggregation for HetroGAT (models.py:380-506 -> PyG 2.0.2 GATConv.forward / message on bipartite
// inputs with add_self_loops=True): softmax over each destination's incoming edges of
//     e_ij = leaky_relu(a_src[j][h] + a_dst[i][h]),   out[i][h][:] = sum_j softmax_j(e_ij) * xs[j][h][:] + bias,
// on the same destination-sorted / source-sorted CSRs as the GIN kernels.  The edge set PyG builds on the fly —
// edges with (source id == destination id) removed, one loop (i, i) appended for every i < min(N_src, N_dst) — is applied
// inside the kernels (skip col == row, one virtual neighbour at the end), so no second adjacency is materialised.
//
// One warp per row, lanes across the H*C feature columns (4 consecutive columns per lane and chunk, up to 4 chunks), a
// head spans C/4 consecutive lanes (C a power of two in [4, 128]).  Neighbours are taken four at a time with all their
// loads issued together (one col latency + one gather latency per four neighbours); the forward is a single pass with a
// running maximum.  No atomics: the forward pass and the two backward
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

// Neighbours of a row in chunks of GAT_U: the real edges in CSR order without those whose id equals the row id
// (remove_self_loops), then the appended loop (add_self_loops).  All ids of a chunk are fetched before anything that
// depends on them, so a chunk costs ONE col latency and ONE gather latency instead of one pair per neighbour.
constexpr int GAT_U = 4;
__device__ __forceinline__ void neighbour_chunk(const int32_t *__restrict__ col, int beg, int deg, bool loop, int row, int base,
                                                int (&id)[GAT_U]) {
#pragma unroll
    for (int u = 0; u < GAT_U; ++u) {
        const int idx = base + u;
        int j = -1;
        if (idx < deg) {
            j = __ldg(col + beg + idx);
            if (j == row) j = -1;
        } else if (idx == deg && loop) {
            j = row;
        }
        id[u] = j;
    }
}

// ---- forward ----------------------------------------------------------------------------------------------
// row_max / row_sum: [num_rows, H] saved for the backward passes (sum BEFORE the +1e-16 of torch_geometric.utils.softmax)
template <int CH>
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
        int hd[CH];
        float ad[CH], mx[CH], sum[CH];
        float4 acc[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int f = (c * 32 + lane) * 4;
            hd[c] = f < sh.hc ? f / sh.c : 0;
            ad[c] = (c < sh.chunks && f < sh.hc) ? __ldg(a_dst + static_cast<int64_t>(row) * sh.heads + hd[c]) : 0.f;
            mx[c] = -INFINITY;
            sum[c] = 0.f;
            acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // one pass with a running maximum (the weights already summed are rescaled when the maximum grows): the result
        // equals max-then-sum up to rounding, and every neighbour's logits and features are read once
        const int deg = end - beg, total = deg + (loop ? 1 : 0);
        for (int base = 0; base < total; base += GAT_U) {
            int id[GAT_U];
            neighbour_chunk(col, beg, deg, loop, row, base, id);
            float ev[GAT_U][CH];
            float4 xv[GAT_U][CH];
#pragma unroll
            for (int u = 0; u < GAT_U; ++u) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int f = (c * 32 + lane) * 4;
                    if (c < sh.chunks && f < sh.hc && id[u] >= 0) {
                        ev[u][c] = lrelu(__ldg(a_src + static_cast<int64_t>(id[u]) * sh.heads + hd[c]) + ad[c], sh.slope);
                        xv[u][c] = ldg4(xs + static_cast<int64_t>(id[u]) * ld_xs + f);
                    } else {
                        ev[u][c] = -INFINITY;
                        xv[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (c >= sh.chunks) continue;
                float m_new = mx[c];
#pragma unroll
                for (int u = 0; u < GAT_U; ++u) m_new = fmaxf(m_new, ev[u][c]);
                if (m_new > mx[c]) {           // (first chunk: mx = -inf, scale = 0, nothing summed yet)
                    const float scale = expf(mx[c] - m_new);
                    sum[c] *= scale;
                    acc[c].x *= scale; acc[c].y *= scale; acc[c].z *= scale; acc[c].w *= scale;
                    mx[c] = m_new;
                }
#pragma unroll
                for (int u = 0; u < GAT_U; ++u) {
                    if (id[u] >= 0) {
                        const float w = expf(ev[u][c] - mx[c]);
                        sum[c] += w;
                        acc[c].x = fmaf(w, xv[u][c].x, acc[c].x);
                        acc[c].y = fmaf(w, xv[u][c].y, acc[c].y);
                        acc[c].z = fmaf(w, xv[u][c].z, acc[c].z);
                        acc[c].w = fmaf(w, xv[u][c].w, acc[c].w);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
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
template <int CH>
__global__ void __launch_bounds__(256)
gat_bwd_dst_kernel(int num_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int n_loop,
                   const float *__restrict__ xs, int64_t ld_xs, const float *__restrict__ a_src, const float *__restrict__ a_dst,
                   const float *__restrict__ row_max, const float *__restrict__ row_sum, const float *__restrict__ g,
                   int64_t ld_g, GatShape sh, float *__restrict__ d_a_dst, float4 *__restrict__ aux) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < num_rows; row += warps) {
        const int beg = rowptr ? __ldg(rowptr + row) : 0, end = rowptr ? __ldg(rowptr + row + 1) : 0;
        const bool loop = row < n_loop;
        int hd[CH];
        float ad[CH], mx[CH], inv[CH], s0[CH], s1[CH], s2[CH];
        float4 gr[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
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
        const int deg = end - beg, total = deg + (loop ? 1 : 0);
        for (int base = 0; base < total; base += GAT_U) {
            int id[GAT_U];
            neighbour_chunk(col, beg, deg, loop, row, base, id);
            float rawv[GAT_U][CH];
            float4 xv[GAT_U][CH];
#pragma unroll
            for (int u = 0; u < GAT_U; ++u) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int f = (c * 32 + lane) * 4;
                    const bool on = c < sh.chunks && f < sh.hc && id[u] >= 0;
                    rawv[u][c] = on ? __ldg(a_src + static_cast<int64_t>(id[u]) * sh.heads + hd[c]) + ad[c] : 0.f;
                    xv[u][c] = on ? ldg4(xs + static_cast<int64_t>(id[u]) * ld_xs + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < GAT_U; ++u) {
                if (id[u] < 0) continue;      // (warp-uniform: the shuffles below are convergent)
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (c < sh.chunks) {
                        const int f = (c * 32 + lane) * 4;
                        const bool on = f < sh.hc;
                        const float alpha = on ? expf(lrelu(rawv[u][c], sh.slope) - mx[c]) * inv[c] : 0.f;
                        const float da = group_sum(dot4(gr[c], xv[u][c]), sh.group);
                        const float lp = rawv[u][c] > 0.f ? 1.f : sh.slope;
                        s0[c] = fmaf(alpha, da, s0[c]);
                        s1[c] = fmaf(lp * alpha, da, s1[c]);
                        s2[c] = fmaf(lp, alpha, s2[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int f = (c * 32 + lane) * 4;
            if (c < sh.chunks && f < sh.hc && (f % sh.c) == 0) {
                const int64_t rh = static_cast<int64_t>(row) * sh.heads + hd[c];
                d_a_dst[rh] = s1[c] - s0[c] * s2[c];
                // what the source pass needs about this destination, in ONE 16-byte record per (row, head):
                // (a_dst, row maximum, 1 / (row sum + 1e-16), softmax dot D) — one gather there instead of four
                aux[rh] = make_float4(ad[c], mx[c], inv[c], s0[c]);
            }
        }
    }
}

// ---- backward, source side (transposed CSR: rows = sources, cols = destinations) -----------------------------------
//   d xs[j][h] = sum_i alpha_ij g[i][h],    d a_src[j][h] = sum_i l'_ij alpha_ij (dalpha_ij - D_i)
template <int CH>
__global__ void __launch_bounds__(256)
gat_bwd_src_kernel(int num_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int n_loop,
                   const float *__restrict__ xs, int64_t ld_xs, const float *__restrict__ a_src,
                   const float4 *__restrict__ aux, const float *__restrict__ g, int64_t ld_g, GatShape sh,
                   float *__restrict__ d_xs, int64_t ld_dxs, float *__restrict__ d_a_src) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < num_rows; row += warps) {
        const int beg = rowptr ? __ldg(rowptr + row) : 0, end = rowptr ? __ldg(rowptr + row + 1) : 0;
        const bool loop = row < n_loop;
        int hd[CH];
        float as[CH], das[CH];
        float4 xr[CH], acc[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int f = (c * 32 + lane) * 4;
            const bool on = c < sh.chunks && f < sh.hc;
            hd[c] = on ? f / sh.c : 0;
            as[c] = on ? __ldg(a_src + static_cast<int64_t>(row) * sh.heads + hd[c]) : 0.f;
            xr[c] = on ? ldg4(xs + static_cast<int64_t>(row) * ld_xs + f) : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            das[c] = 0.f;
        }
        const int deg = end - beg, total = deg + (loop ? 1 : 0);
        for (int base = 0; base < total; base += GAT_U) {
            int id[GAT_U];
            neighbour_chunk(col, beg, deg, loop, row, base, id);
            float4 axv[GAT_U][CH], gvv[GAT_U][CH];
#pragma unroll
            for (int u = 0; u < GAT_U; ++u) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int f = (c * 32 + lane) * 4;
                    const bool on = c < sh.chunks && f < sh.hc && id[u] >= 0;
                    axv[u][c] = on ? __ldg(aux + static_cast<int64_t>(id[u]) * sh.heads + hd[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    gvv[u][c] = on ? ldg4(g + static_cast<int64_t>(id[u]) * ld_g + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < GAT_U; ++u) {
                if (id[u] < 0) continue;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (c < sh.chunks) {
                        const int f = (c * 32 + lane) * 4;
                        const bool on = f < sh.hc;
                        const float4 ax = axv[u][c], gv = gvv[u][c];     // a_dst, max, 1/sum, D
                        const float raw = as[c] + ax.x;
                        const float alpha = on ? expf(lrelu(raw, sh.slope) - ax.y) * ax.z : 0.f;
                        const float da = group_sum(dot4(gv, xr[c]), sh.group);
                        const float lp = raw > 0.f ? 1.f : sh.slope;
                        if (on) das[c] = fmaf(lp * alpha, da - ax.w, das[c]);
                        acc[c].x = fmaf(alpha, gv.x, acc[c].x);
                        acc[c].y = fmaf(alpha, gv.y, acc[c].y);
                        acc[c].z = fmaf(alpha, gv.z, acc[c].z);
                        acc[c].w = fmaf(alpha, gv.w, acc[c].w);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
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
#define HGIN_GAT_FWD(CH)                                                                                               \
    gat_fwd_kernel<CH><<<row_grid(num_rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(                                  \
        static_cast<int>(num_rows), rowptr, col, n_loop, xs, ld_xs, a_src, a_dst, bias, sh, accumulate, out, ld_out, row_max, \
        row_sum)
    if (sh.chunks == 1) HGIN_GAT_FWD(1);
    else if (sh.chunks == 2) HGIN_GAT_FWD(2);
    else HGIN_GAT_FWD(4);
#undef HGIN_GAT_FWD
    HGIN_CHECK_LAUNCH("hgin_gat_fwd");
    return HGIN_OK;
}

extern "C" int32_t hgin_gat_bwd(int64_t num_dst, const int32_t *rowptr_dst, const int32_t *col_dst, int64_t num_src,
                                const int32_t *rowptr_src, const int32_t *col_src, const float *xs, int64_t ld_xs,
                                const float *a_src, const float *a_dst, const float *row_max, const float *row_sum,
                                const float *g, int64_t ld_g, int32_t heads, int32_t channels, float negative_slope,
                                int32_t add_self_loops, float *d_xs, int64_t ld_dxs, float *d_a_src, float *d_a_dst,
                                void *dot_ws, void *stream) {
    GatShape sh;
    const int32_t st = make_shape(heads, channels, negative_slope, &sh, "hgin_gat_bwd");
    if (st != HGIN_OK) return st;
    HGIN_CHECK_ARG(num_dst >= 0 && num_dst < (1ll << 31) && num_src >= 0 && num_src < (1ll << 31), "hgin_gat_bwd: bad row counts");
    HGIN_CHECK_ARG(ld_xs >= sh.hc && ld_g >= sh.hc && ld_dxs >= sh.hc && ld_xs % 4 == 0 && ld_g % 4 == 0 && ld_dxs % 4 == 0,
                   "hgin_gat_bwd: leading dimensions");
    HGIN_CHECK_ARG(aligned16(xs) && aligned16(g) && aligned16(d_xs) && aligned16(dot_ws),
                   "hgin_gat_bwd: 16-byte alignment of xs / g / d_xs / dot_ws");
    HGIN_CHECK_ARG((num_dst == 0 && num_src == 0) || (xs && a_src && a_dst && row_max && row_sum && g && d_xs && d_a_src &&
                                                       d_a_dst && dot_ws),
                   "hgin_gat_bwd: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_loop = add_self_loops ? static_cast<int>(num_dst < num_src ? num_dst : num_src) : 0;
#define HGIN_GAT_BWD(CH)                                                                                                    \
    do {                                                                                                                    \
        if (num_dst > 0)                                                                                                    \
            gat_bwd_dst_kernel<CH><<<row_grid(num_dst), 256, 0, s>>>(static_cast<int>(num_dst), rowptr_dst, col_dst, n_loop, xs, \
                                                                     ld_xs, a_src, a_dst, row_max, row_sum, g, ld_g, sh, d_a_dst, \
                                                                     static_cast<float4 *>(dot_ws));                         \
        if (num_src > 0)                                                                                                    \
            gat_bwd_src_kernel<CH><<<row_grid(num_src), 256, 0, s>>>(static_cast<int>(num_src), rowptr_src, col_src, n_loop, xs, \
                                                                     ld_xs, a_src, static_cast<const float4 *>(dot_ws), g, ld_g, \
                                                                     sh, d_xs, ld_dxs, d_a_src);                             \
    } while (0)
    if (sh.chunks == 1) HGIN_GAT_BWD(1);
    else if (sh.chunks == 2) HGIN_GAT_BWD(2);
    else HGIN_GAT_BWD(4);
#undef HGIN_GAT_BWD
    HGIN_CHECK_LAUNCH("hgin_gat_bwd");
    return HGIN_OK;
}
