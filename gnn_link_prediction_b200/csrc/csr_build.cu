// K0 — destination-sorted CSR from a COO edge list (stable; integer-exact).
//
// Contract (include/hgin.h): perm = argsort(key, stable), rowptr =
// exclusive_cumsum(bincount(key)), col = other[perm].  The reference never builds this (PyG
// keeps COO and scatter-adds with atomics, models.py:208); it exists so that the aggregation
// kernel can sum every row in the CPU reference's order without atomics.
//
// Fast path: the reference emits every relation grouped by ascending source id
// (generateFiles.py:145-181 iterates G.edges; collate keeps that order), so for sort_row = 0 the
// keys are already sorted and the CSR is a copy plus row boundaries.  A first kernel checks
// sortedness on the device; both the fast kernel and the general kernels are always launched
// (static launch sequence: CUDA-graph capturable, no host sync) and each returns at once when the
// flag says the other path applies.
//
// General method (all integer, HBM-bound, four passes over the edge list):
//   1. histogram   cnt[key[e]]++                              (int atomics: exact, order-free)
//   2. scan        rowptr = exclusive prefix sum of cnt       (tile scan, recursive on tile sums)
//   3. fill        seg[rowptr[k] + --cnt[k]] = e              (slot order inside a row is arbitrary)
//   4. row sort    each row's edge ids are sorted ascending (ids are unique, so the result is the
//                  stable order, independent of how the atomics of step 3 interleaved), then
//                  col[slot] = other[id].  One warp per row: rank sort by shuffles for rows <= 32,
//                  rank counting against the row segment for longer rows.
#include "hgin_common.cuh"

namespace hgin {
namespace {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename IndexT>
__global__ void __launch_bounds__(256) csr_histogram(const IndexT *__restrict__ key,
                                                     const IndexT *__restrict__ other, int64_t num_edges,
                                                     int64_t num_rows, int64_t num_cols,
                                                     int32_t *__restrict__ cnt, int32_t *__restrict__ status,
                                                     const int32_t *__restrict__ sorted_flag) {
    if (*sorted_flag) return;   // fast path took this call
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    bool bad = false;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < num_edges; e += stride) {
        const int64_t k = static_cast<int64_t>(key[e]);
        const int64_t o = static_cast<int64_t>(other[e]);
        if (k == -1 && o == -1) continue;   // padding slot (static-shape batches), not an error
        if (k < 0 || k >= num_rows || o < 0 || o >= num_cols) {
            bad = true;
        } else {
            atomicAdd(&cnt[k], 1);
        }
    }
    if (bad) atomicExch(status, 1);
}

// Exclusive scan of one tile per block; tile totals go to `tile_sums`.
__global__ void __launch_bounds__(kScanThreads) scan_tiles(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                           int64_t n, int32_t *__restrict__ tile_sums,
                                                           const int32_t *__restrict__ skip_flag) {
    __shared__ int32_t warp_tot[kScanThreads / 32];
    if (skip_flag && *skip_flag) return;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
    int32_t v[kScanItems];
    int32_t local = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        local += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_tot[lane];
        int32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_tot[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    int32_t run = warp_tot[warp] + incl - local;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_add_offsets(int32_t *__restrict__ out, int64_t n,
                                                                 const int32_t *__restrict__ tile_offsets,
                                                                 const int32_t *__restrict__ skip_flag) {
    if (skip_flag && *skip_flag) return;
    const int32_t off = tile_offsets[blockIdx.x];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) out[base + i] += off;
}

template <typename IndexT>
__global__ void __launch_bounds__(256) csr_fill(const IndexT *__restrict__ key, const IndexT *__restrict__ other,
                                                int64_t num_edges, int64_t num_rows, int64_t num_cols,
                                                const int32_t *__restrict__ rowptr, int32_t *__restrict__ cnt,
                                                int32_t *__restrict__ seg, const int32_t *__restrict__ sorted_flag) {
    if (*sorted_flag) return;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < num_edges; e += stride) {
        const int64_t k = static_cast<int64_t>(key[e]);
        const int64_t o = static_cast<int64_t>(other[e]);
        if (k < 0 || k >= num_rows || o < 0 || o >= num_cols) continue;
        const int32_t slot = atomicSub(&cnt[k], 1) - 1;
        seg[rowptr[k] + slot] = static_cast<int32_t>(e);
    }
}

// G lanes per row (G = 32: long rows; G = 8: ~3-neighbour rows, four rows per warp).
template <typename IndexT, int G>
__global__ void __launch_bounds__(256) csr_sort_rows(const IndexT *__restrict__ other, int64_t num_rows,
                                                     const int32_t *__restrict__ rowptr, const int32_t *__restrict__ seg,
                                                     int32_t *__restrict__ col, int32_t *__restrict__ perm,
                                                     const int32_t *__restrict__ sorted_flag) {
    if (*sorted_flag) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane % G, grp = lane / G;
    const int64_t groups = (static_cast<int64_t>(gridDim.x) * blockDim.x) / G;
    const int64_t iters = (num_rows + groups - 1) / groups;   // warp-uniform trip count for the shuffles
    const int64_t g0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t r = it * groups + g0;
        int32_t beg = 0, len = 0;
        if (r < num_rows) {
            beg = rowptr[r];
            len = rowptr[r + 1] - beg;
        }
        int32_t max_len = len;
#pragma unroll
        for (int o = 16; o >= G; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
        if (max_len <= G) {   // every row of this warp fits its lane group: rank by shuffles
            const int32_t id = sub < len ? seg[beg + sub] : INT32_MAX;
            int32_t rank = 0;
            for (int j = 0; j < max_len; ++j) rank += (__shfl_sync(0xffffffffu, id, grp * G + j) < id) ? 1 : 0;
            if (sub < len) {
                col[beg + rank] = static_cast<int32_t>(other[id]);
                if (perm) perm[beg + rank] = id;
            }
        } else {              // rank counting against the row segment (no shuffles: lanes may diverge)
            for (int32_t i = sub; i < len; i += G) {
                const int32_t id = seg[beg + i];
                int32_t rank = 0;
                for (int32_t j = 0; j < len; ++j) rank += (seg[beg + j] < id) ? 1 : 0;
                col[beg + rank] = static_cast<int32_t>(other[id]);
                if (perm) perm[beg + rank] = id;
            }
        }
    }
}

// flag stays set iff keys are non-decreasing over the valid prefix, padding slots (-1,-1) only trail,
// and every index is in range (out-of-range edges must be DROPPED, which only the general path does).
template <typename IndexT>
__global__ void __launch_bounds__(256) csr_check_sorted(const IndexT *__restrict__ key, const IndexT *__restrict__ other,
                                                        int64_t num_edges, int64_t num_rows, int64_t num_cols,
                                                        int32_t *__restrict__ sorted_flag) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    bool bad = false;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < num_edges; e += stride) {
        const int64_t a = static_cast<int64_t>(key[e]), oa = static_cast<int64_t>(other[e]);
        const bool pad_a = a == -1 && oa == -1;
        if (!pad_a && (a < 0 || a >= num_rows || oa < 0 || oa >= num_cols)) bad = true;
        if (e + 1 < num_edges) {
            const int64_t b = static_cast<int64_t>(key[e + 1]);
            const bool pad_b = b == -1 && static_cast<int64_t>(other[e + 1]) == -1;
            if (pad_a ? !pad_b : (!pad_b && b < a)) bad = true;   // valid after padding, or a descent
        }
    }
    if (bad) atomicExch(sorted_flag, 0);
}

// Sorted keys: col = other (narrowed), rowptr from the key boundaries.  Thread e closes the rows
// (key[e-1], key[e]]; the last valid edge also closes the tail rows.
template <typename IndexT>
__global__ void __launch_bounds__(256) csr_from_sorted(const IndexT *__restrict__ key, const IndexT *__restrict__ other,
                                                       int64_t num_edges, int64_t num_rows, int64_t num_cols,
                                                       int32_t *__restrict__ rowptr, int32_t *__restrict__ col,
                                                       int32_t *__restrict__ perm,
                                                       const int32_t *__restrict__ sorted_flag) {
    if (!*sorted_flag) return;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < num_edges; e += stride) {
        const int64_t k = static_cast<int64_t>(key[e]), o = static_cast<int64_t>(other[e]);
        const bool pad = k == -1 && o == -1;
        if (pad) {
            if (e == 0) for (int64_t r = 0; r <= num_rows; ++r) rowptr[r] = 0;   // nothing but padding
            continue;
        }
        col[e] = static_cast<int32_t>(o);   // in range: csr_check_sorted sends anything else to the general path
        if (perm) perm[e] = static_cast<int32_t>(e);
        const int64_t prev = e > 0 ? static_cast<int64_t>(key[e - 1]) : -1;
        for (int64_t r = prev + 1; r <= k; ++r) rowptr[r] = static_cast<int32_t>(e);
        const bool last = e + 1 == num_edges ||
                          (static_cast<int64_t>(key[e + 1]) == -1 && static_cast<int64_t>(other[e + 1]) == -1);
        if (last) for (int64_t r = k + 1; r <= num_rows; ++r) rowptr[r] = static_cast<int32_t>(e + 1);
    }
}

// Exclusive scan of `n` int32 values (in -> out, may alias), recursing over tile sums carved
// from `scratch`.  Returns false if scratch runs out.
bool exclusive_scan(const int32_t *in, int32_t *out, int64_t n, int32_t *scratch, int64_t scratch_elems,
                    const int32_t *skip_flag, cudaStream_t s) {
    const int64_t tiles = ceil_div(n, kScanTile);
    if (tiles <= 1) {
        scan_tiles<<<1, kScanThreads, 0, s>>>(in, out, n, nullptr, skip_flag);
        return true;
    }
    if (scratch_elems < tiles) return false;
    scan_tiles<<<static_cast<unsigned>(tiles), kScanThreads, 0, s>>>(in, out, n, scratch, skip_flag);
    if (!exclusive_scan(scratch, scratch, tiles, scratch + tiles, scratch_elems - tiles, skip_flag, s)) return false;
    scan_add_offsets<<<static_cast<unsigned>(tiles), kScanThreads, 0, s>>>(out, n, scratch, skip_flag);
    return true;
}

int64_t scan_scratch_elems(int64_t n) {
    int64_t total = 0;
    while (n > kScanTile) {
        n = ceil_div(n, kScanTile);
        total += n;
    }
    return total + 8;   // the int right after these holds the sortedness flag
}

template <typename IndexT>
int32_t csr_build_impl(const IndexT *edge_index, int64_t num_edges, int64_t ld_edge, int32_t sort_row,
                       int64_t num_rows, int64_t num_cols, int32_t *rowptr, int32_t *col, int32_t *perm,
                       int32_t *status, int32_t *cnt, int32_t *seg, int32_t *scratch, int64_t scratch_elems,
                       cudaStream_t s) {
    const IndexT *key = edge_index + (sort_row ? ld_edge : 0);
    const IndexT *other = edge_index + (sort_row ? 0 : ld_edge);
    int32_t *sorted_flag = scratch + scratch_elems;   // one int past the scan scratch
    cudaMemsetAsync(cnt, 0, static_cast<size_t>(num_rows + 1) * sizeof(int32_t), s);
    cudaMemsetAsync(status, 0, sizeof(int32_t), s);
    if (num_edges == 0) {
        cudaMemsetAsync(rowptr, 0, static_cast<size_t>(num_rows + 1) * sizeof(int32_t), s);
        HGIN_CHECK_LAUNCH("hgin_csr_build");
        return HGIN_OK;
    }
    const int edge_grid = grid_for(num_edges, 256 * 4, 8);
    // The sortedness probe is only worth its pass over the edge list where the reference's edge
    // order makes it succeed: builds keyed by SOURCE (sort_row == 0).  Destination-keyed builds go
    // straight to the general path (flag = 0).
    if (sort_row == 0) {
        cudaMemsetAsync(sorted_flag, 0xff, sizeof(int32_t), s);   // non-zero = sorted until a descent is seen
        csr_check_sorted<IndexT><<<edge_grid, 256, 0, s>>>(key, other, num_edges, num_rows, num_cols, sorted_flag);
        csr_from_sorted<IndexT><<<edge_grid, 256, 0, s>>>(key, other, num_edges, num_rows, num_cols, rowptr, col, perm,
                                                           sorted_flag);
    } else {
        cudaMemsetAsync(sorted_flag, 0, sizeof(int32_t), s);
    }
    csr_histogram<IndexT><<<edge_grid, 256, 0, s>>>(key, other, num_edges, num_rows, num_cols, cnt, status, sorted_flag);
    if (!exclusive_scan(cnt, rowptr, num_rows + 1, scratch, scratch_elems, sorted_flag, s))
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_csr_build: scan scratch too small");
    csr_fill<IndexT><<<edge_grid, 256, 0, s>>>(key, other, num_edges, num_rows, num_cols, rowptr, cnt, seg, sorted_flag);
    if (num_edges <= 4 * num_rows)   // short rows: 8 lanes per row, four rows per warp
        csr_sort_rows<IndexT, 8><<<grid_for(num_rows, 32, 8), 256, 0, s>>>(other, num_rows, rowptr, seg, col, perm, sorted_flag);
    else
        csr_sort_rows<IndexT, 32><<<grid_for(num_rows, 8, 8), 256, 0, s>>>(other, num_rows, rowptr, seg, col, perm, sorted_flag);
    HGIN_CHECK_LAUNCH("hgin_csr_build");
    return HGIN_OK;
}

}  // namespace
}  // namespace hgin

extern "C" int64_t hgin_csr_workspace_bytes(int64_t num_edges, int64_t num_rows) {
    using namespace hgin;
    if (num_edges < 0 || num_rows < 0) return -1;
    const int64_t cnt = align_up((num_rows + 1) * 4, 256);
    const int64_t seg = align_up(num_edges * 4, 256);
    const int64_t scr = align_up((scan_scratch_elems(num_rows + 1) + 1) * 4, 256);
    return cnt + seg + scr;
}

extern "C" int32_t hgin_csr_build(const void *edge_index, int32_t index_bytes, int64_t num_edges, int64_t ld_edge,
                                  int32_t sort_row, int64_t num_rows, int64_t num_cols, int32_t *rowptr,
                                  int32_t *col, int32_t *perm, int32_t *status, void *workspace,
                                  int64_t workspace_bytes, void *stream) {
    using namespace hgin;
    HGIN_CHECK_ARG(index_bytes == 4 || index_bytes == 8, "hgin_csr_build: index_bytes must be 4 or 8, got %d", index_bytes);
    HGIN_CHECK_ARG(num_edges >= 0 && num_rows >= 0 && num_cols >= 0, "hgin_csr_build: negative size");
    HGIN_CHECK_ARG(num_edges < INT32_MAX && num_rows < INT32_MAX && num_cols < INT32_MAX,
                   "hgin_csr_build: sizes must fit int32 (edges %lld rows %lld cols %lld)", (long long)num_edges,
                   (long long)num_rows, (long long)num_cols);
    HGIN_CHECK_ARG(sort_row == 0 || sort_row == 1, "hgin_csr_build: sort_row must be 0 or 1");
    HGIN_CHECK_ARG(ld_edge >= num_edges, "hgin_csr_build: ld_edge < num_edges");
    HGIN_CHECK_ARG(rowptr && status && workspace, "hgin_csr_build: null rowptr/status/workspace");
    HGIN_CHECK_ARG(num_edges == 0 || (edge_index && col), "hgin_csr_build: null edge_index/col");
    const int64_t need = hgin_csr_workspace_bytes(num_edges, num_rows);
    if (workspace_bytes < need)
        return fail(HGIN_ERR_WORKSPACE_TOO_SMALL, "hgin_csr_build: workspace %lld < %lld bytes",
                    (long long)workspace_bytes, (long long)need);
    char *ws = static_cast<char *>(workspace);
    int32_t *cnt = reinterpret_cast<int32_t *>(ws);
    ws += align_up((num_rows + 1) * 4, 256);
    int32_t *seg = reinterpret_cast<int32_t *>(ws);
    ws += align_up(num_edges * 4, 256);
    int32_t *scratch = reinterpret_cast<int32_t *>(ws);
    const int64_t scratch_elems = scan_scratch_elems(num_rows + 1);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (index_bytes == 8)
        return csr_build_impl<int64_t>(static_cast<const int64_t *>(edge_index), num_edges, ld_edge, sort_row,
                                       num_rows, num_cols, rowptr, col, perm, status, cnt, seg, scratch,
                                       scratch_elems, s);
    return csr_build_impl<int32_t>(static_cast<const int32_t *>(edge_index), num_edges, ld_edge, sort_row, num_rows,
                                   num_cols, rowptr, col, perm, status, cnt, seg, scratch, scratch_elems, s);
}
