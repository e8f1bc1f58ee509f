// K1 / K4 for LONG rows on block-diagonal batches: output-major gather from source rows staged in shared memory.
//
// The path->link relation of a datanet batch has ~36 neighbours per output row and every 512-byte path row is gathered
// ~2.9 times (once per hop of the path); the row-gather kernel pulls those re-reads through L2 and is bound by the
// L2->SM path (gin_combine.cu, DESIGN.md "Long rows").  A batch is block-diagonal (one block per topology sample,
// contiguous ids on both sides), so here one CTA per SM takes a block at a time and
//   * keeps the block's <= 224 output rows as fp32 accumulators in REGISTERS (warp w owns rows w, w + 32, ...; a lane
//     holds four features of each) — unlike the input-major streaming kernel (gin_scatter_blocks.cuh), whose
//     accumulators live in shared memory and cost a read-modify-write per edge;
//   * streams the block's source rows ONCE, in ascending order, through two ~108 KB shared-memory stages filled by
//     cp.async.bulk (one bulk copy per chunk of 216 fp32 / 432 bf16 rows, completion on an mbarrier), the copy of chunk
//     k + 1 (of this block or the next one) overlapping the gathers of chunk k;
//   * per chunk, every warp visits its rows: the lane-held window of the row's neighbour list (32 indices, refilled
//     asynchronously when consumed) is ballot-tested against the chunk's upper bound, and the neighbours inside the chunk
//     — a prefix, because the list is ascending — are added left to right from shared memory (LDS.128 per lane).
// Every source row crosses HBM -> SM exactly once, the ~2.9 re-reads are shared-memory loads, and the additions happen in
// CSR order: bit-identical to the gather kernel and to the CPU scatter_add_ whenever every row of the output-major CSR
// lists its neighbours in non-decreasing order, which hgin_block_gate verifies on the device (gate[3] == 0; the
// reference ships every relation grouped by source in ascending id order).  When the gate is closed — or a block has more
// output rows than the register tile holds — this kernel returns at once and the gather kernel, launched behind it with
// the inverse gate, does the work: a static, capturable launch sequence.
#pragma once

#include <cuda_bf16.h>

#include "gin_scatter_blocks.cuh"
#include "hgin_common.cuh"
#include "tc_common.cuh"

namespace hgin {
namespace staged {

using scatter::bulk_load;
using scatter::ldg_row4;
using scatter::lds_row4;
using scatter::stg_row4;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int SG_WARPS = 32;
constexpr int SG_THREADS = SG_WARPS * 32;
constexpr int SG_MAXR = 7;                          // output rows per warp
constexpr int SG_CAP_ROWS = SG_WARPS * SG_MAXR;     // 224 output rows per block
constexpr int SG_STAGE_BYTES = 108 * 1024;
constexpr int SG_SMEM = 2 * SG_STAGE_BYTES + 2 * SG_CAP_ROWS * 4 + 64;   // stages, list cursors / ends, barriers

struct SgParams {
    int num_blocks;
    const int64_t *in_ptr;    // [num_blocks + 1] first source row of every block
    const int64_t *out_ptr;   // [num_blocks + 1] first output row of every block
    const int32_t *rowptr;    // CSR_A: rows = output rows, cols = source rows (ascending inside a row)
    const int32_t *col;
    int num_edges;
    const int32_t *gate;      // [4]: see block_gate_kernel
    const void *x_in;         // [N_in, f] contiguous rows
    int f;
    const void *x_self;
    int ld_self;
    const float *eps;
    int self_mode;
    int accumulate;
    void *out;
    int ld_out;
    int in_act;
    const float *in_alpha;
    int self_act;
    const float *self_alpha;
    int debug;                // probe knobs (HGIN_SG_DEBUG): 1 skip the gathers, 2 skip the copies, 4 copies in 16 KB pieces
};

template <typename T, bool PRE>
__global__ void __launch_bounds__(SG_THREADS, 1) stage_blocks_kernel(const SgParams p) {
    if (__ldg(p.gate) != 0 || __ldg(p.gate + 3) != 0 || __ldg(p.gate + 1) > SG_CAP_ROWS) return;

    extern __shared__ __align__(128) uint8_t smem[];
    int *cur_s = reinterpret_cast<int *>(smem + 2 * SG_STAGE_BYTES);   // first neighbour of output row l not yet added
    int *end_s = cur_s + SG_CAP_ROWS;
    uint64_t *full = reinterpret_cast<uint64_t *>(end_s + SG_CAP_ROWS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned all = 0xffffffffu;
    constexpr int es = static_cast<int>(sizeof(T));
    const int f = p.f;
    const int row_bytes = f * es;
    const int chunk_rows = SG_STAGE_BYTES / row_bytes;
    const bool lane_on = lane * 4 < f;
    const T *x_in = static_cast<const T *>(p.x_in);
    const T *x_self = static_cast<const T *>(p.x_self);
    T *out = static_cast<T *>(p.out);
    const float ope = __fadd_rn(1.0f, p.eps ? __ldg(p.eps) : 0.0f);
    const float in_alpha = p.in_act == HGIN_ACT_PRELU ? __ldg(p.in_alpha) : 0.0f;
    const bool pre_self = p.self_act != HGIN_ACT_NONE;
    const float self_alpha = p.self_act == HGIN_ACT_PRELU ? __ldg(p.self_alpha) : 0.0f;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        tc::fence_barrier_init();
    }
    __syncthreads();

    // producer (thread 0): the chunks of this CTA's blocks in order, two in flight
    int pb = blockIdx.x, pk = 0;
    uint32_t pi = 0;
    auto produce = [&]() {
        while (pb < p.num_blocks) {
            const int64_t r_beg = __ldg(p.in_ptr + pb), r_end = __ldg(p.in_ptr + pb + 1);
            const int64_t r0 = r_beg + static_cast<int64_t>(pk) * chunk_rows;
            if (r0 < r_end) {
                const int n = static_cast<int>(r_end - r0 < chunk_rows ? r_end - r0 : chunk_rows);
                const uint32_t bytes = static_cast<uint32_t>(n) * row_bytes;
                const int s = pi & 1;
                if (p.debug & 2) {
                } else if (p.debug & 4) {
                    mbar_expect_tx(&full[s], bytes);
                    for (uint32_t o = 0; o < bytes; o += 16384)
                        bulk_load(smem_u32(smem + s * SG_STAGE_BYTES) + o, reinterpret_cast<const uint8_t *>(x_in + r0 * f) + o,
                                  min(16384u, bytes - o), &full[s]);
                } else {
                    mbar_expect_tx(&full[s], bytes);
                    bulk_load(smem_u32(smem + s * SG_STAGE_BYTES), x_in + r0 * f, bytes, &full[s]);
                }
                ++pi;
                ++pk;
                return;
            }
            pb += gridDim.x;
            pk = 0;
        }
    };
    if (threadIdx.x == 0) {
        produce();
        produce();
    }

    uint32_t ci = 0;
    for (int b = blockIdx.x; b < p.num_blocks; b += gridDim.x) {
        const int in0 = static_cast<int>(__ldg(p.in_ptr + b)), in1 = static_cast<int>(__ldg(p.in_ptr + b + 1));
        const int out0 = static_cast<int>(__ldg(p.out_ptr + b)), nout = static_cast<int>(__ldg(p.out_ptr + b + 1)) - out0;
        float acc[SG_MAXR][4];
        int32_t win[SG_MAXR];  // this lane's entry of the window col[cur + lane] (INT32_MAX past the row's end)
#pragma unroll
        for (int j = 0; j < SG_MAXR; ++j) {
            acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
            const int l = warp + SG_WARPS * j;
            win[j] = INT32_MAX;
            if (l < nout) {      // (cur_s / end_s entries of row l are private to this warp)
                const int beg = __ldg(p.rowptr + out0 + l), end = __ldg(p.rowptr + out0 + l + 1);
                if (lane == 0) {
                    cur_s[l] = beg;
                    end_s[l] = end;
                }
                if (beg + lane < end) win[j] = __ldg(p.col + beg + lane);
            }
        }
        __syncwarp();
        for (int c0 = in0; c0 < in1; c0 += chunk_rows, ++ci) {
            const int c1 = min(c0 + chunk_rows, in1);
            const int s = ci & 1;
            if (!(p.debug & 2)) mbar_wait(&full[s], (ci >> 1) & 1);
            const uint32_t base = smem_u32(smem + s * SG_STAGE_BYTES) + lane * 4 * es - static_cast<uint32_t>(c0) * row_bytes;
            // (the body is kept small on purpose: it is replicated SG_MAXR times — register accumulators need static
            // indices — and a first version with 14 rows per warp and four gathers in flight ran out of instruction cache)
#pragma unroll
            for (int j = 0; j < SG_MAXR; ++j) {
                const int l = warp + SG_WARPS * j;
                if (l < nout && !(p.debug & 1)) {          // warp-uniform
                    while (true) {
                        const int n = __popc(__ballot_sync(all, win[j] < c1));     // ascending list: a prefix of the window
#pragma unroll 1
                        for (int t = 0; t < n; t += 2) {                           // two gathers in flight, added in order
                            const int s0 = __shfl_sync(all, win[j], t), s1 = __shfl_sync(all, win[j], (t + 1) & 31);
                            float v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f};
                            const bool two = t + 1 < n;
                            if (lane_on) {
                                lds_row4<T>(base + static_cast<uint32_t>(s0) * row_bytes, v0);
                                if (two) lds_row4<T>(base + static_cast<uint32_t>(s1) * row_bytes, v1);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float tv = v0[i];
                                if (PRE) tv = tv > 0.f ? tv : in_alpha * tv;
                                acc[j][i] = __fadd_rn(acc[j][i], tv);
                            }
                            if (two) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    float tv = v1[i];
                                    if (PRE) tv = tv > 0.f ? tv : in_alpha * tv;
                                    acc[j][i] = __fadd_rn(acc[j][i], tv);
                                }
                            }
                        }
                        if (n == 0) break;
                        // refill the window behind the consumed prefix (the load lands while other rows are visited)
                        const int cur = cur_s[l] + n;
                        const int end = end_s[l];
                        __syncwarp();
                        if (lane == 0) cur_s[l] = cur;
                        win[j] = INT32_MAX;
                        if (cur + lane < end) win[j] = __ldg(p.col + cur + lane);
                        if (n < 32) break;
                        __syncwarp();
                    }
                }
            }
            __syncthreads();                 // every warp is done with stage s
            if (threadIdx.x == 0) produce();
        }
        // epilogue: out[row] (+)= acc + (1 + eps) * x_self[row]; every load is issued ahead of the first store
        // (in two groups of rows: 64 registers per thread)
#pragma unroll
        for (int g = 0; g < SG_MAXR; g += 4) {
            if (p.self_mode == HGIN_SELF_ADD) {
                float sv[4][4];
#pragma unroll
                for (int j = g; j < g + 4 && j < SG_MAXR; ++j) {
                    const int l = warp + SG_WARPS * j;
                    if (l < nout && lane_on) ldg_row4<T>(x_self + static_cast<int64_t>(out0 + l) * p.ld_self + lane * 4, sv[j - g], false);
                }
#pragma unroll
                for (int j = g; j < g + 4 && j < SG_MAXR; ++j) {
                    if (warp + SG_WARPS * j < nout && lane_on) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float xs = sv[j - g][i];
                            if (pre_self) xs = xs > 0.f ? xs : self_alpha * xs;
                            acc[j][i] = __fadd_rn(acc[j][i], __fmul_rn(ope, xs));
                        }
                    }
                }
            }
            if (p.accumulate) {
                float old[4][4];
#pragma unroll
                for (int j = g; j < g + 4 && j < SG_MAXR; ++j) {
                    const int l = warp + SG_WARPS * j;
                    if (l < nout && lane_on) ldg_row4<T>(out + static_cast<int64_t>(out0 + l) * p.ld_out + lane * 4, old[j - g], true);
                }
#pragma unroll
                for (int j = g; j < g + 4 && j < SG_MAXR; ++j) {
                    if (warp + SG_WARPS * j < nout && lane_on) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[j][i] = __fadd_rn(old[j - g][i], acc[j][i]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SG_MAXR; ++j) {
            const int l = warp + SG_WARPS * j;
            if (l < nout && lane_on) stg_row4(out + static_cast<int64_t>(out0 + l) * p.ld_out + lane * 4, acc[j]);
        }
    }
}

}  // namespace staged
}  // namespace hgin
