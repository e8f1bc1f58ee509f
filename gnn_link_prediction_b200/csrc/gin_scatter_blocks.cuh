// K1 / K4 for LONG rows on block-diagonal batches: input-major streaming ("scatter into shared memory").
//
// The path->link relation of a datanet batch has ~36 (up to hundreds of) neighbours per output row; the row-gather
// kernel pulls every 512-byte input row through L2 ~2.9 times and is bound by the L2->SM slice bandwidth (ncu: 11 TB/s
// of L2 traffic for 4.6 TB/s of DRAM traffic).  A batch is block-diagonal (one block per topology sample, contiguous
// ids on both sides), so a CTA can instead take ONE block at a time, keep the block's <= ~280 output rows as fp32
// accumulators in shared memory, and stream the block's input rows ONCE, in order, through a ring of bulk copies
// (cp.async.bulk + mbarrier): every input row crosses HBM -> SM exactly once and L2 is not re-read.
//
// Order of the additions.  Output row l receives its addends in ascending INPUT row order.  The row-gather kernel adds
// them in the stable edge order of the adjacency it walks (CSR_A, rows = outputs).  The two orders coincide iff every
// row of CSR_A lists its neighbours in non-decreasing order, which hgin_block_gate verifies on the device (the
// reference ships every relation grouped by source in ascending id order, so it holds for its batches); then this
// kernel is bit-identical to the gather kernel and to the CPU scatter_add_.  When the gate is closed (or a block has
// more output rows than fit) this kernel returns at once and the gather kernel, launched behind it with the inverse
// gate, does the work: the launch sequence is static (CUDA-graph capturable).
//
// Work split inside the CTA: 16 consumer warps; warp w OWNS output rows l with l % 16 == w — it alone zeroes, updates
// and finally stores them, so the consumers never synchronise with each other.  Each warp scans the stage's edge list 32
// edges at a time (coalesced; the input row of an edge is found by a 6-step shuffle search over the stage's row
// pointers), ballots the edges whose output row it owns, and performs their read-modify-writes one after the other in
// edge order with the whole warp across the features (float4 per lane).  One extra warp issues the bulk copies.
#pragma once

#include <cuda_bf16.h>

#include "hgin_common.cuh"
#include "tc_common.cuh"

namespace hgin {
namespace scatter {

using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int SB_WARPS = 16;
constexpr int SB_THREADS = (SB_WARPS + 1) * 32;
constexpr int SB_STAGE_ROWS = 32;
constexpr int SB_SMEM = 224 * 1024;
constexpr int SB_RING_TARGET = 80 * 1024;
constexpr int SB_MAX_STAGES = 16;

struct SbParams {
    int num_blocks;
    const int64_t *in_ptr;    // [num_blocks + 1] first INPUT row of every block
    const int64_t *out_ptr;   // [num_blocks + 1] first OUTPUT row of every block
    const int32_t *rowptr;    // CSR_B: rows = input rows, cols = output rows
    const int32_t *col;
    const int32_t *gate;      // [4]: see block_gate_kernel
    int cap_rows;
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
    int stages;
    int stage_bytes;
};

__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void *gptr, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gptr), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T>
__device__ __forceinline__ void lds_row4(uint32_t addr, float (&v)[4]);
template <>
__device__ __forceinline__ void lds_row4<float>(uint32_t addr, float (&v)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
template <>
__device__ __forceinline__ void lds_row4<__nv_bfloat16>(uint32_t addr, float (&v)[4]) {
    uint32_t a, b;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
    v[0] = __uint_as_float(a << 16);
    v[1] = __uint_as_float(a & 0xffff0000u);
    v[2] = __uint_as_float(b << 16);
    v[3] = __uint_as_float(b & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void ldg_row4(const T *p, float (&v)[4], bool coherent);
template <>
__device__ __forceinline__ void ldg_row4<float>(const float *p, float (&v)[4], bool coherent) {
    const float4 t = coherent ? *reinterpret_cast<const float4 *>(p) : __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ldg_row4<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[4], bool coherent) {
    const uint2 t = coherent ? *reinterpret_cast<const uint2 *>(p) : __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(t.x << 16);
    v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16);
    v[3] = __uint_as_float(t.y & 0xffff0000u);
}
__device__ __forceinline__ void stg_row4(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void stg_row4(__nv_bfloat16 *p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 q;
    q.x = *reinterpret_cast<const uint32_t *>(&a);
    q.y = *reinterpret_cast<const uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = q;
}

template <typename T>
__global__ void __launch_bounds__(SB_THREADS, 1) scatter_blocks_kernel(const SbParams p) {
    // gate: streaming is only bit-identical to the gather kernel when CSR_A's rows are ascending, and only possible
    // when every block's output rows fit the accumulator tile
    if (__ldg(p.gate) != 0 || __ldg(p.gate + 3) != 0 || __ldg(p.gate + 1) > p.cap_rows) return;

    extern __shared__ __align__(128) uint8_t smem[];
    float *acc = reinterpret_cast<float *>(smem);
    const int acc_bytes = p.cap_rows * p.f * 4;
    uint8_t *ring = smem + ((acc_bytes + 127) & ~127);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + p.stages * p.stage_bytes);
    uint64_t *empty = full + SB_MAX_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned all = 0xffffffffu;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], SB_WARPS);
        }
        tc::fence_barrier_init();
    }
    __syncthreads();

    const int f = p.f;
    const int row_bytes = f * static_cast<int>(sizeof(T));
    if (warp == SB_WARPS) {
        // ===== producer: one bulk copy per stage (<= 32 consecutive input rows) =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int b = blockIdx.x; b < p.num_blocks; b += gridDim.x) {
                const int64_t r_beg = __ldg(p.in_ptr + b), r_end = __ldg(p.in_ptr + b + 1);
                for (int64_t r0 = r_beg; r0 < r_end; r0 += SB_STAGE_ROWS) {
                    const int n = static_cast<int>(r_end - r0 < SB_STAGE_ROWS ? r_end - r0 : SB_STAGE_ROWS);
                    mbar_wait(&empty[s], ph ^ 1);
                    const uint32_t bytes = static_cast<uint32_t>(n) * row_bytes;
                    mbar_expect_tx(&full[s], bytes);
                    bulk_load(smem_u32(ring + s * p.stage_bytes), static_cast<const T *>(p.x_in) + r0 * f, bytes, &full[s]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const bool on = lane * 4 < f;     // this lane's 4 features exist
    const float ope = __fadd_rn(1.0f, p.eps ? __ldg(p.eps) : 0.0f);
    const float in_alpha = p.in_act == HGIN_ACT_PRELU ? __ldg(p.in_alpha) : 0.0f;
    const float self_alpha = p.self_act == HGIN_ACT_PRELU ? __ldg(p.self_alpha) : 0.0f;
    const uint32_t acc_s = smem_u32(acc);
    for (int r = warp; r < p.cap_rows; r += SB_WARPS)
        if (on) tc::sts_v4(acc_s + (r * f + lane * 4) * 4, 0.f, 0.f, 0.f, 0.f);

    int s = 0;
    uint32_t ph = 0;
    for (int b = blockIdx.x; b < p.num_blocks; b += gridDim.x) {
        const int64_t r_beg = __ldg(p.in_ptr + b), r_end = __ldg(p.in_ptr + b + 1);
        const int64_t l0 = __ldg(p.out_ptr + b);
        const int L = static_cast<int>(__ldg(p.out_ptr + b + 1) - l0);
        for (int64_t r0 = r_beg; r0 < r_end; r0 += SB_STAGE_ROWS) {
            const int n = static_cast<int>(r_end - r0 < SB_STAGE_ROWS ? r_end - r0 : SB_STAGE_ROWS);
            // row pointers of the stage: lane t holds the END of input row r0 + t (lanes >= n repeat the last one)
            const int rp = __ldg(p.rowptr + r0 + (lane < n ? lane : n - 1) + 1);
            const int e0 = __ldg(p.rowptr + r0);
            const int e1 = __shfl_sync(all, rp, n - 1);
            mbar_wait(&full[s], ph);
            const uint32_t stage_s = smem_u32(ring + s * p.stage_bytes);
            for (int c0 = e0; c0 < e1; c0 += 32) {
                const int e = c0 + lane;
                const bool valid = e < e1;
                const int l = valid ? static_cast<int>(__ldg(p.col + e) - l0) : -1;
                // input row of edge e inside the stage = number of rows that end at or before e
                int lo = 0, hi = n;
#pragma unroll
                for (int it = 0; it < 6; ++it) {
                    const int mid = (lo + hi) >> 1;
                    const int v = __shfl_sync(all, rp, mid & 31);
                    if (lo < hi) {
                        if (v <= e) lo = mid + 1;
                        else hi = mid;
                    }
                }
                const bool own = valid && static_cast<unsigned>(l) < static_cast<unsigned>(L) && (l % SB_WARPS) == warp;
                unsigned mask = __ballot_sync(all, own);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int lj = __shfl_sync(all, l, j);
                    const int rj = __shfl_sync(all, lo, j);
                    if (on) {
                        float x[4];
                        lds_row4<T>(stage_s + (rj * f + lane * 4) * static_cast<int>(sizeof(T)), x);
                        const uint32_t a_addr = acc_s + (lj * f + lane * 4) * 4;
                        const float4 a = tc::lds_v4(a_addr);
                        if (p.in_act != HGIN_ACT_NONE) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) x[i] = x[i] > 0.f ? x[i] : in_alpha * x[i];
                        }
                        tc::sts_v4(a_addr, __fadd_rn(a.x, x[0]), __fadd_rn(a.y, x[1]), __fadd_rn(a.z, x[2]), __fadd_rn(a.w, x[3]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        // ===== this warp's output rows of the block: self term, merge, store; accumulators back to zero =====
        for (int r = warp; r < L; r += SB_WARPS) {
            if (!on) continue;
            const int64_t row = l0 + r;
            const uint32_t a_addr = acc_s + (r * f + lane * 4) * 4;
            const float4 a = tc::lds_v4(a_addr);
            float v[4] = {a.x, a.y, a.z, a.w};
            if (p.self_mode == HGIN_SELF_ADD) {
                float sv[4];
                ldg_row4<T>(static_cast<const T *>(p.x_self) + row * p.ld_self + lane * 4, sv, false);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float xs = sv[i];
                    if (p.self_act != HGIN_ACT_NONE) xs = xs > 0.f ? xs : self_alpha * xs;
                    v[i] = __fadd_rn(v[i], __fmul_rn(ope, xs));
                }
            }
            T *o = static_cast<T *>(p.out) + row * p.ld_out + lane * 4;
            if (p.accumulate) {
                float old[4];
                ldg_row4<T>(o, old, true);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = __fadd_rn(old[i], v[i]);
            }
            stg_row4(o, v);
            tc::sts_v4(a_addr, 0.f, 0.f, 0.f, 0.f);
        }
    }
}

// gate[0] += edges of CSR_B (rows = inputs) that leave their block, block tables that do not cover the rows;
// gate[1] = max output rows of a block;  gate[2] = max input rows of a block;
// gate[3] += rows of CSR_A (rows = outputs) whose neighbour list is not non-decreasing.  (gate is zeroed by the caller.)
__global__ void block_gate_kernel(int64_t rows_a, const int32_t *__restrict__ rowptr_a, const int32_t *__restrict__ col_a,
                                  int64_t rows_b, const int32_t *__restrict__ rowptr_b, const int32_t *__restrict__ col_b,
                                  int num_blocks, const int64_t *__restrict__ in_ptr, const int64_t *__restrict__ out_ptr,
                                  int32_t *gate) {
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    int bad = 0, unsorted = 0;
    for (int64_t r = tid; r < rows_a; r += stride) {
        const int b = __ldg(rowptr_a + r), e = __ldg(rowptr_a + r + 1);
        int prev = -1;
        for (int i = b; i < e; ++i) {
            const int c = __ldg(col_a + i);
            unsorted += c < prev;
            prev = c;
        }
    }
    for (int64_t r = tid; r < rows_b; r += stride) {
        int lo = 0, hi = num_blocks;        // block of input row r: last b with in_ptr[b] <= r
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(in_ptr + mid) <= r) lo = mid;
            else hi = mid;
        }
        const int64_t o0 = __ldg(out_ptr + lo), o1 = __ldg(out_ptr + lo + 1);
        const bool inside = r >= __ldg(in_ptr + lo) && r < __ldg(in_ptr + lo + 1);
        const int b = __ldg(rowptr_b + r), e = __ldg(rowptr_b + r + 1);
        if (!inside) bad += (e > b);
        for (int i = b; i < e; ++i) {
            const int c = __ldg(col_b + i);
            bad += (c < o0 || c >= o1);
        }
    }
    for (int64_t i = tid; i < num_blocks; i += stride) {
        atomicMax(gate + 1, static_cast<int>(__ldg(out_ptr + i + 1) - __ldg(out_ptr + i)));
        atomicMax(gate + 2, static_cast<int>(__ldg(in_ptr + i + 1) - __ldg(in_ptr + i)));
    }
    if (tid == 0 && (num_blocks == 0 || __ldg(in_ptr + num_blocks) != rows_b || __ldg(out_ptr + num_blocks) != rows_a ||
                     __ldg(in_ptr) != 0 || __ldg(out_ptr) != 0))
        bad += 1;
    if (bad) atomicAdd(gate, bad);     // integer counters: order-free
    if (unsorted) atomicAdd(gate + 3, unsorted);
}

inline int stage_bytes_for(int f, int elem) { return SB_STAGE_ROWS * f * elem; }
inline int stages_for(int f, int elem) {
    int st = SB_RING_TARGET / stage_bytes_for(f, elem);
    return st < 2 ? 2 : (st > SB_MAX_STAGES ? SB_MAX_STAGES : st);
}
inline int capacity_rows(int f, int elem) {
    const int ring = stages_for(f, elem) * stage_bytes_for(f, elem);
    return (SB_SMEM - ring - 2 * SB_MAX_STAGES * 8 - 256) / (f * 4);
}

}  // namespace scatter
}  // namespace hgin
