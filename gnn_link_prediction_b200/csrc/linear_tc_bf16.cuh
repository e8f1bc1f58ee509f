// Tensor-core dense-layer kernels on bf16 rows (HGIN_DTYPE_BF16 / HGIN_MATH_BF16): tcgen05.mma kind::f16 with
// bf16 operands, fp32 accumulation in TMEM, bf16 activations / gradients in HBM.  Same warp-specialised
// structure as the tf32 kernels of linear_tc.cuh (TMA producer, one MMA-issuing thread, epilogue-operand
// producer, TMEM allocator, two epilogue warpgroups); what changes with 2-byte elements:
//
//   * a 128-byte swizzle row holds 64 elements, so a K-block is 64 wide (K = 128 is TWO 16 KB K-blocks instead
//     of four) and one MMA consumes UMMA_K = 16 elements (the same 32 bytes of start-address advance);
//   * W (<= 128 x 128) takes 32 KB instead of 64 KB, which pays for a 6-deep A ring: three row tiles (96 KB) of
//     operand reads in flight per SM instead of one;
//   * the epilogue works on 64-column chunks (a full 128-byte staging row per tile row): warpgroup g owns
//     chunk g, reads it as two tcgen05.ld.32x32b.x32 halves, does bias / activation / merge / post-activation
//     in fp32 registers, rounds once to bf16 and stores packed 16-byte pieces into the swizzled staging tile,
//     which one TMA store per chunk writes out;
//   * the weight gradient contracts 64 rows per ring stage (8 KB boxes), 4 MMAs of 16 rows each; its operands
//     are MN-major bf16 in the plain SWIZZLE_128B layout (LBO = box pitch, SBO = 1024 B; pinned by
//     tests/test_ops_gpu.py::test_tn_bf16_descriptor_is_exact_layout through hgin_debug_gemm_tn_bf16).
// Every layer is still HBM-bound: the point of bf16 here is half the bytes per row, not tensor throughput.
#pragma once

#include <cuda_bf16.h>

#include "linear_tc.cuh"

namespace hgin {
namespace tcgemm {

using bf16 = __nv_bfloat16;

constexpr int KB16 = 64;              // bf16 per 128-byte swizzle row = one K-block
constexpr int UMMA_K16 = 16;
constexpr int NT16_STAGES = 6;        // A ring, K-block granularity (K = 128: three tiles in flight)
constexpr int CW16 = 64;              // epilogue chunk width (columns) = one 128-byte staging row

__device__ __forceinline__ uint4 lds_v4u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void unpack_bf16x8(uint4 q, float *o) {
    o[0] = __uint_as_float(q.x << 16); o[1] = __uint_as_float(q.x & 0xffff0000u);
    o[2] = __uint_as_float(q.y << 16); o[3] = __uint_as_float(q.y & 0xffff0000u);
    o[4] = __uint_as_float(q.z << 16); o[5] = __uint_as_float(q.z & 0xffff0000u);
    o[6] = __uint_as_float(q.w << 16); o[7] = __uint_as_float(q.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float *o) {
    return make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
}

// Instruction descriptor for kind::f16 with bf16 operands, fp32 accumulate: c_format F32 = 1 at [4,6),
// a/b format BF16 = 1 at [7,10)/[10,13); the remaining fields as in make_idesc_tf32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Nt16Smem {
    static constexpr int kB = 2 * TILE_BYTES;             // W: up to 2 K-blocks of [128 x 128 B]
    static constexpr int kA = NT16_STAGES * TILE_BYTES;   // A ring
    static constexpr int kStage = 4 * TILE_BYTES;         // staging: (out, z) chunk per epilogue group
    static constexpr int kE = 2 * TILE_BYTES;             // epilogue-operand ring (one buffer per group)
    static constexpr int off_b = 0;
    static constexpr int off_a = off_b + kB;
    static constexpr int off_stage = off_a + kA;
    static constexpr int off_e = off_stage + kStage;
    static constexpr int off_small = off_e + kE;
    static constexpr int small_bytes = 128 * 4 /*bias*/ + 128 * 4 * 4 /*w_tail*/ + 32 * 8 /*barriers*/ + 64;
    static constexpr int total = off_small + small_bytes;
    static_assert(total <= 227 * 1024, "gemm_nt_bf16 shared memory exceeds the sm_100 per-CTA limit");
};

// D[M x N] = A[M x K] * B[N x K]^T, A / B bf16 K-major; outputs, the epilogue operand and the post-activation
// rows are bf16; NtParams as for the tf32 kernel (num_kb counts 64-wide K-blocks).
template <int EPI>
__global__ void __launch_bounds__(NT_THREADS, 1)
gemm_nt_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1,
                    const __grid_constant__ CUtensorMap tm_e, const NtParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    uint8_t *smem_b = smem + Nt16Smem::off_b;
    uint8_t *smem_a = smem + Nt16Smem::off_a;
    uint8_t *smem_stage = smem + Nt16Smem::off_stage;
    uint8_t *smem_e = smem + Nt16Smem::off_e;
    float *bias_s = reinterpret_cast<float *>(smem + Nt16Smem::off_small);
    float *wtail_s = bias_s + 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wtail_s + 128 * 4);
    uint64_t *full = bars;                     // [NT16_STAGES]
    uint64_t *empty = bars + NT16_STAGES;      // [NT16_STAGES]
    uint64_t *b_full = bars + 2 * NT16_STAGES; // [1]
    uint64_t *tmem_full = b_full + 1;          // [2]
    uint64_t *tmem_empty = tmem_full + 2;      // [2]
    uint64_t *e_full = tmem_empty + 2;         // [2]
    uint64_t *e_empty = e_full + 2;            // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(e_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nchunks = (p.n + CW16 - 1) / CW16;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_b);
        prefetch_tmap(&tm_o0);
        if (EPI == EPI_FWD && p.want_z) prefetch_tmap(&tm_o1);
        if (p.use_e) prefetch_tmap(&tm_e);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NT16_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);   // one arrive per epilogue warp (2 groups x 4)
            mbar_init(&e_full[i], 1);
            mbar_init(&e_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<TMEM_COLS>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== operand producer =====
        if (lane == 0) {
            const uint32_t b_bytes = static_cast<uint32_t>(p.n) * 128u;
            mbar_expect_tx(b_full, b_bytes * p.num_kb);
            for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(smem_b + kb * TILE_BYTES, &tm_b, b_full, kb * KB16, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], TILE_BYTES);
                    tma_load_2d(smem_a + s * TILE_BYTES, &tm_a, &full[s], kb * KB16, tile * BM);
                    if (++s == NT16_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(BM, p.n, 0, 0);
            mbar_wait(b_full, 0);
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 128;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(smem_a + s * TILE_BYTES);
                    const uint32_t b_base = smem_u32(smem_b + kb * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < KB16 / UMMA_K16; ++k) {
                        umma_bf16(d_tmem, make_smem_desc(a_base + k * UMMA_K16 * 2, 16, 1024),
                                  make_smem_desc(b_base + k * UMMA_K16 * 2, 16, 1024), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[s]);   // ring slot reusable once these MMAs have read it
                    if (++s == NT16_STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp == 2) {
        // ===== epilogue-operand producer: chunk c goes to buffer c & 1, consumed by epilogue group c & 1 =====
        if (lane == 0 && p.use_e) {
            uint32_t ph[2] = {0, 0};
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int c = 0; c < nchunks; ++c) {
                    const int b = c & 1;
                    mbar_wait(&e_empty[b], ph[b] ^ 1);
                    mbar_expect_tx(&e_full[b], TILE_BYTES);
                    tma_load_2d(smem_e + b * TILE_BYTES, &tm_e, &e_full[b], c * CW16, tile * BM);
                    ph[b] ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: group 0 (warps 4-7) takes 64-column chunks 0, 2; group 1 (warps 8-11) chunks 1, 3 =====
        const int grp = (warp - 4) >> 2;
        const int q = (warp - 4) & 3;           // TMEM lane quarter == row block inside the tile
        const int gt = threadIdx.x - 128 - grp * 128;   // 0..127 inside the group
        const int et = threadIdx.x - 128;       // 0..255 over both groups
        const int r = q * 32 + lane;            // row inside the tile owned by this thread
        const int bar_id = EPI_BAR + grp;
        if (EPI == EPI_FWD) {
            for (int i = et; i < 128; i += 256) bias_s[i] = (p.bias && i < p.n) ? __ldg(p.bias + i) : 0.0f;
            for (int i = et; i < 128 * 4; i += 256) {
                const int nn = i >> 2, t = i & 3;
                wtail_s[i] = (nn < p.n && t < p.k2) ? __ldg(p.w_tail + static_cast<int64_t>(nn) * p.ldw + t) : 0.0f;
            }
            named_barrier(EPI_ALL_BAR, 256);
        }
        const float alpha = (p.act == HGIN_ACT_PRELU) ? __ldg(p.alpha) : 0.0f;
        const float self_scale = (EPI == EPI_DX && p.self_eps) ? __fadd_rn(1.0f, __ldg(p.self_eps)) : 1.0f;
        float dot = 0.0f, dot2 = 0.0f;
        const uint32_t sa0 = smem_u32(smem_stage + grp * 2 * TILE_BYTES);   // out / dx
        const uint32_t sa1 = sa0 + TILE_BYTES;                               // z
        const uint32_t eb_ptr = smem_u32(smem_e + grp * TILE_BYTES);
        const int my_chunks = (nchunks - grp + 1) / 2;   // chunks grp, grp + 2, ...
        const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
        const uint32_t rx = static_cast<uint32_t>(r) & 7u;
        int it = 0;
        uint32_t eph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int64_t grow = static_cast<int64_t>(tile) * BM + r;
            float x2v[4] = {0.f, 0.f, 0.f, 0.f};
            if (EPI == EPI_FWD && p.k2 > 0 && grow < p.rows) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t < p.k2) x2v[t] = __ldg(p.x2 + grow * p.ld2 + t);
            }
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tcgen05_fence_after();
            if (my_chunks == 0) {   // nothing to read for this group: release the accumulator at once
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
            for (int c = grp; c < nchunks; c += 2) {
                if (p.use_e) mbar_wait(&e_full[grp], eph);   // chunk c was loaded into buffer c & 1 == grp
                // the group's staging tiles are free once its previous TMA stores have read them
                if (gt == 0) tma_store_wait_read<0>();
                named_barrier(bar_id, 128);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col0 = c * CW16 + h * 32;
                    if (col0 >= p.n) break;                  // (warp-uniform)
                    float v[32];
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + col0, v);
                    if (h == 1 || col0 + 32 >= p.n) {
                        if (c + 2 >= nchunks) {              // this group's last read of the tile: hand the accumulator back
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                        }
                    }
                    float ev[32];
                    if (p.use_e) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            unpack_bf16x8(lds_v4u(eb_ptr + row_off + (((static_cast<uint32_t>(h * 4 + j)) ^ rx) << 4)), ev + j * 8);
                    }
                    float o[32];
                    if (EPI == EPI_FWD) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int nn = col0 + j;
                            float zz = v[j] + bias_s[nn & 127];
                            if (p.k2 > 0) {      // rank-k2 update from the extra fp32 input columns (readout layer 1 only)
                                zz = fmaf(x2v[0], wtail_s[(nn & 127) * 4 + 0], zz);
                                zz = fmaf(x2v[1], wtail_s[(nn & 127) * 4 + 1], zz);
                                zz = fmaf(x2v[2], wtail_s[(nn & 127) * 4 + 2], zz);
                                zz = fmaf(x2v[3], wtail_s[(nn & 127) * 4 + 3], zz);
                            }
                            if (nn >= p.n) zz = 0.0f;
                            v[j] = zz;
                            if (p.want_out) {    // (a lazily activated layer stores z only)
                                float oo = act_forward(zz, p.act, alpha);
                                if (p.use_e) oo += ev[j];
                                o[j] = oo;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j >= p.n) v[j] = 0.0f;            // columns past n hold stale TMEM
                        if (p.use_e == 1) {
                            if (grow < p.rows) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) dot = fmaf(v[j], ev[j], dot);  // OOB columns of e are zero-filled
                            }
                        } else if (p.use_e == 2) {
                            // rows past the end: v == 0 (zero-filled A rows), so they add nothing to the sums
                            if (p.dot2_partials) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) dot2 = fmaf(v[j], act_forward(ev[j], p.act, alpha), dot2);
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float rr = p.self_eps ? __fmul_rn(self_scale, v[j]) : v[j];
                                if (p.act == HGIN_ACT_PRELU && !(ev[j] > 0.f)) dot = fmaf(rr, ev[j], dot);
                                v[j] = act_backward(rr, ev[j], p.act, alpha);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t off = row_off + (((static_cast<uint32_t>(h * 4 + j)) ^ rx) << 4);
                        if (EPI == EPI_FWD) {
                            if (p.want_out) sts_v4u(sa0 + off, pack_bf16x8(o + j * 8));
                            if (p.want_z) sts_v4u(sa1 + off, pack_bf16x8(v + j * 8));
                        } else {
                            sts_v4u(sa0 + off, pack_bf16x8(v + j * 8));
                        }
                    }
                }
                if (p.use_e) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&e_empty[grp]);
                    eph ^= 1;
                }
                fence_proxy_async_smem();
                named_barrier(bar_id, 128);
                if (gt == 0) {
                    if (EPI == EPI_FWD) {
                        if (p.want_out) tma_store_2d(&tm_o0, smem_stage + grp * 2 * TILE_BYTES, c * CW16, tile * BM);
                        if (p.want_z) tma_store_2d(&tm_o1, smem_stage + grp * 2 * TILE_BYTES + TILE_BYTES, c * CW16, tile * BM);
                    } else if (p.want_out) {
                        tma_store_2d(&tm_o0, smem_stage + grp * 2 * TILE_BYTES, c * CW16, tile * BM);
                    }
                    tma_store_commit();
                }
            }
        }
        if (gt == 0) tma_store_wait<0>();
        if (EPI == EPI_DX && (p.dot_partials || p.dot2_partials)) {
            dot = warp_sum(dot);
            dot2 = warp_sum(dot2);
            float *red = bias_s;  // unused by EPI_DX
            if (lane == 0) {
                red[grp * 4 + q] = dot;
                red[8 + grp * 4 + q] = dot2;
            }
            named_barrier(EPI_ALL_BAR, 256);
            if (et == 0 && p.dot_partials)
                p.dot_partials[blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
            if (et == 1 && p.dot2_partials)
                p.dot2_partials[blockIdx.x] = ((red[8] + red[9]) + (red[10] + red[11])) + ((red[12] + red[13]) + (red[14] + red[15]));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) {
        tcgen05_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ---- weight gradient: D[n][k] = sum_m A[m][n] * B[m][k], both operands bf16 MN-major -------------------
constexpr int TN16_ROWS = 64;                            // contraction rows per ring stage
constexpr int TN16_STAGES = 5;
constexpr int TN16_BOX_BYTES = TN16_ROWS * 128;          // one [64 rows x 64 cols] box = 8 KB
constexpr int TN16_OPERAND_BYTES = 2 * TN16_BOX_BYTES;   // up to 128 columns = 2 boxes = 16 KB

struct Tn16Smem {
    static constexpr int stage_bytes = 2 * TN16_OPERAND_BYTES + TN16_BOX_BYTES;    // A, B, the ones box
    static constexpr int off_ring = 0;
    static constexpr int off_small = TN16_STAGES * stage_bytes;
    static constexpr int total = off_small + 32 * 8 + 64 + 1024;
    static_assert(total <= 227 * 1024, "gemm_tn_bf16 shared memory exceeds the sm_100 per-CTA limit");
};

// TnParams as for the tf32 kernel; ones_col needs k % 64 == 0 (the ones box follows B's last 64-column box).
__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *ring = smem + Tn16Smem::off_ring;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Tn16Smem::off_small);
    uint64_t *full = bars;
    uint64_t *empty = bars + TN16_STAGES;
    uint64_t *done = bars + 2 * TN16_STAGES;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t m_beg = static_cast<int64_t>(blockIdx.x) * p.rows_per_cta;
    const int64_t m_end = min(m_beg + p.rows_per_cta, p.rows);
    const int steps = static_cast<int>((m_end - m_beg + TN16_ROWS - 1) / TN16_ROWS);
    const int a_boxes = (p.n + 63) / 64, b_boxes = (p.k + 63) / 64;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < TN16_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<256>(tmem_ptr);
    if (p.ones_col) {   // TMA never writes this box: fill it once per stage with bf16 ones
        for (int i = threadIdx.x; i < TN16_STAGES * (TN16_BOX_BYTES / 4); i += THREADS) {
            const int st = i / (TN16_BOX_BYTES / 4), w = i % (TN16_BOX_BYTES / 4);
            reinterpret_cast<uint32_t *>(ring + st * Tn16Smem::stage_bytes + TN16_OPERAND_BYTES + b_boxes * TN16_BOX_BYTES)[w] = 0x3f803f80u;
        }
        fence_proxy_async_smem();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < steps; ++i) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], (a_boxes + b_boxes) * TN16_BOX_BYTES);
                uint8_t *sa = ring + s * Tn16Smem::stage_bytes;
                uint8_t *sb = sa + TN16_OPERAND_BYTES;
                const int m = static_cast<int>(m_beg + static_cast<int64_t>(i) * TN16_ROWS);
                // CTA ranges are multiples of TN16_ROWS, so boxes never straddle two ranges; the last
                // box of the matrix is zero-filled past `rows` by TMA.
                for (int c = 0; c < a_boxes; ++c) tma_load_2d(sa + c * TN16_BOX_BYTES, &tm_a, &full[s], c * 64, m);
                for (int c = 0; c < b_boxes; ++c) tma_load_2d(sb + c * TN16_BOX_BYTES, &tm_b, &full[s], c * 64, m);
                if (++s == TN16_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.k + (p.ones_col ? 16 : 0), 1, 1);
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < steps; ++i) {
                mbar_wait(&full[s], ph);
                tcgen05_fence_after();
                const uint32_t a_base = smem_u32(ring + s * Tn16Smem::stage_bytes);
                const uint32_t b_base = a_base + TN16_OPERAND_BYTES;
#pragma unroll
                for (int j = 0; j < TN16_ROWS / UMMA_K16; ++j) {
                    umma_bf16(tmem_base, make_smem_desc(a_base + j * p.k_step_bytes, p.lbo, p.sbo, p.layout_type),
                              make_smem_desc(b_base + j * p.k_step_bytes, p.lbo, p.sbo, p.layout_type), idesc,
                              (i | j) != 0);
                }
                umma_commit(&empty[s]);
                if (++s == TN16_STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(done);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int nn = q * 32 + lane;  // output row (column of A) owned by this thread
        float *dst = p.partials + (static_cast<int64_t>(blockIdx.x) * p.n + nn) * p.k;
        const int kch = (p.k + 31) / 32;
        if (steps > 0) {
            mbar_wait(done, 0);
            tcgen05_fence_after();
            for (int c = 0; c < kch; ++c) {
                float v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
                if (nn < p.n) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j < p.k) dst[c * 32 + j] = v[j];
                }
            }
            if (p.ones_col) {
                float v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b_boxes * 64, v);
                if (nn < p.n) p.db_partials[static_cast<int64_t>(blockIdx.x) * p.n + nn] = v[0];
            }
        } else if (nn < p.n) {
            for (int j = 0; j < p.k; ++j) dst[j] = 0.0f;
            if (p.ones_col) p.db_partials[static_cast<int64_t>(blockIdx.x) * p.n + nn] = 0.0f;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) {
        tcgen05_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

}  // namespace tcgemm
}  // namespace hgin
