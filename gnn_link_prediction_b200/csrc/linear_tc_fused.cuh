// Fused dense-layer backward on tensor cores: ONE pass over g, z and x produces
//   dx = (g * act'(z)) W            (input gradient, optional sum(dx * dot_x) = d(eps))
//   dW = (g * act'(z))^T x          (per-CTA partial, reduced afterwards)
//   db, dalpha, the rank-k2 tail of dW
// instead of the three passes dz_prepare -> gemm_nt<EPI_DX> -> gemm_tn, which write dz once and read
// it twice (8 row-sized transfers per layer against 5 here).
//
// Per 128-row tile (persistent CTAs, 384 threads, whole TMEM = 512 columns):
//   warp 0      TMA producer: per 32-column k-block kb the triple (g_kb, z_kb, Wt_kb) into a 2-stage
//               ring, and the x tile [128 x k1] in the MN-major (32-byte swizzle atom) layout
//   warps 4-7   transform: warp 4+kb owns k-block kb.  Lane i owns dz column n = 32 kb + i, walks the
//               128 rows: dz = g * act'(z) is written back in place (the K-major A operand of the dx
//               MMA) and, 32 rows at a time, stored with tcgen05.st into TMEM lane n (the A operand
//               of the dW MMA: A^T lives in tensor memory, so no transposed copy of dz is needed).
//               db / dalpha / tail sums are thread-local because a thread owns its column.
//   warp 1      MMA issuer: D1[m][k] (+)= dz_kb (smem) * Wt_kb (smem) for the 4 k-blocks, then
//               D2[n][k] += dz^T (TMEM) * x (smem, MN-major) over the tile's 128 rows
//   warp 2      TMA producer for dot_x chunks (d(eps))
//   warp 3      TMEM allocation
//   warps 8-11  epilogue: D1 -> registers -> (dot) -> swizzled staging -> TMA store of dx; after the
//               last tile D2 -> partial dW
// Every k-block index has its own full/transformed mbarrier so that no waiter is ever more than one
// phase ahead of its barrier.
#pragma once

#include "linear_tc.cuh"

namespace hgin {
namespace tcgemm {

constexpr int FUSED_THREADS = 384;
constexpr int FUSED_STAGES = 2;
constexpr int FUSED_STAGE_BYTES = 3 * TILE_BYTES;   // g | z | Wt k-block
constexpr int FUSED_EPI_BAR = 2;

struct FusedParams {
    int64_t rows;
    int num_tiles;
    int n;            // dz columns: 32, 64, 96 or 128
    int k1;           // x columns = dx width = dW columns (multiple of 16, <= 128)
    int act;
    const float *alpha;
    const float *x2;  // tail input columns (k2 <= 4) or null
    int64_t ld2;
    int k2;
    int want_dx;
    int use_e;
    int want_sums;
    float *dot_partials;    // [grid]
    float *dw_partials;     // [grid][n][k1]
    float *sum_partials;    // [grid][n][k2 + 1]  (tail columns, then db)
    float *alpha_partials;  // [grid][4]
};

struct FusedSmem {
    static constexpr int off_ring = 0;
    static constexpr int off_h = FUSED_STAGES * FUSED_STAGE_BYTES;          // 96 KB
    static constexpr int off_stage = off_h + 4 * TILE_BYTES;                // + 64 KB
    static constexpr int off_e = off_stage + 2 * TILE_BYTES;                // + 32 KB
    static constexpr int off_small = off_e + 2 * TILE_BYTES;                // + 32 KB = 224 KB
    static constexpr int total = off_small + 512 + 1024;
};

__global__ void __launch_bounds__(FUSED_THREADS, 1)
bwd_fused_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_z,
                 const __grid_constant__ CUtensorMap tm_wt, const __grid_constant__ CUtensorMap tm_h,
                 const __grid_constant__ CUtensorMap tm_dx, const __grid_constant__ CUtensorMap tm_e,
                 const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *ring = smem + FusedSmem::off_ring;
    uint8_t *smem_h = smem + FusedSmem::off_h;
    uint8_t *smem_stage = smem + FusedSmem::off_stage;
    uint8_t *smem_e = smem + FusedSmem::off_e;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + FusedSmem::off_small);
    uint64_t *full = bars;               // [4] indexed by k-block counter & 3: TMA landed
    uint64_t *transformed = bars + 4;    // [4] same indexing: dz written (smem + TMEM)
    uint64_t *empty = bars + 8;          // [2] per stage: dx MMAs done reading
    uint64_t *h_full = bars + 10;
    uint64_t *h_empty = bars + 11;
    uint64_t *a_free = bars + 12;        // dW MMAs done reading the TMEM A operand
    uint64_t *d1_full = bars + 13;       // [2]
    uint64_t *d1_empty = bars + 15;      // [2]
    uint64_t *d2_full = bars + 17;
    uint64_t *e_full = bars + 18;        // [2]
    uint64_t *e_empty = bars + 20;       // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 22);
    float *red = reinterpret_cast<float *>(bars + 24);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = p.n / 32;
    const int nchunks = (p.k1 + 31) / 32;
    const int h_boxes = nchunks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_g);
        prefetch_tmap(&tm_z);
        prefetch_tmap(&tm_wt);
        prefetch_tmap(&tm_h);
        if (p.want_dx) prefetch_tmap(&tm_dx);
        if (p.use_e) prefetch_tmap(&tm_e);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&transformed[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&empty[i], 1);
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 4);
            mbar_init(&e_full[i], 1);
            mbar_init(&e_empty[i], 4);
        }
        mbar_init(h_full, 1);
        mbar_init(h_empty, 1);
        mbar_init(a_free, 1);
        mbar_init(d2_full, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t TM_D1 = tmem_base;          // two accumulators of 128 columns
    const uint32_t TM_D2 = tmem_base + 256;
    const uint32_t TM_A = tmem_base + 384;
    const bool act_on = p.act != HGIN_ACT_NONE;

    if (warp == 0) {
        // ===== operand producer =====
        if (lane == 0) {
            int it = 0;
            int c = 0;   // running k-block counter -> ring stage
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                for (int kb = 0; kb < num_kb; ++kb, ++c) {
                    const int s = c % FUSED_STAGES;
                    uint64_t *fb = &full[c & 3];   // barrier index follows the k-block counter: tied to the stage
                    mbar_wait(&empty[s], ((c / FUSED_STAGES) & 1) ^ 1);
                    uint8_t *slot = ring + s * FUSED_STAGE_BYTES;
                    mbar_expect_tx(fb, (act_on ? 2 : 1) * TILE_BYTES + p.k1 * 128);
                    tma_load_2d(slot, &tm_g, fb, kb * 32, tile * BM);
                    if (act_on) tma_load_2d(slot + TILE_BYTES, &tm_z, fb, kb * 32, tile * BM);
                    tma_load_2d(slot + 2 * TILE_BYTES, &tm_wt, fb, kb * 32, 0);
                }
                // x tile for the dW MMA: needed last in the tile; its buffer frees when the previous
                // tile's dW MMAs have finished
                mbar_wait(h_empty, (it & 1) ^ 1);
                mbar_expect_tx(h_full, h_boxes * TILE_BYTES);
                for (int b = 0; b < h_boxes; ++b) tma_load_2d(smem_h + b * TILE_BYTES, &tm_h, h_full, b * 32, tile * BM);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc_dx = make_idesc_tf32(BM, p.k1, 0, 0);
            const uint32_t idesc_dw = make_idesc_tf32(128, p.k1, 0, 1);   // A from TMEM (K-major), B = x MN-major
            int it = 0;
            int c = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&d1_empty[acc], ((it >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                for (int kb = 0; kb < num_kb; ++kb, ++c) {
                    const int s = c % FUSED_STAGES;
                    mbar_wait(&transformed[c & 3], (c >> 2) & 1);
                    tcgen05_fence_after();
                    const uint32_t a_base = smem_u32(ring + s * FUSED_STAGE_BYTES);
                    const uint32_t b_base = a_base + 2 * TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < KB / UMMA_K; ++k) {
                        umma_tf32(TM_D1 + acc * 128, make_smem_desc(a_base + k * UMMA_K * 4, 16, 1024),
                                  make_smem_desc(b_base + k * UMMA_K * 4, 16, 1024), idesc_dx, (kb | k) != 0);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&d1_full[acc]);
                // dW += dz^T x over this tile's 128 rows: A^T from TMEM (lane = n, column = row m)
                mbar_wait(h_full, it & 1);
                tcgen05_fence_after();
                const uint32_t h_base = smem_u32(smem_h);
#pragma unroll
                for (int j = 0; j < BM / UMMA_K; ++j) {
                    umma_tf32_ts(TM_D2, TM_A + j * UMMA_K,
                                 make_smem_desc(h_base + j * 1024, TILE_BYTES, 512, kLayoutSwizzle128BBase32B), idesc_dw,
                                 (it | j) != 0);
                }
                umma_commit(h_empty);
                umma_commit(a_free);
            }
            umma_commit(d2_full);
        }
    } else if (warp == 2) {
        // ===== dot_x producer =====
        if (lane == 0 && p.use_e) {
            int b = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int cchunk = 0; cchunk < nchunks; ++cchunk) {
                    mbar_wait(&e_empty[b], ph ^ 1);
                    mbar_expect_tx(&e_full[b], TILE_BYTES);
                    tma_load_2d(smem_e + b * TILE_BYTES, &tm_e, &e_full[b], cchunk * 32, tile * BM);
                    if (++b == 2) { b = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== transform: dz = g * act'(z), in place + into TMEM, with the per-column sums =====
        const int kb = warp - 4;
        const int nn = kb * 32 + lane;   // dz column owned by this thread
        const float alpha = p.act == HGIN_ACT_PRELU ? __ldg(p.alpha) : 0.0f;
        float db = 0.0f, dalpha = 0.0f, tail[4] = {0.f, 0.f, 0.f, 0.f};
        if (kb < num_kb) {
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int c = it * num_kb + kb;
                const uint32_t gs = smem_u32(ring + (c % FUSED_STAGES) * FUSED_STAGE_BYTES);
                const uint32_t zs = gs + TILE_BYTES;
                mbar_wait(&full[c & 3], (c >> 2) & 1);
                mbar_wait(a_free, (it & 1) ^ 1);   // previous tile's dW MMAs have consumed the TMEM operand
                tcgen05_fence_after();
                const int64_t row0 = static_cast<int64_t>(tile) * BM;
#pragma unroll 1
                for (int m0 = 0; m0 < BM; m0 += 32) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t off = swz128(m0 + j, lane);
                        float d = lds_f32(gs + off);
                        if (act_on) {
                            const float zv = lds_f32(zs + off);
                            if (p.act == HGIN_ACT_PRELU && !(zv > 0.0f)) dalpha += d * zv;
                            d = act_backward(d, zv, p.act, alpha);
                            sts_f32(gs + off, d);
                        }
                        v[j] = d;
                    }
                    if (p.want_sums) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            db += v[j];
                            if (p.k2 > 0 && row0 + m0 + j < p.rows) {
                                const float *xr = p.x2 + (row0 + m0 + j) * p.ld2;
#pragma unroll
                                for (int t = 0; t < 4; ++t)   // static indices keep tail[] in registers
                                    if (t < p.k2) tail[t] = fmaf(v[j], __ldg(xr + t), tail[t]);
                            }
                        }
                    }
                    tmem_st_32x32(TM_A + (static_cast<uint32_t>(kb * 32) << 16) + m0, v);
                }
                tmem_st_wait();
                fence_proxy_async_smem();      // in-place dz (generic proxy) -> visible to the MMA's async reads
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&transformed[c & 3]);
            }
            if (p.want_sums && nn < p.n) {
                float *dst = p.sum_partials + (static_cast<int64_t>(blockIdx.x) * p.n + nn) * (p.k2 + 1);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (t < p.k2) dst[t] = tail[t];
                dst[p.k2] = db;
            }
        }
        dalpha = warp_sum(dalpha);
        if (lane == 0 && p.alpha_partials) p.alpha_partials[blockIdx.x * 4 + kb] = dalpha;
    } else if (warp >= 8) {
        // ===== epilogue =====
        const int q = warp - 8;
        const int et = threadIdx.x - 256;
        const int r = q * 32 + lane;
        float dot = 0.0f;
        int it = 0, eb = 0;
        uint32_t eph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int64_t grow = static_cast<int64_t>(tile) * BM + r;
            mbar_wait(&d1_full[acc], (it >> 1) & 1);
            tcgen05_fence_after();
            for (int c = 0; c < nchunks; ++c) {
                float v[32];
                tmem_ld_32x32(TM_D1 + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + c * 32, v);
                if (c == nchunks - 1) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d1_empty[acc]);
                }
                if (p.use_e) {
                    mbar_wait(&e_full[eb], eph);
                    const uint32_t eb_ptr = smem_u32(smem_e + eb * TILE_BYTES);
                    if (grow < p.rows) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 t = lds_v4(eb_ptr + swz128(r, j4 * 4));
                            dot = fmaf(v[j4 * 4 + 0], t.x, dot);
                            dot = fmaf(v[j4 * 4 + 1], t.y, dot);
                            dot = fmaf(v[j4 * 4 + 2], t.z, dot);
                            dot = fmaf(v[j4 * 4 + 3], t.w, dot);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&e_empty[eb]);
                    if (++eb == 2) { eb = 0; eph ^= 1; }
                }
                if (p.want_dx) {
                    if (et == 0) tma_store_wait_read<0>();
                    named_barrier(FUSED_EPI_BAR, 128);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        sts_v4(smem_u32(smem_stage) + swz128(r, j4 * 4), v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2],
                               v[j4 * 4 + 3]);
                    fence_proxy_async_smem();
                    named_barrier(FUSED_EPI_BAR, 128);
                    if (et == 0) {
                        tma_store_2d(&tm_dx, smem_stage, c * 32, tile * BM);
                        tma_store_commit();
                    }
                }
            }
        }
        if (et == 0) tma_store_wait<0>();
        // partial dW from D2: this thread owns output row n = r
        mbar_wait(d2_full, 0);
        tcgen05_fence_after();
        for (int c = 0; c < nchunks; ++c) {
            float v[32];
            tmem_ld_32x32(TM_D2 + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
            if (r < p.n) {
                float *dst = p.dw_partials + (static_cast<int64_t>(blockIdx.x) * p.n + r) * p.k1 + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c * 32 + j < p.k1) dst[j] = v[j];
            }
        }
        if (p.dot_partials) {
            dot = warp_sum(dot);
            if (lane == 0) red[q] = dot;
            named_barrier(FUSED_EPI_BAR, 128);
            if (et == 0) p.dot_partials[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace tcgemm
}  // namespace hgin
