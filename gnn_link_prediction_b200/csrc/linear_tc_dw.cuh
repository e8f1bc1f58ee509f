// Weight gradient fused with dz = g * act'(z)  (HGIN_MATH_TF32 backward, first pass).
//
//   dW[n][k] = sum_m dz[m][n] x[m][k],  db, dalpha, tail of dW;  dz is also written out for the
//   input-gradient GEMM that follows (gemm_nt<EPI_DX>).
//
// Replaces dz_prepare_kernel + gemm_tn_kernel: 4 row-sized transfers (read g, z, x; write dz) instead
// of 5.  Unlike the fully fused backward (linear_tc_fused.cuh) no weight matrix sits in shared
// memory, so a 4-stage ring of (g, z) k-blocks (128 KB) + the x tile (64 KB) fits and keeps
// enough bytes in flight.
//
//   warp 0     TMA producer: (g_kb, z_kb) k-blocks [128 rows x 32 cols] into the ring; x tile
//              [128 x k1] in the MN-major (32-byte swizzle atom) layout
//   warps 4-7  transform: warp 4+kb owns k-block kb / TMEM lane quarter kb.  Lane i owns column
//              n = 32 kb + i and walks the 128 rows: dz written in place (then TMA-stored from the
//              ring slot) and stored with tcgen05.st into TMEM as the A^T operand; db / dalpha /
//              tail sums are thread-local
//   warp 1     MMA: D[n][k] += dz^T (TMEM, double-buffered per tile) * x (smem)
//   warp 3     TMEM allocation;  warps 8-11: final D -> partial dW
// Ring stage == k-block counter & 3, so for n in {32, 64, 128} every mbarrier is waited by one
// warp in program order (no waiter can run a phase ahead).
#pragma once

#include "linear_tc.cuh"

namespace hgin {
namespace tcgemm {

constexpr int DW_THREADS = 384;
constexpr int DW_STAGES = 4;
constexpr int DW_STAGE_BYTES = 2 * TILE_BYTES;   // g | z k-block

struct DwParams {
    int64_t rows;
    int64_t rows_per_cta;   // multiple of 128
    int n;                  // dz columns: 32, 64 or 128
    int k1;                 // x columns (multiple of 16, <= 128)
    int act;
    const float *alpha;
    const float *x2;
    int64_t ld2;
    int k2;
    int store_dz;           // 0 when act == NONE (dz == g: the dx GEMM reads g directly)
    float *dw_partials;     // [grid][n][k1]
    float *sum_partials;    // [grid][n][k2 + 1]
    float *alpha_partials;  // [grid][4]
};

struct DwSmem {
    static constexpr int off_ring = 0;
    static constexpr int off_h = DW_STAGES * DW_STAGE_BYTES;   // 128 KB
    static constexpr int off_small = off_h + 4 * TILE_BYTES;   // + 64 KB
    static constexpr int total = off_small + 256;
};

__global__ void __launch_bounds__(DW_THREADS, 1)
dw_fused_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_z,
                const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_dz,
                const DwParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ring = smem + DwSmem::off_ring;
    uint8_t *smem_h = smem + DwSmem::off_h;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + DwSmem::off_small);
    uint64_t *full = bars;           // [4] TMA landed
    uint64_t *empty = bars + 4;      // [4] transform done with the slot (dz store has read it)
    uint64_t *h_full = bars + 8;
    uint64_t *h_empty = bars + 9;
    uint64_t *a_ready = bars + 10;   // [2] all quarters of the TMEM operand written (count = num_kb)
    uint64_t *a_free = bars + 12;    // [2] dW MMAs done reading it
    uint64_t *d_full = bars + 14;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 15);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = p.n / 32;
    const int h_boxes = (p.k1 + 31) / 32;
    const int64_t m_beg = static_cast<int64_t>(blockIdx.x) * p.rows_per_cta;
    const int64_t m_end = min(m_beg + p.rows_per_cta, p.rows);
    const int tiles = static_cast<int>((m_end - m_beg + BM - 1) / BM);
    const bool act_on = p.act != HGIN_ACT_NONE;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_g);
        prefetch_tmap(&tm_z);
        prefetch_tmap(&tm_h);
        prefetch_tmap(&tm_dz);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < DW_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(h_full, 1);
        mbar_init(h_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_ready[i], num_kb);
            mbar_init(&a_free[i], 1);
        }
        mbar_init(d_full, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t TM_D = tmem_base;             // 128 columns
    const uint32_t TM_A = tmem_base + 128;       // 2 x 128 columns

    if (warp == 0) {
        if (lane == 0) {
            int c = 0;
            for (int t = 0; t < tiles; ++t) {
                const int row = static_cast<int>(m_beg) + t * BM;
                for (int kb = 0; kb < num_kb; ++kb, ++c) {
                    const int s = c & 3;
                    mbar_wait(&empty[s], ((c >> 2) & 1) ^ 1);
                    uint8_t *slot = ring + s * DW_STAGE_BYTES;
                    mbar_expect_tx(&full[s], (act_on ? 2 : 1) * TILE_BYTES);
                    tma_load_2d(slot, &tm_g, &full[s], kb * 32, row);
                    if (act_on) tma_load_2d(slot + TILE_BYTES, &tm_z, &full[s], kb * 32, row);
                }
                mbar_wait(h_empty, (t & 1) ^ 1);
                mbar_expect_tx(h_full, h_boxes * TILE_BYTES);
                for (int b = 0; b < h_boxes; ++b) tma_load_2d(smem_h + b * TILE_BYTES, &tm_h, h_full, b * 32, row);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, p.k1, 0, 1);   // A from TMEM (K-major), B = x MN-major
            for (int t = 0; t < tiles; ++t) {
                const int ab = t & 1;
                mbar_wait(&a_ready[ab], (t >> 1) & 1);
                mbar_wait(h_full, t & 1);
                tcgen05_fence_after();
                const uint32_t h_base = smem_u32(smem_h);
#pragma unroll
                for (int j = 0; j < BM / UMMA_K; ++j) {
                    umma_tf32_ts(TM_D, TM_A + ab * 128 + j * UMMA_K,
                                 make_smem_desc(h_base + j * 1024, TILE_BYTES, 512, kLayoutSwizzle128BBase32B), idesc,
                                 (t | j) != 0);
                }
                umma_commit(h_empty);
                umma_commit(&a_free[ab]);
            }
            umma_commit(d_full);
        }
    } else if (warp >= 4 && warp < 8) {
        const int kb = warp - 4;
        const int nn = kb * 32 + lane;
        const float alpha = p.act == HGIN_ACT_PRELU ? __ldg(p.alpha) : 0.0f;
        float db = 0.0f, dalpha = 0.0f, tail[4] = {0.f, 0.f, 0.f, 0.f};
        if (kb < num_kb) {
            for (int t = 0; t < tiles; ++t) {
                const int c = t * num_kb + kb;
                const int s = c & 3;
                const int ab = t & 1;
                const uint32_t gs = smem_u32(ring + s * DW_STAGE_BYTES);
                const uint32_t zs = gs + TILE_BYTES;
                const int64_t row0 = m_beg + static_cast<int64_t>(t) * BM;
                mbar_wait(&full[s], (c >> 2) & 1);
                mbar_wait(&a_free[ab], ((t >> 1) & 1) ^ 1);
                tcgen05_fence_after();
#pragma unroll 1
                for (int m0 = 0; m0 < BM; m0 += 32) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t off = swz128(m0 + j, lane);
                        float d = lds_f32(gs + off);
                        // rows of the tile that belong to the next CTA's range (or lie past the
                        // matrix) must not count: ranges are multiples of 128, so only `rows` clips
                        if (act_on) {
                            const float zv = lds_f32(zs + off);
                            if (p.act == HGIN_ACT_PRELU && !(zv > 0.0f)) dalpha += d * zv;
                            d = act_backward(d, zv, p.act, alpha);
                            sts_f32(gs + off, d);
                        }
                        v[j] = d;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        db += v[j];
                        if (p.k2 > 0 && row0 + m0 + j < p.rows) {
                            const float *xr = p.x2 + (row0 + m0 + j) * p.ld2;
#pragma unroll
                            for (int tt = 0; tt < 4; ++tt)
                                if (tt < p.k2) tail[tt] = fmaf(v[j], __ldg(xr + tt), tail[tt]);
                        }
                    }
                    tmem_st_32x32(TM_A + ab * 128 + (static_cast<uint32_t>(kb * 32) << 16) + m0, v);
                }
                tmem_st_wait();
                tcgen05_fence_before();
                fence_proxy_async_smem();        // in-place dz -> visible to the TMA store
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_ready[ab]);
                    if (p.store_dz) {
                        tma_store_2d(&tm_dz, ring + s * DW_STAGE_BYTES, kb * 32, static_cast<int>(row0));
                        tma_store_commit();
                        tma_store_wait_read<0>();   // the slot may be refilled once the store has read it
                    }
                    mbar_arrive(&empty[s]);
                }
            }
            if (nn < p.n) {
                float *dst = p.sum_partials + (static_cast<int64_t>(blockIdx.x) * p.n + nn) * (p.k2 + 1);
#pragma unroll
                for (int tt = 0; tt < 4; ++tt)
                    if (tt < p.k2) dst[tt] = tail[tt];
                dst[p.k2] = db;
            }
            if (lane == 0 && p.store_dz) tma_store_wait<0>();
        }
        dalpha = warp_sum(dalpha);
        if (lane == 0 && p.alpha_partials) p.alpha_partials[blockIdx.x * 4 + kb] = dalpha;
    } else if (warp >= 8) {
        const int q = warp - 8;
        const int r = q * 32 + lane;
        float *dst = p.dw_partials + (static_cast<int64_t>(blockIdx.x) * p.n + r) * p.k1;
        if (tiles > 0) {
            mbar_wait(d_full, 0);
            tcgen05_fence_after();
            for (int cc = 0; cc < h_boxes; ++cc) {
                float v[32];
                tmem_ld_32x32(TM_D + (static_cast<uint32_t>(q * 32) << 16) + cc * 32, v);
                if (r < p.n) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (cc * 32 + j < p.k1) dst[cc * 32 + j] = v[j];
                }
            }
        } else if (r < p.n) {
            for (int j = 0; j < p.k1; ++j) dst[j] = 0.0f;
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
