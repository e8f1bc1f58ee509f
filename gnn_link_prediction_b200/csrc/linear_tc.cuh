// Tensor-core (tcgen05, kind::tf32, fp32 accumulation in TMEM) kernels for the dense layers whose
// contraction and output widths are genuine GEMM sizes (hidden 128): HGIN_MATH_TF32.
//
//   gemm_nt_kernel<EPI>  D[M x N] = A[M x K] * B[N x K]^T     both operands K-major, M = rows (huge)
//        EPI_FWD : z = D + bias (+ rank-k2 update from x2),  out (+)= act(z)       forward (K2)
//        EPI_DX  : dx = D,  optional  sum(dx * dot_x)                              input grad (K3)
//   gemm_tn_kernel       D[N x K] = A[M x N]^T * B[M x K]     both operands MN-major, split over M
//        partial dW per CTA, reduced deterministically afterwards                   weight grad (K3)
//
// Structure of both (persistent, one CTA per SM):
//   warp 0  TMA producer: operand tiles -> 128B-swizzled shared-memory ring, mbarrier complete_tx
//   warp 1  one elected thread issues tcgen05.mma; tcgen05.commit frees ring slots / publishes TMEM
//   warp 2  TMA producer for the epilogue operand (old `out` when merging, dot_x for d(eps))
//   warp 3  allocates / frees TMEM
//   warps 4-7, 8-11  two epilogue warpgroups on alternating 32-column chunks: tcgen05.ld (warp q of a
//           group owns TMEM lanes 32q..32q+31, i.e. one row per thread) -> bias / PReLU / merge in
//           registers -> the group's own swizzled staging tiles -> TMA store by the group's issuer
//           (the per-tile epilogue, not the MMA, is what competes with HBM time here)
// Every one of these layers is HBM-bound (32 flop/B against a ridge of ~220), so the design goal
// is to keep the TMA queues full: MMA time per 128-row tile is ~0.6 us against ~4 us of HBM time.
#pragma once

#include "hgin_common.cuh"
#include "linear_tc_api.h"
#include "tc_common.cuh"

namespace hgin {
namespace tcgemm {

using namespace tc;

constexpr int BM = 128;               // rows per tile (UMMA M)
constexpr int KB = 32;                // fp32 per 128-byte swizzle row = one K-block
constexpr int UMMA_K = 8;             // tf32
constexpr int TILE_BYTES = BM * 128;  // 16 KB: 128 rows x 128 B
constexpr int NT_STAGES = 4;          // A ring, K-block granularity
constexpr int THREADS = 256;          // gemm_tn_kernel
constexpr int NT_THREADS = 384;       // gemm_nt_kernel: 4 service warps + 2 epilogue warpgroups
constexpr uint32_t TMEM_COLS = 256;   // two 128-column fp32 accumulators
constexpr int EPI_FWD = 0, EPI_DX = 1;
constexpr int EPI_BAR = 1;            // named barrier ids 1, 2: the two epilogue warpgroups (128 threads each)
constexpr int EPI_ALL_BAR = 3;        // both groups (256 threads)

struct NtParams {
    int64_t rows;
    int num_tiles;
    int num_kb;       // ceil(K / 32)
    int n;            // output columns, multiple of 16, <= 128
    // EPI_FWD
    const float *bias;
    const float *alpha;
    int act;
    const float *x2;  // [rows, k2] extra input columns (readout: raw path features), k2 <= 4
    int64_t ld2;
    int k2;
    const float *w_tail;  // &W[0][k1], leading dimension ldw: weights of the x2 columns
    int ldw;
    int want_z;
    int want_out;
    int use_e;        // EPI_FWD: e = previous `out` (HeteroConv sum merge); EPI_DX: 1: e = dot_x,
                      // 2: e = pre-activation z of the layer below -> dx is stored as dx * act'(e)
                      // (act / alpha describe THAT layer) and dot_partials receive sum dx * min(e, 0)
    // EPI_DX
    float *dot_partials;  // [gridDim.x]
    // EPI_DX with use_e == 2, GIN self branch riding on the epilogue (hgin_linear_bwd_post_self): the tile
    // D = dz W is dh_self; it leaves as (1 + eps) * D * act'(e), and dot2_partials receive sum D * act(e)
    // = d(eps) (act(e) is x_dst, the output of the layer below)
    const float *self_eps;
    float *dot2_partials; // [gridDim.x] or NULL
};

// Shared-memory carve-up (dynamic smem, base aligned to 1024 B by the kernel).
struct NtSmem {
    static constexpr int kB = 4 * TILE_BYTES;            // W: up to 4 K-blocks of [128 x 128 B]
    static constexpr int kA = NT_STAGES * TILE_BYTES;    // A ring
    static constexpr int kStage = 4 * TILE_BYTES;        // staging for TMA stores: (out, z) chunk per epilogue group
    static constexpr int kE = 2 * TILE_BYTES;            // epilogue-operand ring
    static constexpr int off_b = 0;
    static constexpr int off_a = off_b + kB;
    static constexpr int off_stage = off_a + kA;
    static constexpr int off_e = off_stage + kStage;
    static constexpr int off_small = off_e + kE;
    static constexpr int small_bytes = 128 * 4 /*bias*/ + 128 * 4 * 4 /*w_tail*/ + 32 * 8 /*barriers*/ + 64;
    // 224 KB of tiles + 2.8 KB: fits the 227 KB limit only without alignment slack, so the kernel
    // declares its dynamic shared memory __align__(1024) (no static shared memory precedes it).
    static constexpr int total = off_small + small_bytes;
    static_assert(total <= 227 * 1024, "gemm_nt shared memory exceeds the sm_100 per-CTA limit");
};

template <int EPI>
__global__ void __launch_bounds__(NT_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1,
               const __grid_constant__ CUtensorMap tm_e, const NtParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;   // SWIZZLE_128B tiles need 1024-byte alignment: guaranteed by the declaration
    uint8_t *smem_b = smem + NtSmem::off_b;
    uint8_t *smem_a = smem + NtSmem::off_a;
    uint8_t *smem_stage = smem + NtSmem::off_stage;
    uint8_t *smem_e = smem + NtSmem::off_e;
    float *bias_s = reinterpret_cast<float *>(smem + NtSmem::off_small);
    float *wtail_s = bias_s + 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wtail_s + 128 * 4);
    uint64_t *full = bars;                   // [NT_STAGES]
    uint64_t *empty = bars + NT_STAGES;      // [NT_STAGES]
    uint64_t *b_full = bars + 2 * NT_STAGES; // [1]
    uint64_t *tmem_full = b_full + 1;        // [2]
    uint64_t *tmem_empty = tmem_full + 2;    // [2]
    uint64_t *e_full = tmem_empty + 2;       // [2]
    uint64_t *e_empty = e_full + 2;          // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(e_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nchunks = (p.n + 31) / 32;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_b);
        prefetch_tmap(&tm_o0);
        if (EPI == EPI_FWD && p.want_z) prefetch_tmap(&tm_o1);
        if (p.use_e) prefetch_tmap(&tm_e);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NT_STAGES; ++i) {
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
            for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(smem_b + kb * TILE_BYTES, &tm_b, b_full, kb * KB, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], TILE_BYTES);
                    tma_load_2d(smem_a + s * TILE_BYTES, &tm_a, &full[s], kb * KB, tile * BM);
                    if (++s == NT_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, p.n, 0, 0);
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
                    for (int k = 0; k < KB / UMMA_K; ++k) {
                        umma_tf32(d_tmem, make_smem_desc(a_base + k * UMMA_K * 4, 16, 1024),
                                  make_smem_desc(b_base + k * UMMA_K * 4, 16, 1024), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[s]);   // ring slot reusable once these MMAs have read it
                    if (++s == NT_STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp == 2) {
        // ===== epilogue-operand producer =====
        if (lane == 0 && p.use_e) {
            uint32_t ph[2] = {0, 0};   // chunk c goes to buffer c & 1, consumed by epilogue group c & 1
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int c = 0; c < nchunks; ++c) {
                    const int b = c & 1;
                    mbar_wait(&e_empty[b], ph[b] ^ 1);
                    mbar_expect_tx(&e_full[b], TILE_BYTES);
                    tma_load_2d(smem_e + b * TILE_BYTES, &tm_e, &e_full[b], c * 32, tile * BM);
                    ph[b] ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: group 0 (warps 4-7) takes chunks 0, 2; group 1 (warps 8-11) chunks 1, 3 =====
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
        float dot2 = 0.0f;
        // this group's staging tiles (st0: out / dx, st1: z) and epilogue-operand buffer
        const uint32_t sa0 = smem_u32(smem_stage + grp * 2 * TILE_BYTES);
        const uint32_t sa1 = sa0 + TILE_BYTES;
        const uint32_t eb_ptr = smem_u32(smem_e + grp * TILE_BYTES);
        const int my_chunks = (nchunks - grp + 1) / 2;   // chunks grp, grp + 2, ...
        float dot = 0.0f;
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
                float v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + c * 32, v);
                if (c + 2 >= nchunks) {          // this group's last chunk: hand the accumulator back
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                float ev[32];
                if (p.use_e) {                   // chunk c was loaded into buffer c & 1 == grp
                    mbar_wait(&e_full[grp], eph);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 t = lds_v4(eb_ptr + swz128(r, j4 * 4));
                        ev[j4 * 4 + 0] = t.x; ev[j4 * 4 + 1] = t.y; ev[j4 * 4 + 2] = t.z; ev[j4 * 4 + 3] = t.w;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&e_empty[grp]);
                    eph ^= 1;
                }
                float o[32];
                if (EPI == EPI_FWD) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int nn = c * 32 + j;
                        float zz = v[j] + bias_s[nn];
                        if (p.k2 > 0) {      // rank-k2 update from the extra input columns (readout layer 1 only)
                            zz = fmaf(x2v[0], wtail_s[nn * 4 + 0], zz);
                            zz = fmaf(x2v[1], wtail_s[nn * 4 + 1], zz);
                            zz = fmaf(x2v[2], wtail_s[nn * 4 + 2], zz);
                            zz = fmaf(x2v[3], wtail_s[nn * 4 + 3], zz);
                        }
                        v[j] = zz;
                        if (p.want_out) {    // (a lazily activated layer stores z only)
                            float oo = act_forward(zz, p.act, alpha);
                            if (p.use_e) oo += ev[j];
                            o[j] = oo;
                        }
                    }
                } else if (p.use_e == 1) {
                    if (grow < p.rows) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) dot = fmaf(v[j], ev[j], dot);  // OOB columns of e are zero-filled
                    }
                } else if (p.use_e == 2) {
                    // rows past the end and OOB columns: v == 0 or e == 0, so they add nothing to dot
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
                // the group's staging tiles are free once its previous TMA stores have read them
                if (gt == 0) tma_store_wait_read<0>();
                named_barrier(bar_id, 128);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const uint32_t off = swz128(r, j4 * 4);
                    if (EPI == EPI_FWD) {
                        if (p.want_out) sts_v4(sa0 + off, o[j4 * 4], o[j4 * 4 + 1], o[j4 * 4 + 2], o[j4 * 4 + 3]);
                        if (p.want_z) sts_v4(sa1 + off, v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                    } else {
                        sts_v4(sa0 + off, v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                    }
                }
                fence_proxy_async_smem();
                named_barrier(bar_id, 128);
                if (gt == 0) {
                    if (EPI == EPI_FWD) {
                        if (p.want_out) tma_store_2d(&tm_o0, smem_stage + grp * 2 * TILE_BYTES, c * 32, tile * BM);
                        if (p.want_z) tma_store_2d(&tm_o1, smem_stage + grp * 2 * TILE_BYTES + TILE_BYTES, c * 32, tile * BM);
                    } else if (p.want_out) {
                        tma_store_2d(&tm_o0, smem_stage + grp * 2 * TILE_BYTES, c * 32, tile * BM);
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

// ---- weight gradient: D[n][k] = sum_m A[m][n] * B[m][k], both operands MN-major -------------------
constexpr int TN_ROWS = 32;                          // contraction rows per ring stage
constexpr int TN_STAGES = 5;
constexpr int TN_BOX_BYTES = TN_ROWS * 128;          // one [32 rows x 32 cols] box = 4 KB
constexpr int TN_OPERAND_BYTES = 4 * TN_BOX_BYTES;   // up to 128 columns = 4 boxes = 16 KB

struct TnParams {
    int64_t rows;
    int64_t rows_per_cta;   // multiple of TN_ROWS
    int n;                  // columns of A = output rows   (<= 128, multiple of 16)
    int k;                  // columns of B = output cols   (<= 128, multiple of 16)
    float *partials;        // [gridDim.x][n][k]
    // db = sum_m dz[m][:] through the tensor core: one extra [32 x 32] box of ones behind B's last
    // box widens the MMA to N = k + 16, so D[:, k] = A^T 1.  Needs k % 32 == 0.
    int ones_col;
    float *db_partials;     // [gridDim.x][n]
    // MN-major operand descriptor fields (defaults in linear_tc.cu; overridable by hgin_debug_gemm_tn)
    int lbo;                // bytes between 32-column boxes of one operand
    int sbo;                // bytes between swizzle atoms along the contraction rows
    int layout_type;        // UMMA layout type
    int k_step_bytes;       // start-address advance per UMMA_K = 8 contraction rows
};

struct TnSmem {
    static constexpr int stage_bytes = 2 * TN_OPERAND_BYTES + TN_BOX_BYTES;    // A, B, the ones box
    static constexpr int off_ring = 0;
    static constexpr int off_small = TN_STAGES * stage_bytes;
    static constexpr int total = off_small + 32 * 8 + 64 + 1024;
};

static __global__ void __launch_bounds__(THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *ring = smem + TnSmem::off_ring;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TnSmem::off_small);
    uint64_t *full = bars;
    uint64_t *empty = bars + TN_STAGES;
    uint64_t *done = bars + 2 * TN_STAGES;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t m_beg = static_cast<int64_t>(blockIdx.x) * p.rows_per_cta;
    const int64_t m_end = min(m_beg + p.rows_per_cta, p.rows);
    const int steps = static_cast<int>((m_end - m_beg + TN_ROWS - 1) / TN_ROWS);
    const int a_boxes = (p.n + 31) / 32, b_boxes = (p.k + 31) / 32;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < TN_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<256>(tmem_ptr);
    if (p.ones_col) {   // TMA never writes this box: fill it once per stage
        for (int i = threadIdx.x; i < TN_STAGES * (TN_BOX_BYTES / 4); i += THREADS) {
            const int st = i / (TN_BOX_BYTES / 4), w = i % (TN_BOX_BYTES / 4);
            reinterpret_cast<float *>(ring + st * TnSmem::stage_bytes + TN_OPERAND_BYTES + b_boxes * TN_BOX_BYTES)[w] = 1.0f;
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
                mbar_expect_tx(&full[s], (a_boxes + b_boxes) * TN_BOX_BYTES);
                uint8_t *sa = ring + s * TnSmem::stage_bytes;
                uint8_t *sb = sa + TN_OPERAND_BYTES;
                const int m = static_cast<int>(m_beg + static_cast<int64_t>(i) * TN_ROWS);
                // CTA ranges are multiples of TN_ROWS, so boxes never straddle two ranges; the last
                // box of the matrix is zero-filled past `rows` by TMA.
                for (int c = 0; c < a_boxes; ++c) tma_load_2d(sa + c * TN_BOX_BYTES, &tm_a, &full[s], c * 32, m);
                for (int c = 0; c < b_boxes; ++c) tma_load_2d(sb + c * TN_BOX_BYTES, &tm_b, &full[s], c * 32, m);
                if (++s == TN_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, p.k + (p.ones_col ? 16 : 0), 1, 1);
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < steps; ++i) {
                mbar_wait(&full[s], ph);
                tcgen05_fence_after();
                const uint32_t a_base = smem_u32(ring + s * TnSmem::stage_bytes);
                const uint32_t b_base = a_base + TN_OPERAND_BYTES;
#pragma unroll
                for (int j = 0; j < TN_ROWS / UMMA_K; ++j) {
                    umma_tf32(tmem_base, make_smem_desc(a_base + j * p.k_step_bytes, p.lbo, p.sbo, p.layout_type),
                              make_smem_desc(b_base + j * p.k_step_bytes, p.lbo, p.sbo, p.layout_type), idesc,
                              (i | j) != 0);
                }
                umma_commit(&empty[s]);
                if (++s == TN_STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(done);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int nn = q * 32 + lane;  // output row (column of A) owned by this thread
        float *dst = p.partials + (static_cast<int64_t>(blockIdx.x) * p.n + nn) * p.k;
        if (steps > 0) {
            mbar_wait(done, 0);
            tcgen05_fence_after();
            for (int c = 0; c < b_boxes; ++c) {
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
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b_boxes * 32, v);
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
