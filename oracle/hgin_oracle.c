/* TEST INFRASTRUCTURE — plain-C CPU restatement of the integer and bit-exact fp32 pieces of the
 * HeteroGIN hot path.  NOT part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg load this library (oracle/c_oracle.py).
 *
 * What it restates (reference file:line):
 *   oracle_csr_build   — the edge order contract of SURVEY §8(a) A0: for a COO list emitted by
 *                        generateFiles.py:145-181 and collated by PyG (dataset.py:242),
 *                        perm = argsort(dst, stable), rowptr = exclusive_cumsum(bincount(dst)),
 *                        col = src[perm].  Pure integer work: must match bit for bit.
 *   oracle_gin_combine — GINConv.forward up to the MLP input (models.py:208-215):
 *                        agg[d,:] = sum over edges e with dst_e == d of x_src[src_e,:], added
 *                        left-to-right in edge order in fp32 (what the CPU reference's
 *                        zeros().scatter_add_() does for this call shape — checked against torch
 *                        in tests/test_oracle.py), then
 *                        concat:  h = [agg | fl(fl(1+eps) * x_dst)]          (models.py:213)
 *                        add:     h = fl(agg + fl(fl(1+eps) * x_dst))        (models.py:215)
 *   oracle_gather_t    — index_select backward (index_add_ in edge order): dx_src[s,:] = sum over
 *                        edges with src_e == s of g[dst_e,:], left-to-right.
 * Parity status: unpinned by the reference's own tests (it has none); pinned by tests/golden/
 * vectors generated from the unmodified reference model (oracle/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: each add/mul rounds once,
 * as the ATen CPU kernels do).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Stable counting sort of edges by dst.  Returns 0, or -1 on an out-of-range index. */
int oracle_csr_build(int64_t num_edges, const int64_t *src, const int64_t *dst,
                     int64_t num_src, int64_t num_dst, int32_t *rowptr, int32_t *col,
                     int32_t *perm /* may be NULL */)
{
    memset(rowptr, 0, (size_t)(num_dst + 1) * sizeof(int32_t));
    for (int64_t e = 0; e < num_edges; ++e) {
        if (dst[e] < 0 || dst[e] >= num_dst || src[e] < 0 || src[e] >= num_src) return -1;
        rowptr[dst[e] + 1] += 1;
    }
    for (int64_t d = 0; d < num_dst; ++d) rowptr[d + 1] += rowptr[d];
    int32_t *cursor = (int32_t *)malloc((size_t)(num_dst > 0 ? num_dst : 1) * sizeof(int32_t));
    if (!cursor) return -2;
    memcpy(cursor, rowptr, (size_t)num_dst * sizeof(int32_t));
    for (int64_t e = 0; e < num_edges; ++e) {
        int32_t slot = cursor[dst[e]]++;
        col[slot] = (int32_t)src[e];
        if (perm) perm[slot] = (int32_t)e;
    }
    free(cursor);
    return 0;
}

/* h[d, 0:F_src] (+)= sequential fp32 sum of neighbour rows; self term per `concat`. */
void oracle_gin_combine(int64_t num_dst, const int32_t *rowptr, const int32_t *col,
                        const float *x_src, int64_t ld_src, int32_t f_src,
                        const float *x_dst, int64_t ld_dst, int32_t f_dst,
                        float eps, int32_t concat, float *h, int64_t ld_h)
{
    volatile float ope = 1.0f + eps; /* fl(1 + eps), models.py:213/215 */
    for (int64_t d = 0; d < num_dst; ++d) {
        float *hd = h + d * ld_h;
        for (int32_t f = 0; f < f_src; ++f) {
            float acc = 0.0f;
            for (int32_t e = rowptr[d]; e < rowptr[d + 1]; ++e)
                acc = acc + x_src[(int64_t)col[e] * ld_src + f];
            hd[f] = acc;
        }
        if (x_dst) {
            if (concat) {
                for (int32_t f = 0; f < f_dst; ++f) hd[f_src + f] = ope * x_dst[d * ld_dst + f];
            } else {
                for (int32_t f = 0; f < f_dst; ++f) {
                    volatile float t = ope * x_dst[d * ld_dst + f];
                    hd[f] = hd[f] + t;
                }
            }
        }
    }
}

/* Transposed gather (backward of index_select): rows of the TRANSPOSED csr are src nodes. */
void oracle_gather_t(int64_t num_src, const int32_t *rowptr_t, const int32_t *col_t,
                     const float *g, int64_t ld_g, int32_t f, float *dx, int64_t ld_dx)
{
    for (int64_t s = 0; s < num_src; ++s)
        for (int32_t k = 0; k < f; ++k) {
            float acc = 0.0f;
            for (int32_t e = rowptr_t[s]; e < rowptr_t[s + 1]; ++e)
                acc = acc + g[(int64_t)col_t[e] * ld_g + k];
            dx[s * ld_dx + k] = acc;
        }
}
