/* hgin.h — C ABI of libhgin.so: the HeteroGIN message-passing hot path on B200 (sm_100a).
 *
 * This is the LOWER drop-in boundary of SURVEY §8(b).  The reference is pure Python and has no
 * FFI of its own; what it binds today are the PyTorch/PyG operator calls listed per entry point
 * below (file:line into /root/reference).  A maintainer replaces those calls with a ctypes stub
 * over this header (INTEGRATION.md shows it); gnn_link_prediction_b200/_lib.py is that stub.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless the name ends in _host; all buffers (inputs,
 *     outputs, saved-for-backward, workspaces) are allocated and owned by the caller.  The
 *     library allocates nothing, frees nothing and keeps no state between calls.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises,
 *     so every entry point is legal inside CUDA-graph capture.
 *   - return value: HGIN_OK (0) or a negative hgin_status; hgin_last_error() returns a
 *     thread-local message for the last failing call.  No exceptions, no aborts.
 *   - matrices are fp32 row-major with an explicit leading dimension `ld*` in ELEMENTS;
 *     indices cross the ABI as int32 CSR (built once per batch by hgin_csr_build).
 */
#ifndef HGIN_H_
#define HGIN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGIN_VERSION 210 /* major*100 + minor */

typedef enum hgin_status {
    HGIN_OK = 0,
    HGIN_ERR_INVALID_ARGUMENT = -1,
    HGIN_ERR_WORKSPACE_TOO_SMALL = -2,
    HGIN_ERR_CUDA = -3,
    HGIN_ERR_UNSUPPORTED = -4
} hgin_status;

/* self-term modes of hgin_gin_combine */
#define HGIN_SELF_NONE 0   /* out = agg                                   */
#define HGIN_SELF_ADD 1    /* out = fl(agg + fl(fl(1+eps) * x_self))      models.py:215 */
#define HGIN_SELF_CONCAT 2 /* out = [agg | fl(fl(1+eps) * x_self)]        models.py:213 */

/* activations of hgin_linear_* */
#define HGIN_ACT_NONE 0
#define HGIN_ACT_PRELU 1 /* single shared slope, torch.nn.PReLU() (models.py:238, config.json:27) */
#define HGIN_ACT_RELU 2
/* the other activations `act = eval(config.MLP_ACT)` / `eval(mlp_head_act)` can name (models.py:301, 330): taken by
 * hgin_act_* and hgin_bn_act_* only (the hgin_linear_* kernels fuse NONE / PRELU / RELU); p0, p1 = module constants */
#define HGIN_ACT_LEAKY_RELU 3 /* p0 = negative_slope */
#define HGIN_ACT_ELU 4        /* p0 = alpha */
#define HGIN_ACT_SIGMOID 5
#define HGIN_ACT_TANH 6
#define HGIN_ACT_GELU 7       /* erf form (approximate='none') */
#define HGIN_ACT_SILU 8
#define HGIN_ACT_SOFTPLUS 9   /* p0 = beta, p1 = threshold */

/* math modes of hgin_linear_* */
#define HGIN_MATH_FP32 0 /* SIMT fp32 FMA: parity mode, rel 1e-5 against the CPU reference */
#define HGIN_MATH_TF32 1 /* tcgen05 kind::tf32 tensor-core tiles, fp32 accumulate in TMEM (16 <= K, N <= 128 per call;
                          * the host side tiles wider layers into 128-blocks of the same kernels) */
#define HGIN_MATH_BF16 2 /* activations / gradients STORED as bf16 (HGIN_DTYPE_BF16 rows), tcgen05 kind::f16 tiles on
                          * bf16 operands, fp32 accumulate; aggregation adds, reductions, loss and Adam in fp32.
                          * BASELINE configs[2] "bf16 MLP GEMMs", parity bar rel 1e-2 */

/* storage types of the row matrices taken by the *_t entry points (weights, biases, reductions and all
 * accumulation stay fp32) */
#define HGIN_DTYPE_F32 0
#define HGIN_DTYPE_BF16 1 /* rows stored as bfloat16: half the bytes per row; leading dimensions in ELEMENTS */

int32_t hgin_version(void);
const char *hgin_last_error(void);

/* ---- K0: destination-sorted CSR (and its transpose) from a COO edge list --------------------
 * Replaces: nothing in the reference (PyG keeps COO and uses atomics); fixes the summation
 * order of `scatter_add_` (models.py:208 -> torch_scatter.scatter) so that K1 is deterministic.
 * Contract, integer-exact:
 *     key  = edge_index[sort_row], other = edge_index[1 - sort_row]
 *     perm = argsort(key, stable);  rowptr = exclusive_cumsum(bincount(key, num_rows));
 *     col  = other[perm]
 * sort_row = 1 gives the forward CSR (rows = destination nodes), sort_row = 0 the transposed
 * CSR used by the backward gather (rows = source nodes).
 *   edge_index : [2, num_edges], index_bytes = 8 (int64, as the reference ships it,
 *                generateFiles.py:172-181) or 4 (int32); row r starts at element r*ld_edge.
 *   rowptr     : int32 [num_rows + 1];  col : int32 [num_edges];  perm : int32 [num_edges] or NULL.
 *   status     : int32 [1], set to 1 if any index is outside [0,num_rows) x [0,num_cols); such
 *                edges are dropped.  The caller reads it when it chooses to (no sync here).
 *                An edge (-1, -1) is a PADDING slot: dropped silently, so a batch can be padded
 *                to a fixed edge count and the whole step replayed as a CUDA graph.
 *   workspace  : hgin_csr_workspace_bytes(num_edges, num_rows) bytes.
 */
int64_t hgin_csr_workspace_bytes(int64_t num_edges, int64_t num_rows);
int32_t hgin_csr_build(const void *edge_index, int32_t index_bytes, int64_t num_edges,
                       int64_t ld_edge, int32_t sort_row, int64_t num_rows, int64_t num_cols,
                       int32_t *rowptr, int32_t *col, int32_t *perm, int32_t *status,
                       void *workspace, int64_t workspace_bytes, void *stream);

/* ---- K1 / K4: segmented neighbour sum over CSR rows, fused with the GIN self term -----------
 * Replaces (forward): GINConv.forward up to the MLP input, models.py:208-215 —
 *   propagate = x_src.index_select(0, edge_index[0]) + scatter_add_ over edge_index[1],
 *   then `(1 + eps) * x_r` and `cat` / in-place `+=`.
 * Replaces (backward): autograd of the same — index_select.backward (index_add_) as a gather
 *   over the TRANSPOSED CSR, fused with `(1+eps) * dh` for the self branch and with the sum over
 *   relations that share a node type (`accumulate`).
 * Per row r (one warp or sub-warp per row; neighbours added left to right in CSR order, fp32,
 * no atomics — bit-exact with the CPU reference):
 *     agg[f]  = sum_{e in [rowptr[r], rowptr[r+1])} x_src[col[e]*ld_src + f]        f < f_src
 *     out row = per `self_mode` above; `accumulate` != 0 adds the result to `out` instead.
 * num_edges: rowptr[num_rows] as known to the host (scheduling hint only: picks lanes per row
 * from the mean row length; pass -1 if unknown).
 * eps: device pointer to the learnable scalar (models.py:191-194) or NULL for eps = 0.
 * HGIN_SELF_ADD needs f_self == f_src.  x_self may be NULL only with HGIN_SELF_NONE.
 */
int32_t hgin_gin_combine(int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                         int64_t num_edges, const float *x_src, int64_t ld_src, int32_t f_src,
                         const float *x_self, int64_t ld_self, int32_t f_self,
                         const float *eps, int32_t self_mode, int32_t accumulate,
                         float *out, int64_t ld_out, void *stream);

/* ---- K1 / K4 for long rows on block-diagonal batches: input-major streaming -------------------------------------------
 * Same result as hgin_gin_combine_t (bit for bit), different schedule (csrc/gin_scatter_blocks.cuh): a batch of topology
 * samples is block-diagonal — block b owns input rows [in_ptr[b], in_ptr[b+1]) and output rows [out_ptr[b], out_ptr[b+1])
 * (PyG's `ptr` vectors of the two node types) — so one CTA per block keeps the block's output rows as fp32 accumulators
 * in shared memory and streams the block's input rows ONCE, in order, through a ring of bulk copies: every input row
 * crosses HBM exactly once instead of ~2.9 times through L2 (path->link rows have ~36 neighbours).
 *   hgin_block_gate: once per batch and relation, gate is int32[4].  gate[0] = containment violations (an edge of the
 *     input-major CSR_B that leaves its block, block tables that do not cover the rows); gate[1] / gate[2] = largest number
 *     of output / input rows of a block; gate[3] = rows of the output-major CSR_A whose neighbours are not in non-decreasing
 *     order (then ascending-input order would differ from the stable edge order of the reference's scatter_add_).
 *   hgin_gin_combine_blocks_t: launches the streaming kernel, which runs only if gate[0] == gate[3] == 0 and gate[1] fits its
 *     accumulator tile, and behind it the gather kernel of hgin_gin_combine_t with the inverse gate: a static, capturable
 *     launch sequence with exactly one of the two doing the work.  (rowptr, col) = CSR_A, rows = outputs, as for
 *     hgin_gin_combine_t; (rowptr_in, col_in) = CSR_B, rows = inputs.  Input rows must be contiguous (ld_src == f_src),
 *     32 <= f_src <= 128; SELF_NONE or SELF_ADD; src_act / self_act as in hgin_gin_combine_pre.  HGIN_ERR_UNSUPPORTED
 *     otherwise (the caller then uses hgin_gin_combine_t).
 */
int32_t hgin_block_gate(int64_t rows_a, const int32_t *rowptr_a, const int32_t *col_a, int64_t rows_b,
                        const int32_t *rowptr_b, const int32_t *col_b, int32_t num_blocks,
                        const int64_t *in_ptr, const int64_t *out_ptr, int32_t *gate, void *stream);
int32_t hgin_gin_combine_blocks_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                                  int64_t num_edges, int64_t num_in, const int32_t *rowptr_in,
                                  const int32_t *col_in, int32_t num_blocks, const int64_t *in_ptr,
                                  const int64_t *out_ptr, const int32_t *gate, const void *x_src,
                                  int64_t ld_src, int32_t f_src, const void *x_self, int64_t ld_self,
                                  const float *eps, int32_t self_mode, int32_t accumulate, void *out,
                                  int64_t ld_out, int32_t src_act, const float *src_alpha, int32_t self_act,
                                  const float *self_alpha, void *stream);

/* ---- K4 with the activation derivative of the layer below ("post-activation") ---------------
 * Replaces: the first op of the NEXT autograd node on the way down, PReLU.backward of the layer
 * whose output this gradient is for (models.py:238 -> at::prelu_backward).  The row result r of
 * hgin_gin_combine is a gradient w.r.t. that layer's output; with post_z (its saved
 * pre-activation, [num_rows, f_src]) the kernel stores  r * act'(post_z)  — i.e. that layer's dz —
 * and reduces  post_dalpha[0] = sum r * min(post_z, 0)  (two-stage, deterministic; NULL to skip).
 * post_ddot (NULL to skip; HGIN_SELF_ADD only):  sum x_self * act(post_z).  In the backward pass the
 * rows of x_self are dh_self and act(post_z) is x_dst, so this is d(eps) = sum dh_self * x_dst of the
 * relation whose self branch rides on the pass (models.py:213/215) without reading x_dst again.
 * The layer below is then called with HGIN_ACT_NONE (hgin_linear_bwd consumes dz in place).
 * post_act == HGIN_ACT_NONE or post_z == NULL: identical to hgin_gin_combine.
 * rowptr == NULL: the relation has no edges; the call reduces to the self term (used for
 * `(1+eps) * dh` when no gather shares the pass).  Not with HGIN_SELF_CONCAT.
 * workspace: hgin_gin_combine_post_workspace_bytes() bytes (per-CTA partials of post_dalpha).
 */
int64_t hgin_gin_combine_post_workspace_bytes(void);
int32_t hgin_gin_combine_post(int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                              int64_t num_edges, const float *x_src, int64_t ld_src, int32_t f_src,
                              const float *x_self, int64_t ld_self, int32_t f_self,
                              const float *eps, int32_t self_mode, int32_t accumulate,
                              float *out, int64_t ld_out, const float *post_z, int64_t ld_post,
                              int32_t post_act, const float *post_alpha, float *post_dalpha,
                              float *post_ddot, void *workspace, int64_t workspace_bytes, void *stream);

/* ---- K1 on pre-activation inputs ---------------------------------------------------------------
 * hgin_gin_combine whose x_src and/or x_self hold the PRE-activation z of the layer that produced
 * them: act(z) (models.py:238, PReLU / ReLU) is applied to every element as it is loaded, bit for
 * bit what that layer's `out` would have held, so the layer below never writes its activated
 * output — one row-sized store less per layer.  src_act / self_act = HGIN_ACT_NONE leaves that
 * input as it is.  Forward twin of hgin_gin_combine_post.
 */
int32_t hgin_gin_combine_pre(int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                             int64_t num_edges, const float *x_src, int64_t ld_src, int32_t f_src,
                             const float *x_self, int64_t ld_self, int32_t f_self,
                             const float *eps, int32_t self_mode, int32_t accumulate,
                             float *out, int64_t ld_out, int32_t src_act, const float *src_alpha,
                             int32_t self_act, const float *self_alpha, void *stream);

/* hgin_gin_combine_t: the three entry points above in one call, on rows of storage type `dtype`
 * (x_src, x_self, out, post_z all share it).  HGIN_DTYPE_BF16: neighbour rows are widened to fp32 as they are
 * loaded, added left to right in fp32 exactly as in the fp32 path, and the finished row is rounded to bf16 once
 * (round-to-nearest-even) on the store — on bf16-representable inputs the result equals the fp32 entry point's
 * result rounded to bf16, bit for bit.  src_act / self_act (pre-activation inputs) and post_z (post-activation)
 * are mutually exclusive as above; pass HGIN_ACT_NONE / NULL to disable either.
 */
int32_t hgin_gin_combine_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                           int64_t num_edges, const void *x_src, int64_t ld_src, int32_t f_src,
                           const void *x_self, int64_t ld_self, int32_t f_self, const float *eps,
                           int32_t self_mode, int32_t accumulate, void *out, int64_t ld_out,
                           int32_t src_act, const float *src_alpha, int32_t self_act,
                           const float *self_alpha, const void *post_z, int64_t ld_post,
                           int32_t post_act, const float *post_alpha, float *post_dalpha,
                           float *post_ddot, void *workspace, int64_t workspace_bytes, void *stream);

/* hgin_gin_combine_staged_t: K1 / K4 for LONG rows on a block-diagonal batch (the path->link aggregation of
 * models.py:211-217 and the backward of link->path: ~36 neighbours per row, every source row gathered ~2.9 times).
 * One CTA per SM takes a block at a time, keeps the block's output rows (<= 224) as fp32 accumulators in registers and
 * streams the block's source rows ONCE, in ascending order, through two shared-memory stages filled by cp.async.bulk;
 * the neighbours of a row that fall inside the staged chunk (a prefix of what is left of its ascending list) are added
 * left to right from shared memory (csrc/gin_stage_blocks.cuh).  Same CSR, same order of additions: bit-identical to
 * hgin_gin_combine_t.  The kernel runs only if gate (hgin_block_gate) reports containment (gate[0] == 0), ascending
 * neighbour lists (gate[3] == 0) and gate[1] <= 224; the kernel of hgin_gin_combine_t follows behind the inverse gate
 * (a static, capturable launch sequence).  Input rows must be contiguous (ld_src == f_src), 16 <= f_src <= 128;
 * SELF_NONE or SELF_ADD; src_act / self_act as in hgin_gin_combine_pre.  HGIN_ERR_UNSUPPORTED otherwise (the caller
 * then uses hgin_gin_combine_t).
 */
int32_t hgin_gin_combine_staged_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                                  int64_t num_edges, int32_t num_blocks, const int64_t *in_ptr,
                                  const int64_t *out_ptr, const int32_t *gate, const void *x_src,
                                  int64_t ld_src, int32_t f_src, const void *x_self, int64_t ld_self,
                                  const float *eps, int32_t self_mode, int32_t accumulate, void *out,
                                  int64_t ld_out, int32_t src_act, const float *src_alpha, int32_t self_act,
                                  const float *self_alpha, void *stream);

/* hgin_gin_combine_table_t: hgin_gin_combine_t for SHORT rows on a block-diagonal batch (the link->path aggregation of
 * models.py:211-217 and the backward of path->link: ~3 neighbours per row, all inside the row's own topology sample).
 * One 1024-thread CTA per SM takes a block at a time, stages the block's source rows (200 link rows = 100 KB fp32) in
 * shared memory once — applying src_act there, once per source element — and every gather is a shared-memory load.
 * Same CSR, same left-to-right additions: rows are bit-identical to hgin_gin_combine_t (post_dalpha / post_ddot are sums
 * over a different CTA partition: equal to fp32 rounding).  The kernel runs only if gate (hgin_block_gate) reports
 * containment (gate[0] == 0) and the source rows of every block fit (gate[2] * f_src * elem <= 220 KB); the kernel of
 * hgin_gin_combine_t follows behind the inverse gate, so the launch sequence is static and capturable.  Shapes without a
 * staged variant (rows that are not short, widths other than 64 / 128, unaligned rows) run hgin_gin_combine_t's kernel only.
 */
int32_t hgin_gin_combine_table_t(int32_t dtype, int64_t num_rows, const int32_t *rowptr, const int32_t *col,
                                 int64_t num_edges, int32_t num_blocks, const int64_t *in_ptr,
                                 const int64_t *out_ptr, const int32_t *gate, const void *x_src, int64_t ld_src,
                                 int32_t f_src, const void *x_self, int64_t ld_self, int32_t f_self,
                                 const float *eps, int32_t self_mode, int32_t accumulate, void *out,
                                 int64_t ld_out, int32_t src_act, const float *src_alpha, int32_t self_act,
                                 const float *self_alpha, const void *post_z, int64_t ld_post,
                                 int32_t post_act, const float *post_alpha, float *post_dalpha,
                                 float *post_ddot, void *workspace, int64_t workspace_bytes, void *stream);

/* ---- K2: dense layer forward  z = [x1 | x2] W^T + b,  out (+)= act(z) ------------------------
 * Replaces: GINLayer.mlp = Linear + PReLU (models.py:236-239, applied at models.py:217), the
 * HeteroConv 'sum' merge over relations with the same destination type (models.py:286-298 ->
 * PyG group(): stack().sum(0)) through `accumulate_out`, and the readout layers including
 * `torch.cat((x_dict['path'], origin_input['path']), 1)` (models.py:366, 373-374) through the
 * two-source input [x1 | x2].
 *   x1 [rows,k1], x2 [rows,k2] (x2 may be NULL with k2 = 0);  W [n, k1+k2] row-major (torch
 *   Linear.weight);  bias [n] or NULL;  alpha: device scalar for PRELU.
 *   z   [rows,n] or NULL — pre-activation, saved for backward;
 *   out [rows,n] or NULL — act(z), overwritten or accumulated.
 * math_mode HGIN_MATH_TF32 runs the tcgen05 kernel when the shapes are GEMM-sized (16 <= k1 <= 128,
 * k1 % 4 == 0, k2 <= 4, n % 16 == 0, n <= 128, rows >= 128, 16-byte aligned rows) and the fp32 SIMT
 * kernel otherwise; workspace: hgin_linear_fwd_workspace_bytes (0 bytes / NULL allowed for FP32).
 */
int64_t hgin_linear_fwd_workspace_bytes(int64_t rows, int32_t k, int32_t n, int32_t math_mode);
int32_t hgin_linear_fwd(int64_t rows, const float *x1, int64_t ld1, int32_t k1,
                        const float *x2, int64_t ld2, int32_t k2, const float *W,
                        const float *bias, int32_t n, int32_t act, const float *alpha,
                        float *z, int64_t ldz, float *out, int64_t ldo, int32_t accumulate_out,
                        void *workspace, int64_t workspace_bytes, int32_t math_mode, void *stream);

/* ---- K3: dense layer backward ---------------------------------------------------------------
 * Replaces: autograd of the above (train.py:43): with dz = g * act'(z),
 *     dx[:, c0:c1] = (dz W)[:, c0:c1]      -> dx   [rows, c1-c0]  (NULL to skip the store)
 *     dW = dz^T [x1 | x2]                  -> dW   [n, k1+k2]
 *     db = sum_rows dz                     -> db   [n]            (NULL to skip)
 *     dalpha = sum g * min(z,0) (PReLU)    -> dalpha [1]          (NULL to skip)
 *     ddot = sum dx[:, c0:c1] * dot_x      -> ddot [1]            (NULL to skip) — this is
 *            d(eps) = sum dh_self * x_dst of GINConv (models.py:213/215).
 * dW/db/dalpha/ddot are OVERWRITTEN (reductions are two-stage and deterministic: per-CTA
 * partials in `workspace`, summed in a fixed order).  Set dW = NULL to skip the weight pass,
 * c1 == c0 to skip the input pass.
 */
int64_t hgin_linear_bwd_workspace_bytes(int64_t rows, int32_t k, int32_t n, int32_t math_mode);
int32_t hgin_linear_bwd(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz,
                        int32_t act, const float *alpha, const float *x1, int64_t ld1, int32_t k1,
                        const float *x2, int64_t ld2, int32_t k2, const float *W, int32_t n,
                        int32_t c0, int32_t c1, float *dx, int64_t lddx, const float *dot_x,
                        int64_t ld_dot, float *ddot, float *dW, float *db, float *dalpha,
                        void *workspace, int64_t workspace_bytes, int32_t math_mode, void *stream);

/* hgin_linear_bwd with act == HGIN_ACT_NONE takes g as dz itself; in HGIN_MATH_TF32 both GEMMs then
 * read it in place and db is produced inside the weight-gradient MMA (a column of ones), so no
 * separate pass over g remains.
 *
 * hgin_linear_bwd_post: the same backward whose input gradient leaves as
 *     dx[:, c0:c1] = (dz W)[:, c0:c1] * act'(post_z),   post_dalpha[0] = sum (dz W) * min(post_z, 0)
 * with post_z [rows, c1-c0] the saved pre-activation of the layer that produced x (see
 * hgin_gin_combine_post).  Fused into the epilogue of the tcgen05 input-gradient kernel when the
 * shapes qualify, one extra elementwise pass otherwise.  No dot_x / ddot on this entry point.
 */
int32_t hgin_linear_bwd_post(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz,
                             int32_t act, const float *alpha, const float *x1, int64_t ld1,
                             int32_t k1, const float *x2, int64_t ld2, int32_t k2, const float *W,
                             int32_t n, int32_t c0, int32_t c1, float *dx, int64_t lddx, float *dW,
                             float *db, float *dalpha, const float *post_z, int64_t ld_post,
                             int32_t post_act, const float *post_alpha, float *post_dalpha,
                             void *workspace, int64_t workspace_bytes, int32_t math_mode,
                             void *stream);

/* hgin_linear_bwd_post_self: hgin_linear_bwd_post for a GIN layer whose ONLY contribution to the gradient of its
 * destination type's input is its own self branch (models.py:215 `out += (1 + eps) * x_r`; e.g. the last GIN layer,
 * whose output feeds the readout and whose source-side gradient is gathered separately).  The input gradient
 * dh = dz W (all k1 columns; no x2) then never reaches memory:
 *     dx          = (1 + eps) * dh * act'(post_z)         -> dz of the layer below (replaces the edgeless
 *                                                            hgin_gin_combine_post pass: 2 row-sized transfers less)
 *     post_dalpha = sum (1 + eps) * dh * min(post_z, 0)   (NULL to skip)
 *     post_ddot   = sum dh * act(post_z) = d(eps)         (NULL to skip; act(post_z) IS x_dst)
 * self_eps: device pointer to eps (NULL = 0).  Only the tensor-core input-gradient kernel carries this epilogue
 * (HGIN_MATH_TF32 / HGIN_MATH_BF16, GEMM-sized shapes): otherwise HGIN_ERR_UNSUPPORTED is returned and nothing is
 * enqueued — the caller then uses hgin_linear_bwd followed by hgin_gin_combine_post.
 */
int32_t hgin_linear_bwd_post_self(int64_t rows, const float *g, int64_t ldg, const float *z, int64_t ldz,
                                  int32_t act, const float *alpha, const float *x1, int64_t ld1,
                                  int32_t k1, const float *W, int32_t n, float *dx, int64_t lddx,
                                  float *dW, float *db, float *dalpha, const float *post_z,
                                  int64_t ld_post, int32_t post_act, const float *post_alpha,
                                  float *post_dalpha, const float *self_eps, float *post_ddot,
                                  void *workspace, int64_t workspace_bytes, int32_t math_mode,
                                  void *stream);

/* ---- K2 / K3 on typed rows --------------------------------------------------------------------------
 * hgin_linear_fwd_t / hgin_linear_bwd_t: hgin_linear_fwd and the hgin_linear_bwd family on row matrices stored as
 * float or bf16.  W, bias, x2 (the <= 4 raw input columns of the readout), every reduction (dW, db, dalpha, ddot)
 * and all accumulation are fp32.  Which matrices share which type:
 *     forward   in_dtype : x1                       out_dtype : z, out
 *     backward  g_dtype  : g, z                     x_dtype   : x1, dx, dot_x, post_z
 * Supported combinations (anything else returns HGIN_ERR_UNSUPPORTED and enqueues nothing):
 *     F32 / F32    the fp32 entry points above (math_mode picks SIMT fp32 or tcgen05 tf32)
 *     BF16 / BF16  GEMM-sized layers (16 <= k1 <= 128, k1 % 16 == 0, k2 <= 4, 16 <= n <= 128, n % 16 == 0,
 *                  rows >= 128, 16-byte aligned rows): tcgen05 kind::f16 on bf16 operands, fp32 accumulate in TMEM
 *     forward F32 -> BF16, backward g BF16 / x F32:  the K <= 8 layers (narrow fp32 input, wide bf16 output)
 *     forward BF16 -> F32, backward g F32 / x BF16:  the n = 1 head (wide bf16 input, one fp32 output column)
 * hgin_linear_bwd_t carries the post-activation (post_z ...: hgin_linear_bwd_post) and the GIN self branch
 * (self_eps, post_ddot: hgin_linear_bwd_post_self) as optional arguments: pass NULL / HGIN_ACT_NONE to disable.
 * Workspaces: hgin_linear_{fwd,bwd}_workspace_bytes with HGIN_MATH_BF16.
 */
int32_t hgin_linear_fwd_t(int32_t in_dtype, int32_t out_dtype, int64_t rows, const void *x1, int64_t ld1,
                          int32_t k1, const float *x2, int64_t ld2, int32_t k2, const float *W,
                          const float *bias, int32_t n, int32_t act, const float *alpha, void *z,
                          int64_t ldz, void *out, int64_t ldo, int32_t accumulate_out, void *workspace,
                          int64_t workspace_bytes, int32_t math_mode, void *stream);
int32_t hgin_linear_bwd_t(int32_t g_dtype, int32_t x_dtype, int64_t rows, const void *g, int64_t ldg,
                          const void *z, int64_t ldz, int32_t act, const float *alpha, const void *x1,
                          int64_t ld1, int32_t k1, const float *x2, int64_t ld2, int32_t k2,
                          const float *W, int32_t n, int32_t c0, int32_t c1, void *dx, int64_t lddx,
                          const void *dot_x, int64_t ld_dot, float *ddot, float *dW, float *db,
                          float *dalpha, const void *post_z, int64_t ld_post, int32_t post_act,
                          const float *post_alpha, float *post_dalpha, const float *self_eps,
                          float *post_ddot, void *workspace, int64_t workspace_bytes, int32_t math_mode,
                          void *stream);

/* ---- loss: sqrt(MAPE) (train.py:12-13, 40-42) -----------------------------------------------
 * hgin_mape_sum:       sums[0] = sum_i |(pred_i - y_i) / y_i|,  sums[1] = n  (fp32, two-stage,
 *                      deterministic).  Multi-GPU: the caller all-reduces `sums`.
 * hgin_sqrt_mape_bwd:  with S = sums[0], N = sums[1] (global), L = sqrt(100 S / N):
 *                      loss_out[0] = 100 S / N, loss_out[1] = L,
 *                      dpred_i = gscale * 50 * sign((pred_i - y_i) / y_i) / (y_i * N * L).
 */
int64_t hgin_reduce_workspace_bytes(int64_t n);
int32_t hgin_mape_sum(int64_t n, const float *pred, const float *y, float *sums, void *workspace,
                      int64_t workspace_bytes, void *stream);
int32_t hgin_sqrt_mape_bwd(int64_t n, const float *pred, const float *y, const float *sums,
                           float gscale, float *loss_out, float *dpred, void *stream);

/* ---- optimizer: Adam / AdamW on one flat fp32 bucket (train.py:44, 140-148) ------------------
 * torch.optim.Adam semantics (amsgrad off, maximize off): `step` is the 1-based step count read
 * from device memory (int32 [1]) so that the call is CUDA-graph replayable; the kernel does not
 * advance it (hgin_increment does).  decoupled != 0 selects AdamW.
 */
int32_t hgin_adam_step(int64_t n, float *param, const float *grad, float *exp_avg,
                       float *exp_avg_sq, const int32_t *step, double lr, double beta1,
                       double beta2, double eps, double weight_decay, int32_t decoupled,
                       void *stream);
int32_t hgin_increment(int32_t *counter, void *stream);

/* ---- device-side batch assembly (collate) ------------------------------------------------------
 * Replaces: PyG's collate — `Batch.from_data_list`, reached through
 * torch_geometric.loader.DataLoader (dataset.py:242-244) — and the per-step `sample.cuda()`
 * (train.py:28).  The dataset is resident in HBM as flat ARENAS, one per field (path.x, path.y,
 * link.x, node.x and, per relation, the per-sample CSR row pointers / columns in LOCAL ids), each
 * with an int64 table ptr[num_samples + 1] of first rows; a batch is assembled on the GPU from the
 * sample ids alone, so a step's host->device traffic is the id list.
 *
 * hgin_collate_offsets: offsets[c][b] (int64 [num_classes][batch + 1]) = exclusive prefix sum over
 *   the batch of (class_ptr[c][ids[b] + 1] - class_ptr[c][ids[b]]), class_ptr being int64
 *   [num_classes][num_samples + 1]; offsets[c][batch] is the batch total.  These are the per-type
 *   node offsets and per-relation edge offsets PyG adds when it concatenates `edge_index`.
 *   status: int32 [1], set to 1 when an id is outside [0, num_samples) (that sample is skipped).
 * hgin_collate_gather: for every field f and batch slot b, rows [ptr[id], ptr[id+1]) of f.src
 *   (width 4-byte elements per row) are copied to f.dst at row offsets[f.size_class][b]; with
 *   f.add_class >= 0 the elements are int32 and offsets[f.add_class][b] is added (CSR columns get
 *   the node offset of the type they index, row pointers the edge offset of their relation);
 *   f.closing_row = 1: every sample's block ends with a closing row (the n+1-th entry of a local
 *   row-pointer array) that is dropped for all but the LAST batch slot.  Integer-exact with the host collate.  `fields_host` is a HOST array (read during the
 *   call); max_words_per_sample sizes the grid (largest rows*width of one sample over the fields).
 */
typedef struct hgin_collate_field {
    const void *src;      /* device: arena of this field */
    void *dst;            /* device: start of this field in the batch */
    const int64_t *ptr;   /* device: int64 [num_samples + 1], first row of every sample in src */
    int32_t width;        /* 4-byte elements per row */
    int32_t size_class;   /* row of `offsets` holding this field's batch row offsets */
    int32_t add_class;    /* >= 0: row of `offsets` added to every (int32) element; -1: plain copy */
    int32_t closing_row;  /* 1: blocks end with a closing row kept only for the last batch slot */
} hgin_collate_field;

int32_t hgin_collate_offsets(int32_t batch, const int32_t *ids, int32_t num_classes,
                             const int64_t *class_ptr, int64_t num_samples, int64_t *offsets,
                             int32_t *status, void *stream);
int32_t hgin_collate_gather(int32_t batch, const int32_t *ids, int64_t num_samples,
                            int32_t num_fields, const hgin_collate_field *fields_host,
                            int32_t num_classes, const int64_t *offsets,
                            int64_t max_words_per_sample, void *stream);

/* Host twin of the two calls above for datasets that stay in HOST memory: every pointer (ids,
 * class_ptr, the fields' src / dst / ptr, offsets) is a host pointer; `dst` normally points into one
 * pinned, packed batch buffer that then crosses PCIe as a single copy.  Runs on `num_threads` host
 * threads (0 = all); no CUDA call, usable without a GPU.  Ids outside [0, num_samples) are an error.
 */
int32_t hgin_host_collate(int32_t batch, const int32_t *ids_host, int64_t num_samples,
                          int32_t num_fields, const hgin_collate_field *fields_host,
                          int32_t num_classes, const int64_t *class_ptr_host,
                          int64_t *offsets_host, int32_t num_threads);

/* ---- queueing-theory baseline (pre-processing features) -----------------------------------------
 * Replaces: QTBaseline.forward + separate_edge_timesteps (models.py:15-158), called once per
 * sample by GNN21Dataset.preprocess (dataset.py:86); its outputs are the `bl_features` columns of
 * link.x and path.x (dataset.py:105-106).  The reference runs it on the CPU as a Python loop over
 * hop positions; here one call handles a whole block-diagonal batch of samples.
 *   path->link relation as two CSRs over the SAME edge list (hgin_csr_build with perm):
 *     rowptr_src/col_src/perm_src : rows = paths, each row the path's links in ROUTE order
 *                                   (perm_src may be NULL when the edge list is already grouped by
 *                                   path, as the reference ships it: edge id == CSR position);
 *     rowptr_dst/perm_dst         : rows = links, perm_dst = edge ids of the incoming edges.
 *   avg_bw          [num_paths]  P[:,1], the traffic a path offers (models.py:94)
 *   capacity_scaled [num_links]  L / 1000 (models.py:73-74);  capacity_raw [num_links]  L
 *   num_iterations  fixed-point iterations (3 in the reference, models.py:43)
 *   path_delay [num_paths]       sum over the path's links of occupancy * 32000 / capacity_raw
 *   link_out   [num_links, 3]    occupancy, rho, pi_0 (as the reference returns them)
 * fp32 like the reference; agrees to rounding (powf), including the NaN pattern on overflow.
 */
int64_t hgin_qt_baseline_workspace_bytes(int64_t num_links, int64_t num_edges);
int32_t hgin_qt_baseline(int64_t num_paths, int64_t num_links, int64_t num_edges,
                         const int32_t *rowptr_src, const int32_t *col_src, const int32_t *perm_src,
                         const int32_t *rowptr_dst, const int32_t *perm_dst, const float *avg_bw,
                         const float *capacity_scaled, const float *capacity_raw,
                         int32_t num_iterations, float *path_delay, float *link_out,
                         void *workspace, int64_t workspace_bytes, void *stream);

/* ---- non-default branches of HetroGIN (row-streaming passes; dtype = HGIN_DTYPE_* of the row matrices) ------------
 *
 * hgin_act_fwd / hgin_act_bwd — replaces: the activation module applied after a readout Linear when it is not one the
 *   linear kernels fuse (`act = eval(act)`, models.py:301, 317-330 -> torch.nn.<Module>.forward / autograd).
 *     out = act(z);      dz = g * act'(z),  dalpha[0] = sum g * min(z, 0)  (NULL to skip; the PReLU slope gradient).
 *   workspace (only with dalpha): hgin_elementwise_workspace_bytes().
 *
 * hgin_dropout — replaces: torch.nn.functional.dropout(x_dict[k], p, training) (models.py:358-359).
 *     out[r][c] = keep(r, c) ? x[r][c] / (1 - p) : 0,  keep = (u(r, c) >= p),  u = Philox4x32-10(counter = (r * ceil(n/4) +
 *     c/4, offset), key = seed) word c%4 scaled to [0, 1).  The backward pass is the same call on the gradient with the same
 *     (seed, offset).  RNG streams differ from torch's by construction: parity is statistical.
 *
 * hgin_bn_* — replaces: torch.nn.BatchNorm1d between Linear and the activation (`mlp_bn`, models.py:303-313).
 *   hgin_bn_stats:     sums[0:n] = column sums of z, sums[n:2n] = column sums of squares, sums[2n] = rows  (fp64,
 *                      two-stage fixed order).  Under data parallelism the caller all-reduces `sums` (2n+1 doubles) so
 *                      that every rank normalises with the statistics of the GLOBAL batch (= the single-process step).
 *   hgin_bn_finalize:  mean, invstd = 1/sqrt(biased var + eps) from sums; running_mean / running_var (NULL to skip)
 *                      updated with `momentum` and the unbiased variance as torch does.  use_running != 0 (eval mode):
 *                      mean / invstd come from the running buffers and nothing is updated.
 *   hgin_bn_act_fwd:   out = act(gamma * (z - mean) * invstd + beta)            (gamma / beta NULL = 1 / 0)
 *   hgin_bn_act_bwd_reduce:  with y recomputed from z and dy = g * act'(y):  sums[0:n] = sum dy (= dbeta),
 *                      sums[n:2n] = sum dy * xhat (= dgamma), sums[2n] = sum g * min(y, 0) (= dalpha).  All-reduced by the
 *                      caller under data parallelism.
 *   hgin_bn_act_bwd_apply:   dz = gamma * invstd * (dy - sums[c]/count - xhat * sums[n+c]/count)  (training != 0; count =
 *                      GLOBAL row count) or gamma * invstd * dy (eval); writes dgamma / dbeta / dalpha (fp32, NULL to skip).
 *   workspace: hgin_bn_workspace_bytes(rows, n).
 *
 * hgin_segment_pool / hgin_readout_tail — replaces: global_mean_pool / global_max_pool of the raw path features over
 *   `path_batch` and the two torch.gather calls that broadcast them back (models.py:347-352), plus the torch.cat of the
 *   constant readout columns (models.py:364-369).  rowptr / col: CSR of the graph ids (hgin_csr_build with
 *   key = path_batch, other = arange; col == NULL means rows are already grouped, row id == CSR position).  Rows are added
 *   left to right in CSR order (bit-exact with the CPU scatter_add_), mean = sum / max(count, 1).
 *     tail[p] = [ origin[p][0:f_origin] | mean[segment_ids[p]] | max[segment_ids[p]] ],  segment ids int64 or int32.
 */
int64_t hgin_elementwise_workspace_bytes(void);
int32_t hgin_act_fwd(int32_t dtype, int64_t rows, int32_t n, const void *z, int64_t ldz, int32_t act,
                     const float *alpha, float p0, float p1, void *out, int64_t ldo, void *stream);
int32_t hgin_act_bwd(int32_t dtype, int64_t rows, int32_t n, const void *g, int64_t ldg, const void *z,
                     int64_t ldz, int32_t act, const float *alpha, float p0, float p1, void *dz,
                     int64_t lddz, float *dalpha, void *workspace, int64_t workspace_bytes, void *stream);
int32_t hgin_dropout(int32_t dtype, int64_t rows, int32_t n, const void *x, int64_t ldx, float p,
                     uint64_t seed, uint64_t offset, void *out, int64_t ldo, void *stream);
int64_t hgin_bn_workspace_bytes(int64_t rows, int32_t n);
int32_t hgin_bn_stats(int32_t dtype, int64_t rows, int32_t n, const void *z, int64_t ldz, double *sums,
                      void *workspace, int64_t workspace_bytes, void *stream);
int32_t hgin_bn_finalize(int32_t n, const double *sums, double eps, double momentum, int32_t use_running,
                         float *mean, float *invstd, float *running_mean, float *running_var, void *stream);
int32_t hgin_bn_act_fwd(int32_t dtype, int64_t rows, int32_t n, const void *z, int64_t ldz,
                        const float *mean, const float *invstd, const float *gamma, const float *beta,
                        int32_t act, const float *alpha, float p0, float p1, void *out, int64_t ldo,
                        void *stream);
int32_t hgin_bn_act_bwd_reduce(int32_t dtype, int64_t rows, int32_t n, const void *g, int64_t ldg,
                               const void *z, int64_t ldz, const float *mean, const float *invstd,
                               const float *gamma, const float *beta, int32_t act, const float *alpha,
                               float p0, float p1, double *sums, void *workspace, int64_t workspace_bytes,
                               void *stream);
int32_t hgin_bn_act_bwd_apply(int32_t dtype, int64_t rows, int32_t n, const void *g, int64_t ldg,
                              const void *z, int64_t ldz, const float *mean, const float *invstd,
                              const float *gamma, const float *beta, int32_t act, const float *alpha,
                              float p0, float p1, const double *sums, double count, int32_t training,
                              void *dz, int64_t lddz, float *dgamma, float *dbeta, float *dalpha,
                              void *stream);
int32_t hgin_segment_pool(int64_t segments, const int32_t *rowptr, const int32_t *col, const float *x,
                          int64_t ldx, int32_t f, float *mean_out, float *max_out, void *stream);
int32_t hgin_readout_tail(int64_t rows, const void *segment_ids, int32_t index_bytes, int64_t segments,
                          const float *origin, int64_t ld_origin, int32_t f_origin, const float *mean_in,
                          const float *max_in, int32_t f, float *tail, int64_t ld_tail, void *stream);

/* ---- graph attention (HetroGAT) --------------------------------------------------------------------------------
 * Replaces: torch_geometric.nn.conv.GATConv.forward from the attention logits on (models.py:417-428 construct it, models.py:466
 * calls it through HeteroConv) — `remove_self_loops` / `add_self_loops`, `propagate` with `message` = x_j * softmax_i(
 * leaky_relu(alpha_j + alpha_i)) and aggr='add', `+ bias` — for bipartite inputs, concat=True, no attention dropout.
 * The dense projections around it (lin_src, and the attention logits a_src = <lin_src(x_src), att_src> per head, a_dst
 * likewise) are hgin_linear_fwd calls.
 *   rowptr / col : destination-sorted CSR of the relation (hgin_csr_build, sort_row = 1); rows = destinations.
 *   xs    [num_src, heads*channels]  projected source features;  a_src [num_src, heads], a_dst [num_rows, heads]
 *   add_self_loops != 0: PyG's bipartite rule — edges whose source id equals their destination id are dropped and one
 *                 loop (i, i) is appended (last in the sum) for every i < min(num_src, num_rows).
 *   out[i] (+)= sum_j w_ij xs[j] / (sum_j w_ij + 1e-16) + bias,   w_ij = exp(e_ij - max_j e_ij)   (accumulate != 0 adds to
 *                 `out`: the HeteroConv sum merge);  row_max / row_sum [num_rows, heads] are saved for hgin_gat_bwd.
 * hgin_gat_bwd: g = d loss / d out.  Destination pass over the same CSR: d_a_dst [num_dst, heads] (dot_ws: 16 * num_dst * heads bytes of
 *   16-byte-aligned scratch: one (a_dst, row max, 1 / row sum, softmax dot) record per destination and head, read by the source pass).  Source pass over the TRANSPOSED CSR (rowptr_src / col_src, sort_row = 0): d_xs [num_src, heads*channels],
 *   d_a_src [num_src, heads].  d bias is the column sum of g (hgin_bn_stats computes it).  No atomics, nothing per edge is
 *   stored, deterministic.  channels: a power of two in [4, 128]; heads * channels <= 512 (HGIN_ERR_UNSUPPORTED otherwise).
 */
int32_t hgin_gat_fwd(int64_t num_rows, const int32_t *rowptr, const int32_t *col, int64_t num_src,
                     const float *xs, int64_t ld_xs, const float *a_src, const float *a_dst,
                     const float *bias, int32_t heads, int32_t channels, float negative_slope,
                     int32_t add_self_loops, int32_t accumulate, float *out, int64_t ld_out,
                     float *row_max, float *row_sum, void *stream);
int32_t hgin_gat_bwd(int64_t num_dst, const int32_t *rowptr_dst, const int32_t *col_dst, int64_t num_src,
                     const int32_t *rowptr_src, const int32_t *col_src, const float *xs, int64_t ld_xs,
                     const float *a_src, const float *a_dst, const float *row_max, const float *row_sum,
                     const float *g, int64_t ld_g, int32_t heads, int32_t channels, float negative_slope,
                     int32_t add_self_loops, float *d_xs, int64_t ld_dxs, float *d_a_src, float *d_a_dst,
                     void *dot_ws, void *stream);

/* ---- the whole train step of config.json's model in three kernels (launch-bound regime) ---------------------------------
 * Replaces: one iteration of train.py:31-44 — HetroGIN.forward (models.py:332-376), mape + sqrt (train.py:12-13, 40-42),
 * loss.backward() — for the model family of config.json: MP_LAYERS = 1 (only link -> path reaches the readout), GIN layer
 * with concat=True, readout Linear-PReLU-Linear-PReLU-Linear with ONE shared PReLU slope, no BatchNorm / global features /
 * dropout.  After the neighbour sum the network is row-local in the path rows, so the forward is recomputed in the
 * backward kernel instead of being stored, and the weight gradients are accumulated in registers per CTA (fixed ownership,
 * no atomics, deterministic).  The optimizer step stays hgin_adam_step.
 *   rowptr / col      destination-sorted CSR of ('link', 'includes', 'path')
 *   x_path / x_link   RAW feature matrices (7 columns in the reference); *_cols_host: HOST arrays naming the f_path / f_link
 *                     columns the model slices out (models.py:333-342), e.g. {0,1,2}
 *   W0 [emb, f_link + f_path], b0, alpha0, eps0: the GIN layer;  W1 [n1, emb + (concat_path ? f_path : 0)], b1, W2 [n2, n1],
 *   b2, W3 [1, n2], b3, alpha_r: the readout.  d*: gradients, same shapes.  sums [2] = (sum |(out - y)/y|, N);
 *   loss_out [2] = (mape, sqrt(mape));  out [num_paths] (NULL to skip): the scores.
 * Limits: f_path, f_link <= 8; emb <= 32; emb + f_path <= 32; n1 <= 128; n2 <= 32 (HGIN_ERR_UNSUPPORTED otherwise).
 * hgin_small_step_phase: the same step split around the loss statistics, for data parallelism (the loss is
 *   sqrt of a GLOBAL mean).  phase 1: forward, `sums` = this rank's (S, N) — the caller all-reduces it;  phase 2: backward
 *   with `sums` holding the global (S, N) and the SAME workspace as the phase-1 call (it holds the per-row pre-activations
 *   the forward kernel left for the backward kernel, 1 KB per path), gradients = this rank's partial sums — the caller
 *   all-reduces the bucket.
 */
int64_t hgin_small_step_workspace_bytes(int64_t num_paths);
int32_t hgin_small_step(int64_t num_paths, const int32_t *rowptr, const int32_t *col, const float *x_path,
                        int64_t ld_path, int32_t f_path, const int32_t *path_cols_host, const float *x_link,
                        int64_t ld_link, int32_t f_link, const int32_t *link_cols_host, const float *y,
                        int32_t emb, int32_t n1, int32_t n2, int32_t concat_path, const float *W0,
                        const float *b0, const float *alpha0, const float *eps0, const float *W1,
                        const float *b1, const float *alpha_r, const float *W2, const float *b2,
                        const float *W3, const float *b3, float *dW0, float *db0, float *dalpha0,
                        float *deps0, float *dW1, float *db1, float *dalpha_r, float *dW2, float *db2,
                        float *dW3, float *db3, float *sums, float *loss_out, float *out, void *workspace,
                        int64_t workspace_bytes, void *stream);
int32_t hgin_small_step_phase(int32_t phase, int64_t num_paths, const int32_t *rowptr, const int32_t *col,
                              const float *x_path, int64_t ld_path, int32_t f_path,
                              const int32_t *path_cols_host, const float *x_link, int64_t ld_link,
                              int32_t f_link, const int32_t *link_cols_host, const float *y, int32_t emb,
                              int32_t n1, int32_t n2, int32_t concat_path, const float *W0, const float *b0,
                              const float *alpha0, const float *eps0, const float *W1, const float *b1,
                              const float *alpha_r, const float *W2, const float *b2, const float *W3,
                              const float *b3, float *dW0, float *db0, float *dalpha0, float *deps0,
                              float *dW1, float *db1, float *dalpha_r, float *dW2, float *db2, float *dW3,
                              float *db3, float *sums, float *loss_out, float *out, void *workspace,
                              int64_t workspace_bytes, void *stream);

/* ---- runtime options --------------------------------------------------------------------------
 * "fused_bwd" (0/1, default 0): HGIN_MATH_TF32 backward through the single-pass fused kernel
 * (csrc/linear_tc_fused.cuh) instead of separate passes.
 * "fused_dw" (0/1, default 0): dz = g*act'(z) fused into the weight-gradient kernel
 * (csrc/linear_tc_dw.cuh) instead of dz_prepare + gemm_tn.  Process-wide, not thread-safe.
 */
int32_t hgin_set_option(const char *name, int32_t value);

/* ---- diagnostics ------------------------------------------------------------------------------
 * out[n,k] = a[rows,n]^T * b[rows,k] through the tcgen05 MN-major weight-gradient kernel with the
 * UMMA shared-memory descriptor fields given explicitly (tests pin the layout with it):
 * tma_swizzle = CUtensorMapSwizzle value, lbo / sbo / k_step_bytes in bytes, layout_type = UMMA
 * layout type.  Pass -1 for any field to use the library default.
 */
int32_t hgin_debug_gemm_tn(int64_t rows, const float *a, int32_t n, const float *b, int32_t k,
                           float *out, void *workspace, int64_t workspace_bytes, int32_t tma_swizzle,
                           int32_t lbo, int32_t sbo, int32_t layout_type, int32_t k_step_bytes,
                           void *stream);

/* The same for the bf16 weight-gradient kernel (a, b: bf16 [rows, n] / [rows, k]; plain SWIZZLE_128B tensor maps). */
int32_t hgin_debug_gemm_tn_bf16(int64_t rows, const void *a, int32_t n, const void *b, int32_t k,
                                float *out, void *workspace, int64_t workspace_bytes, int32_t lbo,
                                int32_t sbo, int32_t layout_type, int32_t k_step_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HGIN_H_ */
