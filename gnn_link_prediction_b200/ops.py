"""Tensor-level wrappers over the C ABI (include/hgin.h).

PyTorch is used here for device memory, streams and nothing else: every function validates its
tensors (CUDA, fp32/int32, unit inner stride), allocates outputs/workspaces with `torch.empty`,
passes raw pointers + the current stream to libhgin.so and raises `HginError` on a non-zero
status.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from .profiling import combine_bytes
from ._lib import (ACT_ELU, ACT_GELU, ACT_LEAKY_RELU, ACT_SIGMOID, ACT_SILU, ACT_SOFTPLUS, ACT_TANH,  # noqa: F401
                   ACT_NONE, ACT_PRELU, ACT_RELU, DTYPE_BF16, DTYPE_F32, MATH_BF16, MATH_FP32, MATH_TF32,  # noqa: F401
                   SELF_ADD, SELF_CONCAT, SELF_NONE, HginError, check)


TIMER = None  # set to a profiling.KernelTimer by bench.py to attribute time per kernel


class _NoRegion:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_REGION = _NoRegion()


def _region(name, **meta):
    return TIMER.region(name, **meta) if TIMER is not None and TIMER.enabled else _NO_REGION


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32_matrix(t, name):
    if t is None:
        return 0, 0
    if not t.is_cuda:
        raise HginError(f"{name}: expected a CUDA tensor (there is no CPU fallback), got device {t.device}")
    if t.dtype != torch.float32 or t.dim() != 2:
        raise HginError(f"{name}: expected a 2-D float32 tensor, got {tuple(t.shape)} {t.dtype}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise HginError(f"{name}: inner stride must be 1, got strides {t.stride()}")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    if ld < t.shape[1]:
        raise HginError(f"{name}: row stride {ld} < width {t.shape[1]} (overlapping rows)")
    return t.data_ptr(), ld


def _ptr(t):
    return 0 if t is None else t.data_ptr()


_DTYPES = {torch.float32: DTYPE_F32, torch.bfloat16: DTYPE_BF16}


def _matrix(t, name, dtype=None):
    """(pointer, leading dimension in elements) of a 2-D row matrix stored as float32 or bfloat16
    (HGIN_DTYPE_*); `dtype` pins the expected torch dtype (all row matrices of one call share it)."""
    if t is None:
        return 0, 0
    if not t.is_cuda:
        raise HginError(f"{name}: expected a CUDA tensor (there is no CPU fallback), got device {t.device}")
    if t.dtype not in _DTYPES or t.dim() != 2:
        raise HginError(f"{name}: expected a 2-D float32 / bfloat16 tensor, got {tuple(t.shape)} {t.dtype}")
    if dtype is not None and t.dtype != dtype:
        raise HginError(f"{name}: dtype {t.dtype} does not match the other row matrices of the call ({dtype})")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise HginError(f"{name}: inner stride must be 1, got strides {t.stride()}")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    if ld < t.shape[1]:
        raise HginError(f"{name}: row stride {ld} < width {t.shape[1]} (overlapping rows)")
    return t.data_ptr(), ld


def _scalar(t, name):
    if t is None:
        return 0
    if not t.is_cuda or t.dtype != torch.float32 or t.numel() != 1:
        raise HginError(f"{name}: expected a 1-element CUDA float32 tensor")
    return t.data_ptr()


class CSR:
    """Row-sorted adjacency: rows own their neighbour lists in stable edge order."""

    __slots__ = ("rowptr", "col", "perm", "status", "num_rows", "num_cols", "num_edges")

    def __init__(self, rowptr, col, perm, status, num_rows, num_cols, num_edges):
        self.rowptr, self.col, self.perm, self.status = rowptr, col, perm, status
        self.num_rows, self.num_cols, self.num_edges = num_rows, num_cols, num_edges

    def validate(self):
        """Synchronising check of the out-of-range flag set by hgin_csr_build."""
        if int(self.status.item()) != 0:
            raise IndexError("edge_index holds node ids outside [0, num_src) x [0, num_dst)")
        return self


def csr_build(edge_index, num_src, num_dst, by="dst", want_perm=False):
    """K0.  edge_index: CUDA int64/int32 [2,E] (row 0 = source ids, row 1 = destination ids).
    by='dst' -> forward CSR (rows = destinations, cols = sources);
    by='src' -> transposed CSR (rows = sources, cols = destinations)."""
    if not edge_index.is_cuda:
        raise HginError("csr_build: edge_index must live on the GPU (there is no CPU fallback)")
    if edge_index.dim() != 2 or edge_index.shape[0] != 2 or edge_index.dtype not in (torch.int64, torch.int32):
        raise HginError(f"csr_build: expected int64/int32 [2,E], got {tuple(edge_index.shape)} {edge_index.dtype}")
    E = edge_index.shape[1]
    if E > 0 and edge_index.stride(1) != 1:
        edge_index = edge_index.contiguous()
    ld_edge = edge_index.stride(0) if E > 0 else 0
    sort_row = 1 if by == "dst" else 0
    rows, cols = (num_dst, num_src) if sort_row else (num_src, num_dst)
    dev = edge_index.device
    lib = _lib.load()
    rowptr = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    perm = torch.empty(E, dtype=torch.int32, device=dev) if want_perm else None
    status = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.hgin_csr_workspace_bytes(E, rows)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    n_kernels = (3 if E > 0 else 0) + (1 if rows + 1 <= 4096 else 3)
    with _region("csr_build", kernels=n_kernels, bytes=E * (2 * edge_index.element_size() + 16) + rows * 12):
        check(lib.hgin_csr_build(edge_index.data_ptr(), edge_index.element_size(), E, max(ld_edge, E), sort_row,
                                 rows, cols, rowptr.data_ptr(), _ptr(col), _ptr(perm), status.data_ptr(),
                                 ws.data_ptr(), ws_bytes, _stream()), "hgin_csr_build")
    return CSR(rowptr, col, perm, status, rows, cols, E)


class PostAct:
    """Activation derivative of the layer BELOW, applied by whichever kernel produces the gradient
    w.r.t. that layer's output (hgin_gin_combine_post / hgin_linear_bwd_post): the producer stores
    dz = grad * act'(z) and fills `dalpha` = sum grad * min(z, 0); the layer below then runs its
    backward with ACT_NONE on dz.  `applied` tells that layer this has happened."""

    __slots__ = ("z", "act", "alpha", "dalpha", "applied", "lazy")

    def __init__(self, z, act, alpha, lazy=False):
        self.z, self.act, self.alpha = z, act, alpha
        self.dalpha, self.applied = None, False
        # lazy: the layer did NOT write its activated output; the tensor handed downstream is z itself and
        # every consumer applies act on load (hgin_gin_combine_pre) — and MUST apply the post-activation
        self.lazy = lazy

    def usable(self):
        return self.z is not None and self.act != ACT_NONE


class _NoEdges:
    """Stand-in adjacency for a self-term-only pass (rowptr = NULL at the ABI)."""

    def __init__(self, num_rows):
        self.num_rows, self.num_cols, self.num_edges, self.rowptr, self.col = num_rows, num_rows, 0, None, None


class StreamPlan:
    """What hgin_gin_combine_blocks_t needs on top of the output-major CSR: the input-major CSR of the same relation, the
    block tables of the two node types (`ptr` vectors of the batch, int64 on the device) and the gate computed once per
    batch by hgin_block_gate (functional.GraphCSR.stream_plan builds and caches it)."""

    __slots__ = ("csr_in", "in_ptr", "out_ptr", "gate", "num_blocks")

    def __init__(self, csr_in, in_ptr, out_ptr, gate):
        self.csr_in, self.in_ptr, self.out_ptr, self.gate = csr_in, in_ptr, out_ptr, gate
        self.num_blocks = in_ptr.numel() - 1


def block_gate(csr_out, csr_in, in_ptr, out_ptr):
    """int32 [4] on the device: containment violations, max output rows of a block, max input rows of a block, output rows
    with a non-ascending neighbour list (include/hgin.h: hgin_block_gate)."""
    for t in (in_ptr, out_ptr):
        if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() and t.dim() == 1):
            raise HginError("block_gate: block tables must be contiguous CUDA int64 vectors")
    if in_ptr.numel() != out_ptr.numel() or in_ptr.numel() < 2:
        raise HginError("block_gate: block tables of the two node types differ in length")
    gate = torch.empty(4, dtype=torch.int32, device=in_ptr.device)
    with _region("csr_build", kernels=1, bytes=8 * csr_out.num_edges):
        check(_lib.load().hgin_block_gate(csr_out.num_rows, _ptr(csr_out.rowptr), _ptr(csr_out.col), csr_in.num_rows,
                                          _ptr(csr_in.rowptr), _ptr(csr_in.col), in_ptr.numel() - 1, in_ptr.data_ptr(),
                                          out_ptr.data_ptr(), gate.data_ptr(), _stream()), "hgin_block_gate")
    return gate


# Opt-in (measured and rejected as the default, DESIGN.md "Long rows"): long-row aggregations of block-diagonal batches
# through the input-major streaming kernel.  Bit-identical to the gather kernel, but every edge costs a 512-byte
# read-modify-write of shared memory (3 x E x F x 4 B = 11 GB per path->link launch against ~36 TB/s of aggregate
# shared-memory bandwidth, a 310 us floor before any latency) where the gather kernel keeps its accumulators in
# registers and is bound by L2->SM bandwidth (336 us): 1.5 ms against 0.34-0.44 ms per launch at Cfg-C.
STREAM_LONG_ROWS = False

# Opt-in (measured, DESIGN.md "Long rows"): long-row aggregations of block-diagonal batches through the staged-source kernel
# (hgin_gin_combine_staged_t): register accumulators, the block's source rows streamed ONCE through two shared-memory stages
# (csrc/gin_stage_blocks.cuh).  Bit-identical; 424 us against 364 us per Cfg-C path->link launch in fp32 and 339 against
# 247 us in bf16 (tools/staged_probe.py): the per-row list walk costs more issue slots than the L2 re-reads it removes.
STAGE_LONG_ROWS = False

# Opt-in (measured at parity, DESIGN.md "Short rows"): short-row aggregations of block-diagonal batches through the
# shared-memory table variant (hgin_gin_combine_table_t) — the source rows of one topology sample staged in shared memory
# once, every gather a shared-memory load.  Bit-identical rows; at Cfg-C 6.05 ms of aggregation per tf32 step against
# 5.92 ms and 4.89 against 4.90 ms in bf16: the contiguous row ranges of the default kernel already keep a sample's 100 KB of
# link rows in L1, and both variants are bound by the self / post / output rows they stream, so the default stays the
# kernel without the extra gate and skipped-twin launches.
TABLE_SHORT_ROWS = False


def gin_combine(csr, x_src, x_self=None, eps=None, self_mode=SELF_NONE, out=None, accumulate=False, post=None,
                want_ddot=False, src_act=None, self_act=None, block_plan=None):
    """K1/K4.  out[r] (+)= sum_{e in row r} x_src[col[e]]  {+ | concat}  (1+eps) * x_self[r].
    csr=None: no edges (self term only; x_src is then only a shape donor).
    post: a PostAct — the stored result is multiplied by act'(post.z) and post.dalpha is filled.
    want_ddot (with post, SELF_ADD): returns (out, ddot) with ddot = sum x_self * act(post.z).
    src_act / self_act: (act, alpha) when x_src / x_self hold PRE-activations (act applied on load).
    block_plan: a StreamPlan — the batch is block-diagonal and this is a long-row aggregation: input-major streaming kernel
    with the gather kernel behind an inverse device-side gate (same bits either way; hgin_gin_combine_blocks_t).
    Rows may be float32 or bfloat16 (all of x_src, x_self, post.z, out alike): bf16 rows are widened on load,
    summed in fp32 in CSR order and rounded once on the store (hgin_gin_combine_t)."""
    if csr is None:
        csr = _NoEdges(x_self.shape[0])
    dt = x_src.dtype
    ps, lds = _matrix(x_src, "gin_combine.x_src")
    pf, ldf = _matrix(x_self, "gin_combine.x_self", dt)
    if x_src.shape[0] != csr.num_cols:
        raise HginError(f"gin_combine: x_src has {x_src.shape[0]} rows, CSR expects {csr.num_cols}")
    if x_self is not None and x_self.shape[0] != csr.num_rows:
        raise HginError(f"gin_combine: x_self has {x_self.shape[0]} rows, CSR has {csr.num_rows}")
    f_src = x_src.shape[1]
    f_self = 0 if x_self is None else x_self.shape[1]
    width = f_src + (f_self if self_mode == SELF_CONCAT else 0)
    if out is None:
        if accumulate:
            raise HginError("gin_combine: accumulate needs an existing `out`")
        out = torch.empty(csr.num_rows, width, dtype=dt, device=x_src.device)
    po, ldo = _matrix(out, "gin_combine.out", dt)
    if tuple(out.shape) != (csr.num_rows, width):
        raise HginError(f"gin_combine: out is {tuple(out.shape)}, expected {(csr.num_rows, width)}")
    es = x_src.element_size()
    alg, comp = combine_bytes(csr.num_rows, csr.num_cols, csr.num_edges, f_src,
                              f_self if self_mode != SELF_NONE else 0, width * (2 if accumulate else 1), es)
    lib = _lib.load()
    post_on = post is not None and post.act != ACT_NONE
    if want_ddot and not post_on:
        raise HginError("gin_combine: want_ddot needs a post-activation")
    pz = ldz = 0
    ddot = ws = None
    ws_bytes = 0
    kernels = 1
    if post_on:
        if src_act is not None or self_act is not None:
            raise HginError("gin_combine: input activations and a post-activation cannot be combined")
        pz, ldz = _matrix(post.z, "gin_combine.post.z", dt)
        if tuple(post.z.shape) != (csr.num_rows, f_src) or self_mode == SELF_CONCAT:
            raise HginError(f"gin_combine: post.z is {tuple(post.z.shape)}, expected {(csr.num_rows, f_src)}")
        want = post.act == ACT_PRELU
        post.dalpha = torch.empty(1, dtype=torch.float32, device=out.device) if want else None
        ddot = torch.empty(1, dtype=torch.float32, device=out.device) if want_ddot else None
        ws_bytes = lib.hgin_gin_combine_post_workspace_bytes()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device)
        extra = es * csr.num_rows * f_src     # the pre-activation rows read on top of the plain pass
        alg, comp = alg + extra, comp + extra
        kernels = 1 + int(want) + int(want_ddot)
    sa, sal = src_act if src_act is not None else (ACT_NONE, None)
    fa, fal = self_act if (self_act is not None and x_self is not None) else (ACT_NONE, None)
    if (block_plan is not None and STAGE_LONG_ROWS and not STREAM_LONG_ROWS and not post_on and self_mode in (SELF_NONE, SELF_ADD)
            and csr.rowptr is not None and csr.num_edges > 8 * csr.num_rows and 16 <= f_src <= 128 and f_src % (16 // es) == 0
            and lds == f_src and ldo % 4 == 0 and (x_self is None or ldf % 4 == 0) and x_src.data_ptr() % 16 == 0
            and out.data_ptr() % 16 == 0 and (x_self is None or x_self.data_ptr() % 16 == 0)):
        with _region("gin_combine", kernels=2, alg_bytes=alg, compulsory_bytes=comp):
            check(lib.hgin_gin_combine_staged_t(_DTYPES[dt], csr.num_rows, _ptr(csr.rowptr), _ptr(csr.col), csr.num_edges,
                                                block_plan.num_blocks, block_plan.in_ptr.data_ptr(),
                                                block_plan.out_ptr.data_ptr(), block_plan.gate.data_ptr(),
                                                ps, lds, f_src, pf, ldf, _scalar(eps, "gin_combine.eps"), self_mode,
                                                1 if accumulate else 0, po, ldo, sa, _scalar(sal, "gin_combine.src_alpha"), fa,
                                                _scalar(fal, "gin_combine.self_alpha"), _stream()), "hgin_gin_combine_staged_t")
        return out
    if (block_plan is not None and STREAM_LONG_ROWS and not post_on and self_mode in (SELF_NONE, SELF_ADD) and csr.rowptr is not None
            and 32 <= f_src <= 128 and f_src % (16 // es) == 0 and x_src.shape[0] > 1 and lds == f_src and ldo % 4 == 0
            and (x_self is None or ldf % 4 == 0) and x_src.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0
            and (x_self is None or x_self.data_ptr() % 16 == 0)):
        cin = block_plan.csr_in
        with _region("gin_combine", kernels=2, alg_bytes=alg, compulsory_bytes=comp):
            check(lib.hgin_gin_combine_blocks_t(_DTYPES[dt], csr.num_rows, _ptr(csr.rowptr), _ptr(csr.col), csr.num_edges,
                                                cin.num_rows, _ptr(cin.rowptr), _ptr(cin.col), block_plan.num_blocks,
                                                block_plan.in_ptr.data_ptr(), block_plan.out_ptr.data_ptr(),
                                                block_plan.gate.data_ptr(),
                                                ps, lds, f_src, pf, ldf, _scalar(eps, "gin_combine.eps"), self_mode,
                                                1 if accumulate else 0, po, ldo, sa, _scalar(sal, "gin_combine.src_alpha"), fa,
                                                _scalar(fal, "gin_combine.self_alpha"), _stream()), "hgin_gin_combine_blocks_t")
        return out
    if (block_plan is not None and TABLE_SHORT_ROWS and csr.rowptr is not None and csr.col is not None and csr.num_rows > 0
            and csr.num_edges <= 8 * csr.num_rows and f_src * es in (256, 512) and self_mode != SELF_CONCAT):
        with _region("gin_combine", kernels=kernels + 1, alg_bytes=alg, compulsory_bytes=comp):
            check(lib.hgin_gin_combine_table_t(_DTYPES[dt], csr.num_rows, _ptr(csr.rowptr), _ptr(csr.col), csr.num_edges,
                                               block_plan.num_blocks, block_plan.in_ptr.data_ptr(),
                                               block_plan.out_ptr.data_ptr(), block_plan.gate.data_ptr(), ps, lds,
                                               f_src, pf, ldf, f_self, _scalar(eps, "gin_combine.eps"), self_mode,
                                               1 if accumulate else 0, po, ldo, sa, _scalar(sal, "gin_combine.src_alpha"), fa,
                                               _scalar(fal, "gin_combine.self_alpha"), pz, ldz,
                                               post.act if post_on else ACT_NONE,
                                               _scalar(post.alpha, "gin_combine.post.alpha") if post_on else 0,
                                               _ptr(post.dalpha) if post_on else 0, _ptr(ddot), _ptr(ws), ws_bytes, _stream()),
                  "hgin_gin_combine_table_t")
        if post_on:
            post.applied = True
        return (out, ddot) if want_ddot else out
    with _region("gin_combine", kernels=kernels, alg_bytes=alg, compulsory_bytes=comp):
        check(lib.hgin_gin_combine_t(_DTYPES[dt], csr.num_rows, _ptr(csr.rowptr), _ptr(csr.col), csr.num_edges, ps, lds,
                                     f_src, pf, ldf, f_self, _scalar(eps, "gin_combine.eps"), self_mode,
                                     1 if accumulate else 0, po, ldo, sa, _scalar(sal, "gin_combine.src_alpha"), fa,
                                     _scalar(fal, "gin_combine.self_alpha"), pz, ldz, post.act if post_on else ACT_NONE,
                                     _scalar(post.alpha, "gin_combine.post.alpha") if post_on else 0,
                                     _ptr(post.dalpha) if post_on else 0, _ptr(ddot), _ptr(ws), ws_bytes, _stream()),
              "hgin_gin_combine_t")
    if post_on:
        post.applied = True
        return (out, ddot) if want_ddot else out
    return out


def _pow2(v):
    return v > 0 and (v & (v - 1)) == 0


def _tc_shape(rows, k1, k2, n):
    return rows >= 128 and 16 <= k1 <= 128 and k1 % 16 == 0 and k2 <= 4 and 16 <= n <= 128 and n % 16 == 0


def typed_fwd_supported(in_dt, out_dt, rows, k1, k2, n):
    """Mirror of hgin_linear_fwd_t's table for calls that involve bf16 rows (include/hgin.h)."""
    if in_dt == torch.bfloat16 and out_dt == torch.bfloat16:
        return _tc_shape(rows, k1, k2, n)
    if in_dt == torch.float32 and out_dt == torch.bfloat16:          # K <= 8 layer
        return k2 == 0 and k1 <= 8 and 4 <= n <= 128 and _pow2(n)
    if in_dt == torch.bfloat16 and out_dt == torch.float32:          # n = 1 head
        return n == 1 and k2 == 0 and 4 <= k1 <= 128 and k1 % 4 == 0 and _pow2(k1 >> 2)
    return True


def typed_bwd_supported(g_dt, x_dt, rows, k1, k2, n, c0, c1, has_dx, has_dot, has_post):
    if g_dt == torch.bfloat16 and x_dt == torch.bfloat16:
        w = c1 - c0
        return _tc_shape(rows, k1, k2, n) and (w == 0 or (16 <= w <= 128 and w % 16 == 0))
    if g_dt == torch.bfloat16 and x_dt == torch.float32:             # K <= 8 layer
        return (k2 == 0 and k1 <= 8 and 4 <= n <= 128 and _pow2(n) and not has_dx and not has_post
                and (c1 == c0 or has_dot) and (c1 - c0 if has_dot else 0) <= 4)
    if g_dt == torch.float32 and x_dt == torch.bfloat16:             # n = 1 head
        return (n == 1 and k2 == 0 and 4 <= k1 <= 128 and k1 % 4 == 0 and _pow2(k1 >> 2) and not has_dot
                and (not has_dx or (c0 == 0 and c1 == k1)) and (c1 == c0 or has_dx))
    return True


def _f32(t):
    return None if t is None else (t if t.dtype == torch.float32 else t.float())


TILE = 128      # the tcgen05 kernels take 16 <= K, N <= 128 (W resident in shared memory); wider layers are tiled here


def _tiled_shape(rows, k1, k2, n, math_mode):
    """Layers wider than one tensor-core tile (hidden 256): run as 128-blocks of the tcgen05 kernels instead of falling to
    the fp32 SIMT engine.  Block partial sums over K are added by a torch op (glue), the activation is a row pass."""
    return (math_mode != MATH_FP32 and rows >= 128 and (k1 > TILE or n > TILE) and k1 % 16 == 0 and n % 16 == 0
            and k1 <= 1024 and n <= 1024 and k2 <= 4)


def _blocks(width):
    return [slice(b, min(b + TILE, width)) for b in range(0, width, TILE)]


def _linear_fwd_tiled(x1, W, bias, x2, act, alpha, want_z, out, accumulate_out, math_mode, want_out, out_dt):
    rows, k1 = x1.shape
    n = W.shape[0]
    z = torch.empty(rows, n, dtype=out_dt, device=x1.device)
    for nb in _blocks(n):
        for i, kb in enumerate(_blocks(k1)):
            first = i == 0
            Wb = W[nb, kb] if (not first or x2 is None) else torch.cat((W[nb, kb], W[nb, k1:]), 1)     # the x2 columns ride on block 0
            _, part = linear_fwd(x1[:, kb], Wb.contiguous(), bias[nb].contiguous() if (first and bias is not None) else None,
                                 x2=x2 if first else None, act=ACT_NONE, want_z=False, out=z[:, nb] if first else None,
                                 math_mode=math_mode, out_dtype=out_dt)
            if not first:
                z[:, nb].add_(part)
    if act == ACT_NONE:
        res = z
    else:
        res = act_fwd(z, act, alpha) if want_out else None
    if want_out and out is not None:
        if accumulate_out:
            out.add_(res)
        else:
            out.copy_(res)
        res = out
    return (z if (want_z or not want_out) and act != ACT_NONE else (z if want_z else None)), res


def _linear_bwd_tiled(g, z, x1, W, x2, act, alpha, want_dx, dot_x, want_dw, want_db, want_dalpha, math_mode):
    rows, k1 = x1.shape
    k2 = 0 if x2 is None else x2.shape[1]
    n = W.shape[0]
    dalpha = None
    if act != ACT_NONE:
        dz, dalpha = act_bwd(g, z, act, alpha, want_dalpha=want_dalpha and act == ACT_PRELU)
        if want_dalpha and dalpha is None:
            dalpha = torch.zeros(1, dtype=torch.float32, device=g.device)
    else:
        dz = g
    db = column_sums(dz) if want_db else None
    dW = torch.empty(n, k1 + k2, dtype=torch.float32, device=g.device) if want_dw else None
    dx = torch.empty(rows, k1, dtype=x1.dtype, device=g.device) if want_dx else None
    for i, kb in enumerate(_blocks(k1)):
        width = kb.stop - kb.start
        for j, nb in enumerate(_blocks(n)):
            tail = x2 if (i == 0 and want_dw) else None
            Wb = W[nb, kb] if tail is None else torch.cat((W[nb, kb], W[nb, k1:]), 1)
            r = linear_bwd(dz[:, nb], None, x1[:, kb], Wb.contiguous(), x2=tail, act=ACT_NONE, dx_cols=(0, width), want_dx=want_dx,
                           want_dw=want_dw, want_db=False, math_mode=math_mode)
            if want_dw:
                dW[nb, kb] = r["dW"][:, :width]
                if tail is not None:
                    dW[nb, k1:] = r["dW"][:, width:]
            if want_dx:
                if j == 0:
                    dx[:, kb] = r["dx"]
                else:
                    dx[:, kb] += r["dx"]
    ddot = None
    if dot_x is not None:
        ddot = (dx.float() * dot_x.float()).sum().reshape(1)
    return {"dx": dx, "ddot": ddot, "dW": dW, "db": db, "dalpha": dalpha}


def linear_fwd(x1, W, bias=None, x2=None, act=ACT_NONE, alpha=None, want_z=True, out=None, accumulate_out=False,
               math_mode=MATH_FP32, want_out=True, out_dtype=None):
    """K2.  z = [x1|x2] W^T + b;  out (+)= act(z).  Returns (z or None, out or None).
    want_out=False: only z is written (the consumers apply act on load, see gin_combine src_act).
    x1 may be float32 or bfloat16 rows; `out_dtype` (default: that of `out`, else of x1) is the storage type of z / out
    (hgin_linear_fwd_t).  A bf16 combination no kernel takes runs in fp32 through casts (odd shapes only)."""
    in_dt = x1.dtype
    out_dt = out_dtype if out_dtype is not None else (out.dtype if out is not None else in_dt)
    rows, k1 = x1.shape
    k2 = 0 if x2 is None else x2.shape[1]
    n = W.shape[0]
    if _tiled_shape(rows, k1, k2, n, math_mode) and (in_dt == out_dt or in_dt == torch.float32):
        return _linear_fwd_tiled(x1, W, bias, x2, act, alpha, want_z, out, accumulate_out, math_mode, want_out, out_dt)
    if (in_dt != torch.float32 or out_dt != torch.float32) and not typed_fwd_supported(in_dt, out_dt, rows, k1, k2, n):
        z32, o32 = linear_fwd(_f32(x1), W, bias, x2=x2, act=act, alpha=alpha, want_z=want_z, out=_f32(out),
                              accumulate_out=accumulate_out, math_mode=math_mode, want_out=want_out)
        if out is not None and o32 is not None and out.dtype != torch.float32:
            out.copy_(o32)
            o32 = out
        cast = (lambda t: None if t is None else (t if t.dtype == out_dt else t.to(out_dt)))
        return cast(z32), cast(o32)
    p1, ld1 = _matrix(x1, "linear_fwd.x1")
    p2, ld2 = _matrix(x2, "linear_fwd.x2", torch.float32)
    if not (W.is_cuda and W.dtype == torch.float32 and W.is_contiguous() and W.shape[1] == k1 + k2):
        raise HginError(f"linear_fwd: W must be contiguous CUDA float32 [n,{k1 + k2}], got {tuple(W.shape)}")
    if x2 is not None and x2.shape[0] != rows:
        raise HginError("linear_fwd: x1 and x2 row counts differ")
    dev = x1.device
    z = torch.empty(rows, n, dtype=out_dt, device=dev) if want_z else None
    if not want_out:
        if out is not None or accumulate_out or not want_z:
            raise HginError("linear_fwd: want_out=False needs want_z and no `out`")
    elif out is None:
        if accumulate_out:
            raise HginError("linear_fwd: accumulate_out needs an existing `out`")
        out = torch.empty(rows, n, dtype=out_dt, device=dev)
    pz, ldz = _matrix(z, "linear_fwd.z", out_dt)
    po, ldo = _matrix(out, "linear_fwd.out", out_dt)
    if out is not None and tuple(out.shape) != (rows, n):
        raise HginError(f"linear_fwd: out is {tuple(out.shape)}, expected {(rows, n)}")
    lib = _lib.load()
    ws_bytes = lib.hgin_linear_fwd_workspace_bytes(rows, k1 + k2, n, math_mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes > 0 else None
    ein, eout = x1.element_size(), (2 if out_dt == torch.bfloat16 else 4)
    with _region("linear_fwd", kernels=1 + (1 if ws_bytes else 0), flops=2 * rows * (k1 + k2) * n,
                 bytes=rows * (ein * k1 + 4 * k2 + eout * n * ((1 if want_z else 0) + (2 if accumulate_out else (1 if want_out else 0))))):
        check(lib.hgin_linear_fwd_t(_DTYPES[in_dt], _DTYPES[out_dt], rows, p1, ld1, k1, p2, ld2, k2, W.data_ptr(), _ptr(bias),
                                    n, act, _scalar(alpha, "linear_fwd.alpha"), pz, ldz, po, ldo, 1 if accumulate_out else 0,
                                    _ptr(ws), ws_bytes, math_mode, _stream()), "hgin_linear_fwd_t")
    return z, out


def linear_bwd(g, z, x1, W, x2=None, act=ACT_NONE, alpha=None, dx_cols=None, want_dx=True, dot_x=None,
               want_dw=True, want_db=True, want_dalpha=False, math_mode=MATH_FP32, post=None, self_eps=None,
               want_self_ddot=False):
    """K3.  Returns dict(dx, ddot, dW, db, dalpha) with None for what was not requested.
    dx_cols=(c0,c1) restricts the input gradient to those columns of [x1|x2].
    act=ACT_NONE: g is dz itself (z may be None).  post: a PostAct for the layer that produced the
    dx columns — dx leaves as that layer's dz and post.dalpha is filled (no dot_x then).
    self_eps (with post, all of x1's columns, no x2; see `post_self_eligible`): the GIN self branch rides on the same
    epilogue — dx = (1 + eps) * (dz W) * act'(post.z), and with want_self_ddot the result carries
    ddot = sum (dz W) * act(post.z) = d(eps) (hgin_linear_bwd_post_self).
    Rows may be float32 or bfloat16: g and z share one type, x1 / dx / dot_x / post.z the other (hgin_linear_bwd_t);
    a bf16 combination no kernel takes runs in fp32 through casts (odd shapes only)."""
    g_dt, x_dt = g.dtype, x1.dtype
    rows, k1 = x1.shape
    k2 = 0 if x2 is None else x2.shape[1]
    k = k1 + k2
    n = W.shape[0]
    if tuple(g.shape) != (rows, n):
        raise HginError(f"linear_bwd: g is {tuple(g.shape)}, expected {(rows, n)}")
    c0, c1 = (0, k) if dx_cols is None else dx_cols
    if not (want_dx or dot_x is not None):
        c1 = c0
    post_on = post is not None and post.act != ACT_NONE and want_dx and c1 > c0
    if (_tiled_shape(rows, k1, k2, n, math_mode) and not post_on and self_eps is None and g_dt == x_dt
            and (c1 == c0 or (c0, c1) == (0, k1))):
        return _linear_bwd_tiled(g, z, x1, W, x2, act, alpha, want_dx and c1 > c0, dot_x, want_dw, want_db, want_dalpha, math_mode)
    if (g_dt != torch.float32 or x_dt != torch.float32) and not typed_bwd_supported(
            g_dt, x_dt, rows, k1, k2, n, c0, c1, want_dx and c1 > c0, dot_x is not None, post_on):
        if self_eps is not None:
            raise HginError("linear_bwd: the self-branch epilogue exists on the tensor-core kernels only")
        post32 = None
        if post_on:
            post32 = PostAct(_f32(post.z), post.act, post.alpha)
        r = linear_bwd(_f32(g), _f32(z), _f32(x1), W, x2=x2, act=act, alpha=alpha, dx_cols=dx_cols, want_dx=want_dx,
                       dot_x=_f32(dot_x), want_dw=want_dw, want_db=want_db, want_dalpha=want_dalpha, math_mode=math_mode,
                       post=post32)
        if post_on:
            post.dalpha, post.applied = post32.dalpha, post32.applied
        if r["dx"] is not None and x_dt != torch.float32:
            r["dx"] = r["dx"].to(x_dt)
        return r
    pg, ldg = _matrix(g, "linear_bwd.g")
    pz, ldz = _matrix(z, "linear_bwd.z", g_dt)
    p1, ld1 = _matrix(x1, "linear_bwd.x1")
    p2, ld2 = _matrix(x2, "linear_bwd.x2", torch.float32)
    pd, ldd = _matrix(dot_x, "linear_bwd.dot_x", x_dt)
    dev = g.device
    dx = torch.empty(rows, c1 - c0, dtype=x_dt, device=dev) if (want_dx and c1 > c0) else None
    pdx, lddx = _matrix(dx, "linear_bwd.dx")
    if dot_x is not None and tuple(dot_x.shape) != (rows, c1 - c0):
        raise HginError(f"linear_bwd: dot_x is {tuple(dot_x.shape)}, expected {(rows, c1 - c0)}")
    ddot = torch.empty(1, dtype=torch.float32, device=dev) if dot_x is not None else None
    dW = torch.empty(n, k, dtype=torch.float32, device=dev) if want_dw else None
    db = torch.empty(n, dtype=torch.float32, device=dev) if want_db else None
    dalpha = torch.empty(1, dtype=torch.float32, device=dev) if want_dalpha else None
    lib = _lib.load()
    ws_bytes = lib.hgin_linear_bwd_workspace_bytes(rows, k, n, math_mode if g_dt == x_dt == torch.float32 else MATH_BF16)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    flops = 2 * rows * n * ((c1 - c0) + (k + 1 if (want_dw or want_db or want_dalpha) else 0))
    reduces = want_dw or want_db or want_dalpha
    n_kernels = ((1 + (1 if ddot is not None else 0)) if c1 > c0 else 0) + \
        ((2 + (1 if (want_dalpha and act == ACT_PRELU) else 0)) if reduces else 0)
    # bytes the passes of this call move (dz pass only with an activation; dx GEMM: dz in, dx out, plus the
    # epilogue operand; dW GEMM: dz and x in; a sums-only pass over dz when the rank-k2 tail of dW is wanted)
    width = c1 - c0
    eg, ex = g.element_size(), x1.element_size()
    has_e = dot_x is not None or post_on
    moved = (3 * n * eg if act != ACT_NONE else 0) \
        + ((n * eg + width * ex * (2 if has_e else 1)) if (width > 0 and dx is not None) else 0) \
        + ((n * eg + k1 * ex + 4 * k2) if want_dw else 0) + (n * eg if (act == ACT_NONE and want_dw and k2 > 0) else 0)
    ppz = ldpz = 0
    sddot = None
    if post_on:
        if dot_x is not None:
            raise HginError("linear_bwd: post-activation and dot_x cannot be combined")
        ppz, ldpz = _matrix(post.z, "linear_bwd.post.z", x_dt)
        if tuple(post.z.shape) != (rows, c1 - c0):
            raise HginError(f"linear_bwd: post.z is {tuple(post.z.shape)}, expected {(rows, c1 - c0)}")
        post.dalpha = torch.empty(1, dtype=torch.float32, device=dev) if post.act == ACT_PRELU else None
        if self_eps is not None:
            if x2 is not None or (c0, c1) != (0, k1):
                raise HginError("linear_bwd: the self-branch epilogue needs dx over all columns of x1 and no x2")
            sddot = torch.empty(1, dtype=torch.float32, device=dev) if want_self_ddot else None
    elif self_eps is not None:
        raise HginError("linear_bwd: self_eps needs a post-activation")
    with _region("linear_bwd", kernels=n_kernels + (1 if post_on else 0) + (1 if sddot is not None else 0), flops=flops,
                 bytes=rows * moved):
        check(lib.hgin_linear_bwd_t(_DTYPES[g_dt], _DTYPES[x_dt], rows, pg, ldg, pz, ldz, act, _scalar(alpha, "linear_bwd.alpha"),
                                    p1, ld1, k1, p2, ld2, k2, W.data_ptr(), n, c0, c1, pdx, lddx, pd, ldd, _ptr(ddot), _ptr(dW),
                                    _ptr(db), _ptr(dalpha), ppz, ldpz, post.act if post_on else ACT_NONE,
                                    _scalar(post.alpha, "linear_bwd.post.alpha") if post_on else 0,
                                    _ptr(post.dalpha) if post_on else 0, _scalar(self_eps, "linear_bwd.self_eps"),
                                    _ptr(sddot), ws.data_ptr(), ws_bytes, math_mode, _stream()), "hgin_linear_bwd_t")
    if post_on:
        post.applied = True
    return {"dx": dx, "ddot": sddot if self_eps is not None else ddot, "dW": dW, "db": db, "dalpha": dalpha}


def post_self_eligible(rows, k, n, math_mode):
    """Shapes whose input gradient runs on the tensor-core kernel that carries the self-branch epilogue
    (csrc/linear_tc.cu bwd_eligible; hgin_linear_bwd_post_self returns HGIN_ERR_UNSUPPORTED otherwise)."""
    return (math_mode != MATH_FP32 and rows >= 128 and 16 <= k <= 128 and k % 16 == 0
            and 16 <= n <= 128 and n % 16 == 0)


# ---- non-default branches of HetroGIN: generic activations, dropout, BatchNorm1d, global pools ---------------------

def _same_shape(a, b, what):
    if tuple(a.shape) != tuple(b.shape) or a.dtype != b.dtype:
        raise HginError(f"{what}: shapes / dtypes differ: {tuple(a.shape)} {a.dtype} vs {tuple(b.shape)} {b.dtype}")


def act_fwd(z, act, alpha=None, p0=0.0, p1=0.0):
    """out = act(z) for any HGIN_ACT_* (models.py:301, 330: `eval(act)` modules the linear kernels do not fuse)."""
    pz, ldz = _matrix(z, "act_fwd.z")
    rows, n = z.shape
    out = torch.empty(rows, n, dtype=z.dtype, device=z.device)
    po, ldo = _matrix(out, "act_fwd.out")
    with _region("act", kernels=1, bytes=2 * rows * n * z.element_size()):
        check(_lib.load().hgin_act_fwd(_DTYPES[z.dtype], rows, n, pz, ldz, act, _scalar(alpha, "act_fwd.alpha"), float(p0),
                                       float(p1), po, ldo, _stream()), "hgin_act_fwd")
    return out


def act_bwd(g, z, act, alpha=None, p0=0.0, p1=0.0, want_dalpha=False):
    """(dz, dalpha or None):  dz = g * act'(z),  dalpha = sum g * min(z, 0)."""
    _same_shape(g, z, "act_bwd")
    pg, ldg = _matrix(g, "act_bwd.g")
    pz, ldz = _matrix(z, "act_bwd.z")
    rows, n = z.shape
    dz = torch.empty(rows, n, dtype=z.dtype, device=z.device)
    pd, ldd = _matrix(dz, "act_bwd.dz")
    lib = _lib.load()
    dalpha = torch.empty(1, dtype=torch.float32, device=z.device) if want_dalpha else None
    ws_bytes = lib.hgin_elementwise_workspace_bytes() if want_dalpha else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device) if ws_bytes else None
    with _region("act", kernels=1 + int(want_dalpha), bytes=3 * rows * n * z.element_size()):
        check(lib.hgin_act_bwd(_DTYPES[z.dtype], rows, n, pg, ldg, pz, ldz, act, _scalar(alpha, "act_bwd.alpha"), float(p0),
                               float(p1), pd, ldd, _ptr(dalpha), _ptr(ws), ws_bytes, _stream()), "hgin_act_bwd")
    return dz, dalpha


def dropout(x, p, seed, offset=0):
    """x * keep / (1 - p) with the Philox mask of (seed, offset); the backward pass is the same call on the gradient."""
    px, ldx = _matrix(x, "dropout.x")
    rows, n = x.shape
    out = torch.empty(rows, n, dtype=x.dtype, device=x.device)
    po, ldo = _matrix(out, "dropout.out")
    with _region("dropout", kernels=1, bytes=2 * rows * n * x.element_size()):
        check(_lib.load().hgin_dropout(_DTYPES[x.dtype], rows, n, px, ldx, float(p), int(seed) & (2 ** 64 - 1),
                                       int(offset) & (2 ** 64 - 1), po, ldo, _stream()), "hgin_dropout")
    return out


def _vec(t, n, name):
    if t is None:
        return 0
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
        raise HginError(f"{name}: expected a contiguous CUDA float32 vector of {n} elements")
    return t.data_ptr()


def bn_stats(z):
    """float64 [2n + 1]: column sums, column sums of squares, row count (all-reduced by the caller under DP)."""
    pz, ldz = _matrix(z, "bn_stats.z")
    rows, n = z.shape
    lib = _lib.load()
    sums = torch.empty(2 * n + 1, dtype=torch.float64, device=z.device)
    ws_bytes = lib.hgin_bn_workspace_bytes(rows, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    with _region("batchnorm", kernels=3, bytes=rows * n * z.element_size()):
        check(lib.hgin_bn_stats(_DTYPES[z.dtype], rows, n, pz, ldz, sums.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
              "hgin_bn_stats")
    return sums


def bn_finalize(n, sums, eps, momentum, running_mean=None, running_var=None, use_running=False):
    """(mean, invstd) fp32 [n]; training mode also updates the running buffers in place."""
    dev = (sums if sums is not None else running_mean).device
    mean = torch.empty(n, dtype=torch.float32, device=dev)
    invstd = torch.empty(n, dtype=torch.float32, device=dev)
    with _region("batchnorm", kernels=1):
        check(_lib.load().hgin_bn_finalize(n, _ptr(sums), float(eps), float(momentum), 1 if use_running else 0, mean.data_ptr(),
                                           invstd.data_ptr(), _vec(running_mean, n, "bn.running_mean"),
                                           _vec(running_var, n, "bn.running_var"), _stream()), "hgin_bn_finalize")
    return mean, invstd


def bn_act_fwd(z, mean, invstd, gamma, beta, act, alpha=None, p0=0.0, p1=0.0):
    pz, ldz = _matrix(z, "bn_act_fwd.z")
    rows, n = z.shape
    out = torch.empty(rows, n, dtype=z.dtype, device=z.device)
    po, ldo = _matrix(out, "bn_act_fwd.out")
    with _region("batchnorm", kernels=1, bytes=2 * rows * n * z.element_size()):
        check(_lib.load().hgin_bn_act_fwd(_DTYPES[z.dtype], rows, n, pz, ldz, _vec(mean, n, "bn.mean"), _vec(invstd, n, "bn.invstd"),
                                          _vec(gamma, n, "bn.weight"), _vec(beta, n, "bn.bias"), act,
                                          _scalar(alpha, "bn_act_fwd.alpha"), float(p0), float(p1), po, ldo, _stream()),
              "hgin_bn_act_fwd")
    return out


def bn_act_bwd_reduce(g, z, mean, invstd, gamma, beta, act, alpha=None, p0=0.0, p1=0.0):
    """float64 [2n + 1]: dbeta, dgamma, dalpha (all-reduced by the caller under DP)."""
    _same_shape(g, z, "bn_act_bwd_reduce")
    pg, ldg = _matrix(g, "bn_act_bwd.g")
    pz, ldz = _matrix(z, "bn_act_bwd.z")
    rows, n = z.shape
    lib = _lib.load()
    sums = torch.empty(2 * n + 1, dtype=torch.float64, device=z.device)
    ws_bytes = lib.hgin_bn_workspace_bytes(rows, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
    with _region("batchnorm", kernels=3, bytes=2 * rows * n * z.element_size()):
        check(lib.hgin_bn_act_bwd_reduce(_DTYPES[z.dtype], rows, n, pg, ldg, pz, ldz, _vec(mean, n, "bn.mean"),
                                         _vec(invstd, n, "bn.invstd"), _vec(gamma, n, "bn.weight"), _vec(beta, n, "bn.bias"), act,
                                         _scalar(alpha, "bn_act_bwd.alpha"), float(p0), float(p1), sums.data_ptr(), ws.data_ptr(),
                                         ws_bytes, _stream()), "hgin_bn_act_bwd_reduce")
    return sums


def bn_act_bwd_apply(g, z, mean, invstd, gamma, beta, act, sums, count, training=True, alpha=None, p0=0.0, p1=0.0,
                     want_dgamma=True, want_dbeta=True, want_dalpha=False):
    """(dz, dgamma, dbeta, dalpha) from the (global) reductions of bn_act_bwd_reduce."""
    pg, ldg = _matrix(g, "bn_act_bwd.g")
    pz, ldz = _matrix(z, "bn_act_bwd.z")
    rows, n = z.shape
    dev = z.device
    dz = torch.empty(rows, n, dtype=z.dtype, device=dev)
    pd, ldd = _matrix(dz, "bn_act_bwd.dz")
    dgamma = torch.empty(n, dtype=torch.float32, device=dev) if want_dgamma else None
    dbeta = torch.empty(n, dtype=torch.float32, device=dev) if want_dbeta else None
    dalpha = torch.empty(1, dtype=torch.float32, device=dev) if want_dalpha else None
    with _region("batchnorm", kernels=1, bytes=3 * rows * n * z.element_size()):
        check(_lib.load().hgin_bn_act_bwd_apply(_DTYPES[z.dtype], rows, n, pg, ldg, pz, ldz, _vec(mean, n, "bn.mean"),
                                                _vec(invstd, n, "bn.invstd"), _vec(gamma, n, "bn.weight"),
                                                _vec(beta, n, "bn.bias"), act, _scalar(alpha, "bn_act_bwd.alpha"), float(p0),
                                                float(p1), sums.data_ptr(), float(count), 1 if training else 0, pd, ldd,
                                                _ptr(dgamma), _ptr(dbeta), _ptr(dalpha), _stream()), "hgin_bn_act_bwd_apply")
    return dz, dgamma, dbeta, dalpha


def global_pool_tail(x, segment_ids, num_segments, origin_cols):
    """models.py:347-352 + the constant columns of the readout input (models.py:364-369):
    [ x[:, :origin_cols] | mean_pool(x, segment_ids)[segment_ids] | max_pool(x, segment_ids)[segment_ids] ] as one fp32
    matrix.  x: raw path features (no gradient), segment_ids: int64/int32 graph id per row."""
    px, ldx = _matrix(x, "global_pool_tail.x", torch.float32)
    rows, f = x.shape
    if not (segment_ids.is_cuda and segment_ids.dim() == 1 and segment_ids.numel() == rows
            and segment_ids.dtype in (torch.int64, torch.int32)):
        raise HginError("global_pool_tail: path_batch must be a CUDA int64/int32 vector with one graph id per path")
    segment_ids = segment_ids.contiguous()
    ids = torch.stack((torch.arange(rows, dtype=segment_ids.dtype, device=x.device), segment_ids))
    csr = csr_build(ids, rows, num_segments, by="dst")
    dev = x.device
    mean = torch.empty(num_segments, f, dtype=torch.float32, device=dev)
    mx = torch.empty(num_segments, f, dtype=torch.float32, device=dev)
    tail = torch.empty(rows, origin_cols + 2 * f, dtype=torch.float32, device=dev)
    lib = _lib.load()
    with _region("global_pool", kernels=2, bytes=rows * 4 * (f + origin_cols + 2 * f)):
        check(lib.hgin_segment_pool(num_segments, csr.rowptr.data_ptr(), _ptr(csr.col), px, ldx, f, mean.data_ptr(),
                                    mx.data_ptr(), _stream()), "hgin_segment_pool")
        check(lib.hgin_readout_tail(rows, segment_ids.data_ptr(), segment_ids.element_size(), num_segments, px, ldx,
                                    origin_cols, mean.data_ptr(), mx.data_ptr(), f, tail.data_ptr(), tail.stride(0), _stream()),
              "hgin_readout_tail")
    return tail, mean, mx, csr


# ---- graph attention (HetroGAT) -------------------------------------------------------------------------------------

def _f32c(t, shape, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
        raise HginError(f"{name}: expected a contiguous CUDA float32 tensor of shape {tuple(shape)}, got {tuple(t.shape)} {t.dtype}")
    return t.data_ptr()


def gat_fwd(csr, xs, a_src, a_dst, bias, heads, channels, negative_slope=0.2, add_self_loops=True, out=None, accumulate=False):
    """PyG GATConv from the attention logits on (include/hgin.h: hgin_gat_fwd).  Returns (out, row_max, row_sum)."""
    hc = heads * channels
    pxs, ldxs = _matrix(xs, "gat_fwd.xs", torch.float32)
    if tuple(xs.shape) != (csr.num_cols, hc):
        raise HginError(f"gat_fwd: xs is {tuple(xs.shape)}, expected {(csr.num_cols, hc)}")
    dev = xs.device
    if out is None:
        if accumulate:
            raise HginError("gat_fwd: accumulate needs an existing `out`")
        out = torch.empty(csr.num_rows, hc, dtype=torch.float32, device=dev)
    po, ldo = _matrix(out, "gat_fwd.out", torch.float32)
    row_max = torch.empty(csr.num_rows, heads, dtype=torch.float32, device=dev)
    row_sum = torch.empty(csr.num_rows, heads, dtype=torch.float32, device=dev)
    with _region("gat", kernels=1, bytes=4 * (csr.num_edges * (hc + 2 * heads) + csr.num_rows * (hc + 3 * heads))):
        check(_lib.load().hgin_gat_fwd(csr.num_rows, _ptr(csr.rowptr), _ptr(csr.col), csr.num_cols, pxs, ldxs,
                                       _f32c(a_src, (csr.num_cols, heads), "gat_fwd.a_src"),
                                       _f32c(a_dst, (csr.num_rows, heads), "gat_fwd.a_dst"), _vec(bias, hc, "gat_fwd.bias"),
                                       heads, channels, float(negative_slope), 1 if add_self_loops else 0,
                                       1 if accumulate else 0, po, ldo, row_max.data_ptr(), row_sum.data_ptr(), _stream()),
              "hgin_gat_fwd")
    return out, row_max, row_sum


def gat_bwd(csr_dst, csr_src, xs, a_src, a_dst, row_max, row_sum, g, heads, channels, negative_slope=0.2,
            add_self_loops=True):
    """Returns (d_xs, d_a_src, d_a_dst) for g = d loss / d out (hgin_gat_bwd)."""
    hc = heads * channels
    pxs, ldxs = _matrix(xs, "gat_bwd.xs", torch.float32)
    pg, ldg = _matrix(g, "gat_bwd.g", torch.float32)
    n_dst, n_src = csr_dst.num_rows, csr_dst.num_cols
    if tuple(g.shape) != (n_dst, hc) or tuple(xs.shape) != (n_src, hc) or csr_src.num_rows != n_src:
        raise HginError("gat_bwd: shapes of g / xs / the transposed CSR do not match the relation")
    dev = xs.device
    d_xs = torch.empty(n_src, hc, dtype=torch.float32, device=dev)
    d_a_src = torch.empty(n_src, heads, dtype=torch.float32, device=dev)
    d_a_dst = torch.empty(n_dst, heads, dtype=torch.float32, device=dev)
    dot_ws = torch.empty(n_dst, heads, 4, dtype=torch.float32, device=dev)      # (a_dst, max, 1/sum, D) records
    with _region("gat", kernels=2, bytes=8 * csr_dst.num_edges * (hc + 4 * heads)):
        check(_lib.load().hgin_gat_bwd(n_dst, _ptr(csr_dst.rowptr), _ptr(csr_dst.col), n_src, _ptr(csr_src.rowptr),
                                       _ptr(csr_src.col), pxs, ldxs, _f32c(a_src, (n_src, heads), "gat_bwd.a_src"),
                                       _f32c(a_dst, (n_dst, heads), "gat_bwd.a_dst"),
                                       _f32c(row_max, (n_dst, heads), "gat_bwd.row_max"),
                                       _f32c(row_sum, (n_dst, heads), "gat_bwd.row_sum"), pg, ldg, heads, channels,
                                       float(negative_slope), 1 if add_self_loops else 0, d_xs.data_ptr(), d_xs.stride(0),
                                       d_a_src.data_ptr(), d_a_dst.data_ptr(), dot_ws.data_ptr(), _stream()), "hgin_gat_bwd")
    return d_xs, d_a_src, d_a_dst


def column_sums(x):
    """fp32 [n] column sums of a row matrix (two-stage fp64 reduction of hgin_bn_stats)."""
    n = x.shape[1]
    return bn_stats(x)[:n].float()


def small_step(csr, x_path, path_cols, x_link, link_cols, y, params, grads, concat_path, want_out=False, phase=0, sums=None,
               workspace=None):
    """One forward + loss + backward of config.json's model family in three kernels (include/hgin.h: hgin_small_step).
    params / grads: dicts with keys W0 b0 a0 eps0 W1 b1 aR W2 b2 W3 b3 (gradients: same shapes, written in place).
    Returns (loss_out [mape, sqrt(mape)], sums [S, N], out or None, workspace).
    phase 1 / 2 (data parallelism, hgin_small_step_phase): 1 = forward only, `sums` = this rank's (S, N), to be all-reduced;
    2 = backward with the global `sums` AND the forward call's `workspace` (it holds the per-row pre-activations) passed
    back in."""
    import ctypes
    for t, name in ((x_path, "x_path"), (x_link, "x_link")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
            raise HginError(f"small_step: {name} must be a CUDA float32 matrix with unit inner stride")
    y = y.reshape(-1)
    if not (y.is_cuda and y.dtype == torch.float32 and y.is_contiguous() and y.numel() == x_path.shape[0]):
        raise HginError("small_step: y must be a contiguous CUDA float32 vector with one entry per path")
    if csr.num_rows != x_path.shape[0] or csr.num_cols != x_link.shape[0]:
        raise HginError("small_step: the CSR does not match the feature matrices")
    keys = ("W0", "b0", "a0", "eps0", "W1", "b1", "aR", "W2", "b2", "W3", "b3")
    for k in keys:
        for d, what in ((params, "parameter"), (grads, "gradient")):
            t = d[k]
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise HginError(f"small_step: {what} {k} must be a contiguous CUDA float32 tensor")
        if params[k].numel() != grads[k].numel():
            raise HginError(f"small_step: gradient of {k} has the wrong size")
    emb, n1, n2 = params["W0"].shape[0], params["W1"].shape[0], params["W2"].shape[0]
    fp, fl = len(path_cols), len(link_cols)
    if params["W0"].shape[1] != fl + fp or params["W1"].shape[1] != emb + (fp if concat_path else 0) \
            or params["W2"].shape[1] != n1 or params["W3"].numel() != n2:
        raise HginError("small_step: parameter shapes do not form the GIN layer + 3-layer readout chain")
    dev = x_path.device
    lib = _lib.load()
    np_ = x_path.shape[0]
    ws_bytes = lib.hgin_small_step_workspace_bytes(np_)
    if phase == 2:
        if sums is None or not (sums.is_cuda and sums.dtype == torch.float32 and sums.numel() == 2):
            raise HginError("small_step: phase 2 needs the all-reduced (S, N) in `sums`")
        if workspace is None or workspace.numel() < ws_bytes:
            raise HginError("small_step: phase 2 needs the workspace of the phase-1 call")
        ws = workspace
    else:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        sums = torch.empty(2, dtype=torch.float32, device=dev)
    loss_out = torch.empty(2, dtype=torch.float32, device=dev)
    out = torch.empty(np_, dtype=torch.float32, device=dev) if want_out else None
    pc = (ctypes.c_int32 * 8)(*(list(path_cols) + [0] * (8 - fp)))
    lc = (ctypes.c_int32 * 8)(*(list(link_cols) + [0] * (8 - fl)))
    args = (np_, _ptr(csr.rowptr), _ptr(csr.col), x_path.data_ptr(), x_path.stride(0), fp, ctypes.cast(pc, ctypes.c_void_p),
            x_link.data_ptr(), x_link.stride(0), fl, ctypes.cast(lc, ctypes.c_void_p), y.data_ptr(), emb, n1, n2,
            1 if concat_path else 0, *[params[k].data_ptr() for k in keys], *[grads[k].data_ptr() for k in keys],
            sums.data_ptr(), loss_out.data_ptr(), _ptr(out), ws.data_ptr(), ws_bytes, _stream())
    with _region("small_step", kernels=3 if phase == 0 else 2, flops=(6 if phase == 0 else (2 if phase == 1 else 6)) * np_ * (
            emb * (fl + fp) + n1 * (emb + fp) + n1 * n2 + n2)):
        if phase == 0:
            check(lib.hgin_small_step(*args), "hgin_small_step")
        else:
            check(lib.hgin_small_step_phase(phase, *args), "hgin_small_step_phase")
    return loss_out, sums, out, ws


def qt_baseline(p_l, avg_bw, capacity, num_paths, num_links, num_iterations=3):
    """Queueing-theory baseline on a (batched) path->link relation.  p_l: CUDA int64/int32 [2,E],
    every path's edges in route order; avg_bw f32 [num_paths]; capacity f32 [num_links] (raw).
    Returns (path_delay f32 [num_paths], link_out f32 [num_links,3] = occupancy, rho, pi_0)."""
    for t, n, name in ((avg_bw, num_paths, "avg_bw"), (capacity, num_links, "capacity")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
            raise HginError(f"qt_baseline: {name} must be a contiguous CUDA float32 vector of {n} elements")
    by_path = csr_build(p_l, num_paths, num_links, by="src", want_perm=True)
    by_link = csr_build(p_l, num_paths, num_links, by="dst", want_perm=True)
    cap_scaled = capacity / 1000                       # models.py:73-74, the same fp32 division
    dev = avg_bw.device
    path_delay = torch.empty(num_paths, dtype=torch.float32, device=dev)
    link_out = torch.empty(num_links, 3, dtype=torch.float32, device=dev)
    lib = _lib.load()
    e = by_path.num_edges
    ws_bytes = lib.hgin_qt_baseline_workspace_bytes(num_links, e)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _region("qt_baseline", kernels=2 + 2 * num_iterations):
        check(lib.hgin_qt_baseline(num_paths, num_links, e, by_path.rowptr.data_ptr(), _ptr(by_path.col),
                                   _ptr(by_path.perm), by_link.rowptr.data_ptr(), _ptr(by_link.perm), avg_bw.data_ptr(),
                                   cap_scaled.data_ptr(), capacity.data_ptr(), num_iterations, path_delay.data_ptr(),
                                   link_out.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "hgin_qt_baseline")
    by_path.validate()
    return path_delay, link_out


def set_option(name, value):
    """Process-wide library option (see include/hgin.h), e.g. set_option("fused_bwd", 1)."""
    check(_lib.load().hgin_set_option(name.encode(), int(value)), "hgin_set_option")


def debug_gemm_tn(a, b, tma_swizzle=-1, lbo=-1, sbo=-1, layout_type=-1, k_step_bytes=-1):
    """Diagnostics: a[rows,n]^T @ b[rows,k] through the tcgen05 MN-major kernel (hgin_debug_gemm_tn)."""
    rows, n = a.shape
    k = b.shape[1]
    lib = _lib.load()
    out = torch.empty(n, k, dtype=torch.float32, device=a.device)
    ws_bytes = lib.hgin_linear_bwd_workspace_bytes(rows, k, n, MATH_TF32)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
    check(lib.hgin_debug_gemm_tn(rows, a.data_ptr(), n, b.data_ptr(), k, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                 tma_swizzle, lbo, sbo, layout_type, k_step_bytes, _stream()), "hgin_debug_gemm_tn")
    return out


def debug_gemm_tn_bf16(a, b, lbo=-1, sbo=-1, layout_type=-1, k_step_bytes=-1):
    """Diagnostics: a[rows,n]^T @ b[rows,k] (bf16 rows) through the bf16 tcgen05 MN-major kernel."""
    rows, n = a.shape
    k = b.shape[1]
    lib = _lib.load()
    out = torch.empty(n, k, dtype=torch.float32, device=a.device)
    ws_bytes = lib.hgin_linear_bwd_workspace_bytes(rows, k, n, MATH_BF16)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
    check(lib.hgin_debug_gemm_tn_bf16(rows, a.data_ptr(), n, b.data_ptr(), k, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                      lbo, sbo, layout_type, k_step_bytes, _stream()), "hgin_debug_gemm_tn_bf16")
    return out


def mape_sum(pred, y):
    """sums = [sum |(pred - y)/y|, n]  (train.py:13 before the mean)."""
    pred = pred.reshape(-1)
    y = y.reshape(-1)
    if not (pred.is_cuda and y.is_cuda and pred.dtype == y.dtype == torch.float32 and pred.numel() == y.numel()):
        raise HginError("mape_sum: pred and y must be CUDA float32 of equal size")
    pred, y = pred.contiguous(), y.contiguous()
    n = pred.numel()
    lib = _lib.load()
    sums = torch.empty(2, dtype=torch.float32, device=pred.device)
    ws_bytes = lib.hgin_reduce_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pred.device)
    with _region("loss", kernels=2):
        check(lib.hgin_mape_sum(n, pred.data_ptr(), y.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws_bytes,
                                _stream()), "hgin_mape_sum")
    return sums


def sqrt_mape_bwd(pred, y, sums, gscale=1.0):
    """Returns (loss_out=[mape, sqrt(mape)], dpred) for the GLOBAL sums (all-reduced by the caller)."""
    shape = pred.shape
    pred = pred.reshape(-1).contiguous()
    y = y.reshape(-1).contiguous()
    n = pred.numel()
    loss_out = torch.empty(2, dtype=torch.float32, device=pred.device)
    dpred = torch.empty(n, dtype=torch.float32, device=pred.device)
    with _region("loss", kernels=1):
        check(_lib.load().hgin_sqrt_mape_bwd(n, pred.data_ptr(), y.data_ptr(), sums.data_ptr(), float(gscale),
                                             loss_out.data_ptr(), dpred.data_ptr(), _stream()),
              "hgin_sqrt_mape_bwd")
    return loss_out, dpred.view(shape)


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
              decoupled=False):
    """In-place Adam/AdamW on flat fp32 buffers; `step` is a CUDA int32[1] holding the 1-based step."""
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
            raise HginError("adam_step: flat contiguous CUDA float32 buffers of equal size expected")
    with _region("adam", kernels=1):
        check(_lib.load().hgin_adam_step(n, param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(),
                                         exp_avg_sq.data_ptr(), step.data_ptr(), lr, beta1, beta2, eps, weight_decay,
                                         1 if decoupled else 0, _stream()), "hgin_adam_step")


def increment(counter):
    with _region("adam", kernels=1):
        check(_lib.load().hgin_increment(counter.data_ptr(), _stream()), "hgin_increment")
