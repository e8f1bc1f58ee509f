"""Autograd glue between the module mirror (models.py) and the CUDA kernels (ops.py).

Two `torch.autograd.Function`s cover the whole hot path:

* `HeteroConvFn`  — one heterogeneous GIN layer over a set of relations (what PyG's
  `HeteroConv({...: GINLayer}, aggr='sum')` does at models.py:356): per relation K1 (neighbour
  sum + self term) -> K2 (Linear + PReLU, accumulated into the destination type's buffer = the
  'sum' merge); backward K3 per relation, then ONE transposed-CSR gather per node type that
  also folds in the `(1+eps)*dh` self branches (K4).
* `LinearActFn`   — one readout layer `act([x1|x2] W^T + b)` (models.py:366-374).

Autograd only sees tensors in / tensors out; saved activations are `h` (MLP input) and `z`
(pre-activation) per relation.  Parameters that feed no live output keep `grad=None`, exactly as
the reference's autograd leaves them.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import ACT_NONE, ACT_PRELU, ACT_RELU, MATH_BF16, MATH_FP32, SELF_ADD, SELF_CONCAT, SELF_NONE, HginError


class GraphCSR:
    """Per-batch adjacency in kernel layout: for every relation a destination-sorted CSR (forward
    aggregation) and, on demand, the source-sorted transpose (backward gather).  Built on the GPU
    by K0 from the COO `edge_index_dict` the reference passes around (train.py:34)."""

    def __init__(self, edge_index_dict, num_nodes, blocks=None):
        self.edge_index_dict = dict(edge_index_dict)
        self.num_nodes = dict(num_nodes)
        self._by_dst, self._by_src = {}, {}
        # block-diagonal batches: node type -> int64 [B + 1] device vector of first rows per sample (PyG's `ptr`)
        self.blocks = dict(blocks) if blocks else {}
        self._plans = {}

    @classmethod
    def from_prebuilt(cls, stores, num_nodes, blocks=None):
        """From CSR tensors already on the device (`data.Batch.from_data_list(csr=True)`): stores[et]
        holds int32 `csr_{dst,src}_{rowptr,col}`; no K0 launch happens for these relations."""
        g = cls({et: st.get("edge_index") for et, st in stores.items()}, num_nodes, blocks)
        for et, st in stores.items():
            n_src, n_dst = num_nodes[et[0]], num_nodes[et[2]]
            e = st["csr_dst_col"].shape[0]
            g._by_dst[et] = ops.CSR(st["csr_dst_rowptr"], st["csr_dst_col"], None, None, n_dst, n_src, e)
            g._by_src[et] = ops.CSR(st["csr_src_rowptr"], st["csr_src_col"], None, None, n_src, n_dst, e)
        return g

    # mapping protocol, so a GraphCSR can be passed wherever an edge_index_dict is expected
    def items(self):
        return self.edge_index_dict.items()

    def keys(self):
        return self.edge_index_dict.keys()

    def __contains__(self, et):
        return et in self.edge_index_dict

    def __getitem__(self, et):
        return self.edge_index_dict[et]

    def fwd(self, et) -> ops.CSR:
        if et not in self._by_dst:
            self._by_dst[et] = ops.csr_build(self.edge_index_dict[et], self.num_nodes[et[0]],
                                             self.num_nodes[et[2]], by="dst")
        return self._by_dst[et]

    def bwd(self, et) -> ops.CSR:
        if et not in self._by_src:
            self._by_src[et] = ops.csr_build(self.edge_index_dict[et], self.num_nodes[et[0]],
                                             self.num_nodes[et[2]], by="src")
        return self._by_src[et]

    def stream_plan(self, et, side):
        """ops.StreamPlan for the aggregation that writes rows of et's destination type (side 'fwd': the forward K1 pass)
        or of its source type (side 'bwd': the transposed K4 gather) — or None when the batch has no block tables.  Short
        rows (~3 neighbours: link->path) take the shared-memory table variant, long rows the (opt-in) streaming kernel;
        ops.gin_combine picks by row length.  The gate is computed once per batch and relation side."""
        key = (tuple(et), side)
        if not (ops.STREAM_LONG_ROWS or ops.STAGE_LONG_ROWS or ops.TABLE_SHORT_ROWS):
            return None                  # both block-diagonal schedules are opt-in (measured: DESIGN.md "Long rows" / "Short rows")
        if key not in self._plans:
            plan = None
            t_out, t_in = (et[2], et[0]) if side == "fwd" else (et[0], et[2])
            p_in, p_out = self.blocks.get(t_in), self.blocks.get(t_out)
            if (p_in is not None and p_out is not None and p_in.is_cuda and p_in.dtype == torch.int64
                    and p_in.numel() == p_out.numel() and p_in.numel() >= 2):
                other = self._by_src if side == "fwd" else self._by_dst
                if et not in other and not torch.is_grad_enabled():
                    return None          # inference: not worth building the second orientation for the gate
                csr_out, csr_in = (self.fwd(et), self.bwd(et)) if side == "fwd" else (self.bwd(et), self.fwd(et))
                long_rows = csr_out.num_edges > 8 * csr_out.num_rows
                if csr_out.num_rows > 0 and ((ops.STREAM_LONG_ROWS or ops.STAGE_LONG_ROWS) if long_rows else ops.TABLE_SHORT_ROWS):
                    plan = ops.StreamPlan(csr_in, p_in.contiguous(), p_out.contiguous(),
                                          ops.block_gate(csr_out, csr_in, p_in.contiguous(), p_out.contiguous()))
            self._plans[key] = plan
        return self._plans[key]

    def validate(self):
        for c in list(self._by_dst.values()) + list(self._by_src.values()):
            if c.status is not None:   # prebuilt CSRs were validated when they were built per sample
                c.validate()
        return self


class RelationSpec:
    """Static description of one relation's GIN layer inside a HeteroConvFn call."""

    __slots__ = ("et", "src", "dst", "concat", "act")

    def __init__(self, et, concat, act):
        self.et, self.src, self.dst = tuple(et), et[0], et[-1]
        self.concat, self.act = bool(concat), act


def _fold_eligible(rows, k, n, math_mode):
    """Shapes for which the backward consumes dz in place (csrc/linear_tc.cu, linear_thin.cu): only
    there does moving act'(z) into the producer of the gradient remove a pass over the rows (and
    lets d(eps) of the layer above come out of the same pass)."""
    if math_mode not in (ops.MATH_TF32, ops.MATH_BF16):
        return False
    tensor_core = rows >= 128 and 16 <= k <= 128 and k % 16 == 0 and 16 <= n <= 128 and n % 16 == 0
    thin = k <= 8 and 4 <= n <= 128 and (n & (n - 1)) == 0      # csrc/linear_thin.cu: reads g only with ACT_NONE
    return tensor_core or thin


class HeteroConvFn(torch.autograd.Function):
    """args: specs, graph, types, math_mode, links_in, links_out, x[types...], then (W, b, alpha, eps)
    per spec.  returns one tensor per destination type, in order of first appearance.

    links_in / links_out (dicts node type -> ops.PostAct, or None) chain consecutive layers INSIDE
    HetroGIN.forward, where every intermediate activation has exactly one consumer: links_out is
    filled with (z, act, alpha) of each output produced by a single relation; the next node receives
    it as links_in and, in its backward, lets the kernel that produces the gradient for that input
    apply act'(z) on the way out (hgin_gin_combine_post), so this node's backward starts from dz."""

    @staticmethod
    def forward(ctx, specs, graph, types, math_mode, links_in, links_out, allow_lazy, *tensors):
        """allow_lazy: the consumer of this node's outputs is another HeteroConvFn of the chain, which can
        read PRE-activations (hgin_gin_combine_pre): an output produced by a single relation is then
        returned as z, its activated copy is never written, and the link says so (`lazy`)."""
        nt = len(types)
        xs = dict(zip(types, tensors[:nt]))
        params = tensors[nt:]
        training = any(ctx.needs_input_grad)
        # HGIN_MATH_BF16: every activation this layer produces is STORED as bf16 (half the bytes per row)
        out_dt = torch.bfloat16 if math_mode == MATH_BF16 else torch.float32
        outs, saved = {}, []
        for i, sp in enumerate(specs):
            W, b, alpha, eps = params[4 * i:4 * i + 4]
            x_src, x_dst = xs[sp.src], xs[sp.dst]
            lz_src, lz_dst = (links_in or {}).get(sp.src), (links_in or {}).get(sp.dst)
            pre = {}
            if lz_src is not None and lz_src.lazy:       # x_src is the pre-activation of the layer below
                pre["src_act"] = (lz_src.act, lz_src.alpha)
            if lz_dst is not None and lz_dst.lazy:
                pre["self_act"] = (lz_dst.act, lz_dst.alpha)
            if x_dst.dtype != x_src.dtype:      # (mixed storage types only arise outside HetroGIN.forward)
                x_dst = x_dst.to(x_src.dtype)
            h = ops.gin_combine(graph.fwd(sp.et), x_src, x_dst, eps, SELF_CONCAT if sp.concat else SELF_ADD,
                                block_plan=None if sp.concat else graph.stream_plan(sp.et, "fwd"), **pre)
            link_ok = (training and links_out is not None and sp.act != ACT_NONE
                       and sum(1 for q in specs if q.dst == sp.dst) == 1
                       and _fold_eligible(h.shape[0], h.shape[1], W.shape[0], math_mode))
            lazy = bool(link_ok and allow_lazy and ctx.needs_input_grad[7 + nt + 4 * i])
            z, o = ops.linear_fwd(h, W, b, act=sp.act, alpha=alpha, want_z=training and sp.act != ACT_NONE,
                                  out=outs.get(sp.dst), accumulate_out=sp.dst in outs, math_mode=math_mode,
                                  want_out=not lazy, out_dtype=out_dt)
            outs[sp.dst] = z if lazy else o
            saved += [h if training else None, z]
            if training and ctx.needs_input_grad[7 + types.index(sp.src)]:
                graph.bwd(sp.et)  # build the transposed CSR alongside the forward work
            if link_ok:
                # (a lazy z is also this node's OUTPUT: the link keeps a detached alias, or output -> grad_fn ->
                # ctx.links_out -> output would be a reference cycle that pins the step's activations until gc runs)
                links_out[sp.dst] = ops.PostAct(z.detach() if lazy else z, sp.act, alpha, lazy=lazy)
        ctx.links_in = dict(links_in) if links_in else {}
        ctx.links_out = dict(links_out) if links_out else {}
        ctx.specs, ctx.graph, ctx.types, ctx.math_mode = specs, graph, types, math_mode
        ctx.out_types = list(outs)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(*tensors, *saved)
        return tuple(outs[t] for t in ctx.out_types)

    @staticmethod
    def backward(ctx, *gouts):
        specs, graph, types = ctx.specs, ctx.graph, ctx.types
        nt, ns = len(types), len(specs)
        tensors = ctx.saved_tensors
        xs = dict(zip(types, tensors[:nt]))
        params = tensors[nt:nt + 4 * ns]
        saved = tensors[nt + 4 * ns:]
        need = ctx.needs_input_grad[7:]
        need_x = dict(zip(types, need[:nt]))
        g_out = dict(zip(ctx.out_types, gouts))
        grads_p = [None] * (4 * ns)

        # Plan of the per-node-type gather passes (K4), fixed before any kernel runs: every live
        # relation contributes a gather (to its source type) and a (1+eps)*dh self branch (to its
        # destination type); one self branch rides on each gather pass, the rest get an edgeless pass.
        live = [i for i, sp in enumerate(specs) if g_out.get(sp.dst) is not None]
        plan = {}
        for t in types:
            gathers = [i for i in live if specs[i].src == t and need_x[t]]
            selfs = [i for i in live if specs[i].dst == t and need_x[t]]
            passes = [(gi, selfs.pop(0) if selfs else None) for gi in gathers] + [(None, si) for si in selfs]
            plan[t] = passes
        # d(eps) of the relation whose self branch rides on the LAST pass of a type comes out of that
        # pass when it applies the post-activation (x_dst = act(z) of the layer below is in registers)
        eps_from_pass = {}
        for t in types:
            post = ctx.links_in.get(t)
            if post is not None and post.usable() and plan[t] and plan[t][-1][1] is not None:
                si = plan[t][-1][1]
                if need[nt + 4 * si + 3] and not specs[si].concat:
                    eps_from_pass[si] = t

        # A relation whose self branch is the ONLY thing type t's gradient consists of (plan == one edgeless
        # pass; e.g. the last GIN layer) hands that pass to the epilogue of its own input-gradient GEMM
        # (hgin_linear_bwd_post_self): dh = dz W never reaches memory.  Its source-side gradient, a gather of dh
        # rows over the transposed CSR, is then taken as  A^T (dz W) = (A^T dz) W : the gather reads dz (which IS in
        # memory) and one source-sized GEMM follows — same bytes for the gather, two row-sized transfers less overall.
        self_in_gemm = {}
        for t in types:
            post = ctx.links_in.get(t)
            if (post is not None and post.usable() and len(plan[t]) == 1 and plan[t][0][0] is None):
                si = plan[t][0][1]
                sp = specs[si]
                h_si, W_si = saved[2 * si], params[4 * si]
                lo = ctx.links_out.get(sp.dst)
                g_is_dz = sp.act == ACT_NONE or (lo is not None and lo.applied)
                # (the gather of this relation must be the only pass of its source type, so that the pre-gathered
                # term needs no pairing with another relation's self branch)
                src_ok = (not need_x[sp.src]) or (g_is_dz and plan[sp.src] == [(si, None)])
                if (not sp.concat and src_ok
                        and ops.post_self_eligible(h_si.shape[0], h_si.shape[1], W_si.shape[0], ctx.math_mode)
                        and tuple(post.z.shape) == tuple(h_si.shape) and post.z.dtype == h_si.dtype
                        and (post.z.stride(0) * post.z.element_size()) % 16 == 0 and post.z.data_ptr() % 16 == 0):
                    self_in_gemm[si] = t

        dh_agg_of, dh_self_of = {}, {}
        for i in live:
            sp = specs[i]
            g = g_out[sp.dst]
            if g.stride(-1) != 1:
                g = g.contiguous()
            W, b, alpha, eps = params[4 * i:4 * i + 4]
            h, z = saved[2 * i], saved[2 * i + 1]
            nW, nb, nalpha, neps = need[nt + 4 * i:nt + 4 * i + 4]
            x_src, x_dst = xs[sp.src], xs[sp.dst]
            fs = x_src.shape[1]
            k = h.shape[1]
            want_agg, want_self = need_x[sp.src], need_x[sp.dst]
            common = dict(act=sp.act, alpha=alpha, math_mode=ctx.math_mode)
            done = ctx.links_out.get(sp.dst)
            done = done if (done is not None and done.applied) else None
            if done is not None:      # g already is dz: the consumer of our output applied act'(z)
                common.update(act=ACT_NONE, alpha=None)
                z, nalpha_here = None, False
            else:
                nalpha_here = bool(nalpha)
                lz = ctx.links_out.get(sp.dst)
                if lz is not None and lz.lazy:
                    raise HginError("a lazily activated output received a gradient that its consumer did not "
                                    "pass through the activation derivative (chain protocol broken)")
            dot_x = x_dst if (neps and i not in eps_from_pass) else None
            lz_dst = ctx.links_in.get(sp.dst)
            if dot_x is not None and lz_dst is not None and lz_dst.lazy:
                # x_dst was handed over as a pre-activation: materialise act(z) for the d(eps) dot product
                dot_x = torch.where(x_dst > 0, x_dst, (lz_dst.alpha * x_dst) if lz_dst.act == ops.ACT_PRELU
                                    else torch.zeros_like(x_dst))
            if i in self_in_gemm:
                post_t = ctx.links_in[self_in_gemm[i]]
                r = ops.linear_bwd(g, z, h, W, dx_cols=(0, k), want_dx=True, want_dw=nW, want_db=bool(nb),
                                   want_dalpha=nalpha_here, post=post_t, self_eps=eps, want_self_ddot=bool(neps),
                                   **common)
                dh_agg, dh_self = None, None
                if want_agg:      # (A^T dz) W, see above; g is dz here
                    gz = ops.gin_combine(graph.bwd(sp.et), g, None, None, SELF_NONE, block_plan=graph.stream_plan(sp.et, "bwd"))
                    dh_agg = ops.linear_bwd(gz, None, x_src, W, act=ACT_NONE, dx_cols=(0, k), want_dx=True, want_dw=False,
                                            want_db=False, math_mode=ctx.math_mode)["dx"]
            elif sp.concat:
                r = ops.linear_bwd(g, z, h, W, dx_cols=(fs, k), want_dx=want_self, dot_x=dot_x,
                                   want_dw=nW, want_db=bool(nb), want_dalpha=nalpha_here, **common)
                dh_self = r["dx"]
                dh_agg = None
                if want_agg:
                    dh_agg = ops.linear_bwd(g, z, h, W, dx_cols=(0, fs), want_dx=True, want_dw=False, want_db=False,
                                            **common)["dx"]
            else:
                r = ops.linear_bwd(g, z, h, W, dx_cols=(0, k), want_dx=want_agg or want_self, dot_x=dot_x,
                                   want_dw=nW, want_db=bool(nb), want_dalpha=nalpha_here, **common)
                dh_agg = dh_self = r["dx"]
            if done is not None and nalpha:
                r["dalpha"] = done.dalpha
            grads_p[4 * i:4 * i + 4] = [r["dW"], r["db"],
                                        None if r["dalpha"] is None else r["dalpha"].view_as(alpha),
                                        None if r["ddot"] is None else r["ddot"].view_as(eps)]
            dh_agg_of[i], dh_self_of[i] = dh_agg, dh_self
            if i in self_in_gemm:
                dh_self_of[i] = r["dx"]      # already (1 + eps) * dh * act'(z below): the finished gradient of type t

        grads_x = []
        done_types = set(self_in_gemm.values())
        for t in types:
            if t in done_types:
                si = plan[t][0][1]
                grads_x.append(dh_self_of[si])
                continue
            post = ctx.links_in.get(t)
            dx = None
            for j, (gi, si) in enumerate(plan[t]):
                last = j == len(plan[t]) - 1
                s_dh = dh_self_of[si] if si is not None else None
                s_eps = params[4 * si + 3] if si is not None else None
                if gi is not None and gi in self_in_gemm:
                    # already gathered ((A^T dz) W) and, by the eligibility rule, the only term of this type
                    dx = dh_agg_of[gi]
                    if post is not None and post.usable():
                        dx = ops.gin_combine(None, dx, dx, None, SELF_ADD, post=post)
                    continue
                csr_t = graph.bwd(specs[gi].et) if gi is not None else None
                src_rows = dh_agg_of[gi] if gi is not None else s_dh
                want_ddot = last and si is not None and eps_from_pass.get(si) == t
                plan_t = graph.stream_plan(specs[gi].et, "bwd") if gi is not None else None
                res = ops.gin_combine(csr_t, src_rows, s_dh, s_eps, SELF_ADD if si is not None else SELF_NONE, out=dx,
                                      accumulate=dx is not None, post=post if last else None, want_ddot=want_ddot,
                                      block_plan=plan_t)
                if want_ddot:
                    dx, ddot = res
                    grads_p[4 * si + 3] = ddot.view_as(s_eps)
                else:
                    dx = res
            grads_x.append(dx)
        return (None, None, None, None, None, None, None, *grads_x, *grads_p)


class LinearActFn(torch.autograd.Function):
    """out = act([x1 | x2] W^T + b)."""

    @staticmethod
    def forward(ctx, x1, x2, W, b, alpha, act, math_mode, link_in=None, link_out=None):
        """link_in: ops.PostAct of the layer that produced x1 (or None); link_out: a list that receives
        this layer's own PostAct for the next layer in the chain (see HeteroConvFn)."""
        training = any(ctx.needs_input_grad)
        # HGIN_MATH_BF16: hidden readout activations are stored as bf16; the n = 1 score column stays fp32
        out_dt = torch.bfloat16 if (math_mode == MATH_BF16 and W.shape[0] > 1) else torch.float32
        z, out = ops.linear_fwd(x1, W, b, x2=x2, act=act, alpha=alpha, want_z=training and act != ACT_NONE,
                                math_mode=math_mode, out_dtype=out_dt)
        ctx.act, ctx.math_mode = act, math_mode
        ctx.link_in, ctx.link_self = link_in, None
        if (training and link_out is not None and z is not None
                and _fold_eligible(x1.shape[0], x1.shape[1], W.shape[0], math_mode)):
            ctx.link_self = ops.PostAct(z, act, alpha)
            link_out.append(ctx.link_self)
        ctx.save_for_backward(x1, x2, W, alpha, z)
        return out

    @staticmethod
    def backward(ctx, g):
        x1, x2, W, alpha, z = ctx.saved_tensors
        n1, n2, nW, nb, nalpha = ctx.needs_input_grad[:5]
        if g.stride(-1) != 1:
            g = g.contiguous()
        k1 = x1.shape[1]
        k = W.shape[1]
        c0, c1 = (0 if n1 else k1), (k if n2 else k1)
        act = ctx.act
        done = ctx.link_self if (ctx.link_self is not None and ctx.link_self.applied) else None
        if done is not None:          # g already is dz
            act, z = ACT_NONE, None
        post = ctx.link_in if (n1 and not n2) else None     # dx must be exactly the gradient of x1
        r = ops.linear_bwd(g, z, x1, W, x2=x2, act=act, alpha=alpha if done is None else None, dx_cols=(c0, c1),
                           want_dx=n1 or n2, want_dw=nW, want_db=bool(nb), want_dalpha=bool(nalpha) and done is None,
                           math_mode=ctx.math_mode, post=post)
        if done is not None and nalpha:
            r["dalpha"] = done.dalpha
        dx = r["dx"]
        dx1 = dx[:, :k1 - c0] if n1 else None
        dx2 = dx[:, k1 - c0:] if n2 else None
        dalpha = None if r["dalpha"] is None else r["dalpha"].view_as(alpha)
        return dx1, dx2, r["dW"], r["db"], dalpha, None, None, None, None


class ActSpec:
    """An activation module resolved to kernel arguments: HGIN_ACT_* code, the learnable PReLU slope (or None) and the
    module's constructor constants p0 / p1.  `fused`: one of the three the linear kernels apply in their epilogue."""

    __slots__ = ("code", "alpha", "p0", "p1")

    def __init__(self, code, alpha=None, p0=0.0, p1=0.0):
        self.code, self.alpha, self.p0, self.p1 = code, alpha, float(p0), float(p1)

    @property
    def fused(self):
        return self.code in (ACT_NONE, ACT_PRELU, ACT_RELU)


def act_spec_of(module) -> ActSpec:
    """Every activation `eval(config.MLP_ACT)` / `eval(mlp_head_act)` (models.py:301, 330) can reasonably name."""
    nn = torch.nn
    if module is None or isinstance(module, nn.Identity):
        return ActSpec(ACT_NONE)
    if isinstance(module, nn.PReLU):
        if module.weight.numel() != 1:
            raise NotImplementedError("only the single-slope torch.nn.PReLU() of the reference is supported")
        return ActSpec(ACT_PRELU, module.weight)
    if isinstance(module, nn.ReLU):
        return ActSpec(ACT_RELU)
    if isinstance(module, nn.LeakyReLU):
        return ActSpec(ops.ACT_LEAKY_RELU, p0=module.negative_slope)
    if isinstance(module, nn.ELU):
        return ActSpec(ops.ACT_ELU, p0=module.alpha)
    if isinstance(module, nn.Sigmoid):
        return ActSpec(ops.ACT_SIGMOID)
    if isinstance(module, nn.Tanh):
        return ActSpec(ops.ACT_TANH)
    if isinstance(module, nn.GELU):
        if getattr(module, "approximate", "none") != "none":
            raise NotImplementedError("torch.nn.GELU(approximate='tanh') has no kernel; the erf form does")
        return ActSpec(ops.ACT_GELU)
    if isinstance(module, nn.SiLU):
        return ActSpec(ops.ACT_SILU)
    if isinstance(module, nn.Softplus):
        return ActSpec(ops.ACT_SOFTPLUS, p0=module.beta, p1=module.threshold)
    raise NotImplementedError(f"activation {type(module).__name__} has no kernel (PReLU, ReLU, LeakyReLU, ELU, Sigmoid, Tanh, "
                              "GELU, SiLU, Softplus and Identity do)")


def activation_of(module):
    """(ACT_* code, slope parameter) for the activations the LINEAR kernels fuse (GIN layers: models.py:236-239)."""
    spec = act_spec_of(module)
    if not spec.fused:
        raise NotImplementedError(f"activation {type(module).__name__} is not fused into the linear kernels")
    return spec.code, spec.alpha


def linear_act_of(seq):
    """Decompose `Sequential(Linear[, act])` (models.py:236-239, 317-330) into kernel arguments."""
    mods = list(seq) if isinstance(seq, torch.nn.Sequential) else [seq]
    if not mods or not isinstance(mods[0], torch.nn.Linear) or len(mods) > 2:
        raise NotImplementedError("fused path expects Sequential(Linear[, activation]); got " + repr(seq))
    act, alpha = activation_of(mods[1] if len(mods) == 2 else None)
    return mods[0].weight, mods[0].bias, act, alpha


def readout_layer_of(seq):
    """`Sequential(Linear[, BatchNorm1d][, act])` (models.py:303-330) -> (Linear, BatchNorm1d or None, ActSpec)."""
    mods = list(seq) if isinstance(seq, torch.nn.Sequential) else [seq]
    if not mods or not isinstance(mods[0], torch.nn.Linear):
        raise NotImplementedError("readout layers are Sequential(Linear[, BatchNorm1d][, activation]); got " + repr(seq))
    rest = mods[1:]
    bn = rest.pop(0) if rest and isinstance(rest[0], torch.nn.BatchNorm1d) else None
    if len(rest) > 1:
        raise NotImplementedError("readout layers are Sequential(Linear[, BatchNorm1d][, activation]); got " + repr(seq))
    return mods[0], bn, act_spec_of(rest[0] if rest else None)


class ActFn(torch.autograd.Function):
    """out = act(z) for an activation the linear kernels do not fuse (hgin_act_fwd / hgin_act_bwd)."""

    @staticmethod
    def forward(ctx, z, alpha, spec):
        ctx.spec = spec
        ctx.save_for_backward(z, alpha)
        return ops.act_fwd(z, spec.code, alpha, spec.p0, spec.p1)

    @staticmethod
    def backward(ctx, g):
        z, alpha = ctx.saved_tensors
        spec = ctx.spec
        if g.stride(-1) != 1 or g.dtype != z.dtype:
            g = g.to(z.dtype).contiguous()
        want_alpha = alpha is not None and ctx.needs_input_grad[1]
        dz, dalpha = ops.act_bwd(g, z, spec.code, alpha, spec.p0, spec.p1, want_dalpha=want_alpha)
        return (dz if ctx.needs_input_grad[0] else None), (None if dalpha is None else dalpha.view_as(alpha)), None


class BatchNormActFn(torch.autograd.Function):
    """out = act(BatchNorm1d(z)) (models.py:303-313).  Training mode normalises with the statistics of the GLOBAL batch:
    under data parallelism the column sums are all-reduced (`comm`), so N ranks reproduce the single-process step on the
    concatenated batch; the running buffers are updated in place as torch.nn.BatchNorm1d does."""

    @staticmethod
    def forward(ctx, z, gamma, beta, alpha, bn, spec, comm):
        n = z.shape[1]
        training = bn.training or (bn.running_mean is None and bn.running_var is None)
        count = z.shape[0]
        if training:
            if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            momentum = bn.momentum
            if momentum is None:      # cumulative moving average (needs the host value of the counter)
                momentum = 1.0 / float(bn.num_batches_tracked)
            sums = ops.bn_stats(z)
            if comm is not None and comm.world > 1:
                comm.all_reduce_sum_(sums)
                count = None          # global count lives in sums[2n] on the device; read lazily below
            update = bn.training and bn.track_running_stats
            mean, invstd = ops.bn_finalize(n, sums, bn.eps, momentum, bn.running_mean if update else None,
                                           bn.running_var if update else None)
            ctx.count_t = sums[2 * n:2 * n + 1] if count is None else None
        else:
            mean, invstd = ops.bn_finalize(n, None, bn.eps, 0.0, bn.running_mean, bn.running_var, use_running=True)
            ctx.count_t = None
        if count is not None and training and count < 2 and bn.training:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(z.shape)}")
        ctx.spec, ctx.training, ctx.count, ctx.comm = spec, training, count, comm
        ctx.save_for_backward(z, gamma, beta, alpha, mean, invstd)
        return ops.bn_act_fwd(z, mean, invstd, gamma, beta, spec.code, alpha, spec.p0, spec.p1)

    @staticmethod
    def backward(ctx, g):
        z, gamma, beta, alpha, mean, invstd = ctx.saved_tensors
        spec = ctx.spec
        if g.stride(-1) != 1 or g.dtype != z.dtype:
            g = g.to(z.dtype).contiguous()
        sums = ops.bn_act_bwd_reduce(g, z, mean, invstd, gamma, beta, spec.code, alpha, spec.p0, spec.p1)
        count = ctx.count
        local = None
        if ctx.training and ctx.comm is not None and ctx.comm.world > 1:
            # dz needs the GLOBAL reductions; the parameter gradients leave as this rank's PARTIAL sums, because the
            # step all-reduces the flat gradient bucket with SUM afterwards (parallel.py)
            local = sums.clone()
            ctx.comm.all_reduce_sum_(sums)
            count = float(ctx.count_t.item())
        nz, ng, nb, na = ctx.needs_input_grad[:4]
        dz, dgamma, dbeta, dalpha = ops.bn_act_bwd_apply(
            g, z, mean, invstd, gamma, beta, spec.code, sums, count, training=ctx.training, alpha=alpha, p0=spec.p0,
            p1=spec.p1, want_dgamma=bool(ng) and gamma is not None, want_dbeta=bool(nb) and beta is not None,
            want_dalpha=bool(na) and alpha is not None)
        if local is not None:
            n = z.shape[1]
            dbeta = None if dbeta is None else local[:n].float()
            dgamma = None if dgamma is None else local[n:2 * n].float()
            dalpha = None if dalpha is None else local[2 * n:2 * n + 1].float()
        return (dz if nz else None, dgamma, dbeta, None if dalpha is None else dalpha.view_as(alpha), None, None, None)


class DropoutFn(torch.autograd.Function):
    """torch.nn.functional.dropout(x, p, training=True) (models.py:358-359) with a Philox mask regenerated in backward."""

    @staticmethod
    def forward(ctx, x, p, seed, offset):
        ctx.args = (p, seed, offset)
        return ops.dropout(x, p, seed, offset)

    @staticmethod
    def backward(ctx, g):
        p, seed, offset = ctx.args
        if g.stride(-1) != 1:
            g = g.contiguous()
        return ops.dropout(g, p, seed, offset), None, None, None


class GATAggregateFn(torch.autograd.Function):
    """PyG GATConv from the attention logits on (hgin_gat_fwd / hgin_gat_bwd): out = [prev +] softmax-weighted sum of the
    projected source rows + bias.  `prev` (optional) is the running HeteroConv sum for this destination type; it is
    updated IN PLACE (the 'sum' merge rides on the kernel's store) and returned."""

    @staticmethod
    def forward(ctx, xs, a_src, a_dst, bias, prev, graph, et, heads, channels, slope, add_self_loops):
        csr = graph.fwd(et)
        out, row_max, row_sum = ops.gat_fwd(csr, xs, a_src, a_dst, bias, heads, channels, slope, add_self_loops,
                                            out=prev, accumulate=prev is not None)
        if prev is not None:
            ctx.mark_dirty(prev)
        if any(ctx.needs_input_grad):
            graph.bwd(et)      # transposed CSR built alongside the forward work
        ctx.graph, ctx.et, ctx.args = graph, et, (heads, channels, slope, add_self_loops)
        ctx.save_for_backward(xs, a_src, a_dst, row_max, row_sum)
        return out

    @staticmethod
    def backward(ctx, g):
        xs, a_src, a_dst, row_max, row_sum = ctx.saved_tensors
        heads, channels, slope, loops = ctx.args
        if g.stride(-1) != 1:
            g = g.contiguous()
        d_xs, d_a_src, d_a_dst = ops.gat_bwd(ctx.graph.fwd(ctx.et), ctx.graph.bwd(ctx.et), xs, a_src, a_dst, row_max, row_sum,
                                             g, heads, channels, slope, loops)
        d_bias = ops.column_sums(g) if ctx.needs_input_grad[3] else None
        d_prev = g if ctx.needs_input_grad[4] else None
        return d_xs, d_a_src, d_a_dst, d_bias, d_prev, None, None, None, None, None, None
