"""Module mirror of the reference's HeteroGIN path — same class names, constructor and `forward`
signatures, attribute tree and `state_dict` keys as `/root/reference/models.py:180-376` (plus the
PyG `HeteroConv` container it instantiates, models.py:286-298) — with every tensor operation
routed to the sm_100a kernels in libhgin.so.  A reference user swaps

    from models import HetroGIN            ->   from gnn_link_prediction_b200.models import HetroGIN

and keeps `train.py`'s step unchanged: `model(sample.x_dict, sample.edge_index_dict,
sample["path"].batch)` (train.py:34) returns `f32[N_path, 1]` with an autograd graph.

Differences from the reference, all deliberate and documented in DESIGN.md:
* relations whose output cannot reach the readout are not evaluated (the reference computes and
  discards them); their parameters keep `grad=None` either way;
* `dropout > 0` draws its masks from a Philox stream of this package (hgin_dropout), not from torch's generator:
  the reference's own CPU and CUDA masks differ from each other in the same way (parity is statistical);
* `mlp_bn=True` under data parallelism normalises with the statistics of the GLOBAL batch (the column sums are
  all-reduced), so N ranks reproduce the single-process step;
* inputs must be CUDA tensors: there is no CPU path.
"""
from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

from . import functional as F_
from . import ops
from .functional import (ActFn, BatchNormActFn, DropoutFn, GATAggregateFn, GraphCSR, HeteroConvFn, LinearActFn,
                         RelationSpec)
from .ops import MATH_BF16, MATH_FP32, MATH_TF32  # noqa: F401


def reset(value: Any):
    """models.py:162-167."""
    if hasattr(value, "reset_parameters"):
        value.reset_parameters()
    else:
        for child in value.children() if hasattr(value, "children") else []:
            reset(child)


class GINConv(torch.nn.Module):
    """models.py:180-228.  `forward(x, edge_index)` with `x` a tensor or a `(x_src, x_dst)` pair and
    `edge_index` a `[2,E]` COO tensor; `nn` must be `Sequential(Linear[, PReLU|ReLU])`."""

    def __init__(self, nn, eps: float = 0.0, train_eps: bool = False, concat: bool = False, **kwargs):
        aggr = kwargs.pop("aggr", "add")
        if aggr not in ("add", "sum"):
            raise NotImplementedError(f"aggr={aggr!r}: the reference path uses sum aggregation (models.py:186)")
        if kwargs:
            raise TypeError(f"unsupported MessagePassing arguments: {sorted(kwargs)}")
        super().__init__()
        self.aggr = "add"
        self.nn = nn
        self.initial_eps = eps
        self.concat = concat
        self.math_mode = MATH_FP32
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))
        self.reset_parameters()

    def reset_parameters(self):
        reset(self.nn)
        self.eps.data.fill_(self.initial_eps)

    def kernel_args(self):
        W, b, act, alpha = F_.linear_act_of(self.nn)
        return W, b, alpha, self.eps, act

    def forward(self, x, edge_index, size=None):
        if isinstance(x, Tensor):
            types, xs, et = ("n",), (x,), ("n", "to", "n")
        else:
            types, xs, et = ("src", "dst"), (x[0], x[1]), ("src", "to", "dst")
            if xs[1] is None:
                raise NotImplementedError("x = (x_src, None) is not used by the reference path")
        num = {t: v.shape[0] for t, v in zip(types, xs)}
        if size is not None and (size[0] not in (None, num[et[0]]) or size[1] not in (None, num[et[2]])):
            raise ValueError(f"size {size} does not match the feature matrices {num}")
        graph = edge_index if isinstance(edge_index, GraphCSR) else GraphCSR({et: edge_index}, num)
        W, b, alpha, eps, act = self.kernel_args()
        (out,) = HeteroConvFn.apply([RelationSpec(et, self.concat, act)], graph, types, self.math_mode, None, None,
                                    False, *xs, W, b, alpha, eps)
        return out

    def __repr__(self):
        return "{}(nn={})".format(self.__class__.__name__, self.nn)


class GINLayer(torch.nn.Module):
    """models.py:231-245: `mlp = Linear(in, out) -> PReLU()`, aliased as `conv.nn`; eps learnable."""

    def __init__(self, in_channels: int, out_channels: int, concat: bool = False) -> None:
        super().__init__()
        self.mlp = torch.nn.Sequential(torch.nn.Linear(in_channels, out_channels), torch.nn.PReLU())
        self.conv = GINConv(self.mlp, eps=0, train_eps=True, concat=concat)

    def forward(self, x, edge_index):
        return self.conv(x, edge_index)


class GATConv(torch.nn.Module):
    """torch_geometric.nn.conv.GATConv (2.0.2) as HetroGAT instantiates it (models.py:417-428): `in_channels` an int (one
    projection shared by source and destination, `lin_dst is lin_src`) or a pair — `(-1, -1)` defers the two weights until
    `materialize()` / the first call, as PyG's lazy Linear does; `concat=True`, `add_self_loops=True`, attention dropout 0.
    Same parameter names and shapes as PyG: `lin_src.weight`, `lin_dst.weight` (glorot), `att_src`, `att_dst`
    `[1, heads, out_channels]` (glorot), `bias [heads * out_channels]` (zeros).  The softmax / aggregation runs in
    hgin_gat_fwd / hgin_gat_bwd, the projections and attention logits in hgin_linear_*."""

    def __init__(self, in_channels, out_channels: int, heads: int = 1, concat: bool = True, negative_slope: float = 0.2,
                 dropout: float = 0.0, add_self_loops: bool = True, bias: bool = True, **kwargs):
        if kwargs.pop("aggr", "add") not in ("add", "sum"):
            raise NotImplementedError("GATConv: only aggr='add' (PyG's default) is implemented")
        if kwargs:
            raise TypeError(f"unsupported MessagePassing arguments: {sorted(kwargs)}")
        if not concat or dropout != 0.0 or not bias:
            raise NotImplementedError("GATConv: the reference uses concat=True, dropout=0.0, bias=True (models.py:417-428)")
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.dropout, self.add_self_loops = negative_slope, dropout, add_self_loops
        self.math_mode = MATH_FP32
        self.lin_src = self.lin_dst = None
        if isinstance(in_channels, int):
            if in_channels > 0:
                self.lin_src = self._linear(in_channels, heads * out_channels)
                self.lin_dst = self.lin_src
        elif in_channels[0] > 0 and in_channels[1] > 0:
            self.lin_src = self._linear(in_channels[0], heads * out_channels)
            self.lin_dst = self._linear(in_channels[1], heads * out_channels)
        self.att_src = torch.nn.Parameter(torch.Tensor(1, heads, out_channels))
        self.att_dst = torch.nn.Parameter(torch.Tensor(1, heads, out_channels))
        self.bias = torch.nn.Parameter(torch.Tensor(heads * out_channels))
        self.reset_parameters()

    @classmethod
    def _linear(cls, n_in, n_out):
        # PyG's Linear(weight_initializer='glorot') draws ONLY the glorot values (once, in its constructor): no default
        # torch init, which would consume the generator, before it
        lin = torch.nn.utils.skip_init(torch.nn.Linear, n_in, n_out, bias=False)
        cls._glorot(lin.weight)
        return lin

    @staticmethod
    def _glorot(t):
        # torch_geometric.nn.inits.glorot; drawn from the CPU generator whatever the device, so that a model built
        # (or materialised) on the GPU is seed-for-seed identical to the reference built on the CPU
        bound = (6.0 / (t.size(-2) + t.size(-1))) ** 0.5
        with torch.no_grad():
            t.copy_(torch.empty(t.shape, dtype=t.dtype).uniform_(-bound, bound))

    def reset_parameters(self):
        # PyG: lin_src.reset_parameters(); lin_dst.reset_parameters() — two draws on the SAME weight when the projection
        # is shared (kept: it is part of the reference's RNG consumption)
        if self.lin_src is not None:
            self._glorot(self.lin_src.weight)
            self._glorot(self.lin_dst.weight)
        self._glorot(self.att_src)
        self._glorot(self.att_dst)
        with torch.no_grad():
            self.bias.zero_()

    def materialize(self, in_src: int, in_dst: int):
        """Create the deferred `(-1, -1)` projections (PyG: first forward call; glorot, source first)."""
        if self.lin_src is None:
            dev = self.att_src.device
            self.lin_src = self._linear(in_src, self.heads * self.out_channels).to(dev)
            if isinstance(self.in_channels, int):
                self.lin_dst = self.lin_src
            else:
                self.lin_dst = self._linear(in_dst, self.heads * self.out_channels).to(dev)
        return self

    def run(self, x_src, x_dst, graph, et, prev=None):
        H, C = self.heads, self.out_channels
        self.materialize(x_src.shape[1], x_dst.shape[1])
        W_src, W_dst = self.lin_src.weight, self.lin_dst.weight
        # attention logits a = <lin(x), att> per head, taken as x V^T with V = att folded into the projection: the [N, H*C]
        # destination projection is never formed (weight-space glue on [H*C, in] tensors, differentiated by autograd)
        V_src = torch.einsum("hc,hci->hi", self.att_src[0], W_src.view(H, C, -1)).contiguous()
        V_dst = torch.einsum("hc,hci->hi", self.att_dst[0], W_dst.view(H, C, -1)).contiguous()
        xs = LinearActFn.apply(x_src, None, W_src, None, None, F_.ACT_NONE, self.math_mode, None, None)
        a_src = LinearActFn.apply(x_src, None, V_src, None, None, F_.ACT_NONE, MATH_FP32, None, None)
        a_dst = LinearActFn.apply(x_dst, None, V_dst, None, None, F_.ACT_NONE, MATH_FP32, None, None)
        return GATAggregateFn.apply(xs, a_src, a_dst, self.bias, prev, graph, et, H, C, float(self.negative_slope),
                                    bool(self.add_self_loops))

    def forward(self, x, edge_index, size=None):
        if isinstance(x, Tensor):
            types, xs, et = ("n",), (x, x), ("n", "to", "n")
        else:
            types, xs, et = ("src", "dst"), (x[0], x[1]), ("src", "to", "dst")
            if xs[1] is None:
                raise NotImplementedError("x = (x_src, None) is not used by the reference path")
        num = {et[0]: xs[0].shape[0], et[2]: xs[1].shape[0]}
        graph = edge_index if isinstance(edge_index, GraphCSR) else GraphCSR({et: edge_index}, num)
        return self.run(xs[0], xs[1], graph, et)

    def __repr__(self):
        return "{}({}, {}, heads={})".format(self.__class__.__name__, self.in_channels, self.out_channels, self.heads)


class HeteroConv(torch.nn.Module):
    """The subset of PyG 2.0.2 `HeteroConv` the reference uses (models.py:286-298, 356): a dict of
    per-relation GIN layers, outputs of relations with the same destination type summed.  All live
    relations of the layer run inside ONE autograd node so that the merge and the per-node-type
    backward gathers are fused."""

    def __init__(self, convs: dict, aggr: str = "sum"):
        super().__init__()
        if aggr not in ("sum", "add"):
            raise NotImplementedError(f"HeteroConv aggr={aggr!r}: the reference uses 'sum' (models.py:290)")
        kinds = {isinstance(v, GATConv) for v in convs.values()}
        for k, v in convs.items():
            if not isinstance(v, (GINLayer, GINConv, GATConv)):
                raise NotImplementedError(f"relation {k}: only GINLayer / GINConv / GATConv modules have kernels")
        if len(kinds) > 1:
            raise NotImplementedError("HeteroConv: GIN and GAT relations cannot be mixed in one layer")
        self.is_gat = kinds == {True}
        self.convs = torch.nn.ModuleDict({"__".join(k): v for k, v in convs.items()})
        self.aggr = aggr
        self.math_mode = MATH_FP32

    def reset_parameters(self):
        for conv in self.convs.values():
            conv.reset_parameters()

    def forward(self, x_dict, edge_index_dict, only=None, chain=None, lazy=False):
        """`only`: optional collection of edge types to evaluate (dead-branch pruning by HetroGIN).
        `chain`: dict node type -> ops.PostAct describing how `x_dict`'s tensors were produced, passed
        by HetroGIN.forward between consecutive layers (every intermediate has one consumer there);
        updated in place to describe this layer's outputs.  `lazy`: the next consumer is another layer of
        the chain, so single-relation outputs may be handed over as pre-activations.  See
        functional.HeteroConvFn."""
        graph = edge_index_dict if isinstance(edge_index_dict, GraphCSR) else GraphCSR(
            edge_index_dict, {t: v.shape[0] for t, v in x_dict.items()})
        if self.is_gat:
            outs = {}
            for et in graph.keys():
                key = "__".join(et)
                if key not in self.convs or (only is not None and tuple(et) not in only):
                    continue
                src, dst = et[0], et[-1]
                outs[dst] = self.convs[key].run(x_dict[src], x_dict[dst], graph, tuple(et), prev=outs.get(dst))
            return outs
        specs, params = [], []
        for et in graph.keys():                      # insertion order of the batch's relations
            key = "__".join(et)
            if key not in self.convs or (only is not None and tuple(et) not in only):
                continue
            m = self.convs[key]
            conv = m.conv if isinstance(m, GINLayer) else m
            W, b, alpha, eps, act = conv.kernel_args()
            specs.append(RelationSpec(et, conv.concat, act))
            params += [W, b, alpha, eps]
        if not specs:
            return {}
        types = tuple(dict.fromkeys(t for sp in specs for t in (sp.src, sp.dst)))
        links_out = {} if chain is not None else None
        outs = HeteroConvFn.apply(specs, graph, types, self.math_mode, dict(chain) if chain else None, links_out,
                                  bool(lazy and chain is not None), *[x_dict[t] for t in types], *params)
        if chain is not None:
            chain.clear()
            chain.update(links_out)
        out_types = list(dict.fromkeys(sp.dst for sp in specs))
        return dict(zip(out_types, outs))


class _HetroBase(torch.nn.Module):
    """What HetroGIN and HetroGAT share in the reference (the two classes repeat it verbatim, models.py:301-376 and
    431-506): the readout construction, the feature slicing, GLOBAL_FEATS, the layer loop with dropout, the readout."""

    RELATIONS = (("path", "uses", "link"), ("link", "includes", "path"),
                 ("link", "connects", "node"), ("node", "has", "link"))

    def _build_readout(self, width, mlp_layers, act, mlp_head_act, mlp_bn):
        self.readout = torch.nn.ModuleList()
        act = eval(act)                                 # one shared activation object (models.py:301)
        F_.act_spec_of(act)                             # fail at construction if no kernel implements it
        for w in mlp_layers:
            if mlp_bn:                                  # models.py:303-313
                self.readout.append(torch.nn.Sequential(torch.nn.Linear(width, w), torch.nn.BatchNorm1d(num_features=w), act))
            else:
                self.readout.append(torch.nn.Sequential(torch.nn.Linear(width, w), act))
            width = w
        if mlp_head_act is None:
            self.readout.append(torch.nn.Sequential(torch.nn.Linear(mlp_layers[-1], 1)))
        else:
            head = eval(mlp_head_act)
            F_.act_spec_of(head)
            self.readout.append(torch.nn.Sequential(torch.nn.Linear(mlp_layers[-1], 1), head))
        self._dropout_calls = 0
        self.communicator = None    # set by TrainStep: BatchNorm statistics are all-reduced over it
        self._edges_validated = False

    def set_math_mode(self, mode):
        """MATH_FP32 (parity), MATH_TF32 (tcgen05 tf32 GEMMs, fp32 activations) or MATH_BF16 (activations and
        gradients stored as bf16, tcgen05 bf16 GEMMs, fp32 accumulation / aggregation adds / loss / optimizer) for
        every dense layer of the model."""
        if mode == MATH_BF16 and any(isinstance(m, GATConv) for m in self.modules()):
            raise NotImplementedError("HetroGAT: the attention kernels take fp32 rows (MATH_FP32 or MATH_TF32)")
        self.math_mode = mode
        for m in self.modules():
            if isinstance(m, (HeteroConv, GINConv, GATConv)):
                m.math_mode = mode
        return self

    def live_relations(self, edge_types):
        """Per layer, the relations whose output can reach the readout (which reads only
        x_dict['path'], models.py:362-371): walk back from {'path'}."""
        present = [tuple(et) for et in edge_types if "__".join(et) in self.convs[0].convs]
        needed = {"path"}
        live = [None] * self.num_layers
        for li in reversed(range(self.num_layers)):
            live[li] = [et for et in present if et[2] in needed]
            needed = {t for et in live[li] for t in (et[0], et[2])}
        return live

    def _validate_once(self, graph, was_coo):
        """PyG raises on an edge_index that points outside its node sets; hgin_csr_build drops such edges and raises a
        device-side flag instead (no sync in the step).  The flag is read back ONCE per model, after the first forward
        call on a COO `edge_index_dict` (one synchronisation), so that a corrupt dataset is reported, not absorbed."""
        if was_coo and not self._edges_validated:
            self._edges_validated = True
            if not torch.cuda.is_current_stream_capturing():
                graph.validate()

    def _slice_inputs(self, x_dict):
        """Feature slicing, rebinding the caller's dict like models.py:333-342."""
        if not self.divided_features:
            if self.bl_features:
                x_dict["path"] = torch.cat([x_dict["path"][:, 0:3], x_dict["path"][:, 6].reshape(-1, 1)], axis=1)
                x_dict["link"] = torch.cat([x_dict["link"][:, 0:3], x_dict["link"][:, 4:7]], axis=1)
            else:   # cat(...)[:, 0:3] of the reference == the first three raw columns: a view (row pitch 7), no copy
                x_dict["path"] = x_dict["path"][:, 0:3]
                x_dict["link"] = x_dict["link"][:, 0:3]
        elif not self.bl_features:
            x_dict["path"] = x_dict["path"][:, 0:6]
            x_dict["link"] = x_dict["link"][:, 0:3]
        return x_dict["path"]

    def _readout_tail(self, origin_path, path_batch, num_graphs):
        """The constant columns of the readout input: the raw path columns (CONCAT_PATH) and, with GLOBAL_FEATS
        (models.py:347-352, 364-369), their per-graph mean / max broadcast back per path."""
        x2 = origin_path if self.concat_path else None
        if self.global_feats:
            if origin_path.requires_grad:
                raise NotImplementedError("global_feats: gradients w.r.t. the raw path features are not propagated")
            if path_batch is None:
                raise ops.HginError("global_feats=True needs `path_batch` (collate with batch_vector=True)")
            if num_graphs is None:
                num_graphs = int(path_batch.max()) + 1 if path_batch.numel() else 0
            x2 = ops.global_pool_tail(origin_path, path_batch, num_graphs, origin_path.shape[1] if self.concat_path else 0)[0]
        return x2

    def _dropout_outputs(self, x_dict):
        """models.py:358-359 on every output of a layer (training mode, p > 0)."""
        seed = int(torch.empty((), dtype=torch.int64).random_())      # torch's CPU generator: follows manual_seed
        for j, k in enumerate(list(x_dict)):
            x_dict[k] = DropoutFn.apply(x_dict[k], float(self.dropout), seed, (self._dropout_calls << 8) | j)
        self._dropout_calls += 1
        return x_dict

    def _run_readout(self, x1, x2, link, chained):
        for i, layer in enumerate(self.readout):
            lin, bn, spec = F_.readout_layer_of(layer)
            nxt = [] if chained else None
            if bn is None and spec.fused:
                x1 = LinearActFn.apply(x1, x2 if i == 0 else None, lin.weight, lin.bias, spec.alpha, spec.code, self.math_mode,
                                       link, nxt)
            else:       # Linear -> [BatchNorm1d] -> activation as separate row passes (models.py:303-330)
                z = LinearActFn.apply(x1, x2 if i == 0 else None, lin.weight, lin.bias, None, F_.ACT_NONE, self.math_mode,
                                      link, None)
                if bn is not None:
                    x1 = BatchNormActFn.apply(z, bn.weight, bn.bias, spec.alpha, bn, spec, self.communicator)
                else:
                    x1 = ActFn.apply(z, spec.alpha, spec)
            link = nxt[0] if nxt else None
        return x1


class HetroGIN(_HetroBase):
    """models.py:248-376 (sic: the reference spells it HetroGIN)."""

    def __init__(self, input_channels: dict, node_embedding_size: int, message_passing_layers: int, dropout: float,
                 concat_path: bool, bl_features: bool, divided_features: bool, global_feats: bool,
                 mlp_layers: list, act, mlp_head_act, mlp_bn: bool):
        super().__init__()
        self.num_layers = message_passing_layers
        self.concat_path = concat_path
        self.bl_features = bl_features
        self.divided_features = divided_features
        self.mlp_layers = mlp_layers
        self.dropout = dropout
        self.global_feats = global_feats
        self.math_mode = MATH_FP32
        self.fold_activation_grad = True   # see forward(); only takes effect for tensor-core-sized layers

        # channel arithmetic, in place on the caller's dict exactly as models.py:260-269 does
        if not self.divided_features:
            input_channels["path"] = input_channels["path"] - 3
            input_channels["link"] = input_channels["link"] - 1
            if not self.bl_features:
                input_channels["path"] = input_channels["path"] - 1
                input_channels["link"] = input_channels["link"] - 3
        elif not self.bl_features:
            input_channels["path"] = input_channels["path"] - 1
            input_channels["link"] = input_channels["link"] - 3
        self.global_feats_size = 8 if global_feats else 0    # models.py:271-274 (sic: 2 x 4 path columns)
        self.concat_size = input_channels["path"] if concat_path else 0

        emb = node_embedding_size
        self.convs = torch.nn.ModuleList()
        # construction order = RNG consumption order of the reference (models.py:286-298)
        self.convs.append(HeteroConv({r: GINLayer(input_channels[r[0]] + input_channels[r[2]], emb, concat=True)
                                      for r in self.RELATIONS}, aggr="sum"))
        for _ in range(self.num_layers - 1):
            self.convs.append(HeteroConv({r: GINLayer(emb, emb) for r in self.RELATIONS}, aggr="sum"))

        self._build_readout(emb + self.concat_size + self.global_feats_size, mlp_layers, act, mlp_head_act, mlp_bn)

    def forward(self, x_dict, edge_index_dict, path_batch=None, num_graphs=None):
        """`num_graphs` (optional, not in the reference signature): the number of graphs in the batch when the caller
        knows it (TrainStep does) — without it GLOBAL_FEATS reads `path_batch.max() + 1` back, as PyG's pools do."""
        drop = self.training and self.dropout > 0
        origin_path = self._slice_inputs(x_dict)
        graph = edge_index_dict if isinstance(edge_index_dict, GraphCSR) else GraphCSR(
            edge_index_dict, {t: v.shape[0] for t, v in x_dict.items()})
        live = self.live_relations(graph.keys())
        # Inside this function every intermediate activation has exactly one consumer, so a layer's
        # backward may hand the layer below its dz instead of g (functional.HeteroConvFn).
        # (a dropout between two layers is a second consumer-side op on every intermediate: no folding then)
        chain = {} if (self.fold_activation_grad and torch.is_grad_enabled() and not drop) else None
        x2 = self._readout_tail(origin_path, path_batch, num_graphs)
        for i in range(self.num_layers):
            x_dict = self.convs[i](x_dict, graph, only=live[i], chain=chain, lazy=i < self.num_layers - 1)
            if drop:
                x_dict = self._dropout_outputs(x_dict)
        self._validate_once(graph, graph is not edge_index_dict)
        link = chain.get("path") if chain is not None else None
        return self._run_readout(x_dict["path"], x2, link, chain is not None)


class HetroGAT(_HetroBase):
    """models.py:380-506.  Layer 0: one GATConv per relation over the raw feature widths (the reference passes the lazy
    `(-1, -1)`; the widths are known here, so the projections are created — in the order the reference's first forward
    call would create them — at the end of the constructor), `heads` heads concatenated; layers >= 1: GATConv(emb, emb)
    with ONE head, exactly as the reference builds them (so MP_LAYERS > 1 only runs with HEADS = 1, models.py:413-428).
    No activation between layers.  Relations whose output cannot reach the readout are not evaluated."""

    def __init__(self, input_channels: dict, node_embedding_size: int, message_passing_layers: int, dropout: float,
                 heads: int, concat_path: bool, bl_features: bool, divided_features: bool, global_feats: bool,
                 mlp_layers: list, act, mlp_head_act, mlp_bn: bool):
        super().__init__()
        self.num_layers = message_passing_layers
        self.dropout = dropout
        self.concat_path = concat_path
        self.mlp_layers = mlp_layers
        self.global_feats = global_feats
        self.heads = heads
        self.bl_features = bl_features
        self.divided_features = divided_features
        self.math_mode = MATH_FP32
        self.global_feats_size = 8 if global_feats else 0
        # models.py:397-408 (unlike HetroGIN, the caller's dict is left untouched)
        cut = {(True, True): (0, 0), (True, False): (1, 3), (False, True): (3, 1), (False, False): (4, 4)}
        p_cut, l_cut = cut[(bool(divided_features), bool(bl_features))]
        widths = {"path": input_channels["path"] - p_cut, "link": input_channels["link"] - l_cut,
                  "node": input_channels["node"]}
        self.concat_size = widths["path"] if concat_path else 0
        emb = node_embedding_size
        self.convs = torch.nn.ModuleList()
        self.convs.append(HeteroConv({r: GATConv((-1, -1), emb, heads=heads, concat=True) for r in self.RELATIONS}, aggr="sum"))
        for _ in range(self.num_layers - 1):
            self.convs.append(HeteroConv({r: GATConv(emb, emb) for r in self.RELATIONS}, aggr="sum"))
        self._build_readout(emb * heads + self.concat_size + self.global_feats_size, mlp_layers, act, mlp_head_act, mlp_bn)
        for r in self.RELATIONS:      # what the first forward call of the reference does, in HeteroConv's relation order
            self.convs[0].convs["__".join(r)].materialize(widths[r[0]], widths[r[2]])

    def forward(self, x_dict, edge_index_dict, path_batch=None, num_graphs=None):
        origin_path = self._slice_inputs(x_dict)
        graph = edge_index_dict if isinstance(edge_index_dict, GraphCSR) else GraphCSR(
            edge_index_dict, {t: v.shape[0] for t, v in x_dict.items()})
        live = self.live_relations(graph.keys())
        x2 = self._readout_tail(origin_path, path_batch, num_graphs)
        for i in range(self.num_layers):
            x_dict = self.convs[i](x_dict, graph, only=live[i])
            if self.training and self.dropout > 0:
                x_dict = self._dropout_outputs(x_dict)
        self._validate_once(graph, graph is not edge_index_dict)
        return self._run_readout(x_dict["path"], x2, None, False)
