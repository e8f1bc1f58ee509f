"""TEST INFRASTRUCTURE — CPU oracle for the HeteroGIN hot path.  NOT part of the product.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; `gnn_link_prediction_b200/` never does.

A restatement, in plain PyTorch CPU ops, of the path
`train.py:27-44 -> HetroGIN.forward (models.py:332-376) -> HeteroConv -> GINLayer -> GINConv
(models.py:180-245)` plus the third-party arithmetic it reaches (PyG 2.0.2 `MessagePassing.
propagate`, `HeteroConv`, torch_scatter `scatter`; none vendored in /root/reference, none pinned
by a lock file — "PytorchGeometric version: 2.0.2" at models.py:182 is the only hint).

Parity status: the reference has no tests, fixtures or golden vectors for this path
 — PARITY IS UNPINNED BY THE REFERENCE'S OWN TESTS.  It is pinned instead by
outputs of the reference itself: `oracle/make_golden.py` imports the UNMODIFIED
`/root/reference/models.py` on `oracle/pyg_shim` in the build container and writes
`tests/golden/*.pt`; `tests/test_oracle.py` checks this file against those vectors (and, when
/root/reference is present, against the live reference model) bit-for-bit.

The module tree mirrors the reference's attribute names so `load_state_dict` accepts a
reference `state_dict()` unchanged (keys listed in SURVEY §8(b)).
"""
from __future__ import annotations

import torch
from torch import nn

# HetroGIN wires exactly these relations, in this order (models.py:286-290).
RELATIONS = (
    ("path", "uses", "link"),
    ("link", "includes", "path"),
    ("link", "connects", "node"),
    ("node", "has", "link"),
)


def scatter_sum(src, index, dim_size):
    """torch_scatter.scatter(src, index, dim=0, dim_size=dim_size, reduce='sum'):
    zeros(dim_size, F).scatter_add_(0, index broadcast to src, src)."""
    out = torch.zeros(dim_size, src.size(1), dtype=src.dtype, device=src.device)
    return out.scatter_add_(0, index.view(-1, 1).expand_as(src), src)


class GINConv(nn.Module):
    """models.py:180-228.  propagate (models.py:208) restated as index_select + scatter_sum."""

    def __init__(self, mlp, eps=0.0, train_eps=False, concat=False):
        super().__init__()
        self.nn = mlp
        self.initial_eps = eps
        self.concat = concat
        if train_eps:
            self.eps = nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))

    def forward(self, x, edge_index):
        if isinstance(x, torch.Tensor):
            x = (x, x)
        x_src, x_dst = x
        msg = x_src.index_select(0, edge_index[0])              # __lift__ + message (models.py:219)
        out = scatter_sum(msg, edge_index[1], x_dst.size(0))    # aggregate, aggr='add' (models.py:186)
        if self.concat:                                         # models.py:212-215
            out = torch.cat((out, (1 + self.eps) * x_dst), 1)
        else:
            out += (1 + self.eps) * x_dst
        return self.nn(out)                                     # models.py:217


class GINLayer(nn.Module):
    """models.py:231-245: mlp = Linear -> PReLU, registered twice (self.mlp and conv.nn)."""

    def __init__(self, in_channels, out_channels, concat=False):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_channels, out_channels), nn.PReLU())
        self.conv = GINConv(self.mlp, eps=0, train_eps=True, concat=concat)

    def forward(self, x, edge_index):
        return self.conv(x, edge_index)


class HeteroConv(nn.Module):
    """PyG 2.0.2 HeteroConv(aggr='sum') as called at models.py:356."""

    def __init__(self, convs):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})

    def forward(self, x_dict, edge_index_dict):
        outs = {}
        for edge_type, edge_index in edge_index_dict.items():
            key = "__".join(edge_type)
            if key not in self.convs:
                continue
            src, _, dst = edge_type
            outs.setdefault(dst, []).append(self.convs[key]((x_dict[src], x_dict[dst]), edge_index))
        return {k: v[0] if len(v) == 1 else torch.stack(v, dim=0).sum(dim=0) for k, v in outs.items()}


class HetroGIN(nn.Module):
    """models.py:248-376: any embedding size / layer count / readout widths, `concat_path` on or off,
    `bl_features` / `divided_features` slicing, `global_feats` (models.py:347-352 with PyG's global_mean_pool /
    global_max_pool restated as scatter reductions), `mlp_bn` (models.py:303-313), any `act` / `mlp_head_act`
    string the reference would `eval`, dropout through torch's own generator."""

    def __init__(self, input_channels, node_embedding_size, message_passing_layers, dropout=0.0,
                 concat_path=True, bl_features=False, divided_features=False, global_feats=False,
                 mlp_layers=(128, 32), act="torch.nn.PReLU()", mlp_head_act=None, mlp_bn=False):
        super().__init__()
        ch = dict(input_channels)                      # channel arithmetic: models.py:260-269
        if not divided_features:
            ch["path"] -= 3
            ch["link"] -= 1
        if not bl_features:
            ch["path"] -= 1
            ch["link"] -= 3
        self.num_layers = message_passing_layers
        self.concat_path = concat_path
        self.bl_features = bl_features
        self.divided_features = divided_features
        self.mlp_layers = list(mlp_layers)
        self.dropout = dropout
        self.global_feats = global_feats
        emb = node_embedding_size
        self.convs = nn.ModuleList()
        self.convs.append(HeteroConv({r: GINLayer(ch[r[0]] + ch[r[2]], emb, concat=True) for r in RELATIONS}))
        for _ in range(self.num_layers - 1):
            self.convs.append(HeteroConv({r: GINLayer(emb, emb) for r in RELATIONS}))
        act = eval(act)                                # ONE activation object, shared (models.py:301)
        self.readout = nn.ModuleList()
        width = emb + (ch["path"] if concat_path else 0) + (8 if global_feats else 0)   # models.py:271-279, 317
        for w in self.mlp_layers:
            if mlp_bn:                                 # models.py:303-313
                self.readout.append(nn.Sequential(nn.Linear(width, w), nn.BatchNorm1d(num_features=w), act))
            else:
                self.readout.append(nn.Sequential(nn.Linear(width, w), act))
            width = w
        if mlp_head_act is None:                       # models.py:326-330
            self.readout.append(nn.Sequential(nn.Linear(width, 1)))
        else:
            self.readout.append(nn.Sequential(nn.Linear(width, 1), eval(mlp_head_act)))

    def forward(self, x_dict, edge_index_dict, path_batch=None):
        x_dict = dict(x_dict)
        p, l = x_dict["path"], x_dict["link"]
        if not self.divided_features:                  # models.py:333-338
            p = torch.cat([p[:, 0:3], p[:, 6].reshape(-1, 1)], dim=1)
            l = torch.cat([l[:, 0:3], l[:, 4:7]], dim=1)
            if not self.bl_features:
                p, l = p[:, 0:3], l[:, 0:3]
        elif not self.bl_features:                     # models.py:339-342
            p, l = p[:, 0:6], l[:, 0:3]
        x_dict["path"], x_dict["link"] = p, l
        origin = dict(x_dict)
        if self.global_feats:                          # models.py:347-352
            gmean = global_mean_pool(origin["path"], path_batch)[path_batch]
            gmax = global_max_pool(origin["path"], path_batch)[path_batch]
        for conv in self.convs:
            x_dict = conv(x_dict, edge_index_dict)
            x_dict = {k: torch.nn.functional.dropout(v, p=self.dropout, training=self.training)
                      for k, v in x_dict.items()}
        parts = [x_dict["path"]] + ([origin["path"]] if self.concat_path else []) + ([gmean, gmax] if self.global_feats else [])
        x = torch.cat(parts, 1) if len(parts) > 1 else parts[0]      # models.py:362-371
        for layer in self.readout:
            x = layer(x)
        return x


def global_mean_pool(x, batch):
    """PyG global_mean_pool = torch_scatter.scatter(x, batch, dim=0, reduce='mean'): scatter_add_, count clamped to 1."""
    size = int(batch.max()) + 1
    out = scatter_sum(x, batch, size)
    cnt = torch.zeros(size, dtype=x.dtype).scatter_add_(0, batch, torch.ones(batch.numel(), dtype=x.dtype))
    return out / cnt.clamp_(min=1).view(-1, 1)


def global_max_pool(x, batch):
    """PyG global_max_pool = torch_scatter.scatter(x, batch, dim=0, reduce='max')."""
    size = int(batch.max()) + 1
    out = torch.full((size, x.size(1)), float("-inf"), dtype=x.dtype)
    return out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax", include_self=True)


class GATConv(nn.Module):
    """PyG 2.0.2 GATConv as HetroGAT uses it (models.py:413-428): bipartite input, `concat=True`, `add_self_loops=True`,
    `negative_slope=0.2`, no attention dropout.  PARITY UNPINNED (PyG is not vendored; restated from the published
    algorithm).  `in_channels`: an int (one shared `lin_src is lin_dst`) or a pair of ints."""

    def __init__(self, in_channels, out_channels, heads=1):
        super().__init__()
        self.heads, self.out_channels = heads, out_channels
        if isinstance(in_channels, int):
            self.lin_src = nn.Linear(in_channels, heads * out_channels, bias=False)
            self.lin_dst = self.lin_src
        else:
            self.lin_src = nn.Linear(in_channels[0], heads * out_channels, bias=False)
            self.lin_dst = nn.Linear(in_channels[1], heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        for t in (self.lin_src.weight, self.lin_dst.weight, self.att_src, self.att_dst):   # glorot
            bound = (6.0 / (t.size(-2) + t.size(-1))) ** 0.5
            t.data.uniform_(-bound, bound)

    def forward(self, x, edge_index):
        x_src, x_dst = x
        H, C = self.heads, self.out_channels
        xs = self.lin_src(x_src).view(-1, H, C)
        xd = self.lin_dst(x_dst).view(-1, H, C)
        a_src = (xs * self.att_src).sum(dim=-1)
        a_dst = (xd * self.att_dst).sum(dim=-1)
        # remove_self_loops / add_self_loops: ids compared across the two node types, loops for i < min(N_src, N_dst)
        keep = edge_index[0] != edge_index[1]
        loops = torch.arange(min(xs.size(0), xd.size(0)), dtype=edge_index.dtype)
        j = torch.cat([edge_index[0][keep], loops])
        i = torch.cat([edge_index[1][keep], loops])
        n = xd.size(0)
        e = torch.nn.functional.leaky_relu(a_src.index_select(0, j) + a_dst.index_select(0, i), 0.2)
        idx = i.view(-1, 1).expand_as(e)
        e_max = torch.full((n, H), float("-inf")).scatter_reduce(0, idx, e, reduce="amax", include_self=True)
        w = (e - e_max.index_select(0, i)).exp()
        w_sum = torch.zeros(n, H).scatter_add_(0, idx, w)
        alpha = w / (w_sum.index_select(0, i) + 1e-16)                      # torch_geometric.utils.softmax
        msg = xs.index_select(0, j) * alpha.unsqueeze(-1)
        out = torch.zeros(n, H, C).scatter_add_(0, i.view(-1, 1, 1).expand_as(msg), msg).view(-1, H * C)
        out += self.bias
        return out


class HetroGAT(nn.Module):
    """models.py:380-506.  Layer 0: GATConv over the raw feature widths (lazy `(-1, -1)` in the reference), `heads` heads
    concatenated; layers >= 1: GATConv(emb, emb) with one head (so MP_LAYERS > 1 only works with HEADS = 1, as in the
    reference).  No activation between layers.  Readout as in HetroGIN."""

    def __init__(self, input_channels, node_embedding_size, message_passing_layers, dropout=0.0, heads=16,
                 concat_path=True, bl_features=False, divided_features=False, global_feats=False,
                 mlp_layers=(128, 32), act="torch.nn.PReLU()", mlp_head_act=None, mlp_bn=False):
        super().__init__()
        ch = dict(input_channels)
        if not divided_features:
            ch["path"] -= 3
            ch["link"] -= 1
        if not bl_features:
            ch["path"] -= 1
            ch["link"] -= 3
        self.num_layers, self.dropout, self.concat_path, self.global_feats = message_passing_layers, dropout, concat_path, global_feats
        self.bl_features, self.divided_features, self.mlp_layers, self.heads = bl_features, divided_features, list(mlp_layers), heads
        emb = node_embedding_size
        self.convs = nn.ModuleList()
        self.convs.append(HeteroConv({r: GATConv((ch[r[0]], ch[r[2]]), emb, heads=heads) for r in RELATIONS}))
        for _ in range(self.num_layers - 1):
            self.convs.append(HeteroConv({r: GATConv(emb, emb) for r in RELATIONS}))
        act = eval(act)
        self.readout = nn.ModuleList()
        width = emb * heads + (ch["path"] if concat_path else 0) + (8 if global_feats else 0)     # models.py:435
        for w in self.mlp_layers:
            if mlp_bn:
                self.readout.append(nn.Sequential(nn.Linear(width, w), nn.BatchNorm1d(num_features=w), act))
            else:
                self.readout.append(nn.Sequential(nn.Linear(width, w), act))
            width = w
        if mlp_head_act is None:
            self.readout.append(nn.Sequential(nn.Linear(width, 1)))
        else:
            self.readout.append(nn.Sequential(nn.Linear(width, 1), eval(mlp_head_act)))

    forward = HetroGIN.forward        # slicing, global feats, layer loop, dropout and readout are the same code path


def mape(preds, actuals):
    """train.py:12-13."""
    return 100.0 * torch.mean(torch.abs((preds - actuals) / actuals))


def train_step(model, opt, batch):
    """One step of train.py:27-44 on an already-resident batch.  Returns (loss_value, out)."""
    opt.zero_grad()
    out = model(batch.x_dict, batch.edge_index_dict, batch["path"].batch)
    label = batch["path"].y.reshape(-1, 1)
    loss_value = mape(out, label)
    loss = torch.sqrt(loss_value)
    loss.backward()
    opt.step()
    return loss_value.detach(), out.detach()


def dense_forward_fp64(model, x_dict, edge_index_dict):
    """Independent cross-check: the same network evaluated with dense fp64
    adjacency matmuls instead of gather/scatter, sharing nothing with the path above except the
    weights.  Supports the default feature slicing only."""
    sd = {k: v.double() for k, v in model.state_dict().items()}
    x = {"path": x_dict["path"][:, 0:3].double(), "link": x_dict["link"][:, 0:3].double(),
         "node": x_dict["node"].double()}
    origin = dict(x)

    def prelu(z, a):
        return torch.where(z > 0, z, a * z)

    for li in range(model.num_layers):
        acc = {}
        for et, ei in edge_index_dict.items():
            key = "__".join(et)
            pre = f"convs.{li}.convs.{key}."
            if pre + "mlp.0.weight" not in sd:
                continue
            src, _, dst = et
            A = torch.zeros(x[dst].size(0), x[src].size(0), dtype=torch.float64)
            A.index_put_((ei[1], ei[0]), torch.ones(ei.size(1), dtype=torch.float64), accumulate=True)
            agg = A @ x[src]
            self_term = (1 + sd[pre + "conv.eps"]) * x[dst]
            h = torch.cat((agg, self_term), 1) if li == 0 else agg + self_term
            z = h @ sd[pre + "mlp.0.weight"].t() + sd[pre + "mlp.0.bias"]
            o = prelu(z, sd[pre + "mlp.1.weight"])
            acc[dst] = o if dst not in acc else acc[dst] + o
        x = acc
    h = torch.cat((x["path"], origin["path"]), 1) if model.concat_path else x["path"]
    n_hidden = len(model.mlp_layers)
    for i in range(n_hidden):
        h = prelu(h @ sd[f"readout.{i}.0.weight"].t() + sd[f"readout.{i}.0.bias"], sd[f"readout.{i}.1.weight"])
    return h @ sd[f"readout.{n_hidden}.0.weight"].t() + sd[f"readout.{n_hidden}.0.bias"]
