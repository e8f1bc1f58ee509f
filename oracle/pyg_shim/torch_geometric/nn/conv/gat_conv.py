"""TEST INFRASTRUCTURE (see ../../../README.md): restatement of torch_geometric.nn.conv.GATConv (2.0.2) and of the
pieces it pulls in (`nn.dense.linear.Linear` with lazy `in_channels=-1` and `weight_initializer='glorot'`,
`nn.inits.glorot / zeros`, `utils.softmax`, `utils.remove_self_loops / add_self_loops`), for the call shapes HetroGAT
uses (models.py:413-428, 466): bipartite `(x_src, x_dst)` inputs with a dense `[2, E]` edge_index, `heads >= 1`,
`concat=True`, `add_self_loops=True`, `negative_slope=0.2`, attention dropout 0.

PARITY UNPINNED: PyG is not vendored by the reference, not installed here and cannot be fetched; this file restates the
published 2.0.2 algorithm from its documentation and source as recalled, and the reference has no tests or golden vectors
for it.  What IS pinned: the reference's own `HetroGAT` module code (unmodified) runs on top of it."""
import math

import torch
import torch.nn.functional as F
from torch.nn.parameter import Parameter, UninitializedParameter

from .message_passing import MessagePassing


def glorot(tensor):
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


class Linear(torch.nn.Module):
    """torch_geometric.nn.dense.linear.Linear: `in_channels = -1` defers the weight to the first forward call."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer, self.bias_initializer = weight_initializer, bias_initializer
        if in_channels > 0:
            self.weight = Parameter(torch.Tensor(out_channels, in_channels))
        else:
            self.weight = UninitializedParameter()
            self._hook = self.register_forward_pre_hook(self.initialize_parameters)
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.in_channels > 0:
            if self.weight_initializer == "glorot":
                glorot(self.weight)
            elif self.weight_initializer is None:
                torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            else:
                raise RuntimeError(self.weight_initializer)
        if self.in_channels > 0 and self.bias is not None:
            if self.bias_initializer == "zeros":
                zeros(self.bias)
            elif self.bias_initializer is None:
                bound = 1 / math.sqrt(self.weight.size(-1))
                torch.nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)

    @torch.no_grad()
    def initialize_parameters(self, module, input):
        if isinstance(self.weight, UninitializedParameter):
            self.in_channels = input[0].size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        self._hook.remove()
        delattr(self, "_hook")


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], edge_attr


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    loop_index = torch.arange(0, num_nodes, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop_index], dim=1), edge_attr


def softmax(src, index, ptr=None, num_nodes=None):
    """torch_geometric.utils.softmax over the groups given by `index` (dim 0)."""
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    src_max = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype).scatter_reduce(
        0, idx, src, reduce="amax", include_self=True)
    src_max = src_max.index_select(0, index)
    out = (src - src_max).exp()
    out_sum = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).scatter_add_(0, idx, out)
    out_sum = out_sum.index_select(0, index)
    return out / (out_sum + 1e-16)


class GATConv(MessagePassing):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(node_dim=0, **kwargs)
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.dropout, self.add_self_loops = negative_slope, dropout, add_self_loops
        if isinstance(in_channels, int):
            self.lin_src = Linear(in_channels, heads * out_channels, bias=False, weight_initializer="glorot")
            self.lin_dst = self.lin_src
        else:
            self.lin_src = Linear(in_channels[0], heads * out_channels, False, weight_initializer="glorot")
            self.lin_dst = Linear(in_channels[1], heads * out_channels, False, weight_initializer="glorot")
        self.att_src = Parameter(torch.Tensor(1, heads, out_channels))
        self.att_dst = Parameter(torch.Tensor(1, heads, out_channels))
        if bias and concat:
            self.bias = Parameter(torch.Tensor(heads * out_channels))
        elif bias and not concat:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_src.reset_parameters()
        self.lin_dst.reset_parameters()
        glorot(self.att_src)
        glorot(self.att_dst)
        zeros(self.bias)

    def forward(self, x, edge_index, size=None):
        H, C = self.heads, self.out_channels
        if isinstance(x, torch.Tensor):
            x_src = x_dst = self.lin_src(x).view(-1, H, C)
        else:
            x_src, x_dst = x
            x_src = self.lin_src(x_src).view(-1, H, C)
            if x_dst is not None:
                x_dst = self.lin_dst(x_dst).view(-1, H, C)
        alpha_src = (x_src * self.att_src).sum(dim=-1)
        alpha_dst = None if x_dst is None else (x_dst * self.att_dst).sum(-1)
        if self.add_self_loops:
            # "only for nodes that appear both as source and target nodes": ids are compared ACROSS node types here
            num_nodes = x_src.size(0)
            if x_dst is not None:
                num_nodes = min(num_nodes, x_dst.size(0))
            num_nodes = min(size) if size is not None else num_nodes
            edge_index, _ = remove_self_loops(edge_index)
            edge_index, _ = add_self_loops(edge_index, num_nodes=num_nodes)
        # propagate(edge_index, x=(x_src, x_dst), alpha=(alpha_src, alpha_dst)) with message() below, aggr='add'
        j, i = edge_index[0], edge_index[1]
        n_dst = x_dst.size(0) if x_dst is not None else x_src.size(0)
        alpha = alpha_src.index_select(0, j)
        if alpha_dst is not None:
            alpha = alpha + alpha_dst.index_select(0, i)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = softmax(alpha, i, None, n_dst)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = x_src.index_select(0, j) * alpha.unsqueeze(-1)
        out = torch.zeros(n_dst, H, C, dtype=msg.dtype).scatter_add_(0, i.view(-1, 1, 1).expand_as(msg), msg)
        if self.concat:
            out = out.view(-1, H * C)
        else:
            out = out.mean(dim=1)
        if self.bias is not None:
            out += self.bias
        return out
