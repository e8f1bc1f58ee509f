"""Restatement of torch_geometric.nn.conv.hetero_conv.HeteroConv (2.0.2); used by the
reference at models.py:286-298 (construction) and models.py:356 (call)."""
from collections import defaultdict
import torch


def group(xs, aggr):
    if len(xs) == 0:
        return None
    elif aggr is None:
        return torch.stack(xs, dim=1)
    elif len(xs) == 1:
        return xs[0]
    else:
        out = torch.stack(xs, dim=0)
        out = getattr(torch, aggr)(out, dim=0)
        out = out[0] if isinstance(out, tuple) else out
        return out


class HeteroConv(torch.nn.Module):
    def __init__(self, convs, aggr="sum"):
        super().__init__()
        self.convs = torch.nn.ModuleDict({"__".join(k): v for k, v in convs.items()})
        self.aggr = aggr

    def reset_parameters(self):
        for conv in self.convs.values():
            conv.reset_parameters()

    def forward(self, x_dict, edge_index_dict, *args_dict, **kwargs_dict):
        out_dict = defaultdict(list)
        for edge_type, edge_index in edge_index_dict.items():
            src, rel, dst = edge_type
            str_edge_type = "__".join(edge_type)
            if str_edge_type not in self.convs:
                continue
            conv = self.convs[str_edge_type]
            if src == dst:
                out = conv(x_dict[src], edge_index)
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index)
            out_dict[dst].append(out)
        for key, value in out_dict.items():
            out_dict[key] = group(value, self.aggr)
        return out_dict
