"""Restatement of torch_geometric.nn.conv.MessagePassing (2.0.2) for the only call shape the
reference uses (models.py:208): `propagate(edge_index: LongTensor[2,E], x=(x_src, x_dst), size=None)`
with `aggr='add'`, flow source_to_target, node_dim=0, and a `message(x_j)` hook."""
import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert aggr in ("add", "sum", "mean")
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        self._msg_params = list(inspect.signature(self.message).parameters)

    def propagate(self, edge_index, size=None, **kwargs):
        assert isinstance(edge_index, torch.Tensor) and edge_index.dim() == 2 and edge_index.size(0) == 2, \
            "shim restates the dense COO path only"
        size = [None, None] if size is None else list(size)
        msg_kwargs = {}
        for name in self._msg_params:
            base, suffix = name[:-2], name[-2:]
            assert suffix in ("_i", "_j"), name
            data = kwargs[base]
            if isinstance(data, (tuple, list)):
                assert len(data) == 2
                for k in (0, 1):
                    if isinstance(data[k], torch.Tensor) and size[k] is None:
                        size[k] = data[k].size(0)
                data = data[0 if suffix == "_j" else 1]
            elif isinstance(data, torch.Tensor):
                size[0] = size[1] = data.size(0)
            # __lift__: j = source = edge_index[0], i = target = edge_index[1]
            idx = edge_index[0] if suffix == "_j" else edge_index[1]
            msg_kwargs[name] = data.index_select(0, idx)
        out = self.message(**msg_kwargs)
        # aggregate
        dim_size = size[1]
        out = scatter(out, edge_index[1], dim=0, dim_size=dim_size,
                      reduce="sum" if self.aggr in ("add", "sum") else self.aggr)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs
