"""TEST INFRASTRUCTURE (see ../README.md): restatement of torch_scatter.scatter (2.0.9).

Used by the reference at models.py:118-119,156 (QTBaseline) and, through PyG's
MessagePassing.aggregate, at models.py:208.
"""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(0, dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        s = scatter_sum(src, index, dim, out, dim_size)
        ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
        cnt = scatter_sum(ones, index, 0, None, s.size(dim)).clamp_(min=1)
        return s / _broadcast(cnt, s, dim)
    raise NotImplementedError(reduce)
