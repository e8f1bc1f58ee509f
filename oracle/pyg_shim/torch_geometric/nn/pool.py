"""global_{mean,max}_pool, used by the reference only when GLOBAL_FEATS (models.py:347-349)."""
import torch


def global_mean_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    out = torch.zeros(size, x.size(1), dtype=x.dtype, device=x.device)
    out.scatter_add_(0, batch.view(-1, 1).expand_as(x), x)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device).scatter_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp_(min=1).view(-1, 1)


def global_max_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    out = torch.full((size, x.size(1)), float("-inf"), dtype=x.dtype, device=x.device)
    return out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax", include_self=True)
