"""Import-only stand-in: GATConv is used by HetroGAT (models.py:380-506), which is out of scope."""
import torch


class GATConv(torch.nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("GATConv is outside the restated HeteroGIN path")
