"""`to_dense_adj` is imported by the reference (models.py:8) but only used by the dead
function `compute_identity` (models.py:170-177)."""
import torch


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
    n = int(edge_index.max()) + 1 if max_num_nodes is None else max_num_nodes
    adj = torch.zeros(1, n, n)
    adj[0].index_put_((edge_index[0], edge_index[1]), torch.ones(edge_index.size(1)), accumulate=True)
    return adj
