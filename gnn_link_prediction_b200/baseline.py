"""Queueing-theory baseline of the reference (`QTBaseline`, models.py:42-158) on the GPU.

Same constructor and `forward(data)` contract as the reference module: `data` is the homogeneous
sample `generateFiles.process_file` stores (`edge_index`, `edge_type`, `type`) with the `P` / `L`
attributes `GNN21Dataset.preprocess` attaches (dataset.py:66-83); the result is
`(delay per path, [occupancy, rho, pi_0] per link)`, the tensors `preprocess` appends to path.x /
link.x as the `bl_features` (dataset.py:86, 105-106).  Differences: tensors come back on the GPU
(the reference pins this module to the CPU), and `forward_hetero` runs any number of samples at
once on a block-diagonal batch — there is no per-sample Python loop over hop positions.
"""
from __future__ import annotations

import torch

from . import ops


class QTBaseline(torch.nn.Module):
    def __init__(self, num_iterations=3, G_dim=4, P_dim=3, L_dim=1, **kwargs):
        super().__init__(**kwargs)
        self.num_iterations = num_iterations
        self.G_dim, self.P_dim, self.L_dim = G_dim, P_dim, L_dim
        self.H = self.H_p = self.H_l = self.H_n = 2          # models.py:50-53 (unused by forward's result)

    def forward(self, data):
        """models.py:55-158 on the reference's homogeneous `Data`."""
        dev = torch.device("cuda", torch.cuda.current_device())
        edge_index = data.edge_index.to(dev).long()
        edge_type = data.edge_type.to(dev)
        node_type = data.type.to(dev)
        is_p, is_l = node_type == 0, node_type == 1
        # local id = rank of a node inside its type (generateFiles.from_networkx numbers them that way)
        local = torch.where(is_p, torch.cumsum(is_p, 0) - 1, torch.cumsum(is_l, 0) - 1)
        sel = (edge_type == 0) & is_p[edge_index[0]]     # path->link half of the type-0 edges: the half that carries traffic
        p_l = torch.stack([local[edge_index[0, sel]], local[edge_index[1, sel]]])
        return self.forward_hetero(p_l, data.P, data.L)

    def forward_hetero(self, p_l, P, L):
        """p_l: [2,E] path->link edges (local or batch-global ids), every path's edges in route
        order; P: [n_paths, 3] = (AvgPktsLambda, PktsGen, AvgBw/1000); L: [n_links, 1] capacities."""
        dev = torch.device("cuda", torch.cuda.current_device())
        P = P.to(dev, torch.float32)
        L = L.to(dev, torch.float32).reshape(-1).contiguous()
        avg_bw = P[:, self.P_dim - 2].contiguous()        # X[:, path_og.stop - 2], models.py:94
        return ops.qt_baseline(p_l.to(dev), avg_bw, L, P.shape[0], L.shape[0], self.num_iterations)
