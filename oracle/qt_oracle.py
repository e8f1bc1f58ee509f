"""TEST INFRASTRUCTURE — CPU restatement of the queueing-theory baseline.  NOT part of the product.

Follows `QTBaseline.forward` and `separate_edge_timesteps` (models.py:15-158) in plain PyTorch
CPU ops, on the hetero view of the sample: only the path->link edges of `edge_type == 0` carry
traffic (the link->path edges of the same type have A = 0 at their sources, models.py:94, so
they add nothing to T at link rows, the only rows that are read).  Pinned by tests/golden/qt_*.pt,
which `oracle/make_golden_qt.py` records from the unmodified reference, and against the live
reference when /root/reference is present (tests/test_oracle.py).  `torch_scatter.scatter(sum)`
on the CPU is `scatter_add_`, i.e. a sequential sum in edge order.
"""
import torch

BUFFER = 32  # models.py:124


def hetero_view(edge_index, edge_type, node_type):
    """(path->link COO in local ids, n_paths, n_links) from the homogeneous graph the reference
    stores (generateFiles.py:186-190, 227-229): local id = rank of the node inside its type."""
    edge_index = edge_index.long()
    is_p, is_l = node_type == 0, node_type == 1
    local = torch.zeros_like(node_type, dtype=torch.long)
    local[is_p] = torch.arange(int(is_p.sum()))
    local[is_l] = torch.arange(int(is_l.sum()))
    sel = (edge_type == 0) & is_p[edge_index[0]]
    src, dst = edge_index[0, sel], edge_index[1, sel]
    return torch.stack([local[src], local[dst]]), int(is_p.sum()), int(is_l.sum())


def qt_baseline(p_l, P, L, num_iterations=3):
    """p_l: int64 [2,E] path->link edges grouped by path in route order; P: f32 [n_p,3]
    (AvgPktsLambda, PktsGen, AvgBw/1000); L: f32 [n_l,1] capacities.
    Returns (path_delay f32 [n_p], link_out f32 [n_l,3] = [occupancy, rho, pi_0])."""
    n_p, n_l = P.shape[0], L.shape[0]
    paths, links = p_l[0], p_l[1]
    # hop position of every edge inside its path (separate_edge_timesteps, models.py:15-39)
    first = torch.ones_like(paths, dtype=torch.bool)
    first[1:] = paths[1:] != paths[:-1]
    start = torch.cummax(torch.where(first, torch.arange(paths.numel()), torch.zeros_like(paths)), 0)[0]
    pos = torch.arange(paths.numel()) - start
    max_pos = int(pos.max()) + 1 if pos.numel() else 0
    A = P[:, 1].clone()                                    # X[:, path_og.stop - 2], models.py:94
    cap = (L / 1000).view(-1)                              # models.py:73-74
    blocking = 0.5 * torch.ones(n_l)                       # models.py:96
    for _ in range(num_iterations):
        T = torch.zeros(n_l)
        traffic = A.clone()
        for k in range(max_pos):                           # update_traffic, models.py:103-120
            if k > 0:
                prev = pos == k - 1
                traffic[paths[prev]] *= (1.0 - blocking[links[prev]])
            cur = pos == k
            T += torch.zeros(n_l).scatter_add_(0, links[cur], traffic[paths[cur]])
        rho = T / cap                                      # update_blocking_probs, models.py:126-134
        num = (1.0 - rho) * torch.pow(rho, BUFFER)
        den = 1.0 - torch.pow(rho, BUFFER + 1)
        blocking = num / (den + 1e-08)
    pi_0 = (1 - rho) / (1 - torch.pow(rho, BUFFER + 1))    # models.py:141-148
    res = 1 * pi_0
    for j in range(32):
        pi_0 = pi_0 * rho
        res = res + (j + 1) * pi_0
    res = res / 32
    x_link = res * 32000.0 / L.view(-1)                    # models.py:153-155
    delay = torch.zeros(n_p).scatter_add_(0, paths, x_link[links])
    return delay, torch.stack([res, rho, pi_0], 1)
