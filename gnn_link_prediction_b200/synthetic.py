"""Seed-pinned synthetic "datanet-shaped" samples.

There is no network in the build/bench environment, so the GNNet-Challenge-2021 dataset
(downloadDataset.py:5-8) cannot be fetched.  This module fabricates samples with the same
SCHEMA and the same EDGE ORDER the reference pipeline produces:

* topology -> tri-partite graph: restates `simulation_to_networkX` (generateFiles.py:21-99):
  nodes `n_i`, links `l_a_b` (ids in first-appearance order: a route may name `l_a_b` before the
  `(a,b)` iteration adds it, generateFiles.py:77 vs :44), paths `p_s_d_f` in `(s,d)` loop order;
* per-relation COO lists: restates `from_networkx` (generateFiles.py:102-190): `G.edges` iterates
  sources in node-insertion order and successors in insertion order, so every relation comes out
  grouped by ascending source id with destinations in insertion order;
* feature schema: `preprocess` (dataset.py:89-117): `path.x f32[Np,7]`, `link.x f32[Nl,7]`,
  `node.x = ones[Nn,3]`, `path.y` = delay label (>0), six relations.

`tests/test_synthetic_vs_reference.py` runs the reference's own two functions on the same
fabricated topology/routing (container only) and compares every relation with `torch.equal`;
`tests/golden/edges_n12.pt` pins the result for the GPU box.
"""
from __future__ import annotations

import numpy as np
import torch

from .data import HeteroData, EDGE_TYPES

Y_MAX = 9.15503  # label range noted at dataset.py:55-56


def random_topology(num_nodes: int, num_links: int, seed: int):
    """Connected undirected graph with `num_links` edges: random spanning tree + random chords.
    Returns a sorted adjacency list."""
    rs = np.random.RandomState(seed)
    assert num_links >= num_nodes - 1
    assert num_links <= num_nodes * (num_nodes - 1) // 2
    label = rs.permutation(num_nodes)
    edges = set()
    for i in range(1, num_nodes):
        j = int(rs.randint(0, i))
        a, b = int(label[i]), int(label[j])
        edges.add((min(a, b), max(a, b)))
    while len(edges) < num_links:
        a, b = (int(v) for v in rs.randint(0, num_nodes, size=2))
        if a != b:
            edges.add((min(a, b), max(a, b)))
    adj = [[] for _ in range(num_nodes)]
    for a, b in sorted(edges):
        adj[a].append(b)
        adj[b].append(a)
    return [sorted(v) for v in adj]


def shortest_path_routing(adj):
    """routes[s][d] = node list of one hop-count-shortest path s -> d (BFS, lowest-index parent)."""
    n = len(adj)
    routes = [[None] * n for _ in range(n)]
    for s in range(n):
        parent = [-1] * n
        parent[s] = s
        frontier = [s]
        while frontier:
            nxt = []
            for u in frontier:
                for v in adj[u]:
                    if parent[v] < 0:
                        parent[v] = u
                        nxt.append(v)
            frontier = nxt
        for d in range(n):
            if d == s:
                continue
            assert parent[d] >= 0, "topology must be connected"
            hops = [d]
            while hops[-1] != s:
                hops.append(parent[hops[-1]])
            routes[s][d] = hops[::-1]
    return routes


def build_relations(adj, routes):
    """The six COO relations of one sample, int64 `[2,E]`, ids and order as the reference emits
    them (see module docstring).  Returns (edge_index dict, n_path, n_link, n_node)."""
    n = len(adj)
    has_edge = [set(v) for v in adj]
    link_id = {}          # (a,b) -> id, first-appearance order
    link_succ_n = {}      # link id -> dst node (set when the (a,b) iteration is reached)
    link_succ_p = {}      # link id -> [path ids] in insertion order
    n_succ_l = [[] for _ in range(n)]
    n_succ_p = [[] for _ in range(n)]
    p_succ_l, p_succ_n = [], []

    def lid(a, b):
        k = link_id.get((a, b))
        if k is None:
            k = link_id[(a, b)] = len(link_id)
            link_succ_p[k] = []
        return k

    for s in range(n):
        for d in range(n):
            if s == d:
                continue
            if d in has_edge[s]:
                k = lid(s, d)
                n_succ_l[s].append(k)
                link_succ_n[k] = d
            route = routes[s][d]
            if route is None:
                continue
            p = len(p_succ_l)
            links, nodes, seen_l, seen_n = [], [], set(), set()
            for h1, h2 in zip(route[:-1], route[1:]):
                for v in (h1, h2):
                    if v not in seen_n:
                        seen_n.add(v)
                        nodes.append(v)
                        n_succ_p[v].append(p)
                k = lid(h1, h2)
                if k not in seen_l:  # DiGraph: a repeated edge is not duplicated
                    seen_l.add(k)
                    links.append(k)
                    link_succ_p[k].append(p)
            p_succ_l.append(links)
            p_succ_n.append(nodes)

    def coo(pairs):
        if not pairs:
            return torch.zeros(2, 0, dtype=torch.int64)
        return torch.tensor(pairs, dtype=torch.int64).t().contiguous()

    n_link = len(link_id)
    rel = {
        ("path", "uses", "link"): coo([(p, k) for p, ks in enumerate(p_succ_l) for k in ks]),
        ("link", "includes", "path"): coo([(k, p) for k in range(n_link) for p in link_succ_p[k]]),
        ("link", "connects", "node"): coo([(k, link_succ_n[k]) for k in range(n_link) if k in link_succ_n]),
        ("node", "has", "link"): coo([(v, k) for v in range(n) for k in n_succ_l[v]]),
        ("path", "is_connected", "node"): coo([(p, v) for p, vs in enumerate(p_succ_n) for v in vs]),
        ("node", "is_used", "path"): coo([(v, p) for v in range(n) for p in n_succ_p[v]]),
    }
    return rel, len(p_succ_l), n_link, n


class Topology:
    """Edge structure of one sample, reusable across samples with different features."""

    def __init__(self, num_nodes=50, num_links=100, seed=0):
        self.adj = random_topology(num_nodes, num_links, seed)
        self.routes = shortest_path_routing(self.adj)
        self.relations, self.n_path, self.n_link, self.n_node = build_relations(self.adj, self.routes)


def make_sample(topology: Topology, seed: int) -> HeteroData:
    """One HeteroData with the schema of dataset.py:89-117.  Features ~ N(0,1) (the reference
    always standardises, dataset.py:165), `node.x = 1` (dataset.py:102), `y ~ U(0.1, Y_MAX)`."""
    g = torch.Generator().manual_seed(seed)
    d = HeteroData()
    d["link"].x = torch.randn(topology.n_link, 7, generator=g)
    d["path"].x = torch.randn(topology.n_path, 7, generator=g)
    d["node"].x = torch.ones(topology.n_node, 3)
    d["path"].y = 0.1 + (Y_MAX - 0.1) * torch.rand(topology.n_path, generator=g)
    for et in EDGE_TYPES:
        d[et].edge_index = topology.relations[et]
    return d


class SyntheticDataset:
    """`num_samples` samples drawn over `num_topologies` distinct 50-node/100-link topologies.
    Seeds follow config.json's SEED=1997 (config.json:2)."""

    def __init__(self, num_samples, num_nodes=50, num_links=100, num_topologies=8, seed=1997):
        self.topologies = [Topology(num_nodes, num_links, seed + 7919 * t)
                           for t in range(min(num_topologies, num_samples))]
        self.num_samples = num_samples
        self.seed = seed
        self._cache = {}

    def __len__(self):
        return self.num_samples

    def __getitem__(self, i):
        if i < 0 or i >= self.num_samples:
            raise IndexError(i)
        if i not in self._cache:
            self._cache[i] = make_sample(self.topologies[i % len(self.topologies)], self.seed + i)
        return self._cache[i]
