"""Synthetic datanet-shaped generator: edge order pinned against the reference's own graph
builder (golden fixtures; live when /root/reference is present), and the collate contract."""
import os
import sys

import pytest
import torch

from conftest import REFERENCE, ROOT, load_golden
from gnn_link_prediction_b200.data import Batch, DataLoader, EDGE_TYPES, CONV_EDGE_TYPES
from gnn_link_prediction_b200.synthetic import SyntheticDataset, Topology, make_sample


@pytest.mark.parametrize("n", [8, 12, 20])
def test_relations_match_reference_fixture(n):
    fx = load_golden(f"edges_n{n}.pt")
    topo = Topology(*fx["spec"])
    assert (topo.n_path, topo.n_link, topo.n_node) == (fx["n_path"], fx["n_link"], fx["n_node"])
    for et in EDGE_TYPES:
        assert topo.relations[et].dtype == torch.int64
        assert torch.equal(topo.relations[et], fx["relations"][et]), et


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="/root/reference not present")
@pytest.mark.parametrize("spec", [(15, 30, 11), (50, 100, 0)])
def test_relations_match_live_reference(spec):
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "pyg_shim"), REFERENCE]
    try:
        import make_golden
        topo = Topology(*spec)
        rel = make_golden.reference_relations(topo)
    finally:
        sys.path[:] = saved
    for et in EDGE_TYPES:
        assert torch.equal(topo.relations[et], rel[et]), et


def test_sample_schema():
    topo = Topology(50, 100, 0)
    s = make_sample(topo, 1997)
    assert (topo.n_path, topo.n_link, topo.n_node) == (2450, 200, 50)
    assert s["path"].x.shape == (2450, 7) and s["link"].x.shape == (200, 7) and s["node"].x.shape == (50, 3)
    assert s["path"].x.dtype == torch.float32 and bool((s["node"].x == 1).all())
    assert bool((s["path"].y > 0).all()) and float(s["path"].y.max()) <= 9.15503
    pl, lp = s[EDGE_TYPES[0]].edge_index, s[EDGE_TYPES[1]].edge_index
    assert bool((pl[0][1:] >= pl[0][:-1]).all())          # grouped by ascending source
    as_set = lambda e: set(map(tuple, e.t().tolist()))
    assert as_set(pl.flip(0)) == as_set(lp)               # l-p is the transpose edge set of p-l


def test_collate_is_block_diagonal_concat():
    ds = SyntheticDataset(3, num_nodes=9, num_links=12, num_topologies=3)
    b = Batch.from_data_list([ds[i] for i in range(3)])
    off = {nt: 0 for nt in ("path", "link", "node")}
    cols = {et: [] for et in EDGE_TYPES}
    for i in range(3):
        s = ds[i]
        for et in EDGE_TYPES:
            shift = torch.tensor([[off[et[0]]], [off[et[2]]]])
            cols[et].append(s[et].edge_index + shift)
        for nt in off:
            off[nt] += s[nt].x.shape[0]
    for et in EDGE_TYPES:
        assert torch.equal(b[et].edge_index, torch.cat(cols[et], 1))
    assert torch.equal(b["path"].x, torch.cat([ds[i]["path"].x for i in range(3)]))
    assert torch.equal(b["path"].y, torch.cat([ds[i]["path"].y for i in range(3)]))
    assert b["path"].batch.tolist() == sum(([i] * ds[i]["path"].x.shape[0] for i in range(3)), [])
    assert list(b.edge_index_dict.keys()) == list(EDGE_TYPES)
    assert set(b.x_dict.keys()) == {"path", "link", "node"}


def test_loader_len_iter_and_narrowing():
    ds = SyntheticDataset(5, num_nodes=8, num_links=9, num_topologies=2)
    dl = DataLoader(ds, batch_size=2, shuffle=False, index_dtype=torch.int32, edge_types=CONV_EDGE_TYPES)
    batches = list(dl)
    assert len(dl) == 3 and len(batches) == 3
    assert batches[-1].num_graphs == 1
    assert list(batches[0].edge_index_dict.keys()) == list(CONV_EDGE_TYPES)
    wide = Batch.from_data_list([ds[0], ds[1]])
    for et in CONV_EDGE_TYPES:
        assert batches[0][et].edge_index.dtype == torch.int32
        assert torch.equal(batches[0][et].edge_index.long(), wide[et].edge_index)


def test_batch_vector_is_derived_lazily_from_ptr():
    ds = SyntheticDataset(3, num_nodes=9, num_links=12, num_topologies=3)
    eager = Batch.from_data_list([ds[i] for i in range(3)])
    lazy = Batch.from_data_list([ds[i] for i in range(3)], batch_vector=False)
    assert "batch" not in lazy["path"]                        # nothing shipped ...
    assert torch.equal(lazy["path"].batch, eager["path"].batch)   # ... derived on first access
    assert torch.equal(lazy["node"]["batch"], eager["node"]["batch"])
    assert lazy.nbytes() > 0
