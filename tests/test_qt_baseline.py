"""Queueing-theory baseline: the CPU oracle against the fixtures recorded from the
unmodified reference (and against the live reference when present); the CUDA path (`-m gpu`)
against the same fixtures, against the oracle on larger random inputs, and batched against
per-sample runs.  fp32 tolerance: rel 1e-5 (powf differs in the last ulp between libm and CUDA),
NaN patterns must agree exactly."""
import os
import sys

import pytest
import torch

from conftest import REFERENCE, ROOT, load_golden
from oracle import qt_oracle

CASES = ["n12", "n8_light", "n20_heavy", "n8_overflow"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_fixture_bit_for_bit(case):
    f = load_golden(f"qt_{case}.pt")
    p_l, n_p, n_l = qt_oracle.hetero_view(f["edge_index"], f["edge_type"], f["type"])
    assert torch.equal(p_l, f["p_l"]) and (n_p, n_l) == (f["P"].shape[0], f["L"].shape[0])
    delay, links = qt_oracle.qt_baseline(p_l, f["P"], f["L"])
    assert torch.allclose(delay, f["out_paths"], rtol=0, atol=0, equal_nan=True)
    assert torch.allclose(links, f["out_links"], rtol=0, atol=0, equal_nan=True)
    if case == "n8_overflow":
        assert bool(torch.isnan(f["out_links"]).any())     # rho^33 overflows fp32: the reference yields NaNs, pinned too


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="/root/reference not present")
def test_oracle_matches_live_reference():
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import make_golden_qt as gen
    data = gen.reference_data((10, 14, 4), seed=7, load=4.0)
    data.P, data.L = data.P.float(), data.L.float()
    torch.clip = gen._clip_keeping_integer_dtype            # torch<=1.9 clamp semantics the reference relies on
    try:
        want_p, want_l = gen.ref_models.QTBaseline()(data)
    finally:
        torch.clip = gen._ORIG_CLIP
    p_l, _, _ = qt_oracle.hetero_view(data.edge_index, data.edge_type, data.type)
    delay, links = qt_oracle.qt_baseline(p_l, data.P, data.L)
    assert torch.equal(delay, want_p) and torch.equal(links, want_l)


class _Data:
    def __init__(self, f):
        self.edge_index, self.edge_type, self.type, self.P, self.L = f["edge_index"], f["edge_type"], f["type"], f["P"], f["L"]


def _close(got, want, rtol=1e-5):
    got = got.cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    ok = ~torch.isnan(want)
    torch.testing.assert_close(got[ok], want[ok], rtol=rtol, atol=rtol * float(want[ok].abs().max()) if ok.any() else 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_baseline_matches_reference_fixture(case):
    from gnn_link_prediction_b200.baseline import QTBaseline
    f = load_golden(f"qt_{case}.pt")
    delay, links = QTBaseline()(_Data(f))                  # the reference's forward(data) contract
    assert delay.is_cuda and delay.shape == f["out_paths"].shape and links.shape == f["out_links"].shape
    _close(delay, f["out_paths"])
    _close(links, f["out_links"])


@pytest.mark.gpu
def test_cuda_baseline_batched_equals_per_sample_and_oracle():
    from gnn_link_prediction_b200.baseline import QTBaseline
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    ds = SyntheticDataset(6, num_nodes=16, num_links=30, num_topologies=3, seed=2)
    samples = [ds[i] for i in range(6)]
    g = torch.Generator().manual_seed(0)
    et = ("path", "uses", "link")
    P = [torch.rand(s["path"]["x"].shape[0], 3, generator=g) * 3 + 0.1 for s in samples]
    L = [torch.rand(s["link"]["x"].shape[0], 1, generator=g) * 30000 + 8000 for s in samples]
    qt = QTBaseline()
    per_sample = [qt.forward_hetero(s[et]["edge_index"], p, l) for s, p, l in zip(samples, P, L)]
    batch = Batch.from_data_list(samples)
    delay, links = qt.forward_hetero(batch[et]["edge_index"], torch.cat(P), torch.cat(L))
    assert torch.equal(delay, torch.cat([d for d, _ in per_sample]))       # samples are independent components
    assert torch.equal(links, torch.cat([l for _, l in per_sample]))
    for s, p, l, (d_gpu, l_gpu) in zip(samples, P, L, per_sample):
        d_ref, l_ref = qt_oracle.qt_baseline(s[et]["edge_index"], p, l)
        _close(d_gpu, d_ref)
        _close(l_gpu, l_ref)
    # deterministic (no atomics), and a different iteration count is honoured
    again = qt.forward_hetero(batch[et]["edge_index"], torch.cat(P), torch.cat(L))
    assert torch.equal(again[0], delay) and torch.equal(again[1], links)
    d1, l1 = QTBaseline(num_iterations=1).forward_hetero(samples[0][et]["edge_index"], P[0], L[0])
    d1_ref, l1_ref = qt_oracle.qt_baseline(samples[0][et]["edge_index"], P[0], L[0], num_iterations=1)
    _close(d1, d1_ref)
    _close(l1, l1_ref)


@pytest.mark.gpu
def test_cuda_baseline_accepts_edges_in_any_order():
    """Route order inside a path is what matters; the edge list itself may be permuted across paths."""
    from gnn_link_prediction_b200.baseline import QTBaseline
    f = load_golden("qt_n12.pt")
    p_l = f["p_l"]
    # interleave the paths' edges while keeping each path's own edges in order (stable shuffle by a random key per path)
    g = torch.Generator().manual_seed(3)
    first = torch.ones(p_l.shape[1], dtype=torch.bool)
    first[1:] = p_l[0, 1:] != p_l[0, :-1]
    pos = torch.arange(p_l.shape[1]) - torch.cummax(torch.where(first, torch.arange(p_l.shape[1]), torch.zeros(1, dtype=torch.long)), 0)[0]
    order = torch.argsort(pos * 1000 + torch.randperm(1000, generator=g)[p_l[0] % 1000], stable=True)
    delay, links = QTBaseline().forward_hetero(p_l[:, order], f["P"], f["L"])
    _close(delay, f["out_paths"])
    _close(links, f["out_links"])
