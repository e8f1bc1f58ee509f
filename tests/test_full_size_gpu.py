"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run
these sizes in seconds): Cfg-C = 1024 topologies per step, hidden 128, 4 GIN layers, tf32 GEMMs;
Cfg-D = 1 M nodes / 20 M edges.  Integer work is checked exactly (sortedness, transposition,
agreement of the two CSR constructions); the fp32 aggregation through inputs on which fp32
addition is exact (small integers), so any summation order must give the same bits as a float64
index_add; the model through batch-composition invariance (samples are independent components of
the batched graph, so a sample's scores cannot depend on what else is in the batch) and
run-to-run determinism of whole train steps."""
import pytest
import torch

from gnn_link_prediction_b200 import ops
from gnn_link_prediction_b200.arena import DeviceDataset, SampleArena
from gnn_link_prediction_b200.data import CONV_EDGE_TYPES
from gnn_link_prediction_b200.synthetic import SyntheticDataset

pytestmark = pytest.mark.gpu

KW = dict(node_embedding_size=128, message_passing_layers=4, dropout=0.0, concat_path=True, bl_features=False,
          divided_features=False, global_feats=False, mlp_layers=[128, 32], act="torch.nn.PReLU()", mlp_head_act=None,
          mlp_bn=False)


@pytest.fixture(scope="module")
def cfgc():
    ds = SyntheticDataset(1024, num_topologies=16, seed=1997)
    samples = [ds[i] for i in range(1024)]
    dev = DeviceDataset(SampleArena.from_samples(samples, keep_coo=False))
    return dev, dev.collate(list(range(1024)))


def _expand(rowptr):
    counts = (rowptr[1:] - rowptr[:-1]).long()
    return torch.repeat_interleave(torch.arange(counts.numel(), device=rowptr.device), counts)


def test_cfgc_adjacency_properties(cfgc):
    _, batch = cfgc
    n = {nt: batch[nt]["x"].shape[0] for nt in ("path", "link", "node")}
    assert (n["path"], n["link"], n["node"]) == (1024 * 2450, 1024 * 200, 1024 * 50)
    for et in CONV_EDGE_TYPES:
        st = batch[et]
        e = st["csr_dst_col"].shape[0]
        for side, rows_t, cols_t in (("dst", et[2], et[0]), ("src", et[0], et[2])):
            rp, col = st[f"csr_{side}_rowptr"], st[f"csr_{side}_col"]
            assert rp.shape[0] == n[rows_t] + 1 and int(rp[0]) == 0 and int(rp[-1]) == e
            assert bool((rp[1:] >= rp[:-1]).all())                                 # sorted rows
            assert int(col.min()) >= 0 and int(col.max()) < n[cols_t]
        # the two orientations hold the same edge multiset
        src_a, dst_a = st["csr_dst_col"].long(), _expand(st["csr_dst_rowptr"])
        src_b, dst_b = _expand(st["csr_src_rowptr"]), st["csr_src_col"].long()
        key_a = torch.sort(src_a * n[et[2]] + dst_a)[0]
        key_b = torch.sort(src_b * n[et[2]] + dst_b)[0]
        assert torch.equal(key_a, key_b)
        # K0 on the COO list in the reference's order (grouped by source = the by-source CSR order)
        # reproduces the collated destination-sorted CSR bit for bit, and vice versa
        coo = torch.stack([src_b, dst_b]).to(torch.int32)
        k0 = ops.csr_build(coo, n[et[0]], n[et[2]], by="dst").validate()
        assert torch.equal(k0.rowptr, st["csr_dst_rowptr"]) and torch.equal(k0.col, st["csr_dst_col"])
        k0t = ops.csr_build(coo, n[et[0]], n[et[2]], by="src").validate()
        assert torch.equal(k0t.rowptr, st["csr_src_rowptr"]) and torch.equal(k0t.col, st["csr_src_col"])
        # samples are separate components: no edge crosses a sample boundary
        ps, pd = batch[et[0]]["ptr"], batch[et[2]]["ptr"]
        assert torch.equal(torch.bucketize(src_b, ps, right=True), torch.bucketize(dst_b, pd, right=True))


@pytest.mark.parametrize("et", [("link", "includes", "path"), ("path", "uses", "link")])
def test_cfgc_aggregation_is_exact_on_integer_features(cfgc, et):
    _, batch = cfgc
    st = batch[et]
    ns, nd = batch[et[0]]["x"].shape[0], batch[et[2]]["x"].shape[0]
    g = torch.Generator(device="cuda").manual_seed(5)
    x_src = torch.randint(-7, 8, (ns, 128), generator=g, device="cuda").float()
    x_dst = torch.randint(-7, 8, (nd, 128), generator=g, device="cuda").float()
    csr = ops.CSR(st["csr_dst_rowptr"], st["csr_dst_col"], None, None, nd, ns, st["csr_dst_col"].shape[0])
    eps = torch.tensor([1.0], device="cuda")                       # (1 + eps) = 2: still exact
    got = ops.gin_combine(csr, x_src, x_dst, eps, ops.SELF_ADD)
    dst = _expand(st["csr_dst_rowptr"])
    want = 2.0 * x_dst
    for c0 in range(0, 128, 32):                                   # float32 index_add is exact on these inputs too
        want[:, c0:c0 + 32].index_add_(0, dst, x_src[st["csr_dst_col"].long(), c0:c0 + 32])
    assert torch.equal(got, want)
    deg = ops.gin_combine(csr, torch.ones(ns, 4, device="cuda"))
    assert torch.equal(deg[:, 0], (st["csr_dst_rowptr"][1:] - st["csr_dst_rowptr"][:-1]).float())
    # transposed pass (backward gather) on the same data: column sums agree exactly
    csr_t = ops.CSR(st["csr_src_rowptr"], st["csr_src_col"], None, None, ns, nd, st["csr_src_col"].shape[0])
    back = ops.gin_combine(csr_t, x_dst)
    assert float(back.double().sum()) == float((x_dst.double() * deg[:, :1].double()).sum())


def test_cfgc_scores_do_not_depend_on_batch_composition(cfgc):
    from gnn_link_prediction_b200.models import HetroGIN, MATH_TF32
    dev, batch = cfgc
    torch.manual_seed(3)
    model = HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **KW).cuda().eval().set_math_mode(MATH_TF32)
    with torch.no_grad():
        full = model(batch.x_dict, batch.graph, None)
        assert full.shape == (1024 * 2450, 1) and bool(torch.isfinite(full).all())
        for ids in ([0, 1, 2], [517, 518, 519, 3], list(range(1000, 1024))):   # >= 128 rows per type: same kernels as the full batch
            sub = dev.collate(ids)
            out = model(sub.x_dict, sub.graph, None)
            ptr = batch["path"]["ptr"]
            rows = torch.cat([full[int(ptr[i]):int(ptr[i + 1])] for i in ids])
            assert torch.equal(out, rows)


def test_cfgc_train_steps_are_deterministic_and_decrease_the_loss(cfgc):
    from gnn_link_prediction_b200.models import HetroGIN, MATH_TF32
    from gnn_link_prediction_b200.train import TrainStep
    dev, batch = cfgc
    runs = []
    for _ in range(2):
        torch.manual_seed(11)
        model = HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **KW).cuda().train().set_math_mode(MATH_TF32)
        step = TrainStep(model, lr=1e-3)
        losses = torch.stack([step(batch).clone() for _ in range(3)])
        runs.append((losses.cpu(), step.flat_p.clone().cpu()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    assert bool(torch.isfinite(runs[0][0]).all()) and float(runs[0][0][-1, 1]) < float(runs[0][0][0, 1])


def test_cfgd_sweep_graph_aggregation_exact_and_transposed():
    """Cfg-D: 900 k path / 80 k link nodes, 9.96 M uniform random edges per direction, F = 128."""
    g = torch.Generator(device="cuda").manual_seed(0)
    n_path, n_link, e = 900_000, 80_000, 9_960_000
    ei = torch.stack([torch.randint(0, n_path, (e,), generator=g, device="cuda"),
                      torch.randint(0, n_link, (e,), generator=g, device="cuda")])
    fwd = ops.csr_build(ei, n_path, n_link, by="dst").validate()
    bwd = ops.csr_build(ei, n_path, n_link, by="src").validate()
    assert int(fwd.rowptr[-1]) == e and int(bwd.rowptr[-1]) == e
    # stable: inside every destination row the sources appear in edge order -> perm is increasing per row
    fwd_p = ops.csr_build(ei, n_path, n_link, by="dst", want_perm=True)
    rows = _expand(fwd_p.rowptr)
    p = fwd_p.perm.long()
    assert torch.equal(ei[1][p], rows) and torch.equal(ei[0][p].to(torch.int32), fwd_p.col)
    same_row = rows[1:] == rows[:-1]
    assert bool((p[1:][same_row] > p[:-1][same_row]).all())
    x = torch.randint(-3, 4, (n_path, 128), generator=g, device="cuda").float()
    got = ops.gin_combine(fwd, x)
    want = torch.zeros(n_link, 128, device="cuda")
    for c0 in range(0, 128, 16):
        want[:, c0:c0 + 16].index_add_(0, ei[1], x[ei[0], c0:c0 + 16])
    assert torch.equal(got, want)
    gl = torch.randint(-3, 4, (n_link, 128), generator=g, device="cuda").float()
    got_t = ops.gin_combine(bwd, gl)
    want_t = torch.zeros(n_path, 128, device="cuda")
    for c0 in range(0, 128, 16):
        want_t[:, c0:c0 + 16].index_add_(0, ei[0], gl[ei[1], c0:c0 + 16])
    assert torch.equal(got_t, want_t)
    # <A x, g> == <x, A^T g> exactly on integers (adjoint identity of forward and backward gathers)
    assert float((got.double() * gl.double()).sum()) == float((x.double() * got_t.double()).sum())
