"""The oracle (oracle/hgin_oracle.py + oracle/hgin_oracle.c) against the golden vectors that
oracle/make_golden.py recorded from the UNMODIFIED reference, and against the live reference
when /root/reference is present.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import MODEL_CASES, REFERENCE, ROOT, build_model, config_to_kwargs, load_golden
from oracle import c_oracle, hgin_oracle


def _oracle_model(fx):
    in_ch = {k: v.shape[1] for k, v in fx["x_dict"].items()}
    m = build_model(hgin_oracle, fx["config"], in_ch)
    m.load_state_dict(fx["state_dict"])
    m.train()
    return m


class _B:  # minimal batch view for hgin_oracle.train_step
    def __init__(self, fx):
        self.x_dict = fx["x_dict"]
        self.edge_index_dict = fx["edge_index_dict"]
        self._p = {"batch": fx["path_batch"], "y": fx["y"]}

    def __getitem__(self, k):
        assert k == "path"
        return type("S", (), self._p)


@pytest.mark.parametrize("case", MODEL_CASES)
def test_port_matches_reference_fixture_bit_exact(case):
    fx = load_golden(f"model_{case}.pt")
    m = _oracle_model(fx)
    out = m(fx["x_dict"], fx["edge_index_dict"], fx["path_batch"])
    loss_value = hgin_oracle.mape(out, fx["y"].reshape(-1, 1))
    torch.sqrt(loss_value).backward()
    assert torch.equal(out, fx["out"])
    assert torch.equal(loss_value, fx["loss_value"])
    named = dict(m.named_parameters())
    assert set(named) == set(fx["grads"])
    for k, g in fx["grads"].items():
        if g is None:
            assert named[k].grad is None, k
        else:
            assert torch.equal(named[k].grad, g), k


@pytest.mark.parametrize("case", MODEL_CASES)
def test_port_adam_trajectory(case):
    fx = load_golden(f"model_{case}.pt")
    m = _oracle_model(fx)
    cfg = fx["config"]
    opt = torch.optim.Adam(m.parameters(), lr=cfg["LEARNING_RATE"], weight_decay=cfg["WEIGHT_DECAY"])
    b = _B(fx)
    losses = [float(hgin_oracle.train_step(m, opt, b)[0]) for _ in fx["losses"]]
    assert losses == fx["losses"]
    for k, v in m.state_dict().items():
        assert torch.equal(v, fx["final_state_dict"][k]), k


@pytest.mark.parametrize("case", ["default", "L3_emb16"])
def test_dense_fp64_cross_check(case):
    """Second, independent restatement (dense adjacency matmul in fp64)."""
    fx = load_golden(f"model_{case}.pt")
    m = _oracle_model(fx)
    ref = hgin_oracle.dense_forward_fp64(m, fx["x_dict"], fx["edge_index_dict"])
    torch.testing.assert_close(fx["out"].double(), ref, rtol=2e-4, atol=2e-5)


def test_state_dict_keys_are_the_reference_keys():
    fx = load_golden("model_default.pt")
    m = _oracle_model(fx)
    assert list(m.state_dict().keys()) == list(fx["state_dict"].keys())


@pytest.mark.parametrize("shape", [(1000, 300, 20000, 128), (50, 2000, 7000, 3), (500, 500, 60000, 8),
                                   (10, 5, 0, 4), (1, 1, 9, 5)])
def test_c_oracle_bit_exact_with_aten_cpu(shape):
    ns, nd, e, f = shape
    g = torch.Generator().manual_seed(e + f)
    ei = torch.stack([torch.randint(0, ns, (e,), generator=g), torch.randint(0, nd, (e,), generator=g)])
    rowptr, col, perm = c_oracle.csr_build(ei.numpy(), ns, nd)
    p = torch.argsort(ei[1], stable=True)
    rp = torch.zeros(nd + 1, dtype=torch.int64)
    rp[1:] = torch.cumsum(torch.bincount(ei[1], minlength=nd), 0)
    assert np.array_equal(rowptr, rp.numpy().astype(np.int32))
    assert np.array_equal(col, ei[0][p].numpy().astype(np.int32))
    assert np.array_equal(perm, p.numpy().astype(np.int32))
    xs = torch.randn(ns, f, generator=g).requires_grad_(True)
    xd = torch.randn(nd, f, generator=g)
    eps = torch.tensor([0.37])
    msg = xs.index_select(0, ei[0])
    agg = torch.zeros(nd, f).scatter_add_(0, ei[1].view(-1, 1).expand_as(msg), msg)
    h_cat = torch.cat((agg, (1 + eps) * xd), 1)
    h_add = agg.clone()
    h_add += (1 + eps) * xd
    xs_np = xs.detach().numpy()
    assert np.array_equal(c_oracle.gin_combine(rowptr, col, xs_np, xd.numpy(), 0.37, True), h_cat.detach().numpy())
    assert np.array_equal(c_oracle.gin_combine(rowptr, col, xs_np, xd.numpy(), 0.37, False), h_add.detach().numpy())
    gout = torch.randn(nd, f, generator=g)
    agg.backward(gout)
    rpt, colt, _ = c_oracle.csr_build(ei.flip(0).contiguous().numpy(), nd, ns)
    assert np.array_equal(c_oracle.gather_t(rpt, colt, gout.numpy()), xs.grad.numpy())


def test_c_oracle_rejects_out_of_range():
    ei = np.array([[0, 5], [0, 1]], dtype=np.int64)
    with pytest.raises(ValueError):
        c_oracle.csr_build(ei, 3, 2)


# ---- live reference (build container only) ------------------------------------------------------
needs_ref = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref_models():
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(ROOT, "oracle", "pyg_shim"), REFERENCE]
    try:
        import models
        yield models
    finally:
        sys.path[:] = saved


@needs_ref
@pytest.mark.parametrize("emb,layers", [(8, 1), (24, 2), (128, 4)])
def test_port_matches_live_reference(ref_models, emb, layers):
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.synthetic import Topology, make_sample
    samples = [make_sample(Topology(9 + i, 12 + 2 * i, i), seed=100 + i) for i in range(3)]
    batch = Batch.from_data_list(samples)
    kw = dict(node_embedding_size=emb, message_passing_layers=layers, dropout=0.0, concat_path=True,
              bl_features=False, divided_features=False, global_feats=False, mlp_layers=[128, 32],
              act="torch.nn.PReLU()", mlp_head_act=None, mlp_bn=False)
    torch.manual_seed(5)
    ref = ref_models.HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **kw)
    mine = hgin_oracle.HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **kw)
    mine.load_state_dict(ref.state_dict())
    y = batch["path"].y.reshape(-1, 1)
    o_ref = ref(batch.x_dict, batch.edge_index_dict, batch["path"].batch)
    o_mine = mine(batch.x_dict, batch.edge_index_dict, batch["path"].batch)
    assert torch.equal(o_ref, o_mine)
    torch.sqrt(hgin_oracle.mape(o_ref, y)).backward()
    torch.sqrt(hgin_oracle.mape(o_mine, y)).backward()
    g_ref = {k: p.grad for k, p in ref.named_parameters()}
    for k, p in mine.named_parameters():
        assert (p.grad is None) == (g_ref[k] is None), k
        if p.grad is not None:
            assert torch.equal(p.grad, g_ref[k]), k
