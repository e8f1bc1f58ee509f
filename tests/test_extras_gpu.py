"""Kernels behind the non-default branches of HetroGIN (models.py:301-330, 347-371) against the torch CPU modules the
reference would run: activations (`eval(act)`), BatchNorm1d (`mlp_bn`), dropout, global mean / max pools."""
import pytest
import torch

from gnn_link_prediction_b200 import functional as F_
from gnn_link_prediction_b200 import ops
from oracle import hgin_oracle

pytestmark = pytest.mark.gpu

ACTS = ["PReLU()", "ReLU()", "LeakyReLU(0.1)", "LeakyReLU()", "ELU()", "ELU(0.5)", "Sigmoid()", "Tanh()", "GELU()", "SiLU()",
        "Softplus()", "Softplus(beta=2.0, threshold=5.0)", "Identity()"]


@pytest.mark.parametrize("shape", [(1000, 32), (257, 7), (64, 1)])
@pytest.mark.parametrize("act", ACTS)
def test_activation_forward_backward_against_torch(act, shape):
    torch.manual_seed(0)
    mod = eval("torch.nn." + act)
    z = (3 * torch.randn(*shape)).requires_grad_(True)
    z.data[0, 0] = 30.0          # beyond the Softplus threshold
    g = torch.randn(*shape)
    out_ref = mod(z)
    out_ref.backward(g)
    spec = F_.act_spec_of(mod)
    alpha = None if spec.alpha is None else spec.alpha.detach().cuda()
    zc = z.detach().cuda()
    out = ops.act_fwd(zc, spec.code, alpha, spec.p0, spec.p1)
    torch.testing.assert_close(out.cpu(), out_ref.detach(), rtol=2e-6, atol=2e-6)   # (1 + erf cancels in the GELU tail)
    dz, dalpha = ops.act_bwd(g.cuda(), zc, spec.code, alpha, spec.p0, spec.p1, want_dalpha=alpha is not None)
    torch.testing.assert_close(dz.cpu(), z.grad, rtol=1e-5, atol=1e-6)
    if alpha is not None:
        torch.testing.assert_close(dalpha.cpu(), mod.weight.grad, rtol=1e-4, atol=1e-5)


def test_activation_bf16_rows_and_strided_views():
    torch.manual_seed(1)
    z = torch.randn(300, 48, device="cuda")
    zb = z.bfloat16()
    out = ops.act_fwd(zb, ops.ACT_GELU)
    assert out.dtype == torch.bfloat16
    torch.testing.assert_close(out.float(), torch.nn.functional.gelu(zb.float()), rtol=8e-3, atol=1e-3)
    view = z[:, 4:20]            # leading dimension 48, 16 columns
    torch.testing.assert_close(ops.act_fwd(view, ops.ACT_TANH), torch.tanh(view), rtol=2e-6, atol=1e-6)
    odd = z[:, 1:8]              # misaligned for 16-byte accesses: scalar path
    torch.testing.assert_close(ops.act_fwd(odd, ops.ACT_SILU), torch.nn.functional.silu(odd), rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("rows,n", [(4096, 32), (1000, 128), (77, 5), (3000, 300)])
@pytest.mark.parametrize("act", ["PReLU()", "LeakyReLU(0.1)", "Identity()"])
def test_batchnorm_act_training_against_torch(rows, n, act):
    """Linear output -> BatchNorm1d (batch statistics, running buffers updated) -> activation, forward and backward."""
    torch.manual_seed(2)
    bn_ref = torch.nn.BatchNorm1d(n)
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.uniform_(-0.5, 0.5)
    act_ref = eval("torch.nn." + act)
    z = (2 * torch.randn(rows, n) + 0.7).requires_grad_(True)
    g = torch.randn(rows, n)
    bn_ref.train()
    out_ref = act_ref(bn_ref(z))
    out_ref.backward(g)

    bn = torch.nn.BatchNorm1d(n).cuda()
    with torch.no_grad():
        bn.weight.copy_(bn_ref.weight), bn.bias.copy_(bn_ref.bias)
    bn.train()
    act_mod = eval("torch.nn." + act).cuda()
    spec = F_.act_spec_of(act_mod)
    zc = z.detach().cuda().requires_grad_(True)
    out = F_.BatchNormActFn.apply(zc, bn.weight, bn.bias, spec.alpha, bn, spec, None)
    out.backward(g.cuda())
    torch.testing.assert_close(out.detach().cpu(), out_ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(zc.grad.cpu(), z.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(bn.weight.grad.cpu(), bn_ref.weight.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bn.bias.grad.cpu(), bn_ref.bias.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bn.running_mean.cpu(), bn_ref.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var.cpu(), bn_ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked) == 1
    if spec.alpha is not None:
        torch.testing.assert_close(act_mod.weight.grad.cpu(), act_ref.weight.grad, rtol=1e-4, atol=1e-4)


def test_batchnorm_eval_mode_uses_running_statistics():
    torch.manual_seed(3)
    n = 24
    bn_ref = torch.nn.BatchNorm1d(n)
    with torch.no_grad():
        bn_ref.running_mean.uniform_(-1, 1), bn_ref.running_var.uniform_(0.5, 2.0)
        bn_ref.weight.uniform_(0.5, 1.5), bn_ref.bias.uniform_(-0.5, 0.5)
    bn_ref.eval()
    z = torch.randn(500, n, requires_grad=True)
    out_ref = torch.relu(bn_ref(z))
    out_ref.sum().backward()
    bn = torch.nn.BatchNorm1d(n).cuda()
    bn.load_state_dict(bn_ref.state_dict())
    bn.eval()
    zc = z.detach().cuda().requires_grad_(True)
    spec = F_.act_spec_of(torch.nn.ReLU())
    out = F_.BatchNormActFn.apply(zc, bn.weight, bn.bias, None, bn, spec, None)
    out.sum().backward()
    torch.testing.assert_close(out.detach().cpu(), out_ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(zc.grad.cpu(), z.grad, rtol=1e-5, atol=1e-6)
    assert torch.equal(bn.running_mean.cpu(), bn_ref.running_mean) and int(bn.num_batches_tracked) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_mask_statistics(p, dtype):
    x = torch.ones(4000, 64, device="cuda", dtype=dtype)
    a = ops.dropout(x, p, seed=1234, offset=0)
    keep = (a != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01
    kept = a[a != 0].float()
    torch.testing.assert_close(kept, torch.full_like(kept, 1 / (1 - p)), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=0)
    assert torch.equal(a, ops.dropout(x, p, seed=1234, offset=0))            # a function of (seed, offset) only
    assert not torch.equal(a, ops.dropout(x, p, seed=1234, offset=1))
    assert not torch.equal(a, ops.dropout(x, p, seed=1235, offset=0))
    # the mask does not depend on the access width: a misaligned view of the same matrix sees the same columns kept
    wide = torch.ones(4000, 65, device="cuda", dtype=dtype)[:, :64]
    # (different leading dimension -> scalar path; same (row, column) -> same decision)
    assert torch.equal(ops.dropout(wide, p, seed=1234, offset=0) != 0, a != 0)
    # per-column and per-row keep rates are flat (no structure in the counter layout)
    assert ((a != 0).float().mean(0) - (1 - p)).abs().max().item() < 0.05
    assert torch.equal(ops.dropout(x, 0.0, seed=1, offset=0), x)
    assert float(ops.dropout(x, 1.0, seed=1, offset=0).abs().sum()) == 0.0


def test_dropout_autograd_reuses_the_forward_mask():
    x = torch.randn(512, 32, device="cuda", requires_grad=True)
    out = F_.DropoutFn.apply(x, 0.4, 99, 7)
    out.backward(torch.ones_like(out))
    mask = (out != 0)
    torch.testing.assert_close(x.grad, mask.float() / 0.6)


@pytest.mark.parametrize("sizes", [[5, 1, 9, 300], [2450] * 6, [1], [3, 0, 4]])
@pytest.mark.parametrize("origin_cols", [0, 4])
def test_global_pools_bit_exact_with_the_oracle(sizes, origin_cols):
    """models.py:347-352: global_mean_pool / global_max_pool over path_batch, gathered back per path — bit-exact with the
    CPU scatter reductions (sequential sums in row order), including an empty graph id in the middle."""
    torch.manual_seed(4)
    f = 4
    batch = torch.cat([torch.full((s,), i, dtype=torch.int64) for i, s in enumerate(sizes)])
    x7 = torch.randn(batch.numel(), 7)
    x = x7[:, :f]
    n_graphs = len(sizes)
    tail, mean, mx, _ = ops.global_pool_tail(x7.cuda()[:, :f], batch.cuda(), n_graphs, origin_cols)
    mean_ref = torch.zeros(n_graphs, f).scatter_add_(0, batch.view(-1, 1).expand(-1, f), x)
    cnt = torch.tensor(sizes, dtype=torch.float32).clamp(min=1).view(-1, 1)
    mean_ref = mean_ref / cnt
    assert torch.equal(mean.cpu(), mean_ref)
    if all(s > 0 for s in sizes):
        assert torch.equal(mean.cpu(), hgin_oracle.global_mean_pool(x, batch))
        assert torch.equal(mx.cpu(), hgin_oracle.global_max_pool(x, batch))
    want = torch.cat(([x[:, :origin_cols]] if origin_cols else []) + [mean_ref[batch], mx.cpu()[batch]], 1)
    assert torch.equal(tail.cpu(), want)
    # int32 graph ids (the narrowed batches of this package) give the same result
    tail32 = ops.global_pool_tail(x7.cuda()[:, :f], batch.cuda().int(), n_graphs, origin_cols)[0]
    assert torch.equal(tail32, tail)


def test_global_pools_unsorted_graph_ids():
    """PyG's pools accept any assignment vector; rows of a graph are then summed in row order (stable)."""
    torch.manual_seed(5)
    batch = torch.randint(0, 7, (5000,))
    x = torch.randn(5000, 4)
    _, mean, mx, _ = ops.global_pool_tail(x.cuda(), batch.cuda(), 7, 0)
    assert torch.equal(mean.cpu(), hgin_oracle.global_mean_pool(x, batch))
    assert torch.equal(mx.cpu(), hgin_oracle.global_max_pool(x, batch))


@pytest.mark.parametrize("heads,c", [(16, 8), (1, 8), (4, 16), (2, 4), (3, 32), (4, 128)])
@pytest.mark.parametrize("n_src,n_dst,e", [(300, 40, 2000), (40, 300, 900), (50, 50, 0), (7, 5, 30)])
def test_gatconv_against_the_oracle(n_src, n_dst, e, heads, c):
    """models.GATConv (hgin_gat_fwd / hgin_gat_bwd + the projection kernels) against the oracle's restatement of PyG's
    GATConv on a random bipartite relation: duplicate edges, edges with src id == dst id (dropped by PyG's
    remove_self_loops), destinations without edges, the appended loops for i < min(N_src, N_dst)."""
    from gnn_link_prediction_b200.models import GATConv
    torch.manual_seed(heads * 100 + c + e)
    ei = torch.stack((torch.randint(0, n_src, (e,)), torch.randint(0, n_dst, (e,))))
    x_src, x_dst = torch.randn(n_src, 5, requires_grad=True), torch.randn(n_dst, 3, requires_grad=True)
    ref = hgin_oracle.GATConv((5, 3), c, heads=heads)
    with torch.no_grad():
        ref.bias.uniform_(-0.5, 0.5)
    mine = GATConv((5, 3), c, heads=heads)
    mine.load_state_dict(ref.state_dict())
    mine.cuda()
    o_ref = ref((x_src, x_dst), ei)
    g = torch.randn_like(o_ref)
    o_ref.backward(g)
    xs_c, xd_c = x_src.detach().cuda().requires_grad_(True), x_dst.detach().cuda().requires_grad_(True)
    o = mine((xs_c, xd_c), ei.cuda())
    o.backward(g.cuda())
    torch.testing.assert_close(o.detach().cpu(), o_ref.detach(), rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(xs_c.grad.cpu(), x_src.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xd_c.grad.cpu(), x_dst.grad, rtol=1e-4, atol=1e-5)
    named = dict(ref.named_parameters())
    for k, p in mine.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), named[k].grad, rtol=1e-4, atol=1e-5 * float(named[k].grad.abs().max()) + 1e-6)
    # run-to-run deterministic (no atomics)
    xs_2 = x_src.detach().cuda().requires_grad_(True)
    mine.zero_grad()
    o2 = mine((xs_2, x_dst.detach().cuda()), ei.cuda())
    o2.backward(g.cuda())
    assert torch.equal(o2, o) and torch.equal(xs_2.grad, xs_c.grad)


def test_gatconv_shared_projection_and_unsupported_widths():
    from gnn_link_prediction_b200.models import GATConv
    torch.manual_seed(9)
    ref = hgin_oracle.GATConv(8, 8)             # layers >= 1 of HetroGAT: one projection shared by both sides
    mine = GATConv(8, 8)
    mine.load_state_dict(ref.state_dict())
    mine.cuda()
    ei = torch.stack((torch.randint(0, 60, (400,)), torch.randint(0, 30, (400,))))
    xs, xd = torch.randn(60, 8, requires_grad=True), torch.randn(30, 8, requires_grad=True)
    o_ref = ref((xs, xd), ei)
    o_ref.sum().backward()
    o = mine((xs.detach().cuda(), xd.detach().cuda()), ei.cuda())
    o.sum().backward()
    torch.testing.assert_close(o.detach().cpu(), o_ref.detach(), rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(mine.lin_src.weight.grad.cpu(), ref.lin_src.weight.grad, rtol=1e-4, atol=1e-5)
    bad = GATConv((3, 3), 6, heads=2).cuda()    # 6 channels per head: not a power of two
    with pytest.raises(ops.HginError):
        bad((torch.randn(4, 3, device="cuda"), torch.randn(4, 3, device="cuda")), torch.zeros(2, 1, dtype=torch.int64, device="cuda"))


def test_batchnorm_on_bf16_rows():
    torch.manual_seed(6)
    n = 64
    z = (1.5 * torch.randn(3000, n) - 0.3).bfloat16()
    g = torch.randn(3000, n).bfloat16()
    bn_ref = torch.nn.BatchNorm1d(n)
    zr = z.float().requires_grad_(True)
    out_ref = torch.nn.functional.leaky_relu(bn_ref(zr), 0.1)
    out_ref.backward(g.float())
    bn = torch.nn.BatchNorm1d(n).cuda()
    spec = F_.act_spec_of(torch.nn.LeakyReLU(0.1))
    zc = z.cuda().requires_grad_(True)
    out = F_.BatchNormActFn.apply(zc, bn.weight, bn.bias, None, bn, spec, None)
    assert out.dtype == torch.bfloat16
    out.backward(g.cuda())
    torch.testing.assert_close(out.float().cpu(), out_ref.detach(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(zc.grad.float().cpu(), zr.grad, rtol=2e-2, atol=2e-3)
    torch.testing.assert_close(bn.weight.grad.cpu(), bn_ref.weight.grad, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(bn.running_var.cpu(), bn_ref.running_var, rtol=1e-4, atol=1e-5)
