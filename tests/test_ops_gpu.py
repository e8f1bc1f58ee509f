"""Kernel-level parity: every C-ABI entry point against the CPU oracle on the same seeded inputs.
Integer work and the fp32 aggregation are compared bit-for-bit; GEMM-shaped work within the fp32
tolerance written in each test."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from gnn_link_prediction_b200 import ops

pytestmark = pytest.mark.gpu


def _rand_edges(ns, nd, e, seed, sort_src=False):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, ns, (e,), generator=g), torch.randint(0, nd, (e,), generator=g)])
    if sort_src and e:
        ei = ei[:, torch.argsort(ei[0], stable=True)]
    return ei


CSR_SHAPES = [(1000, 300, 20000), (50, 2000, 7000), (7, 5, 0), (1, 1, 40), (3000, 10, 50000), (100000, 9000, 400000),
              (5, 70000, 33)]


@pytest.mark.parametrize("ns,nd,e", CSR_SHAPES)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("by", ["dst", "src"])
def test_csr_build_bit_exact(ns, nd, e, dtype, by):
    ei = _rand_edges(ns, nd, e, seed=ns + e, sort_src=(e % 2 == 0))
    csr = ops.csr_build(ei.to(dtype).cuda(), ns, nd, by=by, want_perm=True).validate()
    src = ei if by == "dst" else ei.flip(0).contiguous()
    rows, cols = (nd, ns) if by == "dst" else (ns, nd)
    rowptr, col, perm = c_oracle.csr_build(src.numpy(), cols, rows)
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(csr.col.cpu().numpy(), col)
    assert np.array_equal(csr.perm.cpu().numpy(), perm)
    # and against torch's own stable argsort, independently of the C oracle
    key = ei[1] if by == "dst" else ei[0]
    other = ei[0] if by == "dst" else ei[1]
    p = torch.argsort(key, stable=True)
    assert torch.equal(csr.col.cpu().long(), other[p])


def test_csr_build_long_rows_and_strided_input():
    # one hub row with 5000 entries exercises the rank-counting branch; input is a strided view
    ns, nd = 6000, 40
    ei = _rand_edges(ns, nd, 30000, seed=3)
    ei[1, :5000] = 7
    big = torch.zeros(2, 40000, dtype=torch.int64)
    big[:, :30000] = ei
    dev = big.cuda()[:, :30000]
    csr = ops.csr_build(dev, ns, nd).validate()
    rowptr, col, _ = c_oracle.csr_build(ei.numpy(), ns, nd)
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr) and np.array_equal(csr.col.cpu().numpy(), col)


def test_csr_build_sorted_fast_path_with_gaps_and_padding():
    """Keys already sorted (the reference's by-source order): empty rows at both ends and in the
    middle, duplicate keys, trailing (-1,-1) padding."""
    src = torch.tensor([2, 2, 2, 5, 5, 9, 9, 9, 9, -1, -1], dtype=torch.int32)
    dst = torch.tensor([7, 0, 3, 3, 1, 4, 4, 0, 2, -1, -1], dtype=torch.int32)
    csr = ops.csr_build(torch.stack([src, dst]).cuda(), 12, 8, by="src", want_perm=True).validate()
    assert csr.rowptr.cpu().tolist() == [0, 0, 0, 3, 3, 3, 5, 5, 5, 5, 9, 9, 9]
    assert csr.col.cpu().tolist()[:9] == [7, 0, 3, 3, 1, 4, 4, 0, 2]
    assert csr.perm.cpu().tolist()[:9] == list(range(9))
    # all padding
    pad = torch.full((2, 6), -1, dtype=torch.int64).cuda()
    assert ops.csr_build(pad, 4, 3).validate().rowptr.cpu().tolist() == [0, 0, 0, 0]


def test_csr_build_flags_out_of_range():
    ei = torch.tensor([[0, 5, 1], [0, 1, 9]])
    csr = ops.csr_build(ei.cuda(), 3, 2)
    with pytest.raises(IndexError):
        csr.validate()
    assert csr.rowptr.cpu().tolist() == [0, 1, 1]   # only the valid edge (0 -> 0) is kept


COMBINE_SHAPES = [  # ns, nd, e, f
    (1000, 300, 20000, 128), (50, 2000, 7000, 3), (500, 500, 60000, 8), (10, 5, 0, 4), (300, 200, 9000, 16),
    (300, 200, 9000, 64), (200, 100, 3000, 256), (200, 100, 3000, 5), (200, 100, 3000, 33), (64, 64, 4000, 100),
    (40, 4000, 9000, 32), (2000, 7, 30000, 128)]


@pytest.mark.parametrize("ns,nd,e,f", COMBINE_SHAPES)
@pytest.mark.parametrize("mode", ["add", "concat", "none"])
def test_gin_combine_bit_exact(ns, nd, e, f, mode):
    ei = _rand_edges(ns, nd, e, seed=e + f, sort_src=True)
    g = torch.Generator().manual_seed(f)
    xs = torch.randn(ns, f, generator=g)
    xd = torch.randn(nd, f if mode == "add" else 3, generator=g)
    eps = 0.37
    rowptr, col, _ = c_oracle.csr_build(ei.numpy(), ns, nd)
    want = c_oracle.gin_combine(rowptr, col, xs.numpy(), None if mode == "none" else xd.numpy(), eps, mode == "concat")
    csr = ops.csr_build(ei.cuda(), ns, nd)
    self_mode = {"add": ops.SELF_ADD, "concat": ops.SELF_CONCAT, "none": ops.SELF_NONE}[mode]
    got = ops.gin_combine(csr, xs.cuda(), None if mode == "none" else xd.cuda(),
                          torch.tensor([eps]).cuda() if mode != "none" else None, self_mode)
    assert np.array_equal(got.cpu().numpy(), want)


def test_gin_combine_strided_inputs_and_accumulate():
    # layer-0 layout: 3 columns of a 7-wide row (ld = 7), as models.py:336-338 slices them
    ns, nd, e = 400, 150, 5000
    ei = _rand_edges(ns, nd, e, seed=11, sort_src=True)
    g = torch.Generator().manual_seed(1)
    xs7, xd7 = torch.randn(ns, 7, generator=g), torch.randn(nd, 7, generator=g)
    rowptr, col, _ = c_oracle.csr_build(ei.numpy(), ns, nd)
    want = c_oracle.gin_combine(rowptr, col, xs7[:, 0:3].numpy(), xd7[:, 0:3].numpy(), 0.0, True)
    csr = ops.csr_build(ei.cuda(), ns, nd)
    got = ops.gin_combine(csr, xs7.cuda()[:, 0:3], xd7.cuda()[:, 0:3], torch.zeros(1).cuda(), ops.SELF_CONCAT)
    assert np.array_equal(got.cpu().numpy(), want)
    # accumulate: out += agg, fp32 add of the two bit-exact parts
    base = torch.randn(nd, 6, generator=g)
    acc = base.clone().cuda()
    ops.gin_combine(csr, xs7.cuda()[:, 0:3], xd7.cuda()[:, 0:3], torch.zeros(1).cuda(), ops.SELF_CONCAT, out=acc,
                    accumulate=True)
    assert np.array_equal(acc.cpu().numpy(), base.numpy() + want)


def test_transposed_gather_matches_index_add_backward():
    ns, nd, e, f = 700, 300, 12000, 128
    ei = _rand_edges(ns, nd, e, seed=5, sort_src=True)
    gout = torch.randn(nd, f, generator=torch.Generator().manual_seed(2))
    xs = torch.zeros(ns, f, requires_grad=True)
    msg = xs.index_select(0, ei[0])
    torch.zeros(nd, f).scatter_add_(0, ei[1].view(-1, 1).expand_as(msg), msg).backward(gout)
    csr_t = ops.csr_build(ei.cuda(), ns, nd, by="src")
    got = ops.gin_combine(csr_t, gout.cuda())
    assert torch.equal(got.cpu(), xs.grad)


LINEAR_SHAPES = [  # rows, k1, k2, n
    (1000, 6, 0, 8), (777, 8, 0, 8), (1500, 128, 3, 128), (900, 128, 0, 32), (333, 32, 0, 1), (2000, 11, 0, 128),
    (129, 16, 0, 16), (5, 3, 2, 7), (4100, 64, 0, 64), (300, 200, 0, 150)]
# fp32 GEMM against an fp64 reference: error ~ sqrt(k) * 2^-24 * |x||w|; the CPU reference's own
# sgemm differs from fp64 by the same amount, so this is the "rel 1e-5" bar of the north star.
RTOL, ATOL = 2e-5, 2e-5


@pytest.mark.parametrize("rows,k1,k2,n", LINEAR_SHAPES)
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE, ops.ACT_RELU])
def test_linear_fwd(rows, k1, k2, n, act):
    g = torch.Generator().manual_seed(rows + n)
    x1, x2 = torch.randn(rows, k1, generator=g), (torch.randn(rows, k2, generator=g) if k2 else None)
    W, b = torch.randn(n, k1 + k2, generator=g) / (k1 + k2) ** 0.5, torch.randn(n, generator=g)
    alpha = torch.tensor([0.25])
    x = x1 if x2 is None else torch.cat((x1, x2), 1)
    z_ref = x.double() @ W.double().t() + b.double()
    o_ref = {ops.ACT_PRELU: torch.where(z_ref > 0, z_ref, 0.25 * z_ref), ops.ACT_NONE: z_ref,
             ops.ACT_RELU: z_ref.clamp(min=0)}[act]
    z, o = ops.linear_fwd(x1.cuda(), W.cuda(), b.cuda(), x2=None if x2 is None else x2.cuda(), act=act,
                          alpha=alpha.cuda())
    torch.testing.assert_close(z.cpu().double(), z_ref, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(o.cpu().double(), o_ref, rtol=RTOL, atol=ATOL)
    # merge: out += act(z) (HeteroConv 'sum')
    prev = torch.randn(rows, n, generator=g)
    acc = prev.clone().cuda()
    ops.linear_fwd(x1.cuda(), W.cuda(), b.cuda(), x2=None if x2 is None else x2.cuda(), act=act, alpha=alpha.cuda(),
                   out=acc, accumulate_out=True, want_z=False)
    torch.testing.assert_close(acc.cpu().double(), prev.double() + o_ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("rows,k1,k2,n", LINEAR_SHAPES)
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE])
def test_linear_bwd(rows, k1, k2, n, act):
    g = torch.Generator().manual_seed(rows * 3 + n)
    k = k1 + k2
    x = torch.randn(rows, k, generator=g, dtype=torch.float64, requires_grad=True)
    W = (torch.randn(n, k, generator=g, dtype=torch.float64) / k ** 0.5).requires_grad_(True)
    b = torch.randn(n, generator=g, dtype=torch.float64, requires_grad=True)
    alpha = torch.tensor([0.25], dtype=torch.float64, requires_grad=True)
    dot_x = torch.randn(rows, k, generator=g, dtype=torch.float64)
    z_ref = x @ W.t() + b
    o_ref = torch.nn.functional.prelu(z_ref, alpha) if act == ops.ACT_PRELU else z_ref
    gout = torch.randn(rows, n, generator=g, dtype=torch.float64)
    o_ref.backward(gout)
    f32 = lambda t: t.detach().float().cuda()
    x1, x2 = f32(x[:, :k1]).contiguous(), (f32(x[:, k1:]).contiguous() if k2 else None)
    r = ops.linear_bwd(f32(gout), f32(z_ref), x1, f32(W), x2=x2, act=act, alpha=f32(alpha), dot_x=f32(dot_x),
                       want_dalpha=act == ops.ACT_PRELU)
    scale = rows ** 0.5
    torch.testing.assert_close(r["dx"].cpu().double(), x.grad, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(r["dW"].cpu().double(), W.grad, rtol=RTOL, atol=ATOL * scale)
    torch.testing.assert_close(r["db"].cpu().double(), b.grad, rtol=RTOL, atol=ATOL * scale)
    torch.testing.assert_close(r["ddot"].cpu().double(), (x.grad * dot_x).sum().view(1), rtol=1e-4, atol=ATOL * scale * k ** 0.5)
    if act == ops.ACT_PRELU:
        torch.testing.assert_close(r["dalpha"].cpu().double(), alpha.grad, rtol=1e-4, atol=ATOL * scale * n ** 0.5)
    # column-restricted input gradient (layer-0 self block) without storing dx
    c0 = k // 2
    r2 = ops.linear_bwd(f32(gout), f32(z_ref), x1, f32(W), x2=x2, act=act, alpha=f32(alpha), dx_cols=(c0, k),
                        want_dx=False, dot_x=f32(dot_x[:, c0:]).contiguous(), want_dw=False, want_db=False)
    assert r2["dx"] is None and r2["dW"] is None
    torch.testing.assert_close(r2["ddot"].cpu().double(), (x.grad[:, c0:] * dot_x[:, c0:]).sum().view(1), rtol=1e-4,
                               atol=ATOL * scale * k ** 0.5)


def test_linear_bwd_is_deterministic():
    g = torch.Generator().manual_seed(0)
    rows, k, n = 50000, 128, 128
    x, W = torch.randn(rows, k, generator=g).cuda(), torch.randn(n, k, generator=g).cuda()
    z, gout = torch.randn(rows, n, generator=g).cuda(), torch.randn(rows, n, generator=g).cuda()
    a = torch.tensor([0.25]).cuda()
    r1 = ops.linear_bwd(gout, z, x, W, act=ops.ACT_PRELU, alpha=a, want_dalpha=True)
    r2 = ops.linear_bwd(gout, z, x, W, act=ops.ACT_PRELU, alpha=a, want_dalpha=True)
    for key in ("dx", "dW", "db", "dalpha"):
        assert torch.equal(r1[key], r2[key]), key


@pytest.mark.parametrize("n", [1, 1000, 250000])
def test_sqrt_mape_loss(n):
    g = torch.Generator().manual_seed(n)
    pred = torch.randn(n, 1, generator=g, dtype=torch.float64, requires_grad=True)
    y = 0.1 + 9 * torch.rand(n, generator=g, dtype=torch.float64)
    mape = 100.0 * torch.mean(torch.abs((pred - y.view(-1, 1)) / y.view(-1, 1)))   # train.py:13
    loss = torch.sqrt(mape)                                                      # train.py:42
    loss.backward()
    sums = ops.mape_sum(pred.detach().float().cuda(), y.float().cuda())
    assert float(sums[1]) == n
    loss_out, dpred = ops.sqrt_mape_bwd(pred.detach().float().cuda(), y.float().cuda(), sums)
    torch.testing.assert_close(loss_out.cpu().double(), torch.stack([mape.detach(), loss.detach()]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dpred.cpu().double(), pred.grad, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (False, 0.01), (True, 0.01)])
def test_adam_matches_torch_optim(decoupled, wd):
    g = torch.Generator().manual_seed(4)
    p0 = torch.randn(5000, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref], lr=1e-3, weight_decay=wd)
    p, m, v = p0.clone().cuda(), torch.zeros(5000).cuda(), torch.zeros(5000).cuda()
    step = torch.zeros(1, dtype=torch.int32).cuda()
    for _ in range(6):
        grad = torch.randn(5000, generator=g)
        ref.grad = grad.clone()
        opt.step()
        ops.increment(step)
        ops.adam_step(p, grad.cuda(), m, v, step, 1e-3, 0.9, 0.999, 1e-8, wd, decoupled)
    assert int(step.item()) == 6
    torch.testing.assert_close(p.cpu(), ref.detach(), rtol=1e-6, atol=1e-7)


# ---- tensor-core path (HGIN_MATH_TF32): tcgen05 kind::tf32, fp32 accumulation ---------------------
# tf32 keeps 10 mantissa bits: products carry ~2^-11 relative error, so the bar is the north
# star's "1e-2 under reduced-precision GEMMs" with margin: rtol 5e-3, atol 5e-3 * max|ref|.
TC_SHAPES = [  # rows, k1, k2, n
    (4096, 128, 0, 128), (1000, 128, 3, 128), (130, 128, 0, 32), (20000, 64, 0, 128), (777, 32, 0, 16),
    (5000, 128, 0, 64), (3001, 96, 2, 48)]


def _tc_close(got, want, atol_rel=5e-3):
    torch.testing.assert_close(got.cpu().double(), want, rtol=5e-3, atol=atol_rel * float(want.abs().max()) + 1e-9)


@pytest.mark.parametrize("rows,k1,k2,n", TC_SHAPES)
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE])
def test_linear_fwd_tf32(rows, k1, k2, n, act):
    g = torch.Generator().manual_seed(rows + n)
    x1, x2 = torch.randn(rows, k1, generator=g), (torch.randn(rows, k2, generator=g) if k2 else None)
    W, b = torch.randn(n, k1 + k2, generator=g) / (k1 + k2) ** 0.5, torch.randn(n, generator=g)
    alpha = torch.tensor([0.25])
    x = x1 if x2 is None else torch.cat((x1, x2), 1)
    z_ref = x.double() @ W.double().t() + b.double()
    o_ref = torch.where(z_ref > 0, z_ref, 0.25 * z_ref) if act == ops.ACT_PRELU else z_ref
    z, o = ops.linear_fwd(x1.cuda(), W.cuda(), b.cuda(), x2=None if x2 is None else x2.cuda(), act=act,
                          alpha=alpha.cuda(), math_mode=ops.MATH_TF32)
    _tc_close(z, z_ref)
    _tc_close(o, o_ref)
    prev = torch.randn(rows, n, generator=g)
    acc = prev.clone().cuda()
    ops.linear_fwd(x1.cuda(), W.cuda(), b.cuda(), x2=None if x2 is None else x2.cuda(), act=act, alpha=alpha.cuda(),
                   out=acc, accumulate_out=True, want_z=False, math_mode=ops.MATH_TF32)
    _tc_close(acc, prev.double() + o_ref)


@pytest.fixture(params=["three-pass", "fused-dw", "fused-all"])
def fused_bwd(request):
    """The three tensor-core backward variants: dz_prepare + dx + dW passes (default); dz fused
    into the dW kernel (opt-in); everything in one kernel (opt-in)."""
    ops.set_option("fused_bwd", 1 if request.param == "fused-all" else 0)
    ops.set_option("fused_dw", 1 if request.param == "fused-dw" else 0)
    yield request.param
    ops.set_option("fused_bwd", 0)
    ops.set_option("fused_dw", 0)


@pytest.mark.parametrize("rows,k1,k2,n", [s for s in TC_SHAPES if s[1] % 16 == 0])
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE])
def test_linear_bwd_tf32(rows, k1, k2, n, act, fused_bwd):
    g = torch.Generator().manual_seed(rows * 3 + n)
    k = k1 + k2
    x = torch.randn(rows, k, generator=g, dtype=torch.float64, requires_grad=True)
    W = (torch.randn(n, k, generator=g, dtype=torch.float64) / k ** 0.5).requires_grad_(True)
    b = torch.randn(n, generator=g, dtype=torch.float64, requires_grad=True)
    alpha = torch.tensor([0.25], dtype=torch.float64, requires_grad=True)
    dot_x = torch.randn(rows, k1, generator=g, dtype=torch.float64)
    z_ref = x @ W.t() + b
    o_ref = torch.nn.functional.prelu(z_ref, alpha) if act == ops.ACT_PRELU else z_ref
    gout = torch.randn(rows, n, generator=g, dtype=torch.float64)
    o_ref.backward(gout)
    f32 = lambda t: t.detach().float().cuda()
    x1, x2 = f32(x[:, :k1]).contiguous(), (f32(x[:, k1:]).contiguous() if k2 else None)
    r = ops.linear_bwd(f32(gout), f32(z_ref), x1, f32(W), x2=x2, act=act, alpha=f32(alpha), dx_cols=(0, k1),
                       dot_x=f32(dot_x), want_dalpha=act == ops.ACT_PRELU, math_mode=ops.MATH_TF32)
    _tc_close(r["dx"], x.grad[:, :k1])
    _tc_close(r["dW"], W.grad)
    # act NONE: g is dz, read in place by both GEMMs, and db comes out of the weight-gradient MMA
    # (a column of ones) with tf32-rounded addends when k1 % 32 == 0
    _tc_close(r["db"], b.grad, atol_rel=1e-4 if act == ops.ACT_PRELU else 5e-3)
    _tc_close(r["ddot"], (x.grad[:, :k1] * dot_x).sum().view(1), atol_rel=2e-2)
    if act == ops.ACT_PRELU:
        _tc_close(r["dalpha"], alpha.grad, atol_rel=1e-4)
    r2 = ops.linear_bwd(f32(gout), f32(z_ref), x1, f32(W), x2=x2, act=act, alpha=f32(alpha), dx_cols=(0, k1),
                        math_mode=ops.MATH_TF32)
    for key in ("dx", "dW", "db"):
        assert torch.equal(r[key], r2[key]), f"{key} not deterministic"


@pytest.mark.parametrize("rows,k1,k2,n", [(4096, 128, 0, 128), (1000, 128, 3, 128), (2000, 48, 0, 64), (130, 128, 0, 32),
                                          (500, 64, 0, 16), (300, 8, 0, 16)])
@pytest.mark.parametrize("post_act", [ops.ACT_PRELU, ops.ACT_RELU])
def test_linear_bwd_post_activation_and_dz_in_place(rows, k1, k2, n, post_act):
    """Two chained layers x0 -> (W0, PReLU) -> x -> (W, PReLU) -> out.  The upper layer's backward
    (hgin_linear_bwd_post) must hand the lower layer its dz = dx * act'(z0) plus dalpha0, and the
    lower layer's backward on that dz with ACT_NONE must reproduce autograd (tensor-core shapes and
    the generic elementwise form for shapes the tcgen05 kernel does not take)."""
    g = torch.Generator().manual_seed(rows + k1 + n)
    k = k1 + k2
    d = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float64)
    x0 = d(rows, 32)
    W0 = (d(k1, 32) / 32 ** 0.5).requires_grad_(True)
    b0 = d(k1).requires_grad_(True)
    a0 = torch.tensor([0.25], dtype=torch.float64, requires_grad=True)
    z0 = x0 @ W0.t() + b0
    x1 = torch.nn.functional.prelu(z0, a0) if post_act == ops.ACT_PRELU else torch.relu(z0)
    x2 = d(rows, k2) if k2 else None
    W = (d(n, k) / k ** 0.5).requires_grad_(True)
    b = d(n).requires_grad_(True)
    a = torch.tensor([0.3], dtype=torch.float64, requires_grad=True)
    z = (x1 if x2 is None else torch.cat((x1, x2), 1)) @ W.t() + b
    gout = d(rows, n)
    params = [W, b, a, W0, b0] + ([a0] if post_act == ops.ACT_PRELU else [])
    dz0_ref, *pg = torch.autograd.grad(torch.nn.functional.prelu(z, a), [z0] + params, gout)
    for prm, gr in zip(params, pg):
        prm.grad = gr
    f32 = lambda t: None if t is None else t.detach().float().cuda().contiguous()
    post = ops.PostAct(f32(z0), post_act, f32(a0))
    r = ops.linear_bwd(f32(gout), f32(z), f32(x1), f32(W), x2=f32(x2), act=ops.ACT_PRELU, alpha=f32(a), dx_cols=(0, k1),
                       want_dalpha=True, math_mode=ops.MATH_TF32, post=post)
    assert post.applied
    _tc_close(r["dx"], dz0_ref)
    _tc_close(r["dW"], W.grad)
    _tc_close(r["dalpha"], a.grad, atol_rel=1e-4)
    if post_act == ops.ACT_PRELU:
        _tc_close(post.dalpha, a0.grad, atol_rel=2e-2)
    else:
        assert post.dalpha is None
    # the layer below: its g is already dz
    r0 = ops.linear_bwd(r["dx"], None, f32(x0), f32(W0), act=ops.ACT_NONE, want_dx=False, math_mode=ops.MATH_TF32)
    _tc_close(r0["dW"], W0.grad, atol_rel=1e-2)
    _tc_close(r0["db"], b0.grad, atol_rel=1e-2)
    assert r0["dalpha"] is None and r0["dx"] is None


@pytest.mark.parametrize("rows,k,n", [(5000, 128, 128), (777, 64, 128), (4097, 128, 32)])
@pytest.mark.parametrize("post_act", [ops.ACT_PRELU, ops.ACT_RELU])
def test_linear_bwd_with_the_gin_self_branch_on_its_epilogue(rows, k, n, post_act):
    """hgin_linear_bwd_post_self: dx = (1 + eps) * (dz W) * act'(z0), d(eps) = sum (dz W) * act(z0) and dalpha0 out of
    the input-gradient GEMM == hgin_linear_bwd (dh to memory) followed by the edgeless hgin_gin_combine_post pass,
    and == float64 autograd of  out = ((1 + eps) * act(z0)) W^T."""
    g = torch.Generator().manual_seed(rows + k + n)
    d = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float64)
    z0 = d(rows, k).requires_grad_(True)
    a0 = torch.tensor([0.2], dtype=torch.float64, requires_grad=True)
    eps = torch.tensor([0.37], dtype=torch.float64, requires_grad=True)
    W = (d(n, k) / k ** 0.5).requires_grad_(True)
    x_dst = torch.nn.functional.prelu(z0, a0) if post_act == ops.ACT_PRELU else torch.relu(z0)
    h = (1 + eps) * x_dst
    dz = d(rows, n)
    wanted = [z0, W, eps] + ([a0] if post_act == ops.ACT_PRELU else [])
    grads = torch.autograd.grad(h @ W.t(), wanted, dz)
    f32 = lambda t: t.detach().float().cuda().contiguous()
    post = ops.PostAct(f32(z0), post_act, f32(a0))
    r = ops.linear_bwd(f32(dz), None, f32(h), f32(W), act=ops.ACT_NONE, math_mode=ops.MATH_TF32, post=post,
                       self_eps=f32(eps), want_self_ddot=True)
    assert post.applied and ops.post_self_eligible(rows, k, n, ops.MATH_TF32)
    _tc_close(r["dx"], grads[0])
    _tc_close(r["dW"], grads[1])
    scale = float((f32(dz).abs().mean() * f32(h).abs().mean()) * (rows * k) ** 0.5)     # size of such a sum of products
    assert abs(float(r["ddot"]) - float(grads[2])) <= 2e-2 * scale
    if post_act == ops.ACT_PRELU:
        assert abs(float(post.dalpha) - float(grads[3])) <= 2e-2 * scale
    # the two-pass form it replaces
    post2 = ops.PostAct(f32(z0), post_act, f32(a0))
    dh = ops.linear_bwd(f32(dz), None, f32(h), f32(W), act=ops.ACT_NONE, want_dw=False, want_db=False,
                        math_mode=ops.MATH_TF32)["dx"]
    dx2, ddot2 = ops.gin_combine(None, dh, dh, f32(eps), ops.SELF_ADD, post=post2, want_ddot=True)
    assert torch.equal(r["dx"], dx2)                      # same products, same single rounding of (1 + eps) * dh
    torch.testing.assert_close(r["ddot"], ddot2, rtol=1e-3, atol=1e-3 * scale)
    # shapes outside the tensor-core path are refused, not silently computed some other way
    with pytest.raises(ops.HginError):
        ops.linear_bwd(f32(dz)[:, :8].contiguous(), None, f32(h), f32(W)[:8].contiguous(), act=ops.ACT_NONE,
                       math_mode=ops.MATH_FP32, post=ops.PostAct(f32(z0), post_act, f32(a0)), self_eps=f32(eps))


@pytest.mark.parametrize("ns,nd,e,f", [(3000, 5000, 15000, 128), (700, 90, 4000, 128), (900, 1200, 0, 64), (50, 64, 300, 8),
                                        (40, 33, 100, 5)])
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_RELU])
def test_gin_combine_post_activation(ns, nd, e, f, act):
    """hgin_gin_combine_post == hgin_gin_combine followed by PReLU/ReLU backward of the layer below:
    the stored rows bit for bit, dalpha within fp32 summation-order noise; with and without edges."""
    g = torch.Generator().manual_seed(ns + e + f)
    ei = _rand_edges(ns, nd, e, seed=e + 1).cuda()
    csr = ops.csr_build(ei, ns, nd, by="dst")
    x_src, x_self = torch.randn(ns, f, generator=g).cuda(), torch.randn(nd, f, generator=g).cuda()
    z, eps, alpha = torch.randn(nd, f, generator=g).cuda(), torch.tensor([0.125]).cuda(), torch.tensor([0.2]).cuda()
    prev = torch.randn(nd, f, generator=g).cuda()
    for csr_arg, mode, acc in ((csr, ops.SELF_ADD, False), (csr, ops.SELF_NONE, True), (None, ops.SELF_ADD, False),
                               (None, ops.SELF_ADD, True)):
        xs = x_src if csr_arg is not None else x_self
        plain = ops.gin_combine(csr_arg, xs, x_self if mode != ops.SELF_NONE else None, eps, mode,
                                out=prev.clone() if acc else None, accumulate=acc)
        post = ops.PostAct(z, act, alpha)
        got = ops.gin_combine(csr_arg, xs, x_self if mode != ops.SELF_NONE else None, eps, mode,
                              out=prev.clone() if acc else None, accumulate=acc, post=post,
                              want_ddot=mode == ops.SELF_ADD)
        if mode == ops.SELF_ADD:   # d(eps) = sum x_self * act(z), with act(z) the x_dst of the layer above
            got, ddot = got
            zd = z.double()
            xd = torch.where(zd > 0, zd, 0.2 * zd if act == ops.ACT_PRELU else torch.zeros_like(zd))
            ref_dot = (x_self.double() * xd)
            assert abs(float(ddot) - float(ref_dot.sum())) <= 1e-5 * float(ref_dot.abs().sum()) + 1e-6
        want = torch.where(z > 0, plain, (alpha * plain) if act == ops.ACT_PRELU else torch.zeros_like(plain))
        assert post.applied and torch.equal(got, want)
        if csr_arg is None and not acc:
            assert torch.equal(plain, (1 + eps) * x_self)
        if act == ops.ACT_PRELU:
            ref = (plain.double() * torch.clamp(z.double(), max=0)).sum()
            scale = float((plain.double() * torch.clamp(z.double(), max=0)).abs().sum())
            assert abs(float(post.dalpha) - float(ref)) <= 1e-5 * scale + 1e-6
        else:
            assert post.dalpha is None


@pytest.mark.parametrize("ns,nd,e,f", [(3000, 5000, 15000, 128), (5000, 300, 40000, 128), (900, 1200, 0, 64), (50, 64, 300, 8),
                                        (40, 33, 100, 5)])
@pytest.mark.parametrize("which", ["src", "self", "both"])
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_RELU])
def test_gin_combine_on_pre_activation_inputs(ns, nd, e, f, which, act):
    """hgin_gin_combine_pre(z) == hgin_gin_combine(act(z)) bit for bit (short and long rows, both self modes)."""
    g = torch.Generator().manual_seed(ns + e + f)
    csr = ops.csr_build(_rand_edges(ns, nd, e, seed=e + 2).cuda(), ns, nd, by="dst")
    z_src, z_self = torch.randn(ns, f, generator=g).cuda(), torch.randn(nd, f, generator=g).cuda()
    eps, alpha = torch.tensor([0.3]).cuda(), torch.tensor([0.25]).cuda()
    activate = lambda z: torch.where(z > 0, z, alpha * z if act == ops.ACT_PRELU else torch.zeros_like(z))
    for mode in (ops.SELF_ADD, ops.SELF_CONCAT):
        kw = {}
        if which in ("src", "both"):
            kw["src_act"] = (act, alpha if act == ops.ACT_PRELU else None)
        if which in ("self", "both"):
            kw["self_act"] = (act, alpha if act == ops.ACT_PRELU else None)
        got = ops.gin_combine(csr, z_src, z_self, eps, mode, **kw)
        want = ops.gin_combine(csr, activate(z_src) if "src_act" in kw else z_src,
                               activate(z_self) if "self_act" in kw else z_self, eps, mode)
        assert torch.equal(got, want)


@pytest.mark.parametrize("ns,nd,e,f", [(3000, 5000, 15000, 128), (700, 90, 4000, 128), (900, 1200, 0, 64), (333, 257, 2000, 24),
                                        (50, 64, 300, 16), (40, 33, 100, 5), (2000, 1500, 9000, 256)])
@pytest.mark.parametrize("variant", ["plain", "self_add", "concat", "accumulate", "pre", "post"])
def test_gin_combine_on_bf16_rows_equals_the_fp32_kernel_rounded_once(ns, nd, e, f, variant):
    """HGIN_DTYPE_BF16 rows: every addition is the fp32 addition of the fp32 path (bf16 -> fp32 widening is exact),
    and the finished row is rounded to bf16 once — so the result must equal the fp32 kernel's result on the same
    (bf16-representable) inputs, rounded to bf16, BIT FOR BIT, in every mode."""
    g = torch.Generator().manual_seed(ns + nd + e + f)
    ei = torch.stack([torch.randint(0, ns, (e,), generator=g), torch.randint(0, nd, (e,), generator=g)]).cuda()
    csr = ops.csr_build(ei, ns, nd, by="dst")
    xs = torch.randn(ns, f, generator=g).cuda().to(torch.bfloat16)
    xd = torch.randn(nd, f, generator=g).cuda().to(torch.bfloat16)
    eps = torch.tensor([0.31], device="cuda")
    a = torch.tensor([0.25], device="cuda")
    kw, ref_kw = {}, {}
    mode = ops.SELF_ADD
    if variant == "plain":
        got = ops.gin_combine(csr, xs)
        want = ops.gin_combine(csr, xs.float())
    elif variant == "concat":
        got = ops.gin_combine(csr, xs, xd, eps, ops.SELF_CONCAT)
        want = ops.gin_combine(csr, xs.float(), xd.float(), eps, ops.SELF_CONCAT)
    elif variant == "accumulate":
        old = torch.randn(nd, f, generator=g).cuda().to(torch.bfloat16)
        got = ops.gin_combine(csr, xs, xd, eps, mode, out=old.clone(), accumulate=True)
        want = ops.gin_combine(csr, xs.float(), xd.float(), eps, mode, out=old.float(), accumulate=True)
    elif variant == "pre":
        got = ops.gin_combine(csr, xs, xd, eps, mode, src_act=(ops.ACT_PRELU, a), self_act=(ops.ACT_RELU, None))
        want = ops.gin_combine(csr, xs.float(), xd.float(), eps, mode, src_act=(ops.ACT_PRELU, a), self_act=(ops.ACT_RELU, None))
    elif variant == "post":
        z = torch.randn(nd, f, generator=g).cuda().to(torch.bfloat16)
        p1, p2 = ops.PostAct(z, ops.ACT_PRELU, a), ops.PostAct(z.float(), ops.ACT_PRELU, a)
        got, ddot = ops.gin_combine(csr, xs, xd, eps, mode, post=p1, want_ddot=True)
        want, ddot_ref = ops.gin_combine(csr, xs.float(), xd.float(), eps, mode, post=p2, want_ddot=True)
        scale = float(xd.float().abs().mean()) * (nd * f) ** 0.5
        torch.testing.assert_close(p1.dalpha, p2.dalpha, rtol=1e-4, atol=1e-4 * scale * 10)
        torch.testing.assert_close(ddot, ddot_ref, rtol=1e-4, atol=1e-4 * scale)
    else:
        got = ops.gin_combine(csr, xs, xd, eps, mode)
        want = ops.gin_combine(csr, xs.float(), xd.float(), eps, mode)
    assert got.dtype == torch.bfloat16
    assert torch.equal(got, want.to(torch.bfloat16))


def test_gin_combine_rejects_mixed_row_types():
    csr = ops.csr_build(torch.tensor([[0, 1], [1, 0]], device="cuda"), 2, 2, by="dst")
    x = torch.randn(2, 8, device="cuda")
    with pytest.raises(ops.HginError):
        ops.gin_combine(csr, x.to(torch.bfloat16), x, None, ops.SELF_ADD)


def test_tn_descriptor_default_is_exact_layout():
    """The MN-major descriptor defaults must reproduce a^T b (tools/sweep_tn_descriptor.py finds them)."""
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(4096, 128, generator=g).cuda(), torch.randn(4096, 128, generator=g).cuda()
    ref = a.double().t() @ b.double()
    _tc_close(ops.debug_gemm_tn(a, b), ref.cpu())


# ---- thin-contraction kernels (K <= 8): GIN layer 0 and the emb-8 default config -------------------
@pytest.mark.parametrize("rows,k,n,d0", [(5000, 6, 128, 3), (19600, 6, 8, 3), (1601, 8, 8, 0), (333, 3, 64, 1),
                                         (70000, 6, 128, 3), (1, 6, 16, 3)])
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE])
def test_thin_layer_forward_backward(rows, k, n, d0, act):
    """One pass computes dW, db, dalpha and d(eps) = sum (dz W)[:, d0:] * dot_x without dx."""
    g = torch.Generator().manual_seed(rows + n + k)
    x = torch.randn(rows, k, generator=g, dtype=torch.float64, requires_grad=True)
    W = (torch.randn(n, k, generator=g, dtype=torch.float64) / k ** 0.5).requires_grad_(True)
    b = torch.randn(n, generator=g, dtype=torch.float64, requires_grad=True)
    alpha = torch.tensor([0.25], dtype=torch.float64, requires_grad=True)
    dot_x = torch.randn(rows, k - d0, generator=g, dtype=torch.float64)
    z_ref = x @ W.t() + b
    o_ref = torch.nn.functional.prelu(z_ref, alpha) if act == ops.ACT_PRELU else z_ref
    gout = torch.randn(rows, n, generator=g, dtype=torch.float64)
    o_ref.backward(gout)
    f32 = lambda t: t.detach().float().cuda()
    z, o = ops.linear_fwd(f32(x), f32(W), f32(b), act=act, alpha=f32(alpha))
    torch.testing.assert_close(z.cpu().double(), z_ref.detach(), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(o.cpu().double(), o_ref.detach(), rtol=RTOL, atol=ATOL)
    r = ops.linear_bwd(f32(gout), z, f32(x), f32(W), act=act, alpha=f32(alpha), dx_cols=(d0, k), want_dx=False,
                       dot_x=f32(dot_x), want_dalpha=act == ops.ACT_PRELU)
    scale = rows ** 0.5
    assert r["dx"] is None
    torch.testing.assert_close(r["dW"].cpu().double(), W.grad, rtol=RTOL, atol=ATOL * scale)
    torch.testing.assert_close(r["db"].cpu().double(), b.grad, rtol=RTOL, atol=ATOL * scale)
    torch.testing.assert_close(r["ddot"].cpu().double(), (x.grad[:, d0:] * dot_x).sum().view(1), rtol=1e-4,
                               atol=ATOL * scale * (n * k) ** 0.5)
    if act == ops.ACT_PRELU:
        torch.testing.assert_close(r["dalpha"].cpu().double(), alpha.grad, rtol=1e-4, atol=ATOL * scale * n ** 0.5)
    r2 = ops.linear_bwd(f32(gout), z, f32(x), f32(W), act=act, alpha=f32(alpha), dx_cols=(d0, k), want_dx=False,
                        dot_x=f32(dot_x), want_dalpha=act == ops.ACT_PRELU)
    for key in ("dW", "db", "ddot"):
        assert torch.equal(r[key], r2[key]), key


@pytest.mark.parametrize("k", [32, 128, 24])
def test_head_layer(k):
    """k -> 1 readout head (models.py:328): streaming head kernels when k/4 is a power of two,
    otherwise the SIMT engine with n = 1 on the narrow side of the weight-gradient tile."""
    g = torch.Generator().manual_seed(9)
    rows, n = 30000, 1
    x = torch.randn(rows, k, generator=g, dtype=torch.float64)
    W = torch.randn(n, k, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(n, generator=g, dtype=torch.float64, requires_grad=True)
    gout = torch.randn(rows, n, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    (xr @ W.t() + b).backward(gout)
    f32 = lambda t: t.detach().float().cuda()
    z, o = ops.linear_fwd(f32(x), f32(W), f32(b))
    torch.testing.assert_close(o.cpu().double(), (x @ W.t() + b).detach(), rtol=RTOL, atol=ATOL)
    r = ops.linear_bwd(f32(gout), None, f32(x), f32(W), act=ops.ACT_NONE)
    torch.testing.assert_close(r["dW"].cpu().double(), W.grad, rtol=RTOL, atol=ATOL * rows ** 0.5)
    torch.testing.assert_close(r["db"].cpu().double(), b.grad, rtol=RTOL, atol=ATOL * rows ** 0.5)
    torch.testing.assert_close(r["dx"].cpu().double(), xr.grad, rtol=RTOL, atol=ATOL)


def _block_diagonal_relation(sizes_in, sizes_out, deg, seed, sort_rows=True):
    """A block-diagonal bipartite relation: block b has sizes_in[b] input rows and sizes_out[b] output rows, every input row
    sends `deg`-ish edges to random output rows of ITS block; COO grouped by input row (as the reference ships it)."""
    g = torch.Generator().manual_seed(seed)
    src, dst = [], []
    i0 = o0 = 0
    for ni, no in zip(sizes_in, sizes_out):
        if ni and no:
            k = torch.randint(1, 2 * deg, (ni,), generator=g)
            s = torch.repeat_interleave(torch.arange(ni), k) + i0
            d = torch.randint(0, no, (int(k.sum()),), generator=g) + o0
            src.append(s), dst.append(d)
        i0, o0 = i0 + ni, o0 + no
    ei = torch.stack((torch.cat(src), torch.cat(dst)))
    if not sort_rows:       # shuffle the edge order: stable destination order is then NOT ascending in the input id
        ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
    ptr_in = torch.tensor([0] + list(torch.tensor(sizes_in).cumsum(0)), dtype=torch.int64)
    ptr_out = torch.tensor([0] + list(torch.tensor(sizes_out).cumsum(0)), dtype=torch.int64)
    return ei, ptr_in, ptr_out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("f", [128, 64, 32])
@pytest.mark.parametrize("case", ["plain", "self_add", "accumulate", "pre_act", "unsorted", "oversize_block", "ragged",
                                  "many_blocks"])
@pytest.mark.parametrize("schedule", ["staged", "streaming"])
def test_streaming_long_row_aggregation_is_bit_identical_to_the_gather(case, f, dtype, schedule, monkeypatch):
    """hgin_gin_combine_blocks_t (input-major streaming over a block-diagonal batch, csrc/gin_scatter_blocks.cuh) against
    hgin_gin_combine_t on the same inputs: torch.equal in every mode it takes, including the two gated fall-backs (edge
    order not ascending within output rows; a block with more output rows than the accumulator tile holds)."""
    from gnn_link_prediction_b200.functional import GraphCSR
    # staged: source rows through shared memory, accumulators in registers (hgin_gin_combine_staged_t);
    # streaming: the opt-in input-major schedule with shared-memory accumulators (hgin_gin_combine_blocks_t)
    monkeypatch.setattr(ops, "STREAM_LONG_ROWS", schedule == "streaming")
    monkeypatch.setattr(ops, "STAGE_LONG_ROWS", schedule == "staged")
    if case == "oversize_block":
        sizes_in, sizes_out = [300, 40000, 500], [20, 3000, 30]
    elif case == "ragged":
        sizes_in, sizes_out = [1, 0, 77, 33, 0, 5, 1000], [3, 4, 1, 40, 0, 2, 130]
    elif case == "many_blocks":      # more blocks than SMs: every CTA walks several blocks through its two stages
        sizes_in, sizes_out = [700, 0, 1100, 217] * 100, [100, 3, 224, 9] * 100
    else:
        sizes_in, sizes_out = [700, 650, 900, 31, 64], [60, 50, 80, 7, 33]
    ei, ptr_in, ptr_out = _block_diagonal_relation(sizes_in, sizes_out, 3, 5, sort_rows=case != "unsorted")
    n_in, n_out = int(ptr_in[-1]), int(ptr_out[-1])
    et = ("a", "to", "b")
    graph = GraphCSR({et: ei.cuda()}, {"a": n_in, "b": n_out}, blocks={"a": ptr_in.cuda(), "b": ptr_out.cuda()})
    plan = graph.stream_plan(et, "fwd")
    assert plan is not None
    gate = plan.gate.cpu()
    assert int(gate[0]) == 0 and (int(gate[3]) == 0) == (case != "unsorted")
    assert int(gate[1]) == max(sizes_out) and int(gate[2]) == max(sizes_in)
    torch.manual_seed(1)
    x = torch.randn(n_in, f, device="cuda").to(dtype)
    xs = torch.randn(n_out, f, device="cuda").to(dtype)
    eps = torch.tensor([0.3], device="cuda")
    alpha = torch.tensor([0.25], device="cuda")
    kw = {}
    if case in ("self_add", "accumulate", "pre_act", "ragged", "many_blocks"):
        kw.update(x_self=xs, eps=eps, self_mode=ops.SELF_ADD)
    if case in ("pre_act", "many_blocks"):
        kw.update(src_act=(ops.ACT_PRELU, alpha), self_act=(ops.ACT_RELU, None))

    def run(stream):
        out = None
        if case == "accumulate":
            out = torch.full((n_out, f), 0.5, device="cuda").to(dtype)
        return ops.gin_combine(graph.fwd(et), x, out=out, accumulate=out is not None, block_plan=stream, **kw)

    want = run(None)
    got = run(plan)
    assert torch.equal(got, want)
    # and the transposed direction (the K4 pass): outputs are the `a` rows, inputs the `b` rows
    plan_t = graph.stream_plan(et, "bwd")
    if plan_t is not None:      # only long rows get a plan
        g = torch.randn(n_out, f, device="cuda").to(dtype)
        assert torch.equal(ops.gin_combine(graph.bwd(et), g, block_plan=plan_t), ops.gin_combine(graph.bwd(et), g))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("f", [128, 64])
@pytest.mark.parametrize("case", ["plain", "self_add", "accumulate", "pre_act", "post", "post_relu", "oversize_block", "leaky_block",
                                  "ragged", "many_blocks"])
def test_short_row_table_aggregation_is_bit_identical_to_the_gather(case, f, dtype, monkeypatch):
    """hgin_gin_combine_table_t (source rows of a block staged in shared memory, one CTA per SM) against hgin_gin_combine_t on
    the same inputs: torch.equal on the rows in every mode, including the gated fall-backs (a block whose source rows do not
    fit the table; an edge that leaves its block); the PReLU-slope and eps sums to fp32 rounding."""
    from gnn_link_prediction_b200.functional import GraphCSR
    monkeypatch.setattr(ops, "TABLE_SHORT_ROWS", True)      # opt-in schedule (the global-gather kernel is the default)
    if case == "oversize_block":
        sizes_in, sizes_out = [20, 3000, 30], [300, 4000, 500]
    elif case == "ragged":
        sizes_in, sizes_out = [3, 4, 1, 40, 0, 2, 130], [1, 0, 77, 33, 0, 5, 1000]
    elif case == "many_blocks":
        sizes_in, sizes_out = [200] * 400, [611] * 400
    else:
        sizes_in, sizes_out = [60, 50, 80, 7, 33], [700, 650, 900, 31, 64]
    # relation a -> b where every a row has ~3 b-neighbours of its own block: the transposed gather (rows = a) is the short side
    ei, ptr_a, ptr_b = _block_diagonal_relation(sizes_out, sizes_in, 3, 11)
    if case == "leaky_block":
        ei[1, 5] = int(ptr_b[-1]) - 1
    n_a, n_b = int(ptr_a[-1]), int(ptr_b[-1])
    et = ("a", "to", "b")
    graph = GraphCSR({et: ei.cuda()}, {"a": n_a, "b": n_b}, blocks={"a": ptr_a.cuda(), "b": ptr_b.cuda()})
    plan = graph.stream_plan(et, "bwd")          # outputs = a rows (short), inputs = b rows
    assert plan is not None
    gate = plan.gate.cpu()
    assert (int(gate[0]) == 0) == (case != "leaky_block") and int(gate[2]) == max(sizes_in)
    csr = graph.bwd(et)
    assert csr.num_edges <= 8 * csr.num_rows
    torch.manual_seed(2)
    x = torch.randn(n_b, f, device="cuda").to(dtype)
    xs = torch.randn(n_a, f, device="cuda").to(dtype)
    z = torch.randn(n_a, f, device="cuda").to(dtype)
    eps = torch.tensor([0.3], device="cuda")
    alpha = torch.tensor([0.25], device="cuda")
    kw = {}
    if case in ("self_add", "accumulate", "pre_act", "ragged", "post", "post_relu", "many_blocks"):
        kw.update(x_self=xs, eps=eps, self_mode=ops.SELF_ADD)
    if case in ("pre_act", "many_blocks"):
        kw.update(src_act=(ops.ACT_PRELU, alpha), self_act=(ops.ACT_RELU, None))

    def run(p):
        out = None
        if case == "accumulate":
            out = torch.full((n_a, f), 0.5, device="cuda").to(dtype)
        post = None
        if case == "post":
            post = ops.PostAct(z, ops.ACT_PRELU, alpha)
        if case == "post_relu":
            post = ops.PostAct(z, ops.ACT_RELU, None)
        res = ops.gin_combine(csr, x, out=out, accumulate=out is not None, block_plan=p, post=post,
                              want_ddot=post is not None, **kw)
        if post is not None:
            return res[0], res[1], post.dalpha
        return res, None, None

    want, ddot_w, dal_w = run(None)
    got, ddot_g, dal_g = run(plan)
    assert torch.equal(got, want)
    if ddot_w is not None:
        torch.testing.assert_close(ddot_g, ddot_w, rtol=1e-4, atol=1e-2)
    if dal_w is not None:
        torch.testing.assert_close(dal_g, dal_w, rtol=1e-4, atol=1e-2)


def test_streaming_plan_only_for_long_rows_and_block_tables(monkeypatch):
    from gnn_link_prediction_b200.functional import GraphCSR
    et = ("a", "to", "b")
    ei, ptr_in, ptr_out = _block_diagonal_relation([200, 300], [20, 30], 3, 0)
    off = GraphCSR({et: ei.cuda()}, {"a": 500, "b": 50}, blocks={"a": ptr_in.cuda(), "b": ptr_out.cuda()})
    assert off.stream_plan(et, "fwd") is None                    # opt-in: off by default
    monkeypatch.setattr(ops, "STREAM_LONG_ROWS", True)
    ei, ptr_in, ptr_out = _block_diagonal_relation([200, 300], [20, 30], 3, 0)
    et = ("a", "to", "b")
    no_blocks = GraphCSR({et: ei.cuda()}, {"a": 500, "b": 50})
    assert no_blocks.stream_plan(et, "fwd") is None
    with_blocks = GraphCSR({et: ei.cuda()}, {"a": 500, "b": 50}, blocks={"a": ptr_in.cuda(), "b": ptr_out.cuda()})
    assert with_blocks.stream_plan(et, "fwd") is not None        # ~30 inputs per output row
    assert with_blocks.stream_plan(et, "bwd") is None            # ~3 outputs per input row: the gather kernel's case
    monkeypatch.setattr(ops, "TABLE_SHORT_ROWS", True)
    again = GraphCSR({et: ei.cuda()}, {"a": 500, "b": 50}, blocks={"a": ptr_in.cuda(), "b": ptr_out.cuda()})
    assert again.stream_plan(et, "bwd") is not None              # ... unless the shared-memory table variant is opted in
    # an edge that leaves its block closes the gate (the gather kernel then does the work)
    bad = ei.clone()
    bad[1, 0] = 45
    g = GraphCSR({et: bad.cuda()}, {"a": 500, "b": 50}, blocks={"a": ptr_in.cuda(), "b": ptr_out.cuda()})
    assert int(g.stream_plan(et, "fwd").gate[0]) > 0


@pytest.mark.parametrize("math", ["tf32", "bf16"])
@pytest.mark.parametrize("k,n,k2", [(256, 256, 0), (256, 128, 3), (128, 256, 0), (384, 192, 0)])
def test_layers_wider_than_one_tensor_core_tile_run_as_128_blocks(k, n, k2, math):
    """hidden 256: ops.linear_fwd / linear_bwd tile K and N into 128-blocks of the tcgen05 kernels (instead of falling to
    the fp32 SIMT engine); checked against fp64 at the reduced-precision bar."""
    torch.manual_seed(k + n)
    mode = ops.MATH_TF32 if math == "tf32" else ops.MATH_BF16
    dt = torch.float32 if math == "tf32" else torch.bfloat16
    rows = 1000
    x = torch.randn(rows, k, device="cuda").to(dt)
    x2 = torch.randn(rows, k2, device="cuda") if k2 else None
    W = torch.randn(n, k + k2, device="cuda") / (k ** 0.5)
    b = torch.randn(n, device="cuda")
    alpha = torch.tensor([0.25], device="cuda")
    g = torch.randn(rows, n, device="cuda").to(dt)
    xin = torch.cat((x.double(), x2.double()), 1) if k2 else x.double()
    z_ref = xin @ W.double().t() + b.double()
    o_ref = torch.where(z_ref > 0, z_ref, 0.25 * z_ref)
    z, o = ops.linear_fwd(x, W, b, x2=x2, act=ops.ACT_PRELU, alpha=alpha, want_z=True, math_mode=mode, out_dtype=dt)
    tol = 5e-3 if math == "tf32" else 3e-2

    def rel(a, r):
        return float((a.double() - r).norm() / r.norm())

    assert rel(z, z_ref) < tol and rel(o, o_ref) < tol
    r = ops.linear_bwd(g, z, x, W, x2=x2, act=ops.ACT_PRELU, alpha=alpha, dx_cols=(0, k), want_dx=True, want_dw=True,
                       want_db=True, want_dalpha=True, math_mode=mode)
    dz_ref = g.double() * torch.where(z.double() > 0, 1.0, 0.25)
    assert rel(r["dx"], dz_ref @ W.double()[:, :k]) < tol
    assert rel(r["dW"], dz_ref.t() @ xin) < tol
    assert rel(r["db"], dz_ref.sum(0)) < tol
    da_ref = (g.double() * torch.clamp(z.double(), max=0)).sum()
    assert abs(float(r["dalpha"]) - float(da_ref)) < tol * float((g.double() * z.double()).abs().sum()) ** 0.5 + tol * abs(float(da_ref))
