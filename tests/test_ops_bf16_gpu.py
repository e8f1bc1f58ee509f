"""Dense-layer kernels on bf16 rows (HGIN_DTYPE_BF16 / HGIN_MATH_BF16): hgin_linear_fwd_t / hgin_linear_bwd_t.

* tensor-core kernels (tcgen05 kind::f16, bf16 operands, fp32 accumulate) against a float64 reference evaluated on the
  SAME bf16-rounded operands (x, W rounded to bf16 exactly as the kernel sees them): what remains is fp32 accumulation
  order and the single rounding of the stored result, so the bar is rtol 1e-2 / atol 1e-2 * max|ref| (bf16 keeps 8
  mantissa bits: one rounding is <= 2^-9 relative) — the north star's bound for reduced-precision GEMMs;
* the K <= 8 and n = 1 streaming kernels against the fp32 kernels on the same inputs: same fp32 arithmetic in the same
  order, rounded once on the store -> bit-exact after rounding;
* the MN-major bf16 descriptor of the weight-gradient kernel pinned with integer data (exact in bf16 and fp32)."""
import pytest
import torch

from gnn_link_prediction_b200 import ops

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _close(got, want, rel=1e-2):
    want = want.detach().double().cpu()
    torch.testing.assert_close(got.detach().double().cpu(), want, rtol=rel, atol=rel * float(want.abs().max()) + 1e-9)


def _r(t):
    """bf16 rounding as the kernels see an operand, kept in float64 for the reference."""
    return t.to(BF).double()


TC_SHAPES = [(4096, 128, 0, 128), (1000, 128, 3, 128), (130, 128, 0, 32), (20000, 64, 0, 128), (777, 32, 0, 16),
             (5000, 128, 0, 64), (3001, 96, 2, 48), (129, 16, 0, 112)]


@pytest.mark.parametrize("rows,k1,k2,n", TC_SHAPES)
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE])
def test_linear_fwd_bf16(rows, k1, k2, n, act):
    g = torch.Generator().manual_seed(rows + n)
    x1, x2 = torch.randn(rows, k1, generator=g), (torch.randn(rows, k2, generator=g) if k2 else None)
    W, b = torch.randn(n, k1 + k2, generator=g) / (k1 + k2) ** 0.5, torch.randn(n, generator=g)
    alpha = torch.tensor([0.25])
    z_ref = _r(x1) @ _r(W[:, :k1]).t() + b.double()
    if k2:
        z_ref = z_ref + x2.double() @ W[:, k1:].double().t()        # the raw input columns and their weights stay fp32
    o_ref = torch.where(z_ref > 0, z_ref, 0.25 * z_ref) if act == ops.ACT_PRELU else z_ref
    args = (x1.cuda().to(BF), W.cuda(), b.cuda())
    kw = dict(x2=None if x2 is None else x2.cuda(), act=act, alpha=alpha.cuda(), math_mode=ops.MATH_BF16)
    z, o = ops.linear_fwd(*args, **kw)
    assert z.dtype == BF and o.dtype == BF
    _close(z, z_ref)
    _close(o, o_ref)
    zo, none = ops.linear_fwd(*args, want_out=False, **kw)          # lazily activated layer: z only
    assert none is None and torch.equal(zo, z)
    prev = torch.randn(rows, n, generator=g).to(BF)
    acc = prev.clone().cuda()
    ops.linear_fwd(*args, out=acc, accumulate_out=True, want_z=False, **kw)
    _close(acc, prev.double() + o_ref)
    z2, o2 = ops.linear_fwd(*args, **kw)
    assert torch.equal(z, z2) and torch.equal(o, o2)                # deterministic


@pytest.mark.parametrize("rows,k1,k2,n", TC_SHAPES)
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE])
def test_linear_bwd_bf16(rows, k1, k2, n, act):
    g = torch.Generator().manual_seed(rows * 3 + n)
    k = k1 + k2
    d = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float64)
    x1, x2 = _r(d(rows, k1)), (d(rows, k2).float().double() if k2 else None)
    W = d(n, k) / k ** 0.5
    alpha = torch.tensor([0.25], dtype=torch.float64)
    gout, z = _r(d(rows, n)), _r(d(rows, n))
    dot_x = _r(d(rows, k1))
    if act == ops.ACT_PRELU:
        dz_exact = torch.where(z > 0, gout, alpha * gout)
        dalpha_ref = (gout * torch.where(z > 0, torch.zeros_like(z), z)).sum().view(1)
        dz = _r(dz_exact)                      # the kernel stores dz as bf16 before the two GEMMs read it
    else:
        dz_exact = dz = gout
    dx_ref = dz @ _r(W[:, :k1])
    dW1_ref = dz.t() @ x1
    db_ref = dz_exact.sum(0) if act == ops.ACT_PRELU else dz.sum(0)
    cu = lambda t, dt: None if t is None else t.to(dt).cuda().contiguous()
    r = ops.linear_bwd(cu(gout, BF), cu(z, BF), cu(x1, BF), cu(W, torch.float32), x2=cu(x2, torch.float32), act=act,
                       alpha=cu(alpha, torch.float32), dx_cols=(0, k1), dot_x=cu(dot_x, BF), want_dalpha=act == ops.ACT_PRELU,
                       math_mode=ops.MATH_BF16)
    assert r["dx"].dtype == BF and r["dW"].dtype == torch.float32
    _close(r["dx"], dx_ref)
    _close(r["dW"][:, :k1], dW1_ref, rel=2e-3)
    if k2:
        _close(r["dW"][:, k1:], dz_exact.t() @ x2, rel=2e-3)
    _close(r["db"], db_ref, rel=2e-3)
    scale = float(dx_ref.abs().mean() * dot_x.abs().mean()) * (rows * k1) ** 0.5
    assert abs(float(r["ddot"]) - float((dx_ref * dot_x).sum())) <= 2e-2 * scale
    if act == ops.ACT_PRELU:
        _close(r["dalpha"], dalpha_ref, rel=1e-3)
    r2 = ops.linear_bwd(cu(gout, BF), cu(z, BF), cu(x1, BF), cu(W, torch.float32), x2=cu(x2, torch.float32), act=act,
                        alpha=cu(alpha, torch.float32), dx_cols=(0, k1), math_mode=ops.MATH_BF16)
    for key in ("dx", "dW", "db"):
        assert torch.equal(r[key], r2[key]), f"{key} not deterministic"


@pytest.mark.parametrize("rows,k,n", [(5000, 128, 128), (777, 64, 128), (4097, 128, 32)])
@pytest.mark.parametrize("post_act", [ops.ACT_PRELU, ops.ACT_RELU])
@pytest.mark.parametrize("self_branch", [False, True])
def test_linear_bwd_bf16_post_activation_and_self_branch(rows, k, n, post_act, self_branch):
    """dx leaves the input-gradient GEMM as [(1 + eps) *] (dz W) * act'(z0) in bf16, with dalpha0 and d(eps)."""
    g = torch.Generator().manual_seed(rows + k + n)
    d = lambda *sh: torch.randn(*sh, generator=g, dtype=torch.float64)
    z0, dz, W = _r(d(rows, k)), _r(d(rows, n)), d(n, k) / k ** 0.5
    a0, eps = 0.2, 0.37
    x_dst = torch.where(z0 > 0, z0, (a0 if post_act == ops.ACT_PRELU else 0.0) * z0)
    dh = dz @ _r(W)
    scale_f = float(torch.tensor(1.0 + eps, dtype=torch.float32)) if self_branch else 1.0
    rr = scale_f * dh
    dx_ref = torch.where(z0 > 0, rr, (a0 if post_act == ops.ACT_PRELU else 0.0) * rr)
    cu = lambda t, dt: t.to(dt).cuda().contiguous()
    t32 = lambda v: torch.tensor([v], device="cuda")
    post = ops.PostAct(cu(z0, BF), post_act, t32(a0))
    h = cu(_r(d(rows, k)), BF)
    r = ops.linear_bwd(cu(dz, BF), None, h, cu(W, torch.float32), act=ops.ACT_NONE, math_mode=ops.MATH_BF16, post=post,
                       self_eps=t32(eps) if self_branch else None, want_self_ddot=self_branch)
    assert post.applied and r["dx"].dtype == BF
    _close(r["dx"], dx_ref)
    _close(r["dW"], dz.t() @ h.cpu().double(), rel=2e-3)
    scale = float(dh.abs().mean() * z0.abs().mean()) * (rows * k) ** 0.5
    if post_act == ops.ACT_PRELU:
        want = (rr * torch.where(z0 > 0, torch.zeros_like(z0), z0)).sum()
        assert abs(float(post.dalpha) - float(want)) <= 2e-2 * scale
    if self_branch:
        assert abs(float(r["ddot"]) - float((dh * x_dst).sum())) <= 2e-2 * scale


def test_tn_bf16_descriptor_is_exact_layout():
    """Small integers are exact in bf16 and their products / sums exact in fp32: the MN-major bf16 operand layout of
    the weight-gradient kernel must reproduce a^T b EXACTLY; on failure the message lists the descriptor settings
    that do (sweep over LBO / SBO / layout type / K-step)."""
    g = torch.Generator().manual_seed(1)
    for rows, n, k in [(4096, 128, 128), (1000, 64, 128), (333, 128, 32), (64, 16, 16)]:
        a = torch.randint(-3, 4, (rows, n), generator=g).float()
        b = torch.randint(-3, 4, (rows, k), generator=g).float()
        ref = (a.double().t() @ b.double()).float()
        got = ops.debug_gemm_tn_bf16(a.cuda().to(BF), b.cuda().to(BF)).cpu()
        if not torch.equal(got, ref):
            good = []
            for layout in (2, 1, 4, 6):
                for lbo in (8192, 1024, 128, 16):
                    for sbo in (1024, 512, 2048, 8192, 128):
                        for kstep in (2048, 1024, 4096, 32, 256):
                            try:
                                t = ops.debug_gemm_tn_bf16(a.cuda().to(BF), b.cuda().to(BF), lbo, sbo, layout, kstep).cpu()
                            except ops.HginError:
                                continue
                            if torch.equal(t, ref):
                                good.append((layout, lbo, sbo, kstep))
            pytest.fail(f"default MN-major bf16 descriptor is wrong for {(rows, n, k)}; exact settings "
                        f"(layout, lbo, sbo, k_step): {good}")


# ---- K <= 8 layers and the n = 1 head: bit-exact with the fp32 kernels after one rounding -----------------------
@pytest.mark.parametrize("rows,k,n,d0", [(5000, 6, 128, 3), (19600, 6, 8, 3), (1601, 8, 8, 0), (333, 3, 64, 1), (70000, 6, 128, 3)])
@pytest.mark.parametrize("act", [ops.ACT_PRELU, ops.ACT_NONE])
def test_thin_layer_with_bf16_wide_side(rows, k, n, d0, act):
    g = torch.Generator().manual_seed(rows + n + k)
    x = torch.randn(rows, k, generator=g).cuda()
    W, b = (torch.randn(n, k, generator=g) / k ** 0.5).cuda(), torch.randn(n, generator=g).cuda()
    alpha = torch.tensor([0.25], device="cuda")
    z32, o32 = ops.linear_fwd(x, W, b, act=act, alpha=alpha)
    z, o = ops.linear_fwd(x, W, b, act=act, alpha=alpha, out_dtype=BF, math_mode=ops.MATH_BF16)
    assert z.dtype == BF and torch.equal(z, z32.to(BF)) and torch.equal(o, o32.to(BF))
    prev = torch.randn(rows, n, generator=g).to(BF).cuda()
    acc, acc32 = prev.clone(), prev.float()
    ops.linear_fwd(x, W, b, act=act, alpha=alpha, out=acc, accumulate_out=True, want_z=False, math_mode=ops.MATH_BF16)
    ops.linear_fwd(x, W, b, act=act, alpha=alpha, out=acc32, accumulate_out=True, want_z=False)
    assert torch.equal(acc, acc32.to(BF))
    gout = torch.randn(rows, n, generator=g).to(BF).cuda()
    dot_x = torch.randn(rows, k - d0, generator=g).cuda()
    kw = dict(act=act, alpha=alpha, dx_cols=(d0, k), want_dx=False, dot_x=dot_x, want_dalpha=act == ops.ACT_PRELU)
    r = ops.linear_bwd(gout, z, x, W, math_mode=ops.MATH_BF16, **kw)
    r32 = ops.linear_bwd(gout.float(), z.float(), x, W, **kw)
    # same products; the row blocks (hence the association of the per-thread sums) differ between the two types
    for key in ("dW", "db", "ddot") + (("dalpha",) if act == ops.ACT_PRELU else ()):
        torch.testing.assert_close(r[key], r32[key], rtol=1e-4, atol=1e-4 * float(r32[key].abs().max()) + 1e-6, msg=key)


@pytest.mark.parametrize("k", [32, 128])
@pytest.mark.parametrize("post_act", [ops.ACT_NONE, ops.ACT_PRELU])
def test_head_layer_with_bf16_rows(k, post_act):
    g = torch.Generator().manual_seed(9)
    rows = 30000
    x = torch.randn(rows, k, generator=g).to(BF).cuda()
    W, b = torch.randn(1, k, generator=g).cuda(), torch.randn(1, generator=g).cuda()
    gout = torch.randn(rows, 1, generator=g).cuda()
    z, o = ops.linear_fwd(x, W, b, out_dtype=torch.float32, math_mode=ops.MATH_BF16)
    z32, o32 = ops.linear_fwd(x.float(), W, b)
    assert o.dtype == torch.float32 and torch.equal(o, o32)
    a0 = torch.tensor([0.3], device="cuda")
    zp = torch.randn(rows, k, generator=g).to(BF).cuda()
    post = ops.PostAct(zp, post_act, a0) if post_act != ops.ACT_NONE else None
    post32 = ops.PostAct(zp.float(), post_act, a0) if post_act != ops.ACT_NONE else None
    r = ops.linear_bwd(gout, None, x, W, act=ops.ACT_NONE, math_mode=ops.MATH_BF16, post=post)
    r32 = ops.linear_bwd(gout, None, x.float(), W, act=ops.ACT_NONE, post=post32)
    assert r["dx"].dtype == BF and torch.equal(r["dx"], r32["dx"].to(BF))
    assert torch.equal(r["dW"], r32["dW"]) and torch.equal(r["db"], r32["db"])
    if post is not None:
        assert post.applied and torch.equal(post.dalpha, post32.dalpha)


def test_bf16_combinations_without_a_kernel_run_in_fp32_through_casts():
    """Odd shapes (k1 = 11: the readout of the emb-8 default config) have no bf16 kernel: ops falls back to the fp32
    kernels on cast copies and returns the requested storage type."""
    g = torch.Generator().manual_seed(3)
    rows, k, n = 3000, 11, 128
    x = torch.randn(rows, k, generator=g).to(BF).cuda()
    W, b = torch.randn(n, k, generator=g).cuda(), torch.randn(n, generator=g).cuda()
    z, o = ops.linear_fwd(x, W, b, act=ops.ACT_RELU, math_mode=ops.MATH_BF16)
    z32, o32 = ops.linear_fwd(x.float(), W, b, act=ops.ACT_RELU)
    assert z.dtype == BF and torch.equal(z, z32.to(BF)) and torch.equal(o, o32.to(BF))
    gout = torch.randn(rows, n, generator=g).to(BF).cuda()
    r = ops.linear_bwd(gout, z, x, W, act=ops.ACT_RELU, math_mode=ops.MATH_BF16)
    r32 = ops.linear_bwd(gout.float(), z.float(), x.float(), W, act=ops.ACT_RELU)
    assert r["dx"].dtype == BF and torch.equal(r["dx"], r32["dx"].to(BF)) and torch.equal(r["dW"], r32["dW"])
