"""Parity tests proper: every CUDA operator, called through the C ABI (ctypes → libpcfa_b200.so),
against the oracle on seeded inputs and against the committed reference outputs (tests/golden)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, kw_of

pytestmark = pytest.mark.gpu

TOL = dict(rtol=1e-3, atol_rms=1e-3)       # BASELINE.json: rtol 1e-3 on fp32 cost volumes and flows
TIGHT = dict(rtol=1e-4, atol_rms=1e-4)
# Gradients through the tensor-core backward: the gradient pyramid enters the MMA truncated to TF32 (2^-11 per element,
# rel-L2 3e-4 on the result, scripts/bwd_precision.py), so the element-wise maximum over ~1e4-1e6 entries reaches
# ~4.5 sigma = 1.5e-3 rms.  The fp32 SIMT backward (PCFA_CORR_IMPL=1) is tested at TOL against the same oracle.
GTOL = dict(rtol=1e-3, atol_rms=4e-3)


def cu(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    return t.requires_grad_(True) if grad else t


def npy(t):
    return t.detach().float().cpu().numpy()


# ----------------------------------------------------------------------------------- CorrBlock
@pytest.mark.parametrize("impl", [1, 0])
def test_corrblock_matches_reference_outputs(golden, impl, monkeypatch):
    from pcfa_b200.corr_block import CorrBlock
    monkeypatch.setenv("PCFA_CORR_IMPL", str(impl))
    z = golden("corrblock")
    f1, f2 = cu(z["fmap1"], True), cu(z["fmap2"], True)
    blk = CorrBlock(f1, f2, num_levels=4, radius=4)
    N = f1.shape[0] * f1.shape[2] * f1.shape[3]
    assert [tuple(p.shape) for p in blk.corr_pyramid] == [(N, 1, 16, 24), (N, 1, 8, 12), (N, 1, 4, 6), (N, 1, 2, 3)]
    for l in (1, 2, 3):
        assert_close(npy(blk.corr_pyramid[l]), z[f"level{l}"], what=f"level {l}", **TOL)
    assert_close(npy(blk.corr_pyramid[0])[::37], z["level0_rows"], what="level 0", **TOL)
    out_a = blk(cu(z["coords_a"]))
    out_b = blk(cu(z["coords_b"]))
    assert out_a.is_contiguous() and out_a.dtype == torch.float32
    assert_close(npy(out_a), z["out_a"], what="lookup a", **TOL)
    assert_close(npy(out_b), z["out_b"], what="lookup b (out-of-range taps)", **TOL)
    (out_a * cu(z["gout_a"])).sum().backward(retain_graph=True)
    assert_close(npy(f1.grad), z["g1_a"], what="g fmap1", **TOL)
    assert_close(npy(f2.grad), z["g2_a"], what="g fmap2", **TOL)
    (out_b * cu(z["gout_b"])).sum().backward()
    assert_close(npy(f1.grad), z["g1_ab"], what="g fmap1 accumulated", **TOL)
    assert_close(npy(f2.grad), z["g2_ab"], what="g fmap2 accumulated", **TOL)


@pytest.mark.parametrize("shape,levels,radius", [((2, 32, 18, 27), 4, 4), ((1, 64, 9, 17), 3, 3),
                                                 ((3, 16, 8, 8), 2, 4), ((1, 128, 23, 40), 4, 4),
                                                 # every level width % 4 == 0: the 128-bit lookup kernels
                                                 ((1, 32, 16, 32), 4, 4), ((2, 16, 12, 16), 3, 3), ((1, 16, 17, 48), 3, 4)])
def test_corrblock_vs_oracle_ragged_shapes(shape, levels, radius):
    from oracle import ops as O
    from pcfa_b200.corr_block import CorrBlock
    g = np.random.default_rng(sum(shape))
    B, C, H, W = shape
    f1 = g.standard_normal(shape).astype(np.float32)
    f2 = g.standard_normal(shape).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    coords = (np.stack([xs, ys])[None] + 4 * g.standard_normal((B, 2, H, W))).astype(np.float32)
    t1, t2 = cu(f1, True), cu(f2, True)
    blk = CorrBlock(t1, t2, num_levels=levels, radius=radius)
    pyr = O.corr_pyramid_forward(f1, f2, levels)
    assert_close(npy(blk._flat), pyr, what="pyramid", **TOL)
    out = blk(cu(coords))
    assert_close(npy(out), O.corr_lookup_forward(pyr, coords, levels, radius), what="lookup", **TOL)
    go = g.standard_normal(out.shape).astype(np.float32)
    (out * cu(go)).sum().backward()
    gp = O.corr_lookup_backward(go, coords, levels, radius)
    g1, g2 = O.corr_pyramid_backward(gp, f1, f2, levels)
    assert_close(npy(t1.grad), g1, what="g fmap1", **GTOL)
    assert_close(npy(t2.grad), g2, what="g fmap2", **GTOL)


def test_corrblock_fp32_backward_vs_oracle(monkeypatch):
    """The exact-fp32 SIMT build/backward (impl 1) against the oracle at the cost-volume tolerance."""
    from oracle import ops as O
    from pcfa_b200.corr_block import CorrBlock
    monkeypatch.setenv("PCFA_CORR_IMPL", "1")
    g = np.random.default_rng(5)
    B, C, H, W = 1, 32, 16, 32
    f1 = g.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = g.standard_normal((B, C, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    coords = (np.stack([xs, ys])[None] + 4 * g.standard_normal((B, 2, H, W))).astype(np.float32)
    t1, t2 = cu(f1, True), cu(f2, True)
    out = CorrBlock(t1, t2, num_levels=4, radius=4)(cu(coords))
    go = g.standard_normal(out.shape).astype(np.float32)
    (out * cu(go)).sum().backward()
    g1, g2 = O.corr_pyramid_backward(O.corr_lookup_backward(go, coords, 4, 4), f1, f2, 4)
    assert_close(npy(t1.grad), g1, what="g fmap1 (fp32)", **TOL)
    assert_close(npy(t2.grad), g2, what="g fmap2 (fp32)", **TOL)


def test_corr_static_and_forward_only():
    from pcfa_b200.corr_block import CorrBlock
    g = torch.Generator().manual_seed(0)
    f1 = torch.randn(1, 16, 6, 10, generator=g).cuda()
    f2 = torch.randn(1, 16, 6, 10, generator=g).cuda()
    c = CorrBlock.corr(f1, f2)
    assert c.shape == (1, 6, 10, 1, 6, 10)
    ref = torch.matmul(f1.view(1, 16, 60).transpose(1, 2), f2.view(1, 16, 60)).view(1, 6, 10, 1, 6, 10) / 4.0
    assert_close(npy(c), npy(ref), what="CorrBlock.corr", **TOL)
    with torch.no_grad():
        blk = CorrBlock(f1, f2, num_levels=2, radius=2)
        out = blk(torch.zeros(1, 2, 6, 10).cuda())
    assert out.shape == (1, 2 * 25, 6, 10)


# ----------------------------------------------------------------------------------- SCS / PWC
def test_scs_matches_reference_outputs(golden):
    from pcfa_b200.spatial_correlation_sampler import SpatialCorrelationSampler, spatial_correlation_sample
    z = golden("scs")
    for name in ("pwc", "fn2like", "general", "dilated"):
        kw = kw_of(z, name)
        a, b = cu(z[f"{name}_in1"], True), cu(z[f"{name}_in2"], True)
        out = spatial_correlation_sample(a, b, **kw)
        assert_close(npy(out), z[f"{name}_out"], what=f"scs fwd {name}", **TOL)
        (out * cu(z[f"{name}_gout"])).sum().backward()
        assert_close(npy(a.grad), z[f"{name}_g1"], what=f"scs g1 {name}", **TOL)
        assert_close(npy(b.grad), z[f"{name}_g2"], what=f"scs g2 {name}", **TOL)
        out2 = SpatialCorrelationSampler(**kw)(a.detach(), b.detach())
        assert torch.equal(out2, out.detach())


@pytest.mark.parametrize("C,H,W", [(196, 6, 20), (128, 12, 40), (96, 24, 80), (32, 17, 37)])
def test_pwc_correlate_vs_oracle(C, H, W):
    from oracle import ops as O
    from pcfa_b200.spatial_correlation_sampler import pwc_correlate
    g = np.random.default_rng(C + H)
    a = g.standard_normal((2, C, H, W)).astype(np.float32)
    b = g.standard_normal((2, C, H, W)).astype(np.float32)
    ta, tb = cu(a, True), cu(b, True)
    out = pwc_correlate(ta, tb)
    ref = O.scs_forward(a, b, 1, 9).reshape(2, 81, H, W) / C
    assert_close(npy(out), ref, what="pwc correlate", **TOL)
    go = g.standard_normal(ref.shape).astype(np.float32)
    (out * cu(go)).sum().backward()
    g1, g2 = O.scs_backward(a, b, go.reshape(2, 9, 9, H, W) / C, 1, 9)
    assert_close(npy(ta.grad), g1, what="pwc g1", **TOL)
    assert_close(npy(tb.grad), g2, what="pwc g2", **TOL)


def test_pwc_warp_matches_reference_outputs(golden):
    from pcfa_b200.pwc_warp import pwc_warp
    z = golden("pwc_warp")
    x, f = cu(z["x"], True), cu(z["flow"], True)
    out = pwc_warp(x, f)
    assert_close(npy(out), z["out"], what="warp", **TIGHT)
    (out * cu(z["gout"])).sum().backward()
    assert_close(npy(x.grad), z["gx"], what="warp gx", **TIGHT)
    assert_close(npy(f.grad), z["gflow"], what="warp gflow", rtol=1e-3, atol_rms=1e-3)


# ----------------------------------------------------------------------------------- FlowNet2 ops
@pytest.mark.parametrize("cfg", [dict(pad=20, ks=1, md=20, s1=1, s2=2), dict(pad=4, ks=1, md=4, s1=1, s2=1),
                                 dict(pad=6, ks=3, md=4, s1=1, s2=2), dict(pad=4, ks=1, md=4, s1=2, s2=2)])
def test_fn2_correlation_vs_oracle(cfg):
    from oracle import ops as O
    from pcfa_b200.flownet2_ops import Correlation
    g = np.random.default_rng(cfg["pad"] * 7 + cfg["ks"])
    a = g.standard_normal((2, 12, 14, 19)).astype(np.float32)
    b = g.standard_normal((2, 12, 14, 19)).astype(np.float32)
    ta, tb = cu(a, True), cu(b, True)
    m = Correlation(pad_size=cfg["pad"], kernel_size=cfg["ks"], max_displacement=cfg["md"], stride1=cfg["s1"],
                    stride2=cfg["s2"], corr_multiply=1)
    out = m(ta, tb)
    ref = O.fn2corr_forward(a, b, **cfg)
    assert_close(npy(out), ref, what="fn2 corr fwd", **TOL)
    go = g.standard_normal(ref.shape).astype(np.float32)
    (out * cu(go)).sum().backward()
    g1, g2 = O.fn2corr_backward(a, b, go, **cfg)
    assert_close(npy(ta.grad), g1, what="fn2 corr g1", **TOL)
    assert_close(npy(tb.grad), g2, what="fn2 corr g2", **TOL)


def test_resample2d_and_channelnorm_vs_oracle():
    from oracle import ops as O
    from pcfa_b200.flownet2_ops import ChannelNorm, Resample2d
    g = np.random.default_rng(21)
    img = g.standard_normal((2, 3, 24, 40)).astype(np.float32)
    flow = (6 * g.standard_normal((2, 2, 24, 40))).astype(np.float32)     # leaves the image on all sides
    ti, tf = cu(img, True), cu(flow, True)
    out = Resample2d()(ti, tf)
    assert_close(npy(out), O.resample2d_forward(img, flow), what="resample2d fwd", **TIGHT)
    go = g.standard_normal(img.shape).astype(np.float32)
    (out * cu(go)).sum().backward()
    gi, gf = O.resample2d_backward(img, flow, go)
    assert_close(npy(ti.grad), gi, what="resample2d g img (trunc quirk)", **TIGHT)
    assert_close(npy(tf.grad), gf, what="resample2d g flow", **TIGHT)
    near = Resample2d(bilinear=False)(ti.detach(), tf.detach())
    assert_close(npy(near), O.resample2d_forward(img, flow, bilinear=False), what="nearest", **TIGHT)
    for C in (2, 3):
        x = g.standard_normal((2, C, 24, 40)).astype(np.float32)
        tx = cu(x, True)
        o = ChannelNorm()(tx)
        ref = O.channelnorm_forward(x)
        assert_close(npy(o), ref, what="channelnorm", **TIGHT)
        go = g.standard_normal(ref.shape).astype(np.float32)
        (o * cu(go)).sum().backward()
        assert_close(npy(tx.grad), O.channelnorm_backward(x, ref, go), what="channelnorm bwd", **TIGHT)


# ----------------------------------------------------------------------------------- objective
def test_objective_kernels_match_reference_outputs(golden):
    from pcfa_b200 import objective as J
    z = golden("objective")
    eps = 1e-7
    img1, img2 = cu(z["img1"]), cu(z["img2"])
    pred, target = cu(z["pred"]), cu(z["target"])
    gnet1, gnet2 = cu(z["gnet1"]), cu(z["gnet2"])

    def run(mode, joint, v1, v2, lt, mu, bound):
        fo = J.FusedObjective(lambda a, b: (a.sum() * 0 + b.sum() * 0) + pred, img1, img2, target, mode=mode,
                              joint=joint, pad=(0, 0), eps_box=eps, scale=255.0, delta_bound=bound, mu=mu, loss=lt)
        fo._forward_boxes(v1, v2, True)
        fo._loss(pred, True)
        g1 = torch.empty_like(v1)
        J._box_backward(v1, img1, fo.amax, fo.amin, gnet1, fo.terms, g1, False, mode, eps, 255.0)
        if joint:
            J._box_backward(v1, img2, fo.amax, fo.amin, gnet2, fo.terms, g1, True, mode, eps, 255.0)
            return fo, g1, None
        g2 = torch.empty_like(v2)
        J._box_backward(v2, img2, fo.amax, fo.amin, gnet2, fo.terms, g2, False, mode, eps, 255.0)
        return fo, g1, g2

    w1, w2 = cu(z["cov_w1"]), cu(z["cov_w2"])
    for lt in ("aee", "mse", "cosim"):
        for tag, mu, bound in (("act", 2500. / 0.005, 0.005), ("small", 5.0, 10.0)):
            fo, g1, g2 = run(J.BOX_COV, False, w1, w2, lt, mu, bound)
            np.testing.assert_allclose(float(fo.terms[0]), float(z[f"cov_{lt}_{tag}_loss"]), rtol=1e-4)
            assert_close(npy(fo.gflow), z[f"cov_{lt}_{tag}_gpred"], what=f"gpred {lt} {tag}", **TIGHT)
            assert_close(npy(g1), z[f"cov_{lt}_{tag}_gw1"], what=f"gw1 {lt} {tag}", **TOL)
            assert_close(npy(g2), z[f"cov_{lt}_{tag}_gw2"], what=f"gw2 {lt} {tag}", **TOL)
    assert_close(npy(fo.net_in1), z["cov_net1"], what="cov net_in", **TIGHT)
    assert_close(npy(fo.delta1), z["cov_d1"], what="cov delta", rtol=1e-3, atol_rms=1e-4)

    fo, g1, g2 = run(J.BOX_CLIP, False, cu(z["clip_v1"]), cu(z["clip_v2"]), "aee", 5e5, 0.005)
    np.testing.assert_allclose(float(fo.terms[0]), float(z["clip_loss"]), rtol=1e-4)
    assert_close(npy(g1), z["clip_g1"], what="clip g1", **TIGHT)
    assert_close(npy(g2), z["clip_g2"], what="clip g2", **TIGHT)
    assert_close(npy(fo.net_in1), z["clip_net1"], what="clip net", **TIGHT)

    fo, g1, _ = run(J.BOX_JOINT, True, cu(z["joint_v"]), None, "aee", 5e5, 0.005)
    np.testing.assert_allclose(float(fo.terms[0]), float(z["joint_loss"]), rtol=1e-4)
    assert_close(npy(g1), z["joint_g"], what="joint g", **TIGHT)
    assert_close(npy(fo.net_in2), z["joint_net2"], what="joint net2", **TIGHT)
    assert_close(npy(fo.delta1), z["joint_d"], what="joint delta", **TIGHT)

    fo, g1, g2 = run(J.BOX_UNIVERSAL, False, cu(z["uni_v1"]), cu(z["uni_v2"]), "aee", 5e5, 0.005)
    np.testing.assert_allclose(float(fo.terms[0]), float(z["uni_loss"]), rtol=1e-4)
    assert_close(npy(g1), z["uni_g1"], what="universal g1", **TIGHT)
    assert_close(npy(g2), z["uni_g2"], what="universal g2", **TIGHT)
    fo, g1, _ = run(J.BOX_UNIVERSAL, True, cu(z["uni_v1"]), None, "aee", 5e5, 0.005)
    np.testing.assert_allclose(float(fo.terms[0]), float(z["unij_loss"]), rtol=1e-4)
    assert_close(npy(g1), z["unij_g"], what="universal joint g", **TIGHT)


def test_reference_shaped_objective_functions(golden):
    """extract_deltas / loss_delta_constraint / scaled_input with autograd, as the reference composes them."""
    from pcfa_b200 import objective as J
    z = golden("objective")
    eps = 1e-7
    img1, img2 = cu(z["img1"]), cu(z["img2"])
    a, b = cu(z["cov_w1"], True), cu(z["cov_w2"], True)
    p = cu(z["pred"], True)
    d1, d2 = J.extract_deltas(a, b, img1, img2, "change_of_variables", eps_box=eps)
    loss = J.loss_delta_constraint(p, cu(z["target"]), d1, d2, None, delta_bound=0.005, mu=2500. / 0.005, f_type="aee")
    n1 = J.scaled_input(a, var_change=True, eps_box=eps, make_unit_input=True)
    n2 = J.scaled_input(b, var_change=True, eps_box=eps, make_unit_input=True)
    (loss + (n1 * cu(z["gnet1"])).sum() + (n2 * cu(z["gnet2"])).sum()).backward()
    np.testing.assert_allclose(float(loss), float(z["cov_aee_act_loss"]), rtol=1e-4)
    assert_close(npy(a.grad), z["cov_aee_act_gw1"], what="gw1", **TOL)
    assert_close(npy(b.grad), z["cov_aee_act_gw2"], what="gw2", **TOL)
    assert_close(npy(p.grad), z["cov_aee_act_gpred"], what="gpred", **TIGHT)
    dj = cu(z["joint_v"], True)
    e1, e2 = J.extract_deltas_joint(dj, torch.max(img1, img2), torch.min(img1, img2))
    assert e1 is e2
    assert_close(npy(e1), z["joint_d"], what="joint delta", **TIGHT)
    with pytest.raises(NotImplementedError):
        J.loss_delta_constraint(p, p, d1, d2, f_type="l1")
    with pytest.raises(ValueError):
        J.box_mode("change_of_variables", joint=True)


def test_no_cpu_fallback():
    from pcfa_b200.corr_block import CorrBlock
    from pcfa_b200.spatial_correlation_sampler import spatial_correlation_sample
    x = torch.randn(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        CorrBlock(x, x)
    with pytest.raises(RuntimeError):
        spatial_correlation_sample(x, x, patch_size=3)


# ----------------------------------------------------------------------------------- tcgen05 path
def _build_pyramid(f1, f2, levels, impl):
    from pcfa_b200 import _lib
    from pcfa_b200.corr_block import pyramid_layout
    lib = _lib.load()
    B, C, H, W = f1.shape
    offs, hs, ws = pyramid_layout(B, H, W, levels)
    pyr = torch.full((offs[-1],), float("nan"), device="cuda")
    wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, levels)
    wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    _lib.check(lib.pcfa_corr_pyramid_forward(_lib.ptr(f1), _lib.ptr(f2), _lib.ptr(pyr), _lib.ptr(wsp), wsb,
                                             B, C, H, W, levels, impl, _lib.stream()), "pyramid fwd")
    torch.cuda.synchronize()
    return pyr, offs


@pytest.mark.parametrize("shape,levels", [((1, 64, 16, 24), 4), ((2, 128, 18, 27), 4), ((1, 256, 9, 17), 3),
                                          ((3, 64, 8, 8), 2), ((1, 256, 55, 128), 4), ((2, 256, 46, 62), 4)])
@pytest.mark.parametrize("impl", [2, 3])
def test_allpairs_tcgen05_matches_fp32_simt(shape, levels, impl):
    """bf16x3 tensor-core pyramid (impl 2: one CTA per tile, impl 3: cta_group::2 pairs) vs the exact-fp32
    SIMT pyramid (which the oracle tests pin)."""
    g = torch.Generator().manual_seed(sum(shape))
    f1 = torch.randn(shape, generator=g).cuda() * 3
    f2 = torch.randn(shape, generator=g).cuda() * 3
    tc, offs = _build_pyramid(f1, f2, levels, impl)
    simt, _ = _build_pyramid(f1, f2, levels, 1)
    assert torch.isfinite(tc).all(), "tensor-core path left pyramid cells unwritten"
    for l in range(levels):
        a, b = tc[offs[l]:offs[l + 1]], simt[offs[l]:offs[l + 1]]
        rms = float(b.pow(2).mean().sqrt())
        err = float((a - b).abs().max())
        assert err <= 1e-4 * rms + 1e-4 * float(b.abs().max()), f"level {l}: max err {err} (rms {rms})"


def test_allpairs_tcgen05_vs_oracle_small():
    from oracle import ops as O
    g = np.random.default_rng(9)
    f1 = g.standard_normal((1, 64, 12, 20)).astype(np.float32)
    f2 = g.standard_normal((1, 64, 12, 20)).astype(np.float32)
    tc, _ = _build_pyramid(cu(f1), cu(f2), 4, 2)
    assert_close(npy(tc), O.corr_pyramid_forward(f1, f2, 4), what="tcgen05 pyramid vs oracle", **TIGHT)


def _pyramid_backward(gp, f1, f2, levels, impl):
    from pcfa_b200 import _lib
    lib = _lib.load()
    B, C, H, W = f1.shape
    g1, g2 = torch.full_like(f1, float("nan")), torch.full_like(f2, float("nan"))
    wsb = lib.pcfa_corr_pyramid_workspace_bytes(B, C, H, W, levels)
    wsp = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    _lib.check(lib.pcfa_corr_pyramid_backward(_lib.ptr(gp), _lib.ptr(f1), _lib.ptr(f2), _lib.ptr(g1), _lib.ptr(g2),
                                              _lib.ptr(wsp), wsb, B, C, H, W, levels, impl, _lib.stream()), "pyramid bwd")
    torch.cuda.synchronize()
    return g1, g2


@pytest.mark.parametrize("shape,levels", [((1, 64, 16, 32), 4), ((2, 128, 16, 24), 3), ((1, 256, 12, 20), 2),
                                          ((1, 256, 55, 128), 4), ((2, 256, 46, 64), 4), ((1, 48, 8, 8), 1)])
@pytest.mark.parametrize("impl", [2, 3])
def test_allpairs_backward_tcgen05_matches_fp32_simt(shape, levels, impl):
    """TF32 tensor-core backward (K-major pass I, MN-major pass II) vs the exact-fp32 SIMT backward.
    Expected error: 2^-11 relative truncation of the gradient pyramid, random over the contraction."""
    from pcfa_b200.corr_block import pyramid_layout
    g = torch.Generator().manual_seed(sum(shape) + levels)
    B, C, H, W = shape
    f1 = torch.randn(shape, generator=g).cuda()
    f2 = torch.randn(shape, generator=g).cuda()
    offs, _, _ = pyramid_layout(B, H, W, levels)
    gp = torch.randn(offs[-1], generator=g).cuda()
    t1, t2 = _pyramid_backward(gp, f1, f2, levels, impl)      # 2: one CTA per block pair, 3: cta_group::2 pairs
    s1, s2 = _pyramid_backward(gp, f1, f2, levels, 1)
    for name, a, b in (("grad_fmap1", t1, s1), ("grad_fmap2", t2, s2)):
        assert torch.isfinite(a).all(), f"{name}: unwritten cells"
        rel = float((a - b).norm() / b.norm())
        mx = float((a - b).abs().max() / b.pow(2).mean().sqrt())
        assert rel < 1e-3 and mx < 6e-3, f"{name}: rel L2 {rel:.2e}, max/rms {mx:.2e}"


# ----------------------------------------------------------------------------------- encoder glue (f-4)
@pytest.mark.parametrize("shape", [(2, 8, 37, 53), (1, 16, 32, 64), (1, 4, 220, 512), (3, 5, 1, 7), (2, 96, 20, 24)])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("channels_last", [False, True])
def test_instance_norm_matches_torch(shape, relu, channels_last):
    """Fused instance norm (+ReLU) vs F.instance_norm (+F.relu) on the same GPU: values and input gradient.
    (models/raft/extractor.py:13-55: nn.InstanceNorm2d(affine=False), eps 1e-5, biased variance.)"""
    import torch.nn.functional as F
    from pcfa_b200.instance_norm import instance_norm
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 3 + 1.5).cuda()
    go = torch.randn(shape, generator=g).cuda()
    if channels_last:                                   # [B][H][W][C] memory; C % 4 != 0 falls back to the NCHW kernels
        x = x.contiguous(memory_format=torch.channels_last)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    ya = instance_norm(xa, eps=1e-5, relu=relu)
    if channels_last and shape[1] % 4 == 0 and shape[1] > 1:
        assert ya.is_contiguous(memory_format=torch.channels_last)
    yb = F.instance_norm(xb, eps=1e-5)
    yb = F.relu(yb) if relu else yb
    assert_close(npy(ya), npy(yb), what="instance norm", rtol=1e-5, atol_rms=1e-5)
    (ya * go).sum().backward()
    (yb * go).sum().backward()
    assert_close(npy(xa.grad), npy(xb.grad), what="instance norm grad", rtol=1e-4, atol_rms=1e-4)


@pytest.mark.parametrize("shape", [(1, 128, 55, 128), (2, 96, 7, 9), (3, 5, 3, 3)])
def test_fused_gru_elementwise_matches_torch(shape):
    """gru_gates / gru_blend vs the reference's element-wise composition (models/raft/update.py:16-31), values and
    every input gradient."""
    from pcfa_b200.gru_ops import gru_blend, gru_gates
    g = torch.Generator().manual_seed(sum(shape))
    B, C, H, W = shape
    zr0, h0, qp0 = torch.randn(B, 2 * C, H, W, generator=g).cuda(), torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
    go = torch.randn(shape, generator=g).cuda()
    outs = []
    for fused in (True, False):
        zr, h, qp = (t.clone().requires_grad_(True) for t in (zr0, h0, qp0))
        if fused:
            z, rh = gru_gates(zr, h)
            hn = gru_blend(z, qp + rh, h)                   # rh feeds q_pre like convq([r*h, x]) would
        else:
            z, r = torch.sigmoid(zr[:, :C]), torch.sigmoid(zr[:, C:])
            rh = r * h
            q = torch.tanh(qp + rh)
            hn = (1 - z) * h + z * q
        (hn * go).sum().backward()
        outs.append((hn, zr.grad, h.grad, qp.grad))
    for name, a, b in zip(("h_new", "grad zr", "grad h", "grad q_pre"), outs[0], outs[1]):
        assert_close(npy(a), npy(b), what=name, rtol=1e-5, atol_rms=1e-6)


# ----------------------------------------------------------------------------------- convex up-sampling (f-4)
def _upsample_reference(flow, mask):
    """RAFT.upsample_flow verbatim in torch ops (models/raft/raft.py:72-83)."""
    import torch.nn.functional as F
    N, _, H, W = flow.shape
    mask = mask.view(N, 1, 9, 8, 8, H, W)
    mask = torch.softmax(mask, dim=2)
    up_flow = F.unfold(8 * flow, [3, 3], padding=1)
    up_flow = up_flow.view(N, 2, 9, 1, 1, H, W)
    up_flow = torch.sum(mask * up_flow, dim=2)
    up_flow = up_flow.permute(0, 1, 4, 2, 5, 3)
    return up_flow.reshape(N, 2, 8 * H, 8 * W)


@pytest.mark.parametrize("shape,cl", [((1, 55, 128), True), ((2, 7, 9), True), ((2, 16, 20), False), ((1, 1, 1), True)])
def test_convex_upsample_matches_reference_formula(shape, cl):
    from pcfa_b200.upsample import convex_upsample
    N, H, W = shape
    g = torch.Generator().manual_seed(H * W)
    flow = (3 * torch.randn(N, 2, H, W, generator=g)).cuda().requires_grad_(True)
    raw = (4 * torch.randn(N, 576, H, W, generator=g)).cuda()
    if cl:
        raw = raw.contiguous(memory_format=torch.channels_last)
    raw.requires_grad_(True)
    up = convex_upsample(flow, raw, 0.25)
    f2, r2 = flow.detach().clone().requires_grad_(True), raw.detach().clone().contiguous().requires_grad_(True)
    ref = _upsample_reference(f2, 0.25 * r2)
    assert up.shape == ref.shape and up.is_contiguous()
    assert_close(npy(up), npy(ref), what="convex upsample fwd", **TIGHT)
    go = torch.randn(ref.shape, generator=g).cuda()
    (up * go).sum().backward()
    (ref * go).sum().backward()
    assert_close(npy(flow.grad), npy(f2.grad), what="convex upsample g flow", **TIGHT)
    assert_close(npy(raw.grad), npy(r2.grad), what="convex upsample g mask", **TIGHT)


# ----------------------------------------------------------------------------------- GMA attention softmax (f-2)
@pytest.mark.parametrize("rows,cols", [(7040, 7040), (33, 1000), (5, 16384), (2, 8)])
def test_softmax_rows_f16_matches_fp32_softmax_of_the_fp16_input(rows, cols):
    """models/gma/gma.py:73-74 under fp16 autocast: softmax(sim) is computed in fp32 from the fp16 similarity and cast
    to fp16 by the consumer; the fused kernel must return exactly that rounding (<= 1 fp16 ulp), and its backward
    attn * (g - <attn, g>) from the same fp16 tensors."""
    from pcfa_b200.attention import softmax_rows_f16
    g = torch.Generator().manual_seed(rows + cols)
    sim = (4 * torch.randn(rows, cols, generator=g)).half().cuda().requires_grad_(True)
    attn = softmax_rows_f16(sim)
    ref = torch.softmax(sim.detach().float(), dim=-1)
    assert attn.dtype == torch.float16 and attn.shape == sim.shape
    err = (attn.float() - ref).abs()
    assert float((err - 1e-3 * ref).max()) <= 1e-7, float(err.max())          # fp16 rounding of the fp32 result: 2^-11 relative
    np.testing.assert_allclose(attn.detach().float().sum(-1).cpu().numpy(), 1.0, rtol=2e-3)
    go = torch.randn(rows, cols, generator=g).half().cuda()
    (attn * go).sum().backward()
    a32, g32 = attn.detach().float(), go.float()
    want = a32 * (g32 - (a32 * g32).sum(-1, keepdim=True))
    got = sim.grad.float()
    scale = float(want.abs().max()) + 1e-30
    assert float((got - want).abs().max()) <= 2e-3 * scale


# ----------------------------------------------------------------------------------- convolution epilogue (f-4 glue)
@pytest.mark.parametrize("cl", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("cin,cout,k,stride", [(16, 32, 3, 1), (324, 256, 1, 1), (8, 2, 3, 1), (3, 64, 7, 2)])
def test_conv_act_matches_conv_bias_relu(cl, dtype, cin, cout, k, stride):
    """conv_act (bias-free cuDNN convolution + fused bias/ReLU epilogue, one autograd node) against F.relu(conv(x))."""
    from pcfa_b200.conv_ops import conv_act
    g = torch.Generator().manual_seed(cin * cout + k)
    conv = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2).cuda().to(dtype)
    for p in conv.parameters():
        p.requires_grad = False
    x = torch.randn(2, cin, 20, 24, generator=g).cuda().to(dtype)
    if cl:
        conv = conv.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    for relu, slope in ((True, 0.0), (False, 0.0), (True, 0.1)):          # ReLU, none, LeakyReLU(0.1) (PWCNet / FlowNet2)
        a = x.clone().requires_grad_(True)
        b = x.clone().requires_grad_(True)
        y = conv_act(conv, a, relu, slope=slope)
        ref = conv(b)
        ref = (torch.nn.functional.leaky_relu(ref, slope) if slope else torch.relu(ref)) if relu else ref
        tol = dict(rtol=1e-3, atol_rms=1e-3) if dtype == torch.float32 else dict(rtol=1e-2, atol_rms=1e-2)
        assert_close(npy(y), npy(ref), what="conv_act fwd", **tol)
        go = torch.randn(ref.shape, generator=g).cuda().to(dtype)
        (y * go).sum().backward()
        (ref * go).sum().backward()
        if dtype == torch.float16 and relu:
            # cuDNN's fused epilogue applies bias + ReLU to the fp32 accumulator, the stock sequence rounds the convolution to
            # fp16 first: pre-activations within one fp16 ulp of zero get the other mask, which moves single input-gradient
            # elements by one (weight x upstream gradient) term — a statistical, not an element-wise, criterion
            ga, gb = a.grad.float(), b.grad.float()
            assert float((ga - gb).norm() / gb.norm()) < 2e-2, "conv_act grad rel-L2"
            continue
        assert_close(npy(a.grad), npy(b.grad), what="conv_act grad", **tol)


@pytest.mark.parametrize("shape", [(2, 64, 55, 64), (1, 96, 23, 31), (2, 128, 7, 9)])
@pytest.mark.parametrize("relu", [True, False])
def test_instance_norm_half_input_matches_autocast_semantics(shape, relu):
    """GMA's fp16-autocast encoder: the convolution output is fp16, autocast runs F.instance_norm in fp32 on it and returns
    fp32; the gradient returns to the convolution in fp16.  The *_h kernels read the fp16 tensor directly."""
    import torch.nn.functional as F
    from pcfa_b200.instance_norm import instance_norm
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 3 + 1.5).half().cuda().contiguous(memory_format=torch.channels_last)
    go = torch.randn(shape, generator=g).cuda()
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    ya = instance_norm(xa, eps=1e-5, relu=relu)
    assert ya.dtype == torch.float32 and ya.is_contiguous(memory_format=torch.channels_last)
    yb = F.instance_norm(xb.float(), eps=1e-5)
    yb = F.relu(yb) if relu else yb
    assert_close(npy(ya), npy(yb), what="instance norm (fp16 in)", rtol=1e-5, atol_rms=1e-5)
    (ya * go).sum().backward()
    (yb * go).sum().backward()
    assert xa.grad.dtype == torch.float16
    assert_close(npy(xa.grad), npy(xb.grad), what="instance norm grad (fp16 out)", rtol=2e-3, atol_rms=2e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_gru_step_x_matches_sepconvgru_composition(dtype):
    """One autograd node per SepConvGRU step (gru_ops.gru_step_x: hoisted context share, fused element-wise kernels,
    in-kernel gradient sums, accumulated addend gradients) against the module composition of models/raft/update.py:33-60
    evaluated in fp32 — forward, grad h, grad motion and grad of the context features over TWO chained steps (so the
    accumulate-over-iterations path of _HoistSource is exercised); fp16 storage variant for GMA's autocast."""
    from pcfa_b200.gru_ops import gru_step_x, hoist_sources
    from pcfa_b200.networks.raft import SepConvGRU
    g = torch.Generator().manual_seed(5)
    B, Ch, Ci, Cm, H, W = 1, 128, 128, 128, 12, 20
    gru = SepConvGRU(hidden_dim=Ch, input_dim=Ci + Cm).cuda()
    for p in gru.parameters():
        p.requires_grad = False
    gru.to(memory_format=torch.channels_last)
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)            # noqa: E731
    h0, inp = cl(torch.tanh(torch.randn(B, Ch, H, W, generator=g))), cl(torch.relu(torch.randn(B, Ci, H, W, generator=g)))
    m1, m2 = cl(torch.randn(B, Cm, H, W, generator=g)), cl(torch.randn(B, Cm, H, W, generator=g))
    go = cl(torch.randn(B, Ch, H, W, generator=g))
    # reference: the module, fp32, twice
    ra, ri, rm1, rm2 = (t.clone().requires_grad_(True) for t in (h0, inp, m1, m2))
    ref = gru(gru(ra, torch.cat([ri, rm1], 1)), torch.cat([ri, rm2], 1))
    (ref * go).sum().backward()
    # fused
    a, i_, b1, b2 = (t.clone().to(dtype).requires_grad_(True) for t in (h0, inp, m1, m2))
    with torch.autocast("cuda", enabled=dtype == torch.float16):
        src = hoist_sources(gru.hoisted(i_))
    out = gru_step_x(gru_step_x(a, b1, src, False), b2, src, True)
    assert out.dtype == dtype
    (out.float() * go).sum().backward()
    tol = 2e-3 if dtype == torch.float32 else 3e-2
    rel = lambda x, y: float((x.detach().float() - y.detach()).norm() / y.detach().norm())      # noqa: E731
    assert rel(out, ref) < tol, rel(out, ref)
    for name, got, want in (("h", a.grad, ra.grad), ("m1", b1.grad, rm1.grad), ("m2", b2.grad, rm2.grad), ("inp", i_.grad, ri.grad)):
        assert rel(got, want) < tol, (name, rel(got, want))


# ----------------------------------------------------------------------------------- small fused glue ops (round 2)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("cl", [True, False])
def test_add_relu_matches_torch(dtype, cl):
    from pcfa_b200.conv_ops import add_relu
    g = torch.Generator().manual_seed(5)
    a = torch.randn(2, 64, 22, 32, generator=g).cuda().to(dtype)
    b = torch.randn(2, 64, 22, 32, generator=g).cuda().to(dtype)
    if cl:
        a, b = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
    a1, b1, a2, b2 = (t.clone().requires_grad_(True) for t in (a, b, a, b))
    out, ref = add_relu(a1, b1), torch.relu(a2 + b2)
    assert out.stride() == ref.stride() and torch.equal(out, ref)
    go = torch.randn(out.shape, generator=g).cuda().to(dtype)
    (out * go).sum().backward()
    (ref * go).sum().backward()
    assert torch.equal(a1.grad, a2.grad) and torch.equal(b1.grad, b2.grad)


def test_relu_mask_backward_on_channel_slices_of_a_concatenation():
    """conv_act's backward reads its gradient straight out of a wider channels-last tensor (the gradient of a cat)."""
    from pcfa_b200.conv_ops import conv_act
    g = torch.Generator().manual_seed(9)
    convs = [torch.nn.Conv2d(16, c, 3, padding=1).cuda().to(memory_format=torch.channels_last) for c in (192, 64)]
    for cv in convs:
        for p in cv.parameters():
            p.requires_grad = False
    x = torch.randn(1, 16, 20, 24, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    w = torch.randn(1, 256, 20, 24, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    (torch.cat([conv_act(cv, xa, True) for cv in convs], 1) * w).sum().backward()
    (torch.cat([torch.relu(cv(xb)) for cv in convs], 1) * w).sum().backward()
    assert_close(npy(xa.grad), npy(xb.grad), what="grad through cat slices", **TOL)


def test_flow_step_matches_torch_and_passes_the_gradient():
    from pcfa_b200.conv_ops import flow_step
    g = torch.Generator().manual_seed(2)
    c1 = torch.randn(2, 2, 11, 16, generator=g).cuda()
    c0 = torch.randn(2, 2, 11, 16, generator=g).cuda()
    d8 = torch.randn(2, 8, 11, 16, generator=g).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    new, flow = flow_step(c1, c0, d8[:, :2])
    ref = c1 + d8.detach()[:, :2]
    assert torch.equal(new, ref) and torch.equal(flow, ref - c0)
    new8, flow8 = flow_step(c1, c0, d8[:, :2], 8)                            # zero-padded flow for the tensor-core convf1
    assert torch.equal(new8, ref) and flow8.shape == (2, 8, 11, 16) and flow8.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(flow8[:, :2], ref - c0) and float(flow8[:, 2:].abs().max()) == 0.0
    assert flow.is_contiguous(memory_format=torch.channels_last) and not flow.requires_grad
    go = torch.randn(new.shape, generator=g).cuda()
    (new * go).sum().backward()
    assert torch.equal(d8.grad[:, :2], go) and float(d8.grad[:, 2:].abs().max()) == 0.0
    h8 = d8.detach().half().contiguous(memory_format=torch.channels_last)           # GMA: fp16 flow head
    newh, flowh = flow_step(c1, c0, h8[:, :2], 8)
    refh = c1 + h8[:, :2].float()
    assert torch.equal(newh, refh) and torch.equal(flowh[:, :2], refh - c0) and newh.dtype == torch.float32


def test_dense_conv_cat_matches_torch():
    """x -> cat(LeakyReLU(conv(x)), x) as one node (PWCNet's DenseNet decoder) against the torch composition, incl. the
    zero-padded input channels of the decoder's first concatenation."""
    from pcfa_b200.conv_ops import dense_conv_cat
    g = torch.Generator().manual_seed(4)
    for cin, cpad, cout in ((88, 88, 128), (81, 88, 128), (216, 216, 96)):
        conv = torch.nn.Conv2d(cin, cout, 3, padding=1).cuda().to(memory_format=torch.channels_last)
        for p in conv.parameters():
            p.requires_grad = False
        x = torch.randn(1, cpad, 12, 20, generator=g).cuda()
        x[:, cin:] = 0
        x = x.contiguous(memory_format=torch.channels_last)
        a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        out = dense_conv_cat(conv, a, 0.1)
        ref = torch.cat((torch.nn.functional.leaky_relu(conv(b[:, :cin]), 0.1), b), 1)
        assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
        assert_close(npy(out), npy(ref), what="dense_conv_cat fwd", **TOL)
        go = torch.randn(ref.shape, generator=g).cuda()
        (out * go).sum().backward()
        (ref * go).sum().backward()
        assert_close(npy(a.grad[:, :cin]), npy(b.grad[:, :cin]), what="dense_conv_cat grad", **TOL)


def test_fork_joins_dense_and_sliced_gradients():
    """conv_ops.fork: a tensor feeding a convolution and a later channels-last concatenation gets ONE vectorised gradient join."""
    from pcfa_b200.conv_ops import fork
    from pcfa_b200.gru_ops import cat_channels
    g = torch.Generator().manual_seed(8)
    conv = torch.nn.Conv2d(64, 32, 3, padding=1).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(1, 64, 12, 20, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    other = torch.randn(1, 30, 12, 20, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    grads = []
    for use_fork in (True, False):
        a = x.clone().requires_grad_(True)
        t = a * 1.0                                               # non-leaf, like an encoder feature
        t1, t2 = fork(t) if use_fork else (t, t)
        out = cat_channels([t2, conv(t1), other], True, pad_to=8)
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).cuda().contiguous(memory_format=torch.channels_last)
        (out * w).sum().backward()
        grads.append(a.grad.clone())
    assert_close(npy(grads[0]), npy(grads[1]), what="fork grad", **TIGHT)


def test_add_relu_twin_masks_the_sum_of_both_consumers_gradients():
    from pcfa_b200.conv_ops import add_relu
    g = torch.Generator().manual_seed(6)
    conv = torch.nn.Conv2d(64, 64, 3, padding=1).cuda().to(memory_format=torch.channels_last)
    a = torch.randn(2, 64, 22, 32, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    b = torch.randn(2, 64, 22, 32, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    res = []
    for twin in (True, False):
        a1, b1 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        out = add_relu(a1, b1, twin=twin)
        assert (getattr(out, "_pcfa_twin", None) is not None) == twin
        skip = out._pcfa_twin if twin else out
        z = add_relu(skip, conv(out))                              # the next block: convolution path + skip branch
        (z * z).sum().backward()
        res.append((out.detach().clone(), a1.grad.clone(), b1.grad.clone()))
    assert torch.equal(res[0][0], res[1][0])
    assert_close(npy(res[0][1]), npy(res[1][1]), what="twin grad a", **TIGHT)
    assert_close(npy(res[0][2]), npy(res[1][2]), what="twin grad b", **TIGHT)
    only = add_relu(a.clone().requires_grad_(True), b, twin=True)    # second handle unused: plain mask
    only.sum().backward()
