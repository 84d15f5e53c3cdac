"""Pins oracle/ (the CPU restatement) against vectors produced by the REAL reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle
from conftest import nmax_err

SCAN_CASES = ["n1_l70", "n4_g2_l300", "n16_g4_l600", "n2_z_l129", "n1_g4_l1000", "bf16_n2_l128", "f16_n1_l96"]


@pytest.mark.parametrize("name", SCAN_CASES)
def test_scan_forward_matches_reference(golden_scan, name):
    c = golden_scan.case(name)
    out, last = oracle.selective_scan_oracle(c["u"], c["delta"], c["A"], c["B"], c["C"], c.get("D"), c.get("z"),
                                             c.get("delta_bias"), bool(c["softplus"]), return_last_state=True)
    # fp32 sequential in the same order: agreement to a few ulp; 16-bit cases: the reference rounds `out` to the input type
    tol = {0: 2e-6, 1: 1e-3, 2: 8e-3}[int(c["dtype"])]
    assert nmax_err(out, c["out"]) < tol
    assert nmax_err(last, c["last_state"]) < 2e-6


@pytest.mark.parametrize("name", [n for n in SCAN_CASES if "_z_" not in n])
def test_scan_backward_matches_reference_autograd(golden_scan, name):
    c = golden_scan.case(name)
    r = oracle.selective_scan_oracle_f64(c["u"], c["delta"], c["A"], c["B"], c["C"], c.get("D"), c.get("delta_bias"),
                                         bool(c["softplus"]), dout=c["dout"])
    lo = int(c["dtype"]) != 0   # reference grads of 16-bit leaves are rounded to 16 bit
    tol = 1e-2 if lo else 2e-5
    assert nmax_err(r["out"], c["out"]) < (1e-2 if lo else 1e-5)
    for k in ("du", "ddelta", "dB", "dC"):
        assert nmax_err(r[k], c[k]) < tol, k
    assert nmax_err(r["dA"], c["dA"]) < 2e-5
    if "dD" in c:
        assert nmax_err(r["dD"], c["dD"]) < 2e-5
    if "ddelta_bias" in c:
        assert nmax_err(r["ddelta_bias"], c["ddelta_bias"]) < 2e-5


def test_scan_product_api_case(golden_scan):
    c = golden_scan.case("csms6s")   # csms6s.selective_scan_fn(backend="torch"), fp32 "oflex" output
    out = oracle.selective_scan_oracle(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], None, c["delta_bias"], True)
    assert nmax_err(out, c["out"]) < 2e-6


def _layouts():
    for scans in (0, 1, 2):
        for icf in (1, 0):
            for ocf in (1, 0):
                for obo in (0, 1):
                    yield scans, icf, ocf, obo


@pytest.mark.parametrize("scans,icf,ocf,obo", list(_layouts()))
def test_cross_scan_merge_match_reference(golden_csm, scans, icf, ocf, obo):
    tag = f"s{scans}_i{icf}_o{ocf}_b{obo}"
    x = golden_csm["x4"] if obo else golden_csm["x"]
    H, W = x.shape[-2:]
    if f"scan/{tag}" in golden_csm:
        src = x
        if not icf:
            src = np.transpose(x, (0, 3, 4, 1, 2)) if obo else np.transpose(x, (0, 2, 3, 1))
        y = oracle.cross_scan_oracle(src, bool(icf), bool(ocf), bool(obo), scans)
        ref = golden_csm[f"scan/{tag}"]
        if obo and scans == 2 and not icf:
            pytest.skip("reference torch path indexes the H axis instead of K (csm_triton.py:118-123): not a cross-scan")
        if obo and scans == 1 and icf and not ocf:
            pytest.skip("reference torch path permutes the mis-shaped (B,4,C*H,W) view: result is not a cross-scan")
        if obo and scans == 1 and icf and ocf:
            # reference quirk: cross_scan1b1_fwd uses x.flatten(2, 3) for scans=1 (csm_triton.py:103), i.e. the
            # right data in a (B,4,C*H,W) shape; the Triton path returns (B,4,C,L)
            ref = ref.reshape(y.shape)
        np.testing.assert_array_equal(y, ref)   # pure data movement: bit-exact
    if f"merge/{tag}" in golden_csm:
        ys = golden_csm[f"merge_in/{tag}"]
        ys = ys.reshape(ys.shape[0], 4, -1, H * W) if ocf else ys.reshape(ys.shape[0], H * W, 4, -1)
        m = oracle.cross_merge_oracle(ys, H, W, bool(icf), bool(ocf), bool(obo), scans)
        ref = golden_csm[f"merge/{tag}"]
        assert m.shape == ref.shape
        np.testing.assert_allclose(m, ref, rtol=0, atol=1e-6)


def test_cross_scan_backward_is_merge(golden_csm):
    gy = golden_csm["scan_bwd/gy"]
    H, W = golden_csm["x"].shape[-2:]
    gx = oracle.cross_merge_oracle(gy, H, W)
    np.testing.assert_allclose(gx.reshape(golden_csm["scan_bwd/gx"].shape), golden_csm["scan_bwd/gx"], atol=1e-6)


@pytest.mark.parametrize("tag,kw", [
    ("conv3g", dict(stride=1, padding=1, groups=2)), ("dw3", dict(padding=1, groups=5)),
    ("conv_s2", dict(stride=2, padding=1)), ("pw1", dict()),
])
def test_bayes_conv_matches_reference(golden_bayes, tag, kw):
    c = golden_bayes.case(tag)
    out = oracle.bayes_conv2d_oracle(c["x"], c["mu_weight"], c["rho_weight"], c["eps_weight"], c.get("mu_bias"),
                                     c.get("rho_bias"), c.get("eps_bias"), **kw)
    assert nmax_err(out, c["out"]) < 2e-6
    if "out_det" in c:
        det = oracle.bayes_conv2d_oracle(c["x"], c["mu_weight"], None, None, c.get("mu_bias"), deterministic=True, **kw)
        assert nmax_err(det, c["out_det"]) < 2e-6


def test_bayes_linear_layers_match_reference(golden_bayes):
    c = golden_bayes.case("lin2d")
    out = oracle.bayes_linear2d_oracle(c["x"], c["mu_weight"], c["rho_weight"], c["eps_weight"])
    assert nmax_err(out, c["out"]) < 2e-6
    c = golden_bayes.case("lin")
    out = oracle.bayes_linear_oracle(c["x"], c["mu_weight"], c["rho_weight"], c["eps_weight"], c["mu_bias"],
                                     c["rho_bias"], c["eps_bias"])
    assert nmax_err(out, c["out"]) < 2e-6


def test_bayes_prior_ema_and_kl_match_reference(golden_bayes):
    c = golden_bayes.case("train")
    pm, pr = c["prior_mu0"], None
    rho0 = golden_bayes["rho_init"].reshape(-1)[0]
    pr = np.full_like(pm, rho0)
    pmb, prb = np.zeros(3, np.float32), np.full(3, rho0, np.float32)
    for it in range(3):
        pm, pr, ps = oracle.prior_ema_oracle(pm, pr, c[f"mu_w{it}"], c[f"rho_w{it}"], 0.998, it)
        pmb, prb, psb = oracle.prior_ema_oracle(pmb, prb, c[f"mu_b{it}"], c[f"rho_b{it}"], 0.998, it)
        np.testing.assert_allclose(pm, c[f"prior_mu_w{it}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(pr, c[f"prior_rho_w{it}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ps, c[f"prior_sigma_w{it}"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(pmb, c[f"prior_mu_b{it}"], rtol=1e-6, atol=1e-7)
        kl = oracle.kl_div_oracle(c[f"mu_w{it}"], oracle.softplus_rho(c[f"rho_w{it}"]), pm, ps) + \
            oracle.kl_div_oracle(c[f"mu_b{it}"], oracle.softplus_rho(c[f"rho_b{it}"]), pmb, psb)
        assert abs(kl - float(c[f"kl{it}"])) < 1e-5 * max(1.0, abs(float(c[f"kl{it}"])))
        assert int(c[f"step{it}"]) == it + 1
    assert abs(float(rho0) - np.log(np.expm1(0.05) + 1e-20)) < 1e-6   # conv.py:74


def test_select_matches_python_list_index(golden_select):
    for name in golden_select.cases():
        c = golden_select.case(name)
        assert oracle.select_best_oracle(c["scores"]) == int(c["argmax"]), name
        assert oracle.select_best_oracle(c["scores"], take_min=True) == int(c["argmin"]), name


def _sd(golden, prefix):
    p = prefix + "/"
    return {k[len(p):]: golden[k] for k in golden.z.files if k.startswith(p)}


def test_network_oracle_matches_reference_model(golden_models):
    """oracle/network.py (the CPU baseline of bench.py) against the REAL reference Network: plain weights, Bayesian
    weights in deterministic mode, and the stochastic forward replaying the reference's eps"""
    from oracle import network as onet
    x = golden_models["net/x"]
    y = onet.network_forward(_sd(golden_models, "net/sd_plain"), x)
    assert nmax_err(y.numpy(), golden_models["net/out_det_plain"]) < 1e-5
    sd = _sd(golden_models, "net/sd_bnn")
    y = onet.network_forward(sd, x, deterministic=True)
    assert nmax_err(y.numpy(), golden_models["net/out_det_bnn"]) < 1e-5
    eps = _sd(golden_models, "net/eps")
    y = onet.network_forward(sd, x, eps=eps)
    assert nmax_err(y.numpy(), golden_models["net/out_mc"]) < 1e-5


@pytest.mark.parametrize("C,hidden,H,W,th,tw,pairs", [(8, 12, 11, 37, 4, 16, 5), (4, 8, 3, 5, 4, 128, 16), (6, 16, 9, 33, 2, 8, 16),
                                                      (5, 7, 1, 9, 4, 4, 3)])
def test_tiled_gdmlp_schedule_equals_the_plain_one(C, hidden, H, W, th, tw, pairs):
    """oracle/gdmlp_tiled.py: the fused single-pass schedule planned for x + gdMlp(norm2(x)) (spatial tiles with halo, gate-pair
    groups, K-split project_out) is the same function as the reference's op sequence (vmamba.py:128-133, 1331-1333), which in turn
    is checked against torch's own ops here — ragged tiles, tiles wider than the image, a one-row image"""
    import torch
    import torch.nn.functional as F
    from oracle import gdmlp_tiled as G
    rng = np.random.default_rng(C * 100 + H)
    x = rng.standard_normal((C, H, W))
    gamma, beta = rng.standard_normal(C), rng.standard_normal(C)
    w1, b1 = rng.standard_normal((2 * hidden, C)), rng.standard_normal(2 * hidden)
    wd, bd = rng.standard_normal((2 * hidden, 3, 3)), rng.standard_normal(2 * hidden)
    w2, b2 = rng.standard_normal((C, hidden)), rng.standard_normal(C)
    plain = G.gdmlp_plain(x, gamma, beta, 1e-5, w1, b1, wd, bd, w2, b2)
    t = lambda a: torch.from_numpy(np.asarray(a))
    xn = F.layer_norm(t(x).permute(1, 2, 0), (C,), t(gamma), t(beta), 1e-5).permute(2, 0, 1)[None]
    y = F.conv2d(xn, t(w1)[:, :, None, None], t(b1))
    x1, x2 = F.conv2d(y, t(wd)[:, None], t(bd), padding=1, groups=2 * hidden).chunk(2, dim=1)
    ref = t(x)[None] + F.conv2d(F.gelu(x1) * x2, t(w2)[:, :, None, None], t(b2))
    assert nmax_err(plain, ref[0].numpy()) < 1e-12
    tiled, stats = G.gdmlp_tiled(x, gamma, beta, 1e-5, w1, b1, wd, bd, w2, b2, tile_h=th, tile_w=tw, pairs=pairs)
    assert nmax_err(tiled, plain) < 1e-12
    assert 1.0 <= stats["fc1_recompute"] <= (th + 2) * (tw + 2) / float(th * tw) + 1e-9


@pytest.mark.parametrize("D,H,W,N,R", [(3, 5, 7, 1, 2), (2, 4, 4, 3, 1), (4, 1, 9, 1, 3), (2, 6, 2, 2, 2)])
def test_traversal_aware_ss2d_equals_the_cross_scan_form(D, H, W, N, R):
    """oracle/ss2d_traversal.py: the SS2D core computed from the image and its transpose only (directions 2 / 3 as backward scans,
    nothing flipped, merge = (y0 + y2) + T(y1 + y3)) is the reference's cross_scan -> scan -> cross_merge; that form in turn is
    held against the pinned pieces (cross_scan_oracle, selective_scan_oracle, cross_merge_oracle)"""
    from oracle import ss2d_traversal as T
    rng = np.random.default_rng(D * 1000 + H * 10 + W)
    K = 4
    x = rng.standard_normal((D, H, W))
    xw = rng.standard_normal((K, R + 2 * N, D)) * 0.5
    dtw = rng.standard_normal((K, D, R)) * 0.5
    dtb = rng.standard_normal((K, D)) * 0.5
    A = -np.exp(rng.standard_normal((K, D, N)) * 0.3)
    Dv = rng.standard_normal((K, D))
    a = T.ss2d_cross_scan_form(x, xw, dtw, dtb, A, Dv)
    b = T.ss2d_traversal_aware(x, xw, dtw, dtb, A, Dv)
    assert nmax_err(b, a) < 1e-12
    # the cross-scan form against the pinned oracle pieces (fp32 scan)
    L = H * W
    xs = oracle.cross_scan_oracle(x[None].astype(np.float32))                                  # (1, 4, D, L)
    x_dbl = np.einsum("kcd,kdl->kcl", xw.astype(np.float32), xs[0])
    dts = np.einsum("kdr,krl->kdl", dtw.astype(np.float32), x_dbl[:, :R])
    ys = oracle.selective_scan_oracle(xs.reshape(1, K * D, L), dts.reshape(1, K * D, L), A.reshape(K * D, N).astype(np.float32),
                                      np.ascontiguousarray(x_dbl[None, :, R:R + N]), np.ascontiguousarray(x_dbl[None, :, R + N:]),
                                      Dv.reshape(-1).astype(np.float32), None, dtb.reshape(-1).astype(np.float32), True)
    y = oracle.cross_merge_oracle(np.asarray(ys).reshape(1, K, D, L), H, W)
    assert nmax_err(np.asarray(y).reshape(D, H, W), a) < 2e-5


def test_layernorm2d_oracle_is_the_reference_op_and_its_autograd():
    """oracle.layernorm2d_oracle against the op LayerNorm2d wraps (vmamba.py:58-63), forward and autograd, on the CPU in fp64"""
    import torch
    g = torch.Generator().manual_seed(5)
    for shape, affine in (((2, 7, 5, 3), True), ((1, 40, 4, 6), True), ((3, 5, 2, 2), False)):
        x = torch.randn(*shape, generator=g, dtype=torch.float64, requires_grad=True)
        C = shape[1]
        w = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).requires_grad_() if affine else None
        b = torch.randn(C, generator=g, dtype=torch.float64).requires_grad_() if affine else None
        dy = torch.randn(*shape, generator=g, dtype=torch.float64)
        y = torch.nn.functional.layer_norm(x.permute(0, 2, 3, 1), (C,), w, b, 1e-5).permute(0, 3, 1, 2)
        y.backward(dy)
        o = oracle.layernorm2d_oracle(x.detach().numpy(), None if w is None else w.detach().numpy(), None if b is None else b.detach().numpy(),
                                      1e-5, dy.numpy())
        assert np.allclose(o["y"], y.detach().numpy(), rtol=0, atol=1e-12)
        assert np.allclose(o["dx"], x.grad.numpy(), rtol=0, atol=1e-11)
        if affine:
            assert np.allclose(o["dweight"], w.grad.numpy(), rtol=0, atol=1e-11)
            assert np.allclose(o["dbias"], b.grad.numpy(), rtol=0, atol=1e-11)
