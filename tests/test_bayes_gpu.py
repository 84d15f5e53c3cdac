"""GPU parity of the Bayesian layers: same eps -> same output as the reference layer (golden vectors recorded from
basicsr/bayesian, replayed by injecting the eps the reference left in its eps_* buffers), fp32 tolerance 1e-5."""
import numpy as np
import pytest
import torch

import oracle
from oracle import philox
from conftest import nmax_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _load(layer, c):
    sd = {k: torch.tensor(c[k]) for k in ("mu_weight", "rho_weight", "mu_bias", "rho_bias") if k in c}
    layer.load_state_dict(sd, strict=True)
    return layer.cuda().eval()


def _eps(c):
    e = {"eps_weight": torch.tensor(c["eps_weight"], device="cuda")}
    if "eps_bias" in c:
        e["eps_bias"] = torch.tensor(c["eps_bias"], device="cuda")
    return e


@pytest.mark.parametrize("tag,kw", [
    ("conv3g", dict(in_channels=4, out_channels=6, kernel_size=3, stride=1, padding=1, groups=2, bias=True)),
    ("dw3", dict(in_channels=5, out_channels=5, kernel_size=3, padding=1, groups=5, bias=False)),
    ("conv_s2", dict(in_channels=3, out_channels=4, kernel_size=(3, 2), stride=2, padding=1, bias=True)),
    ("pw1", dict(in_channels=5, out_channels=7, kernel_size=1, bias=True)),
])
def test_conv_golden(golden_bayes, tag, kw):
    from bem_b200 import bayesian
    c = golden_bayes.case(tag)
    layer = _load(bayesian.Conv2dReparameterization(**kw), c)
    x = torch.tensor(c["x"], device="cuda")
    with torch.no_grad():
        out = layer(x, **_eps(c))
    assert nmax_err(out.cpu().numpy(), c["out"]) < TOL
    if "out_det" in c:
        layer.deterministic = True
        with torch.no_grad():
            assert nmax_err(layer(x).cpu().numpy(), c["out_det"]) < TOL


def test_linear_layers_golden(golden_bayes):
    from bem_b200 import bayesian
    c = golden_bayes.case("lin2d")
    layer = _load(bayesian.Linear2dReparameterization(6, 6, bias=False), c)
    with torch.no_grad():
        out = layer(torch.tensor(c["x"], device="cuda"), **_eps(c))
    assert nmax_err(out.cpu().numpy(), c["out"]) < TOL
    c = golden_bayes.case("lin")
    layer = _load(bayesian.LinearReparameterization(5, 3, bias=True), c)
    with torch.no_grad():
        out = layer(torch.tensor(c["x"], device="cuda"), **_eps(c))
    assert nmax_err(out.cpu().numpy(), c["out"]) < TOL


def test_state_dict_contract(golden_bayes):
    """checkpoint keys are exactly mu_* / rho_* (conv.py:55-69: eps / prior buffers are non-persistent)"""
    from bem_b200 import bayesian
    layer = bayesian.Conv2dReparameterization(4, 6, 3, padding=1, groups=2, bias=True)
    assert sorted(layer.state_dict().keys()) == list(golden_bayes["conv3g/state_keys"])
    assert abs(float(layer.rho_weight.flatten()[0]) - float(golden_bayes["rho_init"].flatten()[0])) < 1e-6
    assert {n for n, _ in layer.named_buffers()} == {"eps_weight", "prior_mu_weight", "prior_rho_weight", "eps_bias",
                                                      "prior_mu_bias", "prior_rho_bias"}


def test_training_prior_ema_and_kl(golden_bayes):
    """prior EMA, step counter and kl_loss of the training-mode forward (conv.py:85-104) replayed step by step"""
    from bem_b200 import bayesian
    c = golden_bayes.case("train")
    layer = bayesian.Conv2dReparameterization(3, 3, 1, bias=True, sigma_init=0.05, decay=0.998)
    with torch.no_grad():
        layer.prior_mu_weight.copy_(torch.tensor(c["prior_mu0"]))
    layer = layer.cuda().train()
    x = torch.randn(1, 3, 4, 4, device="cuda")
    for it in range(3):
        with torch.no_grad():
            layer.mu_weight.copy_(torch.tensor(c[f"mu_w{it}"]))
            layer.rho_weight.copy_(torch.tensor(c[f"rho_w{it}"]))
            layer.mu_bias.copy_(torch.tensor(c[f"mu_b{it}"]))
            layer.rho_bias.copy_(torch.tensor(c[f"rho_b{it}"]))
        layer(x)
        np.testing.assert_allclose(layer.prior_mu_weight.cpu().numpy(), c[f"prior_mu_w{it}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(layer.prior_rho_weight.cpu().numpy(), c[f"prior_rho_w{it}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(layer.prior_sigma_weight.cpu().numpy(), c[f"prior_sigma_w{it}"], rtol=1e-5, atol=1e-8)
        assert abs(float(layer.kl_loss()) - float(c[f"kl{it}"])) < 1e-5 * max(1.0, abs(float(c[f"kl{it}"])))
        assert layer.step == int(c[f"step{it}"])


def test_torch_rng_stream_matches_reference_order():
    """default eps source: eps_weight.normal_() then eps_bias.normal_() from torch's global generator (conv.py:107-110),
    so a reference model and this package draw identical noise under the same seed"""
    from bem_b200 import bayesian
    layer = bayesian.Conv2dReparameterization(8, 8, 1, bias=True).cuda().eval()
    x = torch.randn(1, 8, 5, 5, device="cuda")
    torch.manual_seed(123)
    with torch.no_grad():
        out = layer(x)
    torch.manual_seed(123)
    ew = torch.empty_like(layer.eps_weight).normal_()
    eb = torch.empty_like(layer.eps_bias).normal_()
    assert torch.equal(layer.eps_weight, ew) and torch.equal(layer.eps_bias, eb)
    ref = oracle.bayes_conv2d_oracle(x.cpu().numpy(), *(t.detach().cpu().numpy() for t in (
        layer.mu_weight, layer.rho_weight, ew, layer.mu_bias, layer.rho_bias, eb)))
    assert nmax_err(out.cpu().numpy(), ref) < TOL


def test_philox_eps_matches_oracle_and_is_shard_invariant():
    from bem_b200.bayesian import functional as BF
    mu = torch.randn(7, 5, 3, 3, device="cuda")
    rho = torch.full_like(mu, -3.0)
    w, eps = BF.sample_weights(mu, rho, None, n_samples=4, seed=0x1234567890ABCDEF, stream_id=9, sample0=10)
    for s in range(4):
        ref = philox.normal(mu.numel(), 0x1234567890ABCDEF, 9, 10 + s).reshape(mu.shape)
        np.testing.assert_allclose(eps[s].cpu().numpy(), ref, rtol=0, atol=2e-6)
        np.testing.assert_allclose(w[s].cpu().numpy(), oracle.bayes_sample(mu.cpu().numpy(), rho.cpu().numpy(), ref), rtol=0, atol=2e-6)
    # sample 12 drawn alone == third sample of the batch above: world-size / batching invariance
    w1, e1 = BF.sample_weights(mu, rho, None, n_samples=1, seed=0x1234567890ABCDEF, stream_id=9, sample0=12)
    assert torch.equal(e1[0], eps[2]) and torch.equal(w1[0], w[2])
    e = eps.flatten()
    assert abs(float(e.mean())) < 0.08 and abs(float(e.std()) - 1.0) < 0.08


@pytest.mark.parametrize("geom", ["pointwise", "depthwise", "linear2d"])
def test_mc_batched_equals_loop(geom):
    """S samples batched through the grouped kernels == S separate forwards with the same eps"""
    from bem_b200 import bayesian
    S, Bx = 3, 2
    if geom == "pointwise":
        layer = bayesian.Conv2dReparameterization(10, 24, 1, bias=True)
    elif geom == "depthwise":
        layer = bayesian.Conv2dReparameterization(12, 12, 3, padding=1, groups=12, bias=True)
    else:
        layer = bayesian.Linear2dReparameterization(10, 10, bias=False)
    layer = layer.cuda().eval()
    cin = 12 if geom == "depthwise" else 10
    x = torch.randn(S * Bx, cin, 9, 11, device="cuda")
    eps_w = torch.randn((S,) + tuple(layer.eps_weight.shape), device="cuda")
    eps_b = torch.randn((S,) + tuple(layer.eps_bias.shape), device="cuda") if layer.bias else None
    layer.mc_samples = S
    with torch.no_grad():
        batched = layer(x, eps_weight=eps_w, eps_bias=eps_b)
    layer.mc_samples = 1
    for s in range(S):
        with torch.no_grad():
            one = layer(x[s * Bx:(s + 1) * Bx], eps_weight=eps_w[s], eps_bias=None if eps_b is None else eps_b[s])
        assert torch.equal(one, batched[s * Bx:(s + 1) * Bx])


@pytest.mark.parametrize("geom", ["pointwise", "depthwise"])
def test_training_gradients_match_autograd_of_reference_formula(geom):
    """d/d(mu, rho, x) through the kernels == autograd of the reference's eager formula (conv.py:106-114)"""
    import torch.nn.functional as F
    from bem_b200 import bayesian
    if geom == "pointwise":
        layer = bayesian.Conv2dReparameterization(6, 9, 1, bias=True).cuda().train()
        kw = {}
    else:
        layer = bayesian.Conv2dReparameterization(6, 6, 3, padding=1, groups=6, bias=True).cuda().train()
        kw = dict(padding=1, groups=6)
    x = torch.randn(2, 6, 7, 8, device="cuda", requires_grad=True)
    eps_w = torch.randn_like(layer.eps_weight)
    eps_b = torch.randn_like(layer.eps_bias)
    out = layer(x, eps_weight=eps_w, eps_bias=eps_b)
    g = torch.randn_like(out)
    out.backward(g)
    got = [x.grad.clone(), layer.mu_weight.grad.clone(), layer.rho_weight.grad.clone(), layer.mu_bias.grad.clone(),
           layer.rho_bias.grad.clone()]
    x2 = x.detach().clone().requires_grad_()
    mw, rw, mb, rb = (p.detach().clone().requires_grad_() for p in (layer.mu_weight, layer.rho_weight, layer.mu_bias, layer.rho_bias))
    w = mw + torch.log1p(torch.exp(rw)) * eps_w
    b = mb + torch.log1p(torch.exp(rb)) * eps_b
    with torch.backends.cudnn.flags(allow_tf32=False):
        ref = F.conv2d(x2, w, b, **kw)
        ref.backward(g)
    assert nmax_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < TOL
    for a, r in zip(got, (x2.grad, mw.grad, rw.grad, mb.grad, rb.grad)):
        assert nmax_err(a.cpu().numpy(), r.cpu().numpy()) < 2e-5


def test_convert2bnn_selective_and_tools():
    """tools.py:48-84 on the mirror network: only BasicBlock regions are converted; flags and KL helpers work"""
    from bem_b200 import bayesian, network
    net = network.Network(stage=1, n_feat=8, num_blocks=[1, 1, 1], d_state=[1, 1, 1], use_pixelshuffle=True)
    bayesian.convert2bnn_selective(net, {"sigma_init": 0.05, "decay": 0.998, "pretrain": True})
    layers = bayesian.bayesian_layers(net)
    assert len(layers) == 5 * 6                       # 5 VSSBlocks x (in_proj, conv2d, out_proj, project_in, dwconv, project_out)
    assert isinstance(net.first_conv, torch.nn.Conv2d) and isinstance(net.subnets[0].encoder_layers[0][1].reduction, torch.nn.Conv2d)
    bayesian.set_prediction_type(net, deterministic=True)
    assert all(l.deterministic for l in layers)
    bayesian.set_prediction_type(net, deterministic=False)
    assert not any(l.deterministic for l in layers)
    net = net.cuda().train()
    net(torch.rand(1, 3, 16, 16, device="cuda"))
    kl = bayesian.get_kl_loss(net)
    assert torch.isfinite(kl) and float(kl) >= 0


def test_full_size_pointwise_and_depthwise_linearity():
    """600x400 level-0 layer shapes (C=40 -> 8C=320): oracle parity on a pixel subset + linearity in x"""
    from bem_b200.bayesian import functional as BF
    x = torch.randn(1, 40, 400, 600, device="cuda")
    w = torch.randn(1, 320, 40, device="cuda") * 0.1
    b = torch.randn(1, 320, device="cuda")
    out = BF.pointwise_conv(x, w, b, 1)
    ref = torch.einsum("oc,chw->ohw", w[0].double(), x[0, :, :8, :].double()) + b[0].double()[:, None, None]
    assert nmax_err(out[0, :, :8, :].cpu().numpy(), ref.cpu().numpy()) < TOL
    x2 = torch.randn_like(x)
    lin = BF.pointwise_conv(x + x2, w, None, 1) - BF.pointwise_conv(x, w, None, 1) - BF.pointwise_conv(x2, w, None, 1)
    assert float(lin.abs().max()) < 1e-4
    wd = torch.randn(1, 320, 3, 3, device="cuda")
    y = BF.depthwise_conv3x3(out, wd, b, 1)
    refd = torch.nn.functional.conv2d(out[:, :, :16].double(), wd[0].unsqueeze(1).double(), b[0].double(), padding=1, groups=320)
    assert nmax_err(y[0, :, :15].cpu().numpy(), refd[0, :, :15].cpu().numpy()) < TOL


@pytest.mark.parametrize("cin,cout,P,S,Bx", [(40, 40, 1000, 1, 1), (40, 320, 777, 1, 2), (160, 40, 513, 3, 1), (80, 640, 300, 2, 1),
                                              (8, 16, 130, 1, 1), (24, 7, 64, 1, 1), (640, 160, 200, 1, 1), (5, 300, 129, 1, 1)])
def test_tcgen05_pointwise_matches_fp64_and_simt(cin, cout, P, S, Bx):
    """3xTF32 tensor-core contraction: fp32-tier accuracy (1e-5) against an fp64 einsum, for ragged channel counts,
    pixel tails and per-sample weights; and agreement with the fp32 CUDA-core kernel"""
    from bem_b200.bayesian import functional as BF
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout)
    x = torch.randn(S * Bx, cin, P, generator=g).cuda()
    w = (torch.randn(S, cout, cin, generator=g) / cin ** 0.5).cuda()
    b = torch.randn(S, cout, generator=g).cuda()
    out = BF.pointwise_conv(x, w, b, S)
    ref = torch.einsum("soc,sbcp->sbop", w.double(), x.double().view(S, Bx, cin, P)) + b.double()[:, None, :, None]
    assert nmax_err(out.cpu().numpy(), ref.reshape(S * Bx, cout, P).cpu().numpy()) < TOL
    simt = BF.pointwise_conv(x, w, b, S, force_simt=True)
    assert nmax_err(out.cpu().numpy(), simt.cpu().numpy()) < TOL


@pytest.mark.parametrize("cin,cout,H,W", [(40, 40, 20, 30), (160, 40, 9, 13), (80, 640, 8, 8)])
def test_fused_layernorm_pointwise(cin, cout, H, W):
    """LayerNorm2d (vmamba.py:59-64) fused into the staging pass == F.layer_norm followed by the layer"""
    import torch.nn.functional as F
    from bem_b200 import bayesian
    from bem_b200.ss2d import LayerNorm2d
    torch.manual_seed(cin + cout)
    norm = LayerNorm2d(cin).cuda()
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.3)
        norm.bias.normal_(0.0, 0.3)
    layer = bayesian.Conv2dReparameterization(cin, cout, 1, bias=True).cuda().eval()
    x = (torch.randn(2, cin, H, W, device="cuda") * 3 + 1.5)
    eps_w = torch.randn_like(layer.eps_weight)
    eps_b = torch.randn_like(layer.eps_bias)
    with torch.no_grad():
        fused = layer(x, eps_weight=eps_w, eps_bias=eps_b, pre_norm=norm)
        xn = F.layer_norm(x.double().permute(0, 2, 3, 1), (cin,), norm.weight.double(), norm.bias.double(), norm.eps).permute(0, 3, 1, 2)
        wv = (layer.mu_weight.double() + torch.log1p(torch.exp(layer.rho_weight.double())) * eps_w.double()).view(cout, cin)
        bv = layer.mu_bias.double() + torch.log1p(torch.exp(layer.rho_bias.double())) * eps_b.double()
        ref = torch.einsum("oc,bchw->bohw", wv, xn) + bv[None, :, None, None]
        two_step = layer(norm(x), eps_weight=eps_w, eps_bias=eps_b)
    assert nmax_err(fused.cpu().numpy(), ref.cpu().numpy()) < TOL
    assert nmax_err(two_step.cpu().numpy(), ref.cpu().numpy()) < TOL
    lin = bayesian.Linear2dReparameterization(cin, cin, bias=False).cuda().eval()
    lin.deterministic = True
    with torch.no_grad():
        a = lin(x, pre_norm=norm)
        bref = lin(norm(x))
    assert nmax_err(a.cpu().numpy(), bref.cpu().numpy()) < TOL


@pytest.mark.parametrize("C,H,W,S,Bx", [(8, 17, 23, 1, 1), (40, 33, 48, 1, 2), (6, 9, 600, 2, 1), (320, 20, 28, 1, 1)])
@pytest.mark.parametrize("act", [None, "silu", "gelu_gate"])
def test_depthwise_fused_activation(C, H, W, S, Bx, act):
    """depthwise 3x3 + the activation that follows it in the reference (SiLU of SS2D, vmamba.py:708-710; gated GELU of
    gdMlp, vmamba.py:129-131) in one kernel == conv2d then the torch activation, incl. ragged widths and row tails"""
    import torch.nn.functional as F
    from bem_b200.bayesian import functional as BF
    g = torch.Generator(device="cpu").manual_seed(C * 100 + W)
    x = torch.randn(S * Bx, C, H, W, generator=g).cuda()
    w = torch.randn(S, C, 3, 3, generator=g).cuda() / 3
    b = torch.randn(S, C, generator=g).cuda()
    out = BF.depthwise_conv3x3(x, w, b, S, act=act)
    refs = []
    for s in range(S):
        y = F.conv2d(x[s * Bx:(s + 1) * Bx].double(), w[s].unsqueeze(1).double(), b[s].double(), padding=1, groups=C)
        if act == "silu":
            y = F.silu(y)
        elif act == "gelu_gate":
            y1, y2 = y.chunk(2, dim=1)
            y = F.gelu(y1) * y2
        refs.append(y)
    ref = torch.cat(refs, 0)
    assert out.shape == ref.shape
    assert nmax_err(out.cpu().numpy(), ref.cpu().numpy()) < TOL


@pytest.mark.parametrize("cin,cout,H,W,ln", [(40, 40, 20, 30, True), (160, 40, 9, 13, False), (40, 200, 8, 9, True), (24, 7, 5, 4, False)])
def test_pointwise_fused_residual(cin, cout, H, W, ln):
    """skip connection folded into the 1x1 layer's epilogue (vmamba.py:1331-1333): layer(x, residual=r) == r + layer(x),
    on the tensor-core kernel (aligned and unaligned pixel counts) and on the CUDA-core kernel"""
    from bem_b200 import bayesian
    from bem_b200.bayesian import functional as BF
    from bem_b200.ss2d import LayerNorm2d
    torch.manual_seed(cin * 7 + cout)
    layer = bayesian.Conv2dReparameterization(cin, cout, 1, bias=True).cuda().eval()
    norm = LayerNorm2d(cin).cuda() if ln else None
    x = torch.randn(2, cin, H, W, device="cuda")
    r = torch.randn(2, cout, H, W, device="cuda")
    eps_w, eps_b = torch.randn_like(layer.eps_weight), torch.randn_like(layer.eps_bias)
    with torch.no_grad():
        fused = layer(x, eps_weight=eps_w, eps_bias=eps_b, pre_norm=norm, residual=r)
        plain = layer(x, eps_weight=eps_w, eps_bias=eps_b, pre_norm=norm)
    assert nmax_err(fused.cpu().numpy(), (r + plain).cpu().numpy()) < 1e-6
    w = torch.randn(1, cout, cin, device="cuda")
    simt = BF.pointwise_conv(x, w, None, 1, force_simt=True, residual=r)
    assert nmax_err(simt.cpu().numpy(), (r + BF.pointwise_conv(x, w, None, 1, force_simt=True)).cpu().numpy()) < 1e-6


def test_batched_sampling_matches_per_layer_sampling():
    """MCArena (one launch for the whole network) gives every layer bit-identical weights to its own Philox draw"""
    from bem_b200 import bayesian, mc
    from bem_b200.bayesian import functional as BF
    torch.manual_seed(3)
    net = torch.nn.Sequential(bayesian.Conv2dReparameterization(5, 7, 3, padding=1), bayesian.Linear2dReparameterization(7, 300),
                              bayesian.Conv2dReparameterization(300, 300, 3, padding=1, groups=300, bias=False)).cuda().eval()
    bayesian.set_mc_config(net, mc_samples=1, eps_source="philox", seed=99, sample0=0)
    arena = mc.MCArena(net, seed=99)
    for sample in (0, 5, 2 ** 33 + 1):
        arena.draw(sample)
        for L in arena.layers:
            for which in ("weight", "bias"):
                if which == "bias" and not L.bias:
                    continue
                sid = 2 * int(L.layer_id) + (which == "bias")
                w, _ = BF.sample_weights(getattr(L, "mu_" + which), getattr(L, "rho_" + which), None, 1, 99, sid, sample)
                assert torch.equal(w, L._arena_views[which]), (L.layer_id, which, sample)
    arena.sample0.fill_(5)
    arena.draw(None)                                   # sample index from the device word (CUDA-graph replay path)
    ref, _ = BF.sample_weights(arena.layers[1].mu_weight, arena.layers[1].rho_weight, None, 1, 99, 2 * int(arena.layers[1].layer_id), 5)
    assert torch.equal(ref, arena.layers[1]._arena_views["weight"])
    # S weight sets per draw (the S-batched Monte-Carlo forward): set s = the per-layer draw of its own global sample index
    arena4 = mc.MCArena(net, seed=99, n_sets=4)
    ids = [7, 3, 2 ** 33 + 5, 3]
    for mode in ("host", "device"):
        if mode == "host":
            arena4.draw(ids)
        else:
            arena4.set_samples(ids)
            arena4.draw(None)
        for L in arena4.layers:
            for which in ("weight", "bias"):
                if which == "bias" and not L.bias:
                    continue
                v = arena4.views[id(L)][which]
                assert v.shape[0] == 4 and (v.is_contiguous() or v[0].numel() % 4 != 0)   # contiguous for 16-byte multiples (every BEM tensor)
                sid = 2 * int(L.layer_id) + (which == "bias")
                for j, sample in enumerate(ids):
                    w, _ = BF.sample_weights(getattr(L, "mu_" + which), getattr(L, "rho_" + which), None, 1, 99, sid, sample)
                    assert torch.equal(w[0], v[j]), (mode, L.layer_id, which, j)
    arena4.draw(10)                                    # an int: consecutive samples 10, 11, 12, 13
    w, _ = BF.sample_weights(arena4.layers[0].mu_weight, arena4.layers[0].rho_weight, None, 1, 99, 2 * int(arena4.layers[0].layer_id), 12)
    assert torch.equal(w[0], arena4.views[id(arena4.layers[0])]["weight"][2])


def test_constant_weight_pack_cache_tracks_in_place_updates():
    """deterministic 1x1 layers pack their weights once (prepacked = 1 on later calls); an in-place weight / norm update or a
    change of the input's alignment class must trigger a repack"""
    from bem_b200 import network
    from bem_b200.ss2d import LayerNorm2d
    torch.manual_seed(11)
    conv = network.Conv2d(24, 40, 1, bias=True).cuda()
    norm = LayerNorm2d(24).cuda()
    x = torch.randn(2, 24, 12, 16, device="cuda")

    def ref(xx):
        return torch.nn.functional.conv2d(norm(xx).double(), conv.weight.double(), conv.bias.double())

    with torch.no_grad():
        for _ in range(3):                                   # first call packs, the others reuse
            assert nmax_err(conv(x, pre_norm=norm).cpu().numpy(), ref(x).cpu().numpy()) < TOL
        conv.weight.mul_(1.5)                                # optimizer-style in-place update
        assert nmax_err(conv(x, pre_norm=norm).cpu().numpy(), ref(x).cpu().numpy()) < TOL
        norm.weight.add_(0.25)
        assert nmax_err(conv(x, pre_norm=norm).cpu().numpy(), ref(x).cpu().numpy()) < TOL
        x_odd = torch.randn(1, 24, 5, 7, device="cuda")      # 35 pixels: the unaligned kernel with its own tiling
        assert nmax_err(conv(x_odd, pre_norm=norm).cpu().numpy(), ref(x_odd).cpu().numpy()) < TOL
        assert nmax_err(conv(x, pre_norm=norm).cpu().numpy(), ref(x).cpu().numpy()) < TOL


@pytest.mark.parametrize("B,cin,cout,H,W", [(1, 3, 40, 20, 28), (2, 40, 3, 9, 13), (1, 5, 7, 6, 600), (1, 3, 3, 1, 4), (1, 1, 17, 5, 5)])
def test_conv3x3_direct_matches_conv2d(B, cin, cout, H, W):
    """bem_conv3x3 (the network's 3x3 stems, UNet_arch.py:423-431) == F.conv2d in fp64, incl. ragged widths and tiny images"""
    import torch.nn.functional as F
    from bem_b200 import network
    from bem_b200.bayesian import functional as BF
    g = torch.Generator(device="cpu").manual_seed(cin * 50 + cout)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()
    b = torch.randn(cout, generator=g).cuda()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    assert nmax_err(BF.conv3x3_direct(x, w, b).cpu().numpy(), ref.cpu().numpy()) < TOL
    assert nmax_err(BF.conv3x3_direct(x, w, None).cpu().numpy(), (ref - b.double()[None, :, None, None]).cpu().numpy()) < TOL
    conv = network.Conv2d(cin, cout, 3, 1, 1).cuda()
    with torch.no_grad():
        conv.weight.copy_(w)
        conv.bias.copy_(b)
        y = conv(x)
    if min(cin, cout) <= 8:
        assert nmax_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL


@pytest.mark.parametrize("cin,cout,H,W,nslopes", [(160, 320, 10, 12, 1), (80, 80, 7, 5, 1), (24, 40, 8, 8, 40)])
def test_pointwise_fused_prelu(cin, cout, H, W, nslopes):
    """nn.PReLU after a 1x1 conv (DualUpSample, UNet_arch.py:113-135) folded into the epilogue, tensor-core kernels (aligned and
    unaligned inputs) and CUDA-core kernel; single slope and per-channel slopes"""
    import torch.nn.functional as F
    from bem_b200 import network
    from bem_b200.bayesian import functional as BF
    torch.manual_seed(cin + cout)
    conv = network.Conv2d(cin, cout, 1, bias=True).cuda()
    act = torch.nn.PReLU(nslopes).cuda()
    with torch.no_grad():
        act.weight.uniform_(0.05, 0.4)
    x = torch.randn(2, cin, H, W, device="cuda")
    with torch.no_grad():
        y = conv(x, post_prelu=act)
        ref = F.prelu(F.conv2d(x.double(), conv.weight.double(), conv.bias.double()), act.weight.double())
        simt = BF.pointwise_conv(x, conv.weight.view(1, cout, cin), conv.bias.view(1, -1), 1, force_simt=True, prelu=act.weight)
    assert nmax_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL
    assert nmax_err(simt.cpu().numpy(), ref.cpu().numpy()) < TOL
