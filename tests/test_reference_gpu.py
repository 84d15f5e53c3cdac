"""The sm_100a path against the UNMODIFIED reference running on the same GPU (staged by oracle/make_ref.py):

  * bem_scan_fwd / bem_scan_bwd  vs  the reference CUDA extension selective_scan_cuda_oflex (recompiled for sm_100a) on the
    reference's own grid and with its own tolerances (kernels/selective_scan/test_selective_scan.py:372-502), and both
    against selective_scan_ref on the short lengths;
  * cross_scan_fn / cross_merge_fn  vs  the reference's Triton kernels (basicsr/vmamba/models/csm_triton.py:278-505), bit-exact;
  * the Bayesian layers  vs  basicsr/bayesian on the GPU with the eps the reference drew;
  * bem_b200.patch.install() on the REAL reference models: the reference's stage-1 `Network` (basicsr/archs/UNet_arch.py),
    Bayesian-converted by the reference's own tools, unpatched (reference extension + Triton + eager layers) vs patched.

/root/reference is never read here; everything comes from oracle/_ref (git-ignored, travels with the snapshot).
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import nmax_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_loader as R  # noqa: E402

if not R.available():
    pytest.skip(R.why_unavailable(), allow_module_level=True)


def _bem():
    import bem_b200
    return bem_b200


# ---------------------------------------------------------------------------------------------------------------------
# selective scan vs the reference CUDA extension
# ---------------------------------------------------------------------------------------------------------------------
def _grid_inputs(seqlen, itype, varBC_groups, has_D, has_delta_bias, dstate=1, dim=768, batch_size=2):
    """test_selective_scan.py:404-441, same seed and draw order"""
    device = "cuda"
    torch.random.manual_seed(0)
    A = (-0.5 * torch.rand(dim, dstate, device=device, dtype=torch.float32))
    shp = (batch_size, dstate, seqlen) if varBC_groups == 1 else (batch_size, varBC_groups, dstate, seqlen)
    B = torch.randn(*shp, device=device, dtype=itype)
    C = torch.randn(*shp, device=device, dtype=itype)
    D = torch.randn(dim, device=device, dtype=torch.float32) if has_D else None
    bias = (0.5 * torch.rand(dim, device=device, dtype=torch.float32)) if has_delta_bias else None
    u = torch.randn(batch_size, dim, seqlen, device=device, dtype=itype)
    delta = (0.5 * torch.rand(batch_size, dim, seqlen, device=device, dtype=itype))
    return dict(u=u, delta=delta, A=A, B=B, C=C, D=D, delta_bias=bias)


def _run_api(fn, inp, softplus, g=None):
    leaves = {k: (v.detach().clone().requires_grad_() if v is not None else None) for k, v in inp.items()}
    out, state = fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"], z=None,
                    delta_bias=leaves["delta_bias"], delta_softplus=softplus, return_last_state=True)
    if g is None:
        g = torch.randn_like(out)
    out.backward(g)
    return out.detach(), state.detach(), {k: (v.grad if v is not None else None) for k, v in leaves.items()}, g


def _ref_test_api():
    """the reference's mamba-style wrapper around ITS extension: build_selective_scan_fn(selective_scan_cuda_oflex, "ssoflex")
    (test_selective_scan.py:18-165), taken from the staged file by AST like selective_scan_ref"""
    return R.build_selective_scan_fn()(R.oflex_ext(), mode="ssoflex")


@pytest.mark.parametrize("itype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("seqlen", [64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("varBC_groups", [1, 2])
@pytest.mark.parametrize("opts", [(False, False, False), (True, True, True), (True, False, True), (False, True, False)],
                         ids=["plain", "D_bias_softplus", "D_softplus", "bias"])
def test_scan_vs_reference_extension_on_its_grid(itype, seqlen, varBC_groups, opts):
    """the reference's grid (test_selective_scan.py:372-392: dim 768, batch 2, dstate 1, fp32/fp16/bf16 x 7 lengths x
    varBC_groups x D / bias / softplus) with ours in the seat of `selective_scan_fn` and the reference extension as
    `selective_scan_ref`, held to the tolerances the reference holds itself to (:398-401, :490-502)"""
    has_D, has_bias, softplus = opts
    rtol, atol = (6e-4, 2e-3) if itype == torch.float32 else (3e-3, 5e-3)
    if itype == torch.bfloat16:
        rtol, atol = 3e-2, 5e-2
    rtolw, atolw = 1e-3, 1e-3
    inp = _grid_inputs(seqlen, itype, varBC_groups, has_D, has_bias)
    ours = _bem().selective_scan_fn_test_api
    ref = _ref_test_api()
    out, state, gr, g = _run_api(ours, inp, softplus)
    out_r, state_r, gr_r, _ = _run_api(ref, inp, softplus, g)
    assert out.dtype == out_r.dtype and out.shape == out_r.shape
    assert torch.allclose(out, out_r, rtol=rtol, atol=atol)
    assert torch.allclose(state, state_r, rtol=rtol, atol=atol)
    assert torch.allclose(gr["u"], gr_r["u"], rtol=rtol * 2, atol=atol * 2)
    assert torch.allclose(gr["delta"], gr_r["delta"], rtol=rtol * 5, atol=atol * 10)
    assert torch.allclose(gr["A"], gr_r["A"], rtol=rtolw, atol=atolw * 5)
    assert torch.allclose(gr["B"], gr_r["B"], rtol=rtol, atol=atol)
    assert torch.allclose(gr["C"], gr_r["C"], rtol=rtol, atol=atol)
    if has_D:
        assert torch.allclose(gr["D"], gr_r["D"], rtol=rtolw, atol=atolw)
    if has_bias:
        assert torch.allclose(gr["delta_bias"], gr_r["delta_bias"], rtol=rtolw, atol=atolw)
    if itype == torch.float32:   # north_star: 1e-5 relative in fp32 against the reference CUDA extension is NOT attainable by
        # construction (the extension is built with --use_fast_math and is itself ~1e-6..1e-4 from selective_scan_ref);
        # what is checked: ours is at least as close to the extension as the extension's own distance from the oracle allows
        assert nmax_err(out.cpu().numpy(), out_r.cpu().numpy()) < 2e-4


@pytest.mark.parametrize("seqlen", [64, 256, 512])
@pytest.mark.parametrize("dstate", [1, 4, 16])
def test_scan_three_way_fp32(seqlen, dstate):
    """ours, the reference extension and the reference's selective_scan_ref (Python loop, on the GPU) on one input:
    ours must be within 1e-5 of selective_scan_ref (north_star) — and is closer to it than the reference extension is"""
    inp = _grid_inputs(seqlen, torch.float32, 2, True, True, dstate=dstate, dim=96)
    ours = _bem().selective_scan_fn_test_api
    ext = _ref_test_api()
    sref = R.selective_scan_ref()
    out, state, gr, g = _run_api(ours, inp, True)
    out_e, state_e, gr_e, _ = _run_api(ext, inp, True, g)
    out_s, state_s, gr_s, _ = _run_api(lambda *a, **k: sref(*a, **k), inp, True, g)
    e_ours = nmax_err(out.cpu().numpy(), out_s.cpu().numpy())
    e_ext = nmax_err(out_e.cpu().numpy(), out_s.cpu().numpy())
    assert e_ours < 1e-5, (e_ours, e_ext)
    assert nmax_err(state.cpu().numpy(), state_s.cpu().numpy()) < 1e-5
    for k in ("u", "delta", "A", "B", "C", "D", "delta_bias"):
        eo = nmax_err(gr[k].cpu().numpy(), gr_s[k].cpu().numpy())
        assert eo < 2e-5, (k, eo, nmax_err(gr_e[k].cpu().numpy(), gr_s[k].cpu().numpy()))


def test_extension_module_is_a_drop_in_for_the_reference_wrapper():
    """the reference's own wrapper class (build_selective_scan_fn) runs unchanged on this package's module object in place
    of the pybind module: same positional fwd / bwd signature, same return lists (selective_scan_oflex.cpp:157-358)"""
    wrapped_ours = R.build_selective_scan_fn()(_bem().selective_scan_cuda_oflex, mode="ssoflex")
    inp = _grid_inputs(1024, torch.float32, 2, True, True, dim=64)
    out, state, gr, g = _run_api(wrapped_ours, inp, True)
    out_r, state_r, gr_r, _ = _run_api(_ref_test_api(), inp, True, g)
    assert torch.allclose(out, out_r, rtol=6e-4, atol=2e-3)
    for k in gr:
        assert torch.allclose(gr[k], gr_r[k], rtol=3e-3, atol=2e-2), k


def test_csms6s_product_api_vs_reference_on_gpu():
    """csms6s.selective_scan_fn (csms6s.py:116-130) with the reference extension vs bem_b200.selective_scan_fn, BEM shapes"""
    ref = R.csms6s(True)
    assert ref.WITH_SELECTIVESCAN_OFLEX
    bem = _bem()
    for (Bn, KD, L) in ((1, 160, 6000), (2, 320, 1500), (8, 640, 256)):
        torch.manual_seed(L)
        u = torch.randn(Bn, KD, L, device="cuda")
        delta = 0.5 * torch.rand(Bn, KD, L, device="cuda")
        A = -torch.rand(KD, 1, device="cuda") - 0.5
        Bm = torch.randn(Bn, 4, 1, L, device="cuda")
        Cm = torch.randn(Bn, 4, 1, L, device="cuda")
        D = torch.randn(KD, device="cuda")
        bias = 0.5 * torch.rand(KD, device="cuda")
        a = bem.selective_scan_fn(u, delta, A, Bm, Cm, D, bias, True, True)
        b = ref.selective_scan_fn(u, delta, A, Bm, Cm, D, bias, True, True)
        assert a.dtype == b.dtype == torch.float32
        assert torch.allclose(a, b, rtol=6e-4, atol=2e-3)
        assert nmax_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# cross scan / merge vs the reference Triton kernels
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 40, 56, 57), (1, 40, 100, 150), (1, 80, 64, 64), (1, 8, 33, 31)])
@pytest.mark.parametrize("scans", [0, 1, 2])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_cross_scan_merge_vs_reference_triton(shape, scans, dtype):
    """csm_triton.py:491-505 API, Triton path (force_torch=False) on the GPU; 56 x 57 is the reference's own check shape
    (csm_triton.py:510-640). Data movement and a 4-term sum in the tensor's dtype: bit-exact."""
    ref = R.csm_triton()
    bem = _bem()
    Bn, Cc, H, W = shape
    torch.manual_seed(0)
    x = torch.randn(Bn, Cc, H, W, device="cuda", dtype=dtype)
    a = bem.cross_scan_fn(x, True, True, False, scans)
    b = ref.cross_scan_fn(x, True, True, False, scans, force_torch=False)
    assert a.shape == b.shape and torch.equal(a, b)
    ys = torch.randn(Bn, 4, Cc, H, W, device="cuda", dtype=dtype)
    m = bem.cross_merge_fn(ys, True, True, False, scans)
    mr = ref.cross_merge_fn(ys, True, True, False, scans, force_torch=False)
    assert m.shape == mr.shape
    if dtype == torch.float32:
        # Triton sums the four directions in its own order; ours keeps the torch path's (y0 + y2) + (y1 + y3)
        # (csm_triton.py:60-62). Equal up to one rounding of a 4-term fp32 sum.
        assert torch.allclose(m, mr, rtol=0, atol=4e-6 * float(ys.abs().max()))
        mt = ref.cross_merge_fn(ys, True, True, False, scans, force_torch=True)
        assert torch.equal(m, mt)
    else:
        assert torch.allclose(m.float(), mr.float(), rtol=2e-3, atol=2e-2)
    # one_by_one traversal of per-direction channel blocks
    x4 = torch.randn(Bn, 4, Cc, H, W, device="cuda", dtype=dtype)
    a4 = bem.cross_scan_fn(x4, True, True, True, scans)
    b4 = ref.cross_scan_fn(x4, True, True, True, scans, force_torch=False)
    assert torch.equal(a4, b4)


# ---------------------------------------------------------------------------------------------------------------------
# Bayesian layers vs basicsr/bayesian on the GPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["pw", "dw", "lin2d"])
def test_bayesian_layers_vs_reference_layers_same_eps(kind):
    """run the reference layer on the GPU, read the eps it left in its buffers (bayesian/conv.py:107,110) and feed it to ours"""
    refb = R.bayesian()
    from bem_b200 import bayesian as ours
    torch.manual_seed(11)
    if kind == "pw":
        args = dict(in_channels=40, out_channels=320, kernel_size=1, bias=True)
        r, o = refb.Conv2dReparameterization(**args), ours.Conv2dReparameterization(**args)
        x = torch.randn(1, 40, 60, 100, device="cuda")
    elif kind == "dw":
        args = dict(in_channels=320, out_channels=320, kernel_size=3, padding=1, groups=320, bias=True)
        r, o = refb.Conv2dReparameterization(**args), ours.Conv2dReparameterization(**args)
        x = torch.randn(1, 320, 60, 100, device="cuda")
    else:
        r, o = refb.Linear2dReparameterization(40, 40, bias=False), ours.Linear2dReparameterization(40, 40, bias=False)
        x = torch.randn(2, 40, 30, 50, device="cuda")
    r = r.cuda().eval()
    with torch.no_grad():
        for p in r.parameters():
            p.add_(0.05 * torch.randn_like(p))
    o.load_state_dict(r.state_dict(), strict=True)
    o = o.cuda().eval()
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
        yr = r(x)
        inj = {"weight": r.eps_weight.clone()}
        if r.bias:
            inj["bias"] = r.eps_bias.clone()
        o._injected_eps = inj
        yo = o(x)
    assert nmax_err(yo.cpu().numpy(), yr.cpu().numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# patch.install() on the real reference models
# ---------------------------------------------------------------------------------------------------------------------
def _collect_eps(net):
    eps = {}
    for n, m in net.named_modules():
        if hasattr(m, "deterministic") and hasattr(m, "eps_weight"):
            eps[n] = {"weight": m.eps_weight.detach().clone()}
            if getattr(m, "bias", False) and getattr(m, "eps_bias", None) is not None:
                eps[n]["bias"] = m.eps_bias.detach().clone()
    return eps


def test_patch_install_on_the_reference_network_end_to_end():
    """The reference's own stage-1 `Network` (UNet_arch.build_model: n_feat 40, blocks [2,2,2]) converted by the reference's
    convert2bnn_selective, run (1) unpatched: reference CUDA extension + Triton cross scan/merge + eager Bayesian layers, and
    (2) after bem_b200.patch.install(): same reference model code, same weights, same eps, every hot-path op on libbem_b200."""
    import bem_b200
    unet = R.unet_arch(True)
    refb = R.bayesian()
    vm = R.vmamba(True)
    assert vm.selective_scan_fn.__module__.endswith("csms6s")
    torch.manual_seed(21)
    net = unet.build_model()
    refb.convert2bnn_selective(net, {"sigma_init": 0.05, "decay": 0.998, "pretrain": False})
    net = net.cuda().eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.02 * torch.randn_like(p))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(1, 3, 96, 144, device="cuda")
    with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
        refb.set_prediction_type(net, deterministic=True)
        y_det_ref = net(x)[-1].clone()
        refb.set_prediction_type(net, deterministic=False)
        y_mc_ref = net(x)[-1].clone()
    eps = _collect_eps(net)
    assert len(eps) == 60

    patched = bem_b200.patch.install()
    try:
        assert "bayesian" in patched and any(p.endswith("vmamba") for p in patched)
        assert vm.selective_scan_fn is bem_b200.selective_scan_fn
        assert sys.modules["bayesian"] is bem_b200.bayesian
        net2 = unet.build_model()
        sys.modules["bayesian"].convert2bnn_selective(net2, {"sigma_init": 0.05, "decay": 0.998, "pretrain": False})
        net2.load_state_dict(sd, strict=True)
        net2 = net2.cuda().eval()
        mods = dict(net2.named_modules())
        assert sorted(eps) == sorted(n for n, m in mods.items() if hasattr(m, "deterministic"))
        _lib = bem_b200._lib
        _lib.profile.reset()
        with torch.no_grad(), torch.backends.cudnn.flags(allow_tf32=False):
            bem_b200.bayesian.set_prediction_type(net2, deterministic=True)
            y_det = net2(x)[-1]
            for n, e in eps.items():
                mods[n]._injected_eps = e
            bem_b200.bayesian.set_prediction_type(net2, deterministic=False)
            y_mc = net2(x)[-1]
        assert _lib.profile.launches > 100          # the patched model really ran on libbem_b200.so
        assert nmax_err(y_det.cpu().numpy(), y_det_ref.cpu().numpy()) < 1e-4
        assert nmax_err(y_mc.cpu().numpy(), y_mc_ref.cpu().numpy()) < 1e-4
    finally:
        bem_b200.patch.uninstall()
    assert sys.modules["bayesian"] is refb and vm.selective_scan_fn is not bem_b200.selective_scan_fn


# ---------------------------------------------------------------------------------------------------------------------
# no-reference scorer and LinearReparameterization on the 1x1 kernel (a14)
# ---------------------------------------------------------------------------------------------------------------------
def _textured(seed, H=400, W=600):
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, H // 8 + 2, W // 8 + 2, generator=g)
    up = torch.nn.functional.interpolate(base, size=(H, W), mode="bicubic", align_corners=False)
    return (up + 0.03 * torch.randn(1, 3, H, W, generator=g)).clamp(0, 1)[0]


def test_niqe_scorer_vs_reference_numpy_niqe_and_selection():
    """bem_b200.NiqeScorer (float64 on the device, batched) against basicsr/metrics/niqe.py::calculate_niqe called exactly as
    Enhancement/eval.py:250 calls it (`calculate_niqe(pred*255, crop_border=0)` on the (H, W, 3) float prediction), on textured
    600x400 images and on a 300x200 one; then the selection of eval.py:272-275 (`index(min(...))`) on the device scores."""
    import bem_b200
    niqe_ref = R.niqe_module().calculate_niqe
    scorer = bem_b200.NiqeScorer(os.path.join(R.ROOT, "basicsr/metrics/niqe_pris_params.npz"))
    preds = torch.stack([_textured(s) for s in range(8)])
    ref = [niqe_ref(p.permute(1, 2, 0).numpy() * 255, crop_border=0) for p in preds]
    ours = scorer(preds.cuda())
    assert ours.dtype == torch.float32 and ours.shape == (8,)
    rel = np.abs(ours.cpu().numpy().astype(np.float64) - np.array(ref)) / np.array(ref)
    assert rel.max() < 1e-3, rel
    idx, val = bem_b200.mc.select_best(ours, take_min=True)
    assert int(idx) == ref.index(min(ref))
    small = torch.stack([_textured(20 + s, 200, 300) for s in range(3)])
    ref_s = [niqe_ref(p.permute(1, 2, 0).numpy() * 255, crop_border=0) for p in small]
    rel = np.abs(scorer(small.cuda()).cpu().numpy().astype(np.float64) - np.array(ref_s)) / np.array(ref_s)
    assert rel.max() < 1e-3, rel
    with pytest.raises(RuntimeError):
        scorer(preds)                              # CUDA only, like every operator of the package


def test_mc_infer_with_the_niqe_scorer_selects_the_reference_minimum():
    """the MC loop end to end on the device: MCSampler predictions of a small Bayesian network, scored by NiqeScorer, selected by
    bem_select_best (take_min) — same index as the reference's numpy NIQE over the same predictions"""
    import bem_b200
    from bem_b200 import mc, network
    torch.manual_seed(0)
    net = network.build_bayesian_model().cuda().eval()
    x = torch.rand(1, 3, 192, 288, device="cuda")
    scorer = bem_b200.NiqeScorer(os.path.join(R.ROOT, "basicsr/metrics/niqe_pris_params.npz"))
    sampler = mc.MCSampler(net, seed=5, arena=True, graph=False)
    res = mc.mc_infer(sampler, x, 6, score_fn=scorer, take_min=True)
    preds = sampler.sample(x, list(range(6)))
    niqe_ref = R.niqe_module().calculate_niqe
    ref = [niqe_ref(p.permute(1, 2, 0).cpu().numpy() * 255, crop_border=0) for p in preds]
    assert res["index"] == ref.index(min(ref))
    assert torch.equal(res["best"], preds[res["index"]])
    assert np.allclose(res["scores"].cpu().numpy(), np.array(ref), rtol=1e-3)


@pytest.mark.parametrize("S", [1, 3])
def test_linear_reparameterization_on_the_pointwise_kernel_vs_reference(S):
    """LinearReparameterization (linear.py:106-203): the reference layer on the GPU and ours with the same eps; ours runs the
    contraction on bem_bayes_pointwise (rows as pixels), also with S weight samples per call"""
    refb = R.bayesian()
    import bem_b200
    from bem_b200 import bayesian as ours
    torch.manual_seed(3)
    r = refb.LinearReparameterization(40, 24, bias=True).cuda().eval()
    with torch.no_grad():
        r.mu_bias.normal_()
    o = ours.LinearReparameterization(40, 24, bias=True)
    o.load_state_dict(r.state_dict(), strict=True)
    o = o.cuda().eval()
    x = torch.randn(S, 6, 10, 40, device="cuda")
    outs, epsw, epsb = [], [], []
    with torch.no_grad():
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            for s in range(S):
                outs.append(r(x[s]))
                epsw.append(r.eps_weight.clone())
                epsb.append(r.eps_bias.clone())
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        bem_b200._lib.profile.reset()
        ours.set_mc_config(o, mc_samples=S)
        y = o(x.reshape(S * 6, 10, 40), eps_weight=torch.stack(epsw), eps_bias=torch.stack(epsb))
        ours.set_mc_config(o, mc_samples=1)
    assert bem_b200._lib.profile.launches >= 2            # sample kernel(s) + the pointwise kernel
    assert nmax_err(y.reshape(S, 6, 10, 24).cpu().numpy(), torch.stack(outs).cpu().numpy()) < 1e-5


def test_patch_install_on_the_reference_training_model_forward_and_backward():
    """BASELINE configs[4]: the reference's own DecompDualBranchDDWavelet (Options/DecompDualBranch2DDWavelet_4.yml: 18 VSSBlocks,
    d_state 1, frozen quaternion / wavelet decomposition in front), training mode, forward + L1 loss + backward on 2 x 6 x 64 x 64:
    unpatched (reference CUDA extension + Triton traversal) vs after bem_b200.patch.install() (this package's scan forward /
    backward and traversal kernels under the reference's autograd graph), same weights. Outputs and every parameter gradient."""
    import bem_b200
    torch.manual_seed(31)
    net = R.train_model(True).train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(2, 6, 64, 64, device="cuda")
    tgt = torch.rand(2, 3, 64, 64, device="cuda")

    def step(model):
        model.zero_grad(set_to_none=True)
        with torch.backends.cudnn.flags(allow_tf32=False):
            out = model(x)[-1]
            loss = torch.nn.functional.l1_loss(out, tgt)
            loss.backward()
        return out.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out_ref, g_ref = step(net)
        bem_b200.patch.install()
        try:
            net2 = R.train_model(True).train()          # built after the patch: forward_core binds the patched forward_corev2
            net2.load_state_dict(sd, strict=True)
            bem_b200._lib.profile.reset()
            out, g = step(net2)
            launches = bem_b200._lib.profile.launches
        finally:
            bem_b200.patch.uninstall()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert launches >= 18 * 4          # per VSSBlock: cross_scan, scan fwd, cross_merge forward + scan bwd (+ traversal adjoints) backward
    assert out.shape == out_ref.shape == (2, 3, 64, 64)
    assert nmax_err(out.cpu().numpy(), out_ref.cpu().numpy()) < 1e-4
    assert set(g) == set(g_ref) and len(g) > 300
    worst = max((nmax_err(g[n].cpu().numpy(), g_ref[n].cpu().numpy()), n) for n in g if float(g_ref[n].abs().max()) > 0)
    assert worst[0] < 2e-3, worst       # the reference extension is built with --use_fast_math and sums dB / dC with fp32 atomics
