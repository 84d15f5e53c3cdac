"""GPU parity of the sm_100a selective scan against the oracle and the reference's golden vectors.

Bars (BASELINE.json north_star / SURVEY 8c): fp32 -> normalised max error <= 1e-5 on out, last_state and every gradient
(measured against the fp64 oracle), plus the reference's own allclose tolerances (test_selective_scan.py:398-401,
490-502) as a secondary gate; fp16 / bf16 -> 1e-2.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import nmax_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
LOW_TOL = 1e-2


def _bem():
    import bem_b200
    return bem_b200


def make_inputs(Bt, KD, N, G, L, dtype, has_D=True, has_bias=True, seed=0, bc3d=False, dev="cuda"):
    """distributions of test_selective_scan.py:406-441"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = -0.5 * torch.rand(KD, N, generator=g)
    bshape = (Bt, N, L) if bc3d else (Bt, G, N, L)
    Bm = torch.randn(*bshape, generator=g).to(dtype)
    Cm = torch.randn(*bshape, generator=g).to(dtype)
    D = torch.randn(KD, generator=g) if has_D else None
    bias = 0.5 * torch.rand(KD, generator=g) if has_bias else None
    u = torch.randn(Bt, KD, L, generator=g).to(dtype)
    delta = (0.5 * torch.rand(Bt, KD, L, generator=g)).to(dtype)
    dout = torch.randn(Bt, KD, L, generator=g)
    mv = lambda t: None if t is None else t.to(dev)
    return dict(u=mv(u), delta=mv(delta), A=mv(A), B=mv(Bm), C=mv(Cm), D=mv(D), delta_bias=mv(bias), dout=mv(dout))


def run_case(inp, softplus, tol, check_bwd=True, ref_tols=None):
    bem = _bem()
    dev = inp["u"].device
    leaves = {k: (v.clone().requires_grad_() if v is not None and k != "dout" else v) for k, v in inp.items()}
    out, last = bem.selective_scan_fn_test_api(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                                               None, leaves["delta_bias"], softplus, return_last_state=True)
    assert bem._lib.scan_error_word(dev) == 0
    o = oracle.selective_scan_oracle_f64(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"],
                                         softplus, dout=inp["dout"] if check_bwd else None)
    errs = {"out": nmax_err(out.detach().float().cpu().numpy(), o["out"]),
            "last_state": nmax_err(last.cpu().numpy(), o["last_state"])}
    if check_bwd:
        out.backward(inp["dout"].to(out.dtype))
        assert bem._lib.scan_error_word(dev) == 0
        pairs = [("du", "u"), ("ddelta", "delta"), ("dA", "A"), ("dB", "B"), ("dC", "C")]
        if inp["D"] is not None:
            pairs.append(("dD", "D"))
        if inp["delta_bias"] is not None:
            pairs.append(("ddelta_bias", "delta_bias"))
        for gk, lk in pairs:
            errs[gk] = nmax_err(leaves[lk].grad.float().cpu().numpy(), o[gk])
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, f"normalised max errors over {tol}: {bad} (all: {errs})"
    return errs


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["n1_l70", "n4_g2_l300", "n16_g4_l600", "n2_z_l129", "n1_g4_l1000", "bf16_n2_l128", "f16_n1_l96"])
def test_golden_reference_vectors(golden_scan, name):
    """the committed outputs of the REAL selective_scan_ref (+ its autograd) on the same inputs"""
    bem = _bem()
    c = golden_scan.case(name)
    dt = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}[int(c["dtype"])]
    tol = FP32_TOL if dt == torch.float32 else LOW_TOL
    t = lambda k, d=torch.float32: torch.tensor(c[k], device="cuda").to(d) if k in c else None
    u, delta, Bm, Cm = (t(k, dt).requires_grad_() for k in ("u", "delta", "B", "C"))
    A = t("A").requires_grad_()
    D = t("D").requires_grad_() if "D" in c else None
    bias = t("delta_bias").requires_grad_() if "delta_bias" in c else None
    z = t("z", dt).requires_grad_() if "z" in c else None
    out, last = bem.selective_scan_fn_test_api(u, delta, A, Bm, Cm, D, z, bias, bool(c["softplus"]), return_last_state=True)
    assert out.dtype == dt
    assert nmax_err(out.detach().float().cpu().numpy(), c["out"]) < tol
    assert nmax_err(last.detach().cpu().numpy(), c["last_state"]) < tol
    out.backward(t("dout", dt))
    gtol = 3e-5 if dt == torch.float32 else 2e-2
    for gk, leaf in (("du", u), ("ddelta", delta), ("dA", A), ("dB", Bm), ("dC", Cm), ("dD", D), ("ddelta_bias", bias), ("dz", z)):
        if leaf is not None:
            assert nmax_err(leaf.grad.float().cpu().numpy(), c[gk]) < gtol, gk


def test_product_api_golden(golden_scan):
    """csms6s.selective_scan_fn contract: 4-D B/C, fp32 'oflex' output (csms6s.py:116-130)"""
    bem = _bem()
    c = golden_scan.case("csms6s")
    t = lambda k: torch.tensor(c[k], device="cuda")
    out = bem.selective_scan_fn(t("u"), t("delta"), t("A"), t("B"), t("C"), t("D"), t("delta_bias"), True, True)
    assert out.dtype == torch.float32
    assert nmax_err(out.cpu().numpy(), c["out"]) < FP32_TOL


@pytest.mark.parametrize("L", [1, 31, 64, 128, 256, 383, 384, 385, 512, 1024, 2048, 4096, 5000])
@pytest.mark.parametrize("N,G", [(1, 1), (1, 2), (4, 2), (16, 4)])
def test_fp32_seqlens_and_states(L, N, G):
    """seqlen grid of test_selective_scan.py:376 extended with ragged / multi-chunk / unaligned lengths"""
    inp = make_inputs(2, 24 * G, N, G, L, torch.float32, seed=L + N)
    run_case(inp, True, FP32_TOL)


@pytest.mark.parametrize("has_D", [False, True])
@pytest.mark.parametrize("has_bias", [False, True])
@pytest.mark.parametrize("softplus", [False, True])
@pytest.mark.parametrize("bc3d", [False, True])
def test_fp32_option_grid(has_D, has_bias, softplus, bc3d):
    """has_delta_bias x delta_softplus x has_D x B/C rank (test_selective_scan.py:378-386)"""
    inp = make_inputs(2, 40, 1, 1, 1000, torch.float32, has_D, has_bias, seed=7, bc3d=bc3d)
    run_case(inp, softplus, FP32_TOL)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("L", [64, 512, 513, 2048])
@pytest.mark.parametrize("N", [1, 2])
def test_low_precision_inputs(dtype, L, N):
    inp = make_inputs(2, 32, N, 2, L, dtype, seed=3)
    bem = _bem()
    out = bem.selective_scan_fn(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True, True)
    assert out.dtype == torch.float32      # oflex: fp32 out for 16-bit in (selective_scan_oflex.cpp:219)
    out2 = bem.selective_scan_fn(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True, False)
    assert out2.dtype == dtype
    run_case(inp, True, LOW_TOL)


def test_reference_allclose_tolerances():
    """the reference's own gate (test_selective_scan.py:398-401, 490-502) at its shape B=2, dim=768, dstate=1"""
    bem = _bem()
    for L in (64, 1024, 4096):
        inp = make_inputs(2, 768, 1, 2, L, torch.float32, seed=L)
        leaves = {k: (v.clone().requires_grad_() if k != "dout" else v) for k, v in inp.items()}
        out, last = bem.selective_scan_fn_test_api(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"],
                                                   leaves["D"], None, leaves["delta_bias"], True, return_last_state=True)
        out.backward(inp["dout"])
        o = oracle.selective_scan_oracle_f64(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"],
                                             inp["delta_bias"], True, dout=inp["dout"])
        rtol, atol, rtolw, atolw = 6e-4, 2e-3, 1e-3, 1e-3
        ac = lambda a, b, r, t: np.allclose(a.detach().cpu().numpy(), b, rtol=r, atol=t)
        assert ac(out, o["out"], rtol, atol) and ac(last, o["last_state"], rtol, atol)
        assert ac(leaves["u"].grad, o["du"], rtol * 2, atol * 2)
        assert ac(leaves["delta"].grad, o["ddelta"], rtol * 5, atol * 10)
        assert ac(leaves["A"].grad, o["dA"], rtolw, atolw * 5)
        assert ac(leaves["B"].grad, o["dB"], rtol, atol) and ac(leaves["C"].grad, o["dC"], rtol, atol)
        assert ac(leaves["D"].grad, o["dD"], rtolw, atolw) and ac(leaves["delta_bias"].grad, o["ddelta_bias"], rtolw, atolw)


def test_strided_inputs_and_extension_module_contract():
    """fwd/bwd accept any batch / dim strides with last-dim stride 1 (selective_scan_oflex.cpp:181-182) and return the
    reference's lists; x keeps `last_state = x[:, :, -1, 1::2]` (test_selective_scan.py:79)"""
    bem = _bem()
    ext = bem.selective_scan_cuda_oflex
    Bt, KD, N, G, L = 2, 16, 2, 2, 900
    inp = make_inputs(Bt, KD, N, G, L, torch.float32, seed=11)
    big_u = torch.randn(Bt, KD + 3, L + 8, device="cuda")
    u = big_u[:, 1:KD + 1, 4:L + 4]            # strided AND 16-byte aligned only on some rows
    u.copy_(inp["u"])
    xdbl = torch.randn(Bt, G, 3 + 2 * N, L, device="cuda")
    Bv, Cv = xdbl[:, :, 3:3 + N], xdbl[:, :, 3 + N:]
    Bv.copy_(inp["B"])
    Cv.copy_(inp["C"])
    out, x = ext.fwd(u, inp["delta"], inp["A"], Bv, Cv, inp["D"], inp["delta_bias"], True, 1, True)
    CL = bem.chunk_len(torch.float32)
    assert x.shape == (Bt, KD, (L + CL - 1) // CL, 2 * N) and x.dtype == torch.float32
    o = oracle.selective_scan_oracle_f64(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"],
                                         True, dout=inp["dout"])
    assert nmax_err(out.cpu().numpy(), o["out"]) < FP32_TOL
    assert nmax_err(x[:, :, -1, 1::2].cpu().numpy(), o["last_state"]) < FP32_TOL
    res = ext.bwd(u, inp["delta"], inp["A"], Bv, Cv, inp["D"], inp["delta_bias"], inp["dout"], x, True, 1)
    assert len(res) == 7
    for t, k in zip(res, ("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias")):
        assert nmax_err(t.cpu().numpy(), o[k]) < FP32_TOL, k
    # absent D / delta_bias -> None gradients (selective_scan_oflex.cpp:329-332)
    res = ext.bwd(u, inp["delta"], inp["A"], Bv, Cv, None, None, inp["dout"], x, True, 1)
    assert res[5] is None and res[6] is None


def test_carries_match_sequential_states():
    """x[b, d, c, 2n+1] is the recurrence state at the end of chunk c; x[..., 2n] the running decay"""
    bem = _bem()
    Bt, KD, N, G, L = 1, 8, 2, 1, 1500
    inp = make_inputs(Bt, KD, N, G, L, torch.float32, seed=5)
    _, x = bem.selective_scan_cuda_oflex.fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"],
                                             inp["delta_bias"], True, 1, True)
    CL = bem.chunk_len(torch.float32)
    for c in range(x.shape[2]):
        end = min(L, (c + 1) * CL)
        o = oracle.selective_scan_oracle_f64(inp["u"][..., :end], inp["delta"][..., :end], inp["A"], inp["B"][..., :end],
                                             inp["C"][..., :end], inp["D"], inp["delta_bias"], True)
        assert nmax_err(x[:, :, c, 1::2].cpu().numpy(), o["last_state"]) < FP32_TOL
        dl = torch.nn.functional.softplus(inp["delta"][..., :end].double() + inp["delta_bias"].double()[None, :, None])
        decay = torch.exp(dl.sum(-1)[..., None] * inp["A"].double()[None])
        assert nmax_err(x[:, :, c, 0::2].cpu().numpy(), decay.cpu().numpy()) < 1e-4


def test_error_contract():
    """argument violations raise RuntimeError like TORCH_CHECK (selective_scan_oflex.cpp:166-216)"""
    bem = _bem()
    inp = make_inputs(1, 8, 1, 1, 64, torch.float32)
    f = bem.selective_scan_cuda_oflex.fwd
    with pytest.raises(RuntimeError):
        f(inp["u"].cpu(), inp["delta"], inp["A"], inp["B"], inp["C"], None, None, True, 1, True)          # CPU tensor
    with pytest.raises(RuntimeError):
        f(inp["u"].half(), inp["delta"], inp["A"], inp["B"], inp["C"], None, None, True, 1, True)         # dtype mismatch
    with pytest.raises(RuntimeError):
        f(inp["u"], inp["delta"], inp["A"].half(), inp["B"], inp["C"], None, None, True, 1, True)         # A not fp32
    with pytest.raises(RuntimeError):
        f(inp["u"].transpose(1, 2).contiguous().transpose(1, 2), inp["delta"], inp["A"], inp["B"], inp["C"], None, None, True, 1, True)
    with pytest.raises(RuntimeError):
        f(inp["u"], inp["delta"], inp["A"], inp["B"][:, :, :, :32], inp["C"], None, None, True, 1, True)  # shape
    with pytest.raises(RuntimeError):
        bem.selective_scan_fn(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], backend="torch")      # no fallback backends


@pytest.mark.parametrize("shape", ["BEM-I-L2", "BEM-T-L0", "C1"])
def test_full_size_shapes(shape):
    """BASELINE shapes: config 1 and two BEM level shapes, forward + backward against the fp64 oracle"""
    cfg = {"BEM-I-L2": (1, 640, 1, 4, 15000), "BEM-T-L0": (8, 160, 1, 4, 4096), "C1": (1, 384, 16, 4, 4096)}[shape]
    inp = make_inputs(*cfg, torch.float32, seed=1)
    run_case(inp, True, FP32_TOL)


def test_full_size_forward_bem_i_l0_linearity():
    """BEM-I level 0 (B1 KD160 L240000): oracle parity on the forward, and the size-independent property that the scan
    is linear in u for fixed delta/B/C (out(u1 + u2) - D-term consistency)"""
    bem = _bem()
    inp = make_inputs(1, 160, 1, 4, 240000, torch.float32, seed=2)
    args = (inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True, True)
    out1 = bem.selective_scan_fn(inp["u"], *args)
    ref = oracle.selective_scan_oracle(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None,
                                       inp["delta_bias"], True)
    assert nmax_err(out1.cpu().numpy(), ref) < FP32_TOL
    u2 = torch.randn_like(inp["u"])
    out2 = bem.selective_scan_fn(u2, *args)
    out12 = bem.selective_scan_fn(inp["u"] + u2, *args)
    assert nmax_err((out1 + out2).cpu().numpy(), out12.cpu().numpy()) < 2e-5
    assert bem._lib.scan_error_word(inp["u"].device) == 0


@pytest.mark.parametrize("cfg", [(1, 160, 1, 4, 24000), (2, 40, 1, 1, 1536), (1, 64, 4, 2, 3000), (1, 32, 1, 1, 384)])
def test_forward_is_bit_reproducible(cfg):
    """the look-back combines tile aggregates in a fixed order, so repeated launches agree bit for bit (the reference
    forward is sequential and therefore deterministic too); doubles as a race detector"""
    bem = _bem()
    inp = make_inputs(*cfg, torch.float32, seed=9)
    args = (inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True, 1, True)
    out0, x0 = bem.selective_scan_cuda_oflex.fwd(*args)
    for _ in range(30):
        out, x = bem.selective_scan_cuda_oflex.fwd(*args)
        assert torch.equal(out, out0) and torch.equal(x, x0)
    assert bem._lib.scan_error_word(inp["u"].device) == 0


def test_scan_workspace_rearms_itself_across_launches_and_graph_replays():
    """The look-back workspace is never cleared between launches (descriptor epoch + ticket re-armed by the kernel): back to
    back launches of different shapes on one workspace, and replays of a captured CUDA graph (identical kernel arguments
    every time), must all give the result of a fresh launch, bit for bit, forward and backward."""
    import bem_b200
    from bem_b200 import selective_scan as ss
    torch.manual_seed(5)
    dev = "cuda"

    def mk(Bn, KD, L, N=1, G=2):
        u = torch.randn(Bn, KD, L, device=dev)
        dl = 0.5 * torch.randn(Bn, KD, L, device=dev)
        A = -torch.rand(KD, N, device=dev) - 0.2
        Bm = torch.randn(Bn, G, N, L, device=dev)
        Cm = torch.randn(Bn, G, N, L, device=dev)
        D = torch.randn(KD, device=dev)
        bias = torch.randn(KD, device=dev)
        return u, dl, A, Bm, Cm, D, bias

    big, small = mk(1, 16, 20000), mk(2, 8, 3000)
    ref_big = ss.fwd(*big, True, 1, True)
    ref_small = ss.fwd(*small, True, 1, True)
    for _ in range(3):   # alternate shapes: every launch sees the other shape's stale descriptors
        a = ss.fwd(*small, True, 1, True)
        b = ss.fwd(*big, True, 1, True)
        assert torch.equal(a[0], ref_small[0]) and torch.equal(b[0], ref_big[0]) and torch.equal(b[1], ref_big[1])
    dout = torch.randn_like(ref_big[0])
    ref_bwd = ss.bwd(*big, dout, ref_big[1], True, 1)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ss.fwd(*big, True, 1, True)
        ss.bwd(*big, dout, ref_big[1], True, 1)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_g = ss.fwd(*big, True, 1, True)
        bwd_g = ss.bwd(*big, dout, out_g[1], True, 1)
    for _ in range(4):
        out_g[0].zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out_g[0], ref_big[0])
        assert torch.equal(bwd_g[0], ref_bwd[0]) and torch.equal(bwd_g[1], ref_bwd[1])
    assert bem_b200._lib.scan_error_word(torch.device("cuda", torch.cuda.current_device())) == 0


@pytest.mark.parametrize("Bn,G,Dg,L,R", [(1, 4, 40, 2400, 3), (2, 2, 10, 777, 5), (1, 1, 3, 70, 1), (1, 4, 16, 5000, 8)])
def test_scan_fused_dt_proj_matches_projection_then_scan(Bn, G, Dg, L, R):
    """dt_rank > 0: the scan kernel forms delta = dt_projs_weight . dt_lowrank itself (SS2Dv2.forward_corev2,
    vmamba.py:660-661 runs the grouped conv1d first). Same result as projecting in fp64 and scanning, on out and carries;
    incl. unaligned / ragged rows, strided low-rank views and more than one batch."""
    from bem_b200 import selective_scan as ss
    g = torch.Generator(device="cpu").manual_seed(L + R)
    KD, N = G * Dg, 1
    u = torch.randn(Bn, KD, L, generator=g).cuda()
    x_dbl = torch.randn(Bn, G, R + 2 * N, L, generator=g).cuda()           # as x_proj leaves it: dt | B | C per direction
    dtl, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=2)                      # strided views, no copy
    Wdt = (torch.randn(KD, R, generator=g) * 0.5).cuda()
    A = (-torch.rand(KD, N, generator=g) - 0.2).cuda()
    D = torch.randn(KD, generator=g).cuda()
    bias = torch.randn(KD, generator=g).cuda()
    out, xc = ss.fwd(u, dtl, A, Bm, Cm, D, bias, True, 1, True, dt_weight=Wdt)
    delta = torch.einsum("bgrl,gdr->bgdl", dtl.double(), Wdt.double().view(G, Dg, R)).reshape(Bn, KD, L)
    ref, xref = ss.fwd(u, delta.float().contiguous(), A, Bm.contiguous(), Cm.contiguous(), D, bias, True, 1, True)
    assert nmax_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    assert nmax_err(xc.cpu().numpy(), xref.cpu().numpy()) < 1e-5
    out2, _ = ss.fwd(u, dtl, A, Bm, Cm, D, bias, True, 1, True, dt_weight=Wdt)
    assert torch.equal(out, out2)                                           # deterministic


def test_classic_schedule_still_matches_the_deferred_one():
    """fp32 / dstate-1 launches take the deferred-finish kernel (384-position tiles); the classic schedule (768-position
    tiles) still serves 16-bit inputs and general dstate, and this shape under BEM_FWD_CLASSIC=1. The two associate the
    recurrence differently across tiles, so they are compared at the fp32 parity tolerance, out and carries."""
    import os
    import subprocess
    import sys
    import tempfile
    from bem_b200 import selective_scan as ss
    torch.manual_seed(21)
    Bn, KD, G, N, L = 2, 16, 2, 1, 5000
    u = torch.randn(Bn, KD, L, device="cuda")
    dl = 0.5 * torch.randn(Bn, KD, L, device="cuda")
    A = -torch.rand(KD, N, device="cuda") - 0.2
    Bm = torch.randn(Bn, G, N, L, device="cuda")
    Cm = torch.randn(Bn, G, N, L, device="cuda")
    D = torch.randn(KD, device="cuda")
    bias = torch.randn(KD, device="cuda")
    out, x = ss.fwd(u, dl, A, Bm, Cm, D, bias, True, 1, True)
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "io.pt")
        torch.save({k: v.cpu() for k, v in dict(u=u, dl=dl, A=A, Bm=Bm, Cm=Cm, D=D, bias=bias).items()}, f)
        code = ("import torch, sys; sys.path.insert(0, %r); from bem_b200 import selective_scan as ss; t = torch.load(%r);"
                "t = {k: v.cuda() for k, v in t.items()};"
                "o, x = ss.fwd(t['u'], t['dl'], t['A'], t['Bm'], t['Cm'], t['D'], t['bias'], True, 1, True);"
                "torch.save({'o': o.cpu(), 'x': x.cpu()}, %r)") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), f, f + ".out")
        env = dict(os.environ, BEM_FWD_CLASSIC="1")
        subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
        ref = torch.load(f + ".out")
    assert nmax_err(out.cpu().numpy(), ref["o"].numpy()) < 2e-6
    assert nmax_err(x.cpu().numpy(), ref["x"].numpy()) < 2e-6


@pytest.mark.parametrize("lo,hi", [(-14.0, -6.0), (-9.5, -2.0)])
def test_softplus_keeps_relative_accuracy_near_the_dt_floor(lo, hi):
    """delta + bias in [-14, -6]: dt = softplus(.) between 8e-7 and 2.5e-3, the dt_init_floor regime of mamba_init.dt_init
    (vmamba.py:224-249). log1p(exp(x)) must stay accurate RELATIVE to dt (a `1 + z` formulation loses up to 1 % there);
    out, ddelta and dA scale with dt, so all of them are held to the fp32 bar against the fp64 oracle."""
    inp = make_inputs(2, 16, 1, 2, 1500, torch.float32, seed=11)
    g = torch.Generator(device="cpu").manual_seed(5)
    inp["delta"] = (lo + (hi - lo) * torch.rand(2, 16, 1500, generator=g)).cuda()
    inp["delta_bias"] = torch.zeros(16).cuda()
    inp["A"] = -(1.0 + 50.0 * torch.rand(16, 1, generator=g)).cuda()     # delta * A still moves the state at dt ~ 1e-4
    run_case(inp, True, FP32_TOL)


@pytest.mark.parametrize("shape", ["BEM-I-L0", "BEM-I-L1"])
def test_full_size_backward_bem_inference_levels(shape):
    """BEM-I level 0 / 1 (B1 KD160 L240000, KD320 L60000: the 600x400 workload's own scans): forward AND every gradient
    against the fp64 oracle at full size (round 1 held the level-0 shape forward-only)"""
    cfg = {"BEM-I-L0": (1, 160, 1, 4, 240000), "BEM-I-L1": (1, 320, 1, 4, 60000)}[shape]
    inp = make_inputs(*cfg, torch.float32, seed=4)
    run_case(inp, True, FP32_TOL)


@pytest.mark.parametrize("cfg", [(1, 640, 1, 4, 129600), (1, 384, 16, 4, 129600)], ids=["HD-KD640-N1", "HD-KD384-N16"])
def test_full_size_hd_bf16_forward(cfg):
    """BASELINE configs[3]: the 1920x1080 long-sequence scans (L = 129600 per direction) with bf16 inputs, fp32 'oflex' output,
    against the fp64 oracle evaluated on the same bf16-rounded inputs; bar 1e-2 (north_star), measured ~1e-5"""
    bem = _bem()
    inp = make_inputs(*cfg, torch.bfloat16, seed=6)
    out = bem.selective_scan_fn(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True, True)
    assert out.dtype == torch.float32 and bem._lib.scan_error_word(out.device) == 0
    o = oracle.selective_scan_oracle_f64(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True)
    err = nmax_err(out.cpu().numpy(), o["out"])
    assert err < LOW_TOL, err
    assert err < 1e-4, err      # the inputs are bf16, the arithmetic is fp32: far inside the bf16 bar
    out_b = bem.selective_scan_fn(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True, False)
    assert out_b.dtype == torch.bfloat16 and nmax_err(out_b.float().cpu().numpy(), o["out"]) < LOW_TOL


# ---------------------------------------------------------------------------------------------------------------------
# row-sequential kernels for dstate >= 2 (scan_rows.cu). The dispatcher picks them when B * KD fills the machine; the env
# knob forces them here so that the small shapes exercise every code path (partial chunks, misaligned rows, states per warp).
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [1, 37, 384, 385, 1000, 2049, 4096])
@pytest.mark.parametrize("N,G", [(2, 1), (3, 1), (8, 2), (16, 4), (24, 2)])
def test_row_kernels_fp32(monkeypatch, L, N, G):
    monkeypatch.setenv("BEM_SCAN_ROWS", "1")
    inp = make_inputs(2, 8, N, G, L, torch.float32, seed=L + N)
    run_case(inp, True, FP32_TOL)


@pytest.mark.parametrize("has_D,has_bias,softplus", [(False, False, False), (True, False, True), (False, True, True), (True, True, False)])
def test_row_kernels_option_grid(monkeypatch, has_D, has_bias, softplus):
    monkeypatch.setenv("BEM_SCAN_ROWS", "1")
    inp = make_inputs(2, 12, 16, 2, 777, torch.float32, has_D=has_D, has_bias=has_bias, seed=9)
    run_case(inp, softplus, FP32_TOL)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("L", [64, 513, 2048])
@pytest.mark.parametrize("N", [2, 16])
def test_row_kernels_low_precision(monkeypatch, dtype, L, N):
    monkeypatch.setenv("BEM_SCAN_ROWS", "1")
    inp = make_inputs(2, 8, N, 2, L, dtype, seed=3)
    run_case(inp, True, LOW_TOL)


def test_row_kernels_agree_with_the_look_back_kernels(monkeypatch):
    """same inputs through both organisations (the dispatcher switches between them on the row count): outputs, carries and
    gradients agree within the fp32 tier"""
    bem = _bem()
    inp = make_inputs(2, 16, 16, 4, 3000, torch.float32, seed=21)
    res = {}
    for rows in ("0", "1"):
        monkeypatch.setenv("BEM_SCAN_ROWS", rows)
        out, x = bem.selective_scan_cuda_oflex.fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True, 1, True)
        grads = bem.selective_scan_cuda_oflex.bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"],
                                                  inp["dout"], x, True, 1)
        res[rows] = [out, x] + list(grads)
    for a, b in zip(res["0"], res["1"]):
        assert nmax_err(b.float().cpu().numpy(), a.double().cpu().numpy()) < FP32_TOL   # two fp32 association orders + atomics


@pytest.mark.parametrize("N", [32, 64, 256])
def test_large_dstate_runs_on_the_row_kernels(N):
    """dstate up to the reference's MAX_DSTATE 256 (selective_scan_oflex.cpp:190)"""
    inp = make_inputs(1, 6, N, 1, 700, torch.float32, seed=N)
    run_case(inp, True, FP32_TOL)
