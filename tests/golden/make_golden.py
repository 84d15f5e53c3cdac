#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference at /root/reference (build container only).

The reference is Python, so it is imported here (CPU) and its outputs on small seeded inputs are frozen as
fixtures; /root/reference does not exist on the GPU box, so tests only ever read the .npz files.
Run:  python tests/golden/make_golden.py
Shims follow SURVEY.md Appendix B (timm / fvcore stubs, bare `basicsr` namespace, AST-extracted selective_scan_ref).
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange, repeat

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_selective_scan_ref():
    tree = ast.parse(open(f"{REF}/kernels/selective_scan/test_selective_scan.py").read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "selective_scan_ref")
    ns = dict(torch=torch, F=F, rearrange=rearrange, repeat=repeat)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "selective_scan_ref", "exec"), ns)
    return ns["selective_scan_ref"]


def path_load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def install_shims():
    tl = types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = nn.init.trunc_normal_

    class DropPath(nn.Module):
        def __init__(self, p=0.):
            super().__init__()
            self.drop_prob = p

        def forward(self, x):
            return x
    tl.DropPath = DropPath
    for name, mod in {"timm": types.ModuleType("timm"), "timm.models": types.ModuleType("timm.models"),
                      "timm.models.layers": tl}.items():
        sys.modules[name] = mod
    fv = types.ModuleType("fvcore.nn")
    for n in ("FlopCountAnalysis", "flop_count_str", "flop_count", "parameter_count"):
        setattr(fv, n, None)
    sys.modules["fvcore"] = types.ModuleType("fvcore")
    sys.modules["fvcore.nn"] = fv
    pkg = types.ModuleType("basicsr")
    pkg.__path__ = [f"{REF}/basicsr"]
    sys.modules["basicsr"] = pkg
    sys.path.insert(0, f"{REF}/basicsr")


def npy(t):
    return None if t is None else t.detach().float().cpu().numpy().copy()   # copy: later in-place edits must not leak in


# ---------------------------------------------------------------------------------------------------
def gen_scan(selective_scan_ref, csms6s):
    cases = [
        # name, B, KD, N, G(0 = 3-D B/C), L, dtype, has_D, has_bias, softplus, has_z
        ("n1_l70", 2, 8, 1, 0, 70, torch.float32, True, True, True, False),
        ("n4_g2_l300", 1, 8, 4, 2, 300, torch.float32, False, False, False, False),
        ("n16_g4_l600", 1, 8, 16, 4, 600, torch.float32, True, True, True, False),
        ("n2_z_l129", 1, 4, 2, 1, 129, torch.float32, True, False, True, True),
        ("n1_g4_l1000", 1, 8, 1, 4, 1000, torch.float32, True, True, True, False),
        ("bf16_n2_l128", 1, 4, 2, 1, 128, torch.bfloat16, True, True, True, False),
        ("f16_n1_l96", 2, 4, 1, 2, 96, torch.float16, True, True, True, False),
    ]
    out = {}
    for (name, Bt, KD, N, G, L, dt, has_D, has_bias, sp, has_z) in cases:
        torch.manual_seed(0)   # distributions follow test_selective_scan.py:406-441
        A = (-0.5 * torch.rand(KD, N)).requires_grad_()
        bshape = (Bt, N, L) if G == 0 else (Bt, G, N, L)
        Bm = torch.randn(*bshape, dtype=dt, requires_grad=True)
        Cm = torch.randn(*bshape, dtype=dt, requires_grad=True)
        D = torch.randn(KD, requires_grad=True) if has_D else None
        z = torch.randn(Bt, KD, L, dtype=dt, requires_grad=True) if has_z else None
        bias = (0.5 * torch.rand(KD)).requires_grad_() if has_bias else None
        u = torch.randn(Bt, KD, L, dtype=dt, requires_grad=True)
        delta = (0.5 * torch.rand(Bt, KD, L, dtype=dt)).requires_grad_()
        o, last = selective_scan_ref(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=sp,
                                     return_last_state=True)
        g = torch.randn_like(o)
        o.backward(g)
        rec = dict(u=npy(u), delta=npy(delta), A=npy(A), B=npy(Bm), C=npy(Cm), D=npy(D), z=npy(z), delta_bias=npy(bias),
                   softplus=np.array(sp), out=npy(o), last_state=npy(last), dout=npy(g),
                   du=npy(u.grad), ddelta=npy(delta.grad), dA=npy(A.grad), dB=npy(Bm.grad), dC=npy(Cm.grad),
                   dD=npy(D.grad) if has_D else None, ddelta_bias=npy(bias.grad) if has_bias else None,
                   dz=npy(z.grad) if has_z else None,
                   dtype=np.array({torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dt]))
        for k, v in rec.items():
            if v is not None:
                out[f"{name}/{k}"] = v
    # product API (csms6s.selective_scan_fn, torch backend, oflex float output) on a 4-D B/C case
    torch.manual_seed(1)
    Bt, K, Dc, N, L = 1, 4, 2, 2, 50
    u = torch.randn(Bt, K * Dc, L)
    delta = 0.5 * torch.rand(Bt, K * Dc, L)
    A = -0.5 * torch.rand(K * Dc, N)
    Bm = torch.randn(Bt, K, N, L)
    Cm = torch.randn(Bt, K, N, L)
    D = torch.randn(K * Dc)
    bias = 0.5 * torch.rand(K * Dc)
    o = csms6s.selective_scan_fn(u, delta, A, Bm, Cm, D, bias, True, True, backend="torch")
    for k, v in dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=D, delta_bias=bias, out=o).items():
        out[f"csms6s/{k}"] = npy(v)
    np.savez_compressed(f"{OUT}/scan.npz", **out)
    print("scan.npz", len(out), "arrays")


def gen_csm(csm):
    out = {}
    torch.manual_seed(2)
    Bt, Cc, H, W = 2, 3, 5, 7
    x = torch.randn(Bt, Cc, H, W)
    x4 = torch.randn(Bt, 4, Cc, H, W)
    out["x"] = npy(x)
    out["x4"] = npy(x4)
    for scans in (0, 1, 2):
        for icf in (True, False):
            for ocf in (True, False):
                for obo in (False, True):
                    tag = f"s{scans}_i{int(icf)}_o{int(ocf)}_b{int(obo)}"
                    src = x4 if obo else x
                    if not icf:
                        src = src.permute(0, 3, 4, 1, 2).contiguous() if obo else src.permute(0, 2, 3, 1).contiguous()
                    try:
                        y = csm.cross_scan_fn(src, icf, ocf, obo, scans, force_torch=True)
                        out[f"scan/{tag}"] = npy(y)
                    except Exception as e:  # some layout combinations are broken in the reference's torch path
                        print("cross_scan", tag, "reference raised", type(e).__name__)
                    # merge: sequence-side input laid out per out_channel_first
                    ys = torch.randn(Bt, 4, Cc, H, W) if ocf else torch.randn(Bt, H, W, 4, Cc)
                    try:
                        m = csm.cross_merge_fn(ys, icf, ocf, obo, scans, force_torch=True)
                        out[f"merge_in/{tag}"] = npy(ys)
                        out[f"merge/{tag}"] = npy(m)
                    except Exception as e:
                        print("cross_merge", tag, "reference raised", type(e).__name__)
    # autograd pairing: d(cross_scan)/dx == cross_merge  (csm_triton.py:207-225)
    xg = x.clone().requires_grad_()
    y = csm.cross_scan_fn(xg, True, True, False, 0, force_torch=True)
    gy = torch.randn_like(y)
    y.backward(gy)
    out["scan_bwd/gy"] = npy(gy)
    out["scan_bwd/gx"] = npy(xg.grad)
    np.savez_compressed(f"{OUT}/csm.npz", **out)
    print("csm.npz", len(out), "arrays")


def gen_bayes(bayesian):
    out = {}

    def dump(tag, layer, x, y):
        out[f"{tag}/x"] = npy(x)
        out[f"{tag}/out"] = npy(y)
        for n in ("mu_weight", "rho_weight", "eps_weight", "mu_bias", "rho_bias", "eps_bias",
                  "prior_mu_weight", "prior_rho_weight", "prior_mu_bias", "prior_rho_bias"):
            if hasattr(layer, n) and getattr(layer, n) is not None and torch.is_tensor(getattr(layer, n)):
                out[f"{tag}/{n}"] = npy(getattr(layer, n))

    torch.manual_seed(3)
    # grouped 3x3 with bias, eval-mode stochastic forward (conv.py:106-114)
    conv = bayesian.Conv2dReparameterization(4, 6, 3, stride=1, padding=1, groups=2, bias=True, sigma_init=0.05)
    conv.rho_weight.data.uniform_(-4, -1)
    conv.rho_bias.data.uniform_(-4, -1)
    conv.mu_bias.data.normal_()
    conv.eval()
    x = torch.randn(2, 4, 6, 5)
    dump("conv3g", conv, x, conv(x))
    conv.deterministic = True
    out["conv3g/out_det"] = npy(conv(x))
    # depthwise 3x3 no bias (SS2D.conv2d after conversion), strided dilated general conv
    dw = bayesian.Conv2dReparameterization(5, 5, 3, padding=1, groups=5, bias=False)
    dw.eval()
    x = torch.randn(1, 5, 7, 6)
    dump("dw3", dw, x, dw(x))
    gen = bayesian.Conv2dReparameterization(3, 4, (3, 2), stride=2, padding=1, dilation=1, groups=1, bias=True)
    gen.eval()
    x = torch.randn(1, 3, 9, 8)
    dump("conv_s2", gen, x, gen(x))
    # 1x1 conv with bias (gdMlp.project_in)
    pw = bayesian.Conv2dReparameterization(5, 7, 1, bias=True)
    pw.mu_bias.data.normal_()
    pw.eval()
    x = torch.randn(2, 5, 4, 3)
    dump("pw1", pw, x, pw(x))
    # Linear2d (in_proj / out_proj), no bias
    l2 = bayesian.Linear2dReparameterization(6, 6, bias=False)
    l2.eval()
    x = torch.randn(2, 6, 3, 4)
    dump("lin2d", l2, x, l2(x))
    # Linear with bias
    ln = bayesian.LinearReparameterization(5, 3, bias=True)
    ln.mu_bias.data.normal_()
    ln.eval()
    x = torch.randn(4, 2, 5)
    dump("lin", ln, x, ln(x))
    # training mode: prior EMA + step + kl (conv.py:85-104)
    tr = bayesian.Conv2dReparameterization(3, 3, 1, bias=True, sigma_init=0.05, decay=0.998)
    tr.train()
    out["train/mu0"] = npy(tr.mu_weight)
    out["train/prior_mu0"] = npy(tr.prior_mu_weight)
    x = torch.randn(1, 3, 4, 4)
    for it in range(3):
        with torch.no_grad():
            tr.mu_weight.add_(0.1 * torch.randn_like(tr.mu_weight))
            tr.rho_weight.add_(0.1 * torch.randn_like(tr.rho_weight))
            tr.mu_bias.add_(0.1 * torch.randn_like(tr.mu_bias))
        out[f"train/mu_w{it}"] = npy(tr.mu_weight)
        out[f"train/rho_w{it}"] = npy(tr.rho_weight)
        out[f"train/mu_b{it}"] = npy(tr.mu_bias)
        out[f"train/rho_b{it}"] = npy(tr.rho_bias)
        y = tr(x)
        out[f"train/prior_mu_w{it}"] = npy(tr.prior_mu_weight)
        out[f"train/prior_rho_w{it}"] = npy(tr.prior_rho_weight)
        out[f"train/prior_sigma_w{it}"] = npy(tr.prior_sigma_weight)
        out[f"train/prior_mu_b{it}"] = npy(tr.prior_mu_bias)
        out[f"train/kl{it}"] = npy(tr.kl_loss())
        out[f"train/step{it}"] = np.array(tr.step)
    # state_dict keys (checkpoint contract)
    out["conv3g/state_keys"] = np.array(sorted(conv.state_dict().keys()))
    out["rho_init"] = npy(bayesian.Conv2dReparameterization(1, 1, 1, sigma_init=0.05).rho_weight)
    np.savez_compressed(f"{OUT}/bayes.npz", **out)
    print("bayes.npz", len(out), "arrays")


def gen_models(bayesian):
    """SS2D core, VSSBlock and a small stage-1 Network, deterministic and Bayesian with captured eps."""
    from basicsr.vmamba.models import vmamba
    from basicsr.archs.UNet_arch import Network
    out = {}
    torch.manual_seed(4)
    # ---- SS2D core (forward_corev2 cross2d, vmamba.py:656-698) via a VSSBlock's op ----
    for tag, dim, ds, H, W in (("ss2d_n1", 8, 1, 6, 5), ("ss2d_n4", 16, 4, 4, 7)):
        blk = vmamba.VSSBlock(hidden_dim=dim, drop_path=0, norm_layer=vmamba.LayerNorm2d, channel_first=True,
                              ssm_d_state=ds, ssm_ratio=1, ssm_dt_rank="auto", ssm_act_layer=nn.SiLU, ssm_conv=3,
                              ssm_conv_bias=False, ssm_drop_rate=0, ssm_init="v0", forward_type="v05_noz",
                              mlp_ratio=4, mlp_act_layer=nn.GELU, mlp_drop_rate=0.0, mlp_type="gdmlp")
        blk.eval()
        with torch.no_grad():   # de-trivialise the init (Ds = 1, LN weight = 1 ...)
            for p in blk.parameters():
                p.add_(0.05 * torch.randn_like(p))
        x = torch.randn(2, dim, H, W)
        with torch.no_grad():
            core = blk.op.forward_core(x)
            ss2d = blk.op(x)
            full = blk(x)
        out[f"{tag}/x"] = npy(x)
        out[f"{tag}/core"] = npy(core)
        out[f"{tag}/op"] = npy(ss2d)
        out[f"{tag}/block"] = npy(full)
        for k, v in blk.state_dict().items():
            out[f"{tag}/sd/{k}"] = npy(v)
    # ---- stage-1 Network (UNet_arch.py:365-474), small width ----
    torch.manual_seed(5)
    net = Network(stage=1, n_feat=8, num_blocks=[1, 1, 1], d_state=[1, 1, 1], ssm_ratio=1, mlp_ratio=4,
                  mlp_type="gdmlp", use_pixelshuffle=True)
    net.eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.02 * torch.randn_like(p))
    x = torch.rand(1, 3, 16, 24)
    with torch.no_grad():
        y = net(x)[-1]
    out["net/x"] = npy(x)
    out["net/out_det_plain"] = npy(y)
    for k, v in net.state_dict().items():
        out[f"net/sd_plain/{k}"] = npy(v)
    # Bayesian conversion exactly as ConditionGenerator does it (condition_generator_model.py:51-59)
    bayesian.convert2bnn_selective(net, {"sigma_init": 0.05, "decay": 0.998, "pretrain": True})
    net.eval()
    names = [n for n, m in net.named_modules() if hasattr(m, "deterministic")]
    out["net/bnn_layers"] = np.array(names)
    out["net/bnn_types"] = np.array([type(dict(net.named_modules())[n]).__name__ for n in names])
    bayesian.set_prediction_type(net, deterministic=True)
    with torch.no_grad():
        out["net/out_det_bnn"] = npy(net(x)[-1])
    bayesian.set_prediction_type(net, deterministic=False)
    torch.manual_seed(6)
    with torch.no_grad():
        y = net(x)[-1]
    out["net/out_mc"] = npy(y)
    mods = dict(net.named_modules())
    for n in names:
        out[f"net/eps/{n}.eps_weight"] = npy(mods[n].eps_weight)
        if mods[n].bias:
            out[f"net/eps/{n}.eps_bias"] = npy(mods[n].eps_bias)
    for k, v in net.state_dict().items():
        out[f"net/sd_bnn/{k}"] = npy(v)
    np.savez_compressed(f"{OUT}/models.npz", **out)
    print("models.npz", len(out), "arrays")


def gen_select():
    """`lst.index(max(lst))` (Enhancement/eval.py:270-274) on tie / NaN / single-element cases."""
    rng = np.random.RandomState(7)
    cases = [
        [0.3, 0.9, 0.9, 0.1], [1.0], [0.5, 0.5, 0.5], [-1.0, -2.0, -0.5, -0.5],
        [float("nan"), 1.0, 2.0], [1.0, float("nan"), 2.0], [2.0, float("nan"), 1.0],
        [float("inf"), 1.0, float("inf")], [float("-inf"), float("-inf")],
        list(rng.rand(100).astype(np.float32)), list(np.round(rng.rand(200) * 5).astype(np.float32)),
    ]
    out = {}
    for i, c in enumerate(cases):
        c = [float(np.float32(v)) for v in c]
        out[f"c{i}/scores"] = np.array(c, np.float32)
        out[f"c{i}/argmax"] = np.array(c.index(max(c)))
        out[f"c{i}/argmin"] = np.array(c.index(min(c)))
    np.savez_compressed(f"{OUT}/select.npz", **out)
    print("select.npz", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    os.chdir(REF)
    install_shims()
    import bayesian  # noqa: E402  (tools.py:1 assumes a top-level `bayesian` package)
    selective_scan_ref = load_selective_scan_ref()
    csms6s = path_load("ref_csms6s", f"{REF}/basicsr/vmamba/models/csms6s.py")
    csm = path_load("ref_csm_triton", f"{REF}/basicsr/vmamba/models/csm_triton.py")
    gen_scan(selective_scan_ref, csms6s)
    gen_csm(csm)
    gen_bayes(bayesian)
    gen_select()
    gen_models(bayesian)
