"""Import the staged, UNMODIFIED reference (oracle/_ref/, produced by oracle/make_ref.py).  TEST / BENCH INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs use this module; the product package
never imports it. Nothing here reads /root/reference: the GPU box does not have it.

Shims follow SURVEY.md Appendix B: `timm.models.layers` / `fvcore.nn` stubs (absent from the image; every BEM arch uses
drop_path = 0), a bare `basicsr` namespace package (basicsr/__init__.py star-imports the whole trainer), the staged
basicsr/ directory on sys.path (basicsr/bayesian/tools.py:1 does `import bayesian`), and `selective_scan_ref` taken out of
kernels/selective_scan/test_selective_scan.py:168-234 by AST (the module imports two CUDA extensions at import time).
"""
from __future__ import annotations

import ast
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ROOT = os.path.join(REF_DIR, "reference")
EXT_SO = os.path.join(REF_DIR, "selective_scan_cuda_oflex.so")

_state = {"shims": False}


def available() -> bool:
    return os.path.isdir(os.path.join(ROOT, "basicsr", "vmamba")) and os.path.exists(EXT_SO)


def why_unavailable() -> str:
    return (f"{REF_DIR} is not staged: run `python oracle/make_ref.py` in the build container "
            "(__graft_entry__.build() does it when /root/reference is present)")


def install_shims():
    if _state["shims"]:
        return
    import torch.nn as nn
    tl = types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = nn.init.trunc_normal_

    class DropPath(nn.Module):   # timm 0.4.12 semantics at drop_prob = 0 (all BEM archs)
        def __init__(self, p=0.):
            super().__init__()
            self.drop_prob = p

        def forward(self, x):
            return x
    tl.DropPath = DropPath
    for name, mod in {"timm": types.ModuleType("timm"), "timm.models": types.ModuleType("timm.models"),
                      "timm.models.layers": tl}.items():
        sys.modules.setdefault(name, mod)
    fv = types.ModuleType("fvcore.nn")
    for n in ("FlopCountAnalysis", "flop_count_str", "flop_count", "parameter_count"):
        setattr(fv, n, None)
    sys.modules.setdefault("fvcore", types.ModuleType("fvcore"))
    sys.modules.setdefault("fvcore.nn", fv)
    pkg = types.ModuleType("basicsr")
    pkg.__path__ = [os.path.join(ROOT, "basicsr")]
    sys.modules["basicsr"] = pkg
    if os.path.join(ROOT, "basicsr") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "basicsr"))
    _state["shims"] = True


def _path_load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def selective_scan_ref():
    """the reference's fp32 sequential oracle, kernels/selective_scan/test_selective_scan.py:168-234"""
    import torch
    import torch.nn.functional as F
    from einops import rearrange, repeat
    path = os.path.join(ROOT, "kernels/selective_scan/test_selective_scan.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "selective_scan_ref")
    ns = dict(torch=torch, F=F, rearrange=rearrange, repeat=repeat)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "selective_scan_ref", "exec"), ns)
    return ns["selective_scan_ref"]


def build_selective_scan_fn():
    """the reference's mamba-style autograd wrapper factory, test_selective_scan.py:18-165 (module global SSOFLEX_FLOAT = True)"""
    import torch
    from einops import rearrange
    path = os.path.join(ROOT, "kernels/selective_scan/test_selective_scan.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "build_selective_scan_fn")
    ns = dict(torch=torch, rearrange=rearrange, SSOFLEX_FLOAT=True)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "build_selective_scan_fn", "exec"), ns)
    return ns["build_selective_scan_fn"]


def oflex_ext():
    """the reference's CUDA extension `selective_scan_cuda_oflex` (pybind fwd / bwd, cusoflex/selective_scan_oflex.cpp:360-363),
    compiled unmodified for sm_100a by oracle/make_ref.py. Registered under its own name so csms6s.py:9-10 finds it."""
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    mod = sys.modules.get("_ref_selective_scan_cuda_oflex")
    if mod is None:
        loader = importlib.machinery.ExtensionFileLoader("selective_scan_cuda_oflex", EXT_SO)
        spec = importlib.util.spec_from_file_location("selective_scan_cuda_oflex", EXT_SO, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        sys.modules["_ref_selective_scan_cuda_oflex"] = mod
    return mod


def csms6s(with_ext=True):
    """basicsr/vmamba/models/csms6s.py with the reference extension as its oflex backend (or torch-only)."""
    name = "ref_csms6s" if with_ext else "ref_csms6s_torch"
    if name in sys.modules:
        return sys.modules[name]
    saved = sys.modules.get("selective_scan_cuda_oflex")
    if with_ext:
        sys.modules["selective_scan_cuda_oflex"] = oflex_ext()
    else:
        sys.modules["selective_scan_cuda_oflex"] = None   # import -> ImportError -> torch path
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                mod = _path_load(name, os.path.join(ROOT, "basicsr/vmamba/models/csms6s.py"))
    finally:
        if saved is None:
            sys.modules.pop("selective_scan_cuda_oflex", None)
        else:
            sys.modules["selective_scan_cuda_oflex"] = saved
    return mod


def csm_triton():
    """basicsr/vmamba/models/csm_triton.py: Triton cross_scan_fn / cross_merge_fn (and their torch twins)."""
    if "ref_csm_triton" in sys.modules:
        return sys.modules["ref_csm_triton"]
    return _path_load("ref_csm_triton", os.path.join(ROOT, "basicsr/vmamba/models/csm_triton.py"))


def bayesian():
    """the reference's top-level `bayesian` package (basicsr/bayesian)"""
    install_shims()
    mod = sys.modules.get("_ref_bayesian")
    if mod is None:
        cur = sys.modules.pop("bayesian", None)
        for k in [k for k in sys.modules if k.startswith("bayesian.")]:
            sys.modules.pop(k)
        mod = importlib.import_module("bayesian")
        if not os.path.abspath(mod.__file__).startswith(ROOT):
            raise RuntimeError(f"`bayesian` resolved to {mod.__file__}, not the staged reference")
        sys.modules["_ref_bayesian"] = mod
        if cur is not None and cur is not mod:
            pass   # the reference's package stays registered as `bayesian` until bem_b200.patch.install() replaces it
    return mod


class _ext_registered:
    """while reference modules are being imported, `import selective_scan_cuda_oflex` (csms6s.py:9-10) finds the reference
    extension (or fails, with_ext=False -> the pure-PyTorch scan); the import-time prints of the reference are swallowed"""

    def __init__(self, with_ext):
        self.with_ext = with_ext

    def __enter__(self):
        import contextlib
        import io
        self.saved = sys.modules.get("selective_scan_cuda_oflex")
        sys.modules["selective_scan_cuda_oflex"] = oflex_ext() if self.with_ext else None
        self.stack = contextlib.ExitStack()
        w = warnings.catch_warnings()
        self.stack.enter_context(w)
        warnings.simplefilter("ignore")
        self.stack.enter_context(contextlib.redirect_stdout(io.StringIO()))
        return self

    def __exit__(self, *exc):
        self.stack.close()
        if self.saved is None:
            sys.modules.pop("selective_scan_cuda_oflex", None)
        else:
            sys.modules["selective_scan_cuda_oflex"] = self.saved
        return False


def vmamba(with_ext=True):
    """basicsr/vmamba/models/vmamba.py imported as basicsr.vmamba.models.vmamba, its scan backend = the reference extension
    (with_ext=False: the reference's pure-PyTorch selective_scan_torch, csms6s.py:29-72)"""
    install_shims()
    name = "basicsr.vmamba.models.vmamba"
    if name in sys.modules:
        return sys.modules[name]
    with _ext_registered(with_ext):
        return importlib.import_module(name)


def unet_arch(with_ext=True):
    """basicsr/archs/UNet_arch.py (stage-1 `Network`, build_model). basicsr/archs/__init__.py imports every *_arch.py, some
    of which import a second copy of vmamba as top-level `vmamba.models.vmamba`; both copies see the same extension."""
    vmamba(with_ext)
    with _ext_registered(with_ext):
        return importlib.import_module("basicsr.archs.UNet_arch")


def arch(module: str, with_ext=True):
    """any basicsr/archs/<module>.py, e.g. DecompDualBranchDDWavelet_arch (relative checkpoint paths: chdir to ROOT first)"""
    vmamba(with_ext)
    with _ext_registered(with_ext):
        return importlib.import_module("basicsr.archs." + module)


def niqe_module():
    """basicsr/metrics/niqe.py (calculate_niqe + niqe_pris_params.npz next to it)"""
    install_shims()
    if "basicsr.metrics" not in sys.modules:   # bare namespace: basicsr/metrics/__init__.py imports skimage (absent here)
        pkg = types.ModuleType("basicsr.metrics")
        pkg.__path__ = [os.path.join(ROOT, "basicsr", "metrics")]
        sys.modules["basicsr.metrics"] = pkg
    return importlib.import_module("basicsr.metrics.niqe")


def train_model(with_ext=True, device="cuda"):
    """the reference's configs[4] network exactly as Options/DecompDualBranch2DDWavelet_4.yml:54-68 builds it
    (basicsr/archs/DecompDualBranchDDWavelet_arch.py:147; its frozen decomposition net loads basicsr/QD/checkpoints/model4_999.pth
    through a path relative to the reference root, :67, so the working directory is switched for the construction)"""
    import torch
    mod = arch("DecompDualBranchDDWavelet_arch", with_ext)
    cwd = os.getcwd()
    real_load = torch.load
    os.chdir(ROOT)
    torch.load = lambda f, *a, **k: real_load(f, *a, **{**k, "map_location": device})
    try:
        net = mod.DecompDualBranchDDWavelet(in_channels=6, out_channels=3, n_feat=40, d_state=[1, 1, 1], ssm_ratio=1, mlp_ratio=4,
                                            mlp_type="gdmlp", use_pixelshuffle=True, drop_path=0.0, sam=False, stage=1,
                                            num_blocks=[2, 2, 2], decomp_model="model4")
    finally:
        torch.load = real_load
        os.chdir(cwd)
    return net.to(device)
