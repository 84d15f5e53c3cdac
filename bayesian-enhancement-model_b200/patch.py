"""Install the sm_100a operators into an imported copy of the reference, so its models run unmodified on them.

    import bem_b200
    bem_b200.patch.install()            # patches whatever reference modules are already in sys.modules
    bem_b200.patch.install(vmamba=mod)  # or pass modules explicitly
    bem_b200.patch.uninstall()          # restores what install() replaced

What gets replaced:
  * `csms6s.selective_scan_fn` and the name imported into vmamba.py (basicsr/vmamba/models/vmamba.py:27-30)
  * `csm_triton.cross_scan_fn` / `cross_merge_fn` and the names imported into vmamba.py (:22-25)
  * `SS2D.forward_corev2` (vmamba.py:547-698) by `ss2d.forward_corev2_patched` (one x_proj launch on the un-scanned x + the
    fused bem_ss2d_fwd call at inference; the reference op sequence on this package's kernels when a gradient is needed)
  * `LayerNorm2d.forward` (vmamba.py:58-63) by the channel-first LayerNorm kernels, forward and backward (layernorm.py;
    `install(layernorm=False)` leaves it alone)
  * the top-level `bayesian` package that basicsr/bayesian/tools.py:1 and the model wrappers import
"""
from __future__ import annotations

import sys

_MISSING = object()
_undo: list = []   # (object, attribute, previous value) of the last install(), newest last


def _set(obj, name, value):
    _undo.append((obj, name, vars(obj).get(name, _MISSING) if isinstance(obj, type) else getattr(obj, name, _MISSING)))
    setattr(obj, name, value)


def uninstall():
    """put back everything install() replaced (module attributes, SS2D.forward_corev2, sys.modules entries)"""
    while _undo:
        obj, name, old = _undo.pop()
        if obj is sys.modules:
            if old is _MISSING:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
        elif old is _MISSING:
            try:
                delattr(obj, name)
            except AttributeError:
                pass
        else:
            setattr(obj, name, old)


def install(vmamba=None, csms6s=None, csm_triton=None, replace_bayesian=True, layernorm=True):
    from . import bayesian as _bayes
    from .layernorm import layernorm2d_forward_patched
    from .csm import cross_merge_fn, cross_scan_fn
    from .selective_scan import SelectiveScanCuda, selective_scan_cuda_oflex, selective_scan_fn
    from .ss2d import forward_corev2_patched

    patched = []
    mods = dict(sys.modules)
    for name, mod in mods.items():
        if mod is None:
            continue
        base = name.rsplit(".", 1)[-1]
        if (csms6s is None and base == "csms6s") or mod is csms6s:
            _set(mod, "selective_scan_fn", selective_scan_fn)
            _set(mod, "SelectiveScanCuda", SelectiveScanCuda)
            _set(mod, "selective_scan_cuda_oflex", selective_scan_cuda_oflex)
            _set(mod, "WITH_SELECTIVESCAN_OFLEX", True)
            patched.append(name)
        if (csm_triton is None and base == "csm_triton") or mod is csm_triton:
            _set(mod, "cross_scan_fn", cross_scan_fn)
            _set(mod, "cross_merge_fn", cross_merge_fn)
            patched.append(name)
        if (vmamba is None and base == "vmamba" and hasattr(mod, "SS2D")) or mod is vmamba:
            _set(mod, "selective_scan_fn", selective_scan_fn)
            _set(mod, "cross_scan_fn", cross_scan_fn)
            _set(mod, "cross_merge_fn", cross_merge_fn)
            for cls_name in ("SS2Dv2", "SS2D"):
                cls = getattr(mod, cls_name, None)
                if cls is not None and "forward_corev2" in vars(cls):
                    _set(cls, "forward_corev2", forward_corev2_patched)
            ln = getattr(mod, "LayerNorm2d", None)
            if layernorm and ln is not None and "forward" in vars(ln):
                _set(ln, "forward", layernorm2d_forward_patched)
            patched.append(name)

    def set_module(name, value):
        _undo.append((sys.modules, name, sys.modules.get(name, _MISSING)))
        sys.modules[name] = value
    if replace_bayesian:
        set_module("bayesian", _bayes)
        patched.append("bayesian")
    if "selective_scan_cuda_oflex" not in sys.modules:
        set_module("selective_scan_cuda_oflex", selective_scan_cuda_oflex)
    return patched
