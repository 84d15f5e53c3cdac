"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference algorithms on the hot path. Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
(``bem_b200``) never does and fails loudly when its CUDA library is missing.

Pinned against the real reference (imported in the build container from /root/reference) by
``tests/golden/make_golden.py``; the resulting vectors live in ``tests/golden/*.npz`` and are checked by
``tests/test_oracle_golden.py`` on every CPU run.

Contents
  scan_oracle.c          selective_scan_ref restated in C (fp32 reference order + fp64 fwd/bwd)
  selective_scan_oracle  numpy front-end with the signature of selective_scan_ref
                         (kernels/selective_scan/test_selective_scan.py:168-234)
  cross_scan_oracle / cross_merge_oracle   numpy restatement of csm_triton.py:22-179 (torch fall-backs)
  bayes_*                numpy restatement of basicsr/bayesian/{conv,linear,base_layer}.py
  select_best_oracle     Enhancement/eval.py:270-274
  philox                 the counter-based eps generator the CUDA sample kernel implements
  network                stage-1 Network forward (basicsr/archs/UNet_arch.py:365-474) on CPU for the CPU baseline
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_BUILD, "libscan_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile scan_oracle.c with gcc (-O2 -fopenmp, strict IEEE: no -ffast-math)."""
    src = os.path.join(_HERE, "scan_oracle.c")
    os.makedirs(_BUILD, exist_ok=True)
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", _LIB_PATH, src, "-lm"]
        subprocess.check_call(cmd)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        i = ctypes.c_int
        _lib.oracle_scan_fwd_f32.argtypes = [fp, fp, fp, fp, fp, fp, fp, i, i, i, i, i, i, fp, fp]
        _lib.oracle_scan_fwd_f32.restype = None
        _lib.oracle_scan_f64.argtypes = [fp, fp, fp, fp, fp, fp, fp, i, i, i, i, i, i, dp, dp, fp,
                                         dp, dp, dp, dp, dp, dp, dp]
        _lib.oracle_scan_f64.restype = None
    return _lib


def _f32(a):
    """contiguous fp32 numpy view/copy of a numpy array or torch tensor (bf16/fp16 are widened exactly)."""
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().float().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ct=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def _norm_bc(Bm, batch):
    """(B,N,L) -> (B,1,N,L) like test_selective_scan.py:36-41."""
    Bm = _f32(Bm)
    if Bm.ndim == 3:
        Bm = Bm[:, None]
    assert Bm.ndim == 4 and Bm.shape[0] == batch
    return np.ascontiguousarray(Bm)


def selective_scan_oracle(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                          return_last_state=False):
    """fp32, reference operation order. Same signature/semantics as selective_scan_ref
    (test_selective_scan.py:168-234) for real A and variable B/C; returns fp32 numpy (the caller casts)."""
    lib = _load()
    u_, d_, A_ = _f32(u), _f32(delta), _f32(A)
    Bt, KD, L = u_.shape
    N = A_.shape[1]
    B_, C_ = _norm_bc(B, Bt), _norm_bc(C, Bt)
    G = B_.shape[1]
    assert KD % G == 0 and B_.shape == (Bt, G, N, L) and C_.shape == B_.shape
    D_, db_ = _f32(D), _f32(delta_bias)
    out = np.empty((Bt, KD, L), np.float32)
    last = np.empty((Bt, KD, N), np.float32)
    lib.oracle_scan_fwd_f32(_ptr(u_), _ptr(d_), _ptr(A_), _ptr(B_), _ptr(C_), _ptr(D_), _ptr(db_),
                            int(bool(delta_softplus)), Bt, KD, L, N, G, _ptr(out), _ptr(last))
    if z is not None:  # out * silu(z), test_selective_scan.py:233-234
        z_ = _f32(z)
        out = out * (z_ / (1.0 + np.exp(-z_)))
    return (out, last) if return_last_state else out


def selective_scan_oracle_f64(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, dout=None):
    """fp64 forward (+ backward when dout is given). Returns a dict of float64 numpy arrays:
    out, last_state [, du, ddelta, dA, dB, dC, dD, ddelta_bias]."""
    lib = _load()
    u_, d_, A_ = _f32(u), _f32(delta), _f32(A)
    Bt, KD, L = u_.shape
    N = A_.shape[1]
    squeeze = hasattr(B, "ndim") and B.ndim == 3 or (hasattr(B, "dim") and B.dim() == 3)
    B_, C_ = _norm_bc(B, Bt), _norm_bc(C, Bt)
    G = B_.shape[1]
    D_, db_ = _f32(D), _f32(delta_bias)
    out = np.empty((Bt, KD, L), np.float64)
    last = np.empty((Bt, KD, N), np.float64)
    res = {"out": out, "last_state": last}
    dp = ctypes.c_double
    if dout is None:
        lib.oracle_scan_f64(_ptr(u_), _ptr(d_), _ptr(A_), _ptr(B_), _ptr(C_), _ptr(D_), _ptr(db_),
                            int(bool(delta_softplus)), Bt, KD, L, N, G, _ptr(out, dp), _ptr(last, dp),
                            None, None, None, None, None, None, None, None)
        return res
    g_ = _f32(dout)
    du = np.empty((Bt, KD, L), np.float64)
    dd = np.empty((Bt, KD, L), np.float64)
    dA = np.empty((KD, N), np.float64)
    dB = np.empty((Bt, G, N, L), np.float64)
    dC = np.empty((Bt, G, N, L), np.float64)
    dD = np.empty((KD,), np.float64) if D_ is not None else None
    dbias = np.empty((KD,), np.float64) if db_ is not None else None
    lib.oracle_scan_f64(_ptr(u_), _ptr(d_), _ptr(A_), _ptr(B_), _ptr(C_), _ptr(D_), _ptr(db_),
                        int(bool(delta_softplus)), Bt, KD, L, N, G, _ptr(out, dp), _ptr(last, dp), _ptr(g_),
                        _ptr(du, dp), _ptr(dd, dp), _ptr(dA, dp), _ptr(dB, dp), _ptr(dC, dp),
                        _ptr(dD, dp), _ptr(dbias, dp))
    if squeeze:
        dB, dC = dB[:, 0], dC[:, 0]
    res.update(du=du, ddelta=dd, dA=dA, dB=dB, dC=dC, dD=dD, ddelta_bias=dbias)
    return res


# ---------------------------------------------------------------------------------------------------
# CrossScan / CrossMerge — csm_triton.py:22-179
# ---------------------------------------------------------------------------------------------------
def cross_scan_oracle(x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
    """numpy restatement of cross_scan_fwd / cross_scan1b1_fwd (csm_triton.py:22-53, 88-131).
    Returns (B,4,C,L) or (B,L,4,C)."""
    x = np.asarray(x)
    if one_by_one:
        xs = x if in_channel_first else np.transpose(x, (0, 3, 4, 1, 2))      # (B,4,C,H,W)
    else:
        xi = x if in_channel_first else np.transpose(x, (0, 3, 1, 2))          # (B,C,H,W)
        xs = np.broadcast_to(xi[:, None], (xi.shape[0], 4) + xi.shape[1:])
    Bt, _, Cc, H, W = xs.shape
    L = H * W
    y = np.empty((Bt, 4, Cc, L), x.dtype)
    if scans == 0:
        y[:, 0] = xs[:, 0].reshape(Bt, Cc, L)
        y[:, 1] = np.transpose(xs[:, 1], (0, 1, 3, 2)).reshape(Bt, Cc, L)
        y[:, 2] = xs[:, 2].reshape(Bt, Cc, L)[..., ::-1]
        y[:, 3] = np.transpose(xs[:, 3], (0, 1, 3, 2)).reshape(Bt, Cc, L)[..., ::-1]
    elif scans == 1:
        for k in range(4):
            y[:, k] = xs[:, k].reshape(Bt, Cc, L)
    elif scans == 2:
        y[:, 0] = xs[:, 0].reshape(Bt, Cc, L)
        y[:, 1] = xs[:, 1].reshape(Bt, Cc, L)
        y[:, 2] = xs[:, 2].reshape(Bt, Cc, L)[..., ::-1]
        y[:, 3] = xs[:, 3].reshape(Bt, Cc, L)[..., ::-1]
    else:
        raise ValueError(scans)
    if not out_channel_first:
        y = np.ascontiguousarray(np.transpose(y, (0, 3, 1, 2)))
    return y


def cross_merge_oracle(ys, H, W, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
    """numpy restatement of cross_merge_fwd / cross_merge1b1_fwd (csm_triton.py:56-85, 134-179).
    ys: (B,4,C,L) (out_channel_first) or (B,L,4,C). Returns (B,C,L)|(B,L,C) or, one_by_one, (B,4,C,L)|(B,L,4,C).
    NOTE the reference's naming: for merge, `out_channel_first` describes the SEQUENCE-side input and
    `in_channel_first` the IMAGE-side output (csm_triton.py:56-85)."""
    ys = np.asarray(ys)
    if not out_channel_first:
        ys = np.transpose(ys, (0, 2, 3, 1))           # (B,4,C,L)
    Bt, K, Cc, L = ys.shape
    assert K == 4 and L == H * W
    parts = np.empty((Bt, 4, Cc, L), ys.dtype)
    unT = lambda a: np.transpose(a.reshape(Bt, Cc, W, H), (0, 1, 3, 2)).reshape(Bt, Cc, L)
    if scans == 0:
        parts[:, 0] = ys[:, 0]
        parts[:, 1] = unT(ys[:, 1])
        parts[:, 2] = ys[:, 2][..., ::-1]
        parts[:, 3] = unT(ys[:, 3][..., ::-1])
    elif scans == 1:
        parts[:] = ys
    elif scans == 2:
        parts[:, 0] = ys[:, 0]
        parts[:, 1] = ys[:, 1]
        parts[:, 2] = ys[:, 2][..., ::-1]
        parts[:, 3] = ys[:, 3][..., ::-1]
    else:
        raise ValueError(scans)
    if one_by_one:
        y = parts
        if not in_channel_first:
            y = np.transpose(y, (0, 3, 1, 2))          # (B,L,4,C)
    else:
        # reference association: (y0 + y2) + (y1 + y3)  (csm_triton.py:60-62)
        if scans == 0:
            y = (parts[:, 0] + parts[:, 2]) + (parts[:, 1] + parts[:, 3])
        elif scans == 1:
            y = ys.sum(1)
        else:
            y = ((parts[:, 0] + parts[:, 2]) + (parts[:, 1] + parts[:, 3]))
        if not in_channel_first:
            y = np.transpose(y, (0, 2, 1))             # (B,L,C)
    return np.ascontiguousarray(y)


# ---------------------------------------------------------------------------------------------------
# Bayesian layers — basicsr/bayesian/{conv,linear,base_layer}.py
# ---------------------------------------------------------------------------------------------------
def softplus_rho(rho):
    """sigma = log1p(exp(rho))  (conv.py:106)"""
    rho = np.asarray(rho, np.float32)
    return np.log1p(np.exp(rho)).astype(np.float32)


def bayes_sample(mu, rho, eps):
    """w = mu + sigma * eps  (conv.py:106-107); eps may carry a leading sample axis."""
    return (np.asarray(mu, np.float32) + softplus_rho(rho) * np.asarray(eps, np.float32)).astype(np.float32)


def conv2d_oracle(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
    """Direct numpy conv2d (cross-correlation, zero padding) — what F.conv2d computes (conv.py:114)."""
    x = np.asarray(x, np.float32)
    w = np.asarray(w, np.float32)
    st = (stride, stride) if np.isscalar(stride) else tuple(stride)
    pd = (padding, padding) if np.isscalar(padding) else tuple(padding)
    dl = (dilation, dilation) if np.isscalar(dilation) else tuple(dilation)
    Bt, Cin, H, W = x.shape
    Cout, Cg, KH, KW = w.shape
    assert Cin == Cg * groups and Cout % groups == 0
    xp = np.pad(x, ((0, 0), (0, 0), (pd[0], pd[0]), (pd[1], pd[1])))
    Ho = (H + 2 * pd[0] - dl[0] * (KH - 1) - 1) // st[0] + 1
    Wo = (W + 2 * pd[1] - dl[1] * (KW - 1) - 1) // st[1] + 1
    out = np.zeros((Bt, Cout, Ho, Wo), np.float64)
    og = Cout // groups
    for g in range(groups):
        xg = xp[:, g * Cg:(g + 1) * Cg]
        wg = w[g * og:(g + 1) * og].astype(np.float64)
        for i in range(KH):
            for j in range(KW):
                patch = xg[:, :, i * dl[0]: i * dl[0] + (Ho - 1) * st[0] + 1: st[0],
                           j * dl[1]: j * dl[1] + (Wo - 1) * st[1] + 1: st[1]]
                out[:, g * og:(g + 1) * og] += np.einsum("bchw,oc->bohw", patch.astype(np.float64), wg[:, :, i, j])
    if b is not None:
        out += np.asarray(b, np.float64)[None, :, None, None]
    return out.astype(np.float32)


def bayes_conv2d_oracle(x, mu_w, rho_w, eps_w, mu_b=None, rho_b=None, eps_b=None, stride=1, padding=0,
                        dilation=1, groups=1, deterministic=False):
    """Conv2dReparameterization forward (conv.py:91-128) with explicit eps."""
    if deterministic:
        w, b = np.asarray(mu_w, np.float32), (None if mu_b is None else np.asarray(mu_b, np.float32))
    else:
        w = bayes_sample(mu_w, rho_w, eps_w)
        b = None if mu_b is None else bayes_sample(mu_b, rho_b, eps_b)
    return conv2d_oracle(x, w, b, stride, padding, dilation, groups)


def bayes_linear2d_oracle(x, mu_w, rho_w, eps_w, mu_b=None, rho_b=None, eps_b=None, deterministic=False):
    """Linear2dReparameterization forward (linear.py:67-104): 1x1 conv with weight[:, :, None, None]."""
    mw = np.asarray(mu_w, np.float32)[:, :, None, None]
    if deterministic:
        return conv2d_oracle(x, mw, mu_b)
    w = bayes_sample(mu_w, rho_w, eps_w)[:, :, None, None]
    b = None if mu_b is None else bayes_sample(mu_b, rho_b, eps_b)
    return conv2d_oracle(x, w, b)


def bayes_linear_oracle(x, mu_w, rho_w, eps_w, mu_b=None, rho_b=None, eps_b=None, deterministic=False):
    """LinearReparameterization forward (linear.py:165-203): F.linear on the last axis."""
    x = np.asarray(x, np.float32)
    if deterministic:
        w, b = np.asarray(mu_w, np.float32), mu_b
    else:
        w = bayes_sample(mu_w, rho_w, eps_w)
        b = None if mu_b is None else bayes_sample(mu_b, rho_b, eps_b)
    out = x.astype(np.float64) @ w.astype(np.float64).T
    if b is not None:
        out = out + np.asarray(b, np.float64)
    return out.astype(np.float32)


def kl_div_oracle(mu_q, sigma_q, mu_p, sigma_p):
    """BaseLayer_.kl_div (base_layer.py:26-39)."""
    mu_q, sigma_q, mu_p, sigma_p = (np.asarray(a, np.float64) for a in (mu_q, sigma_q, mu_p, sigma_p))
    kl = np.log(sigma_p) - np.log(sigma_q) + (sigma_q ** 2 + (mu_q - mu_p) ** 2) / (2 * sigma_p ** 2) - 0.5
    return float(kl.mean())


def prior_ema_oracle(prior_mu, prior_rho, mu, rho, decay, step):
    """training-mode prior update (conv.py:92-104): decay' = min(decay, (1+step)/(10+step))."""
    d = min(decay, (1 + step) / (10 + step))
    pm = d * np.asarray(prior_mu, np.float32) + (1 - d) * np.asarray(mu, np.float32)
    pr = d * np.asarray(prior_rho, np.float32) + (1 - d) * np.asarray(rho, np.float32)
    return pm.astype(np.float32), pr.astype(np.float32), softplus_rho(pr)


def select_best_oracle(scores, take_min=False):
    """`lst.index(max(lst))` / `lst.index(min(lst))` exactly as Python evaluates it (Enhancement/eval.py:270-274)."""
    lst = [float(s) for s in scores]
    v = min(lst) if take_min else max(lst)
    # list.index uses `is` before `==`, so a NaN extremum (only possible at position 0) is found at 0
    for i, s in enumerate(lst):
        if s is v or s == v:
            return i
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm2d (basicsr/vmamba/models/vmamba.py:58-63: permute to channels-last, F.layer_norm over C, permute back) and its
# gradients, fp64 numpy. Checker for csrc/ln2d.cu (tests only).
# ---------------------------------------------------------------------------------------------------------------------
def layernorm2d_oracle(x, weight=None, bias=None, eps=1e-5, dy=None):
    """x: (B, C, *spatial). Returns {"y"} and, when dy is given, {"dx", "dweight", "dbias"} — the biased variance and
    rstd = 1/sqrt(var + eps) of torch.nn.functional.layer_norm."""
    x = np.asarray(x, dtype=np.float64)
    C = x.shape[1]
    shp = (1, C) + (1,) * (x.ndim - 2)
    w = np.ones(C) if weight is None else np.asarray(weight, dtype=np.float64)
    b = np.zeros(C) if bias is None else np.asarray(bias, dtype=np.float64)
    mean = x.mean(axis=1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xh = (x - mean) * rstd
    out = {"y": xh * w.reshape(shp) + b.reshape(shp)}
    if dy is not None:
        g = np.asarray(dy, dtype=np.float64)
        gw = g * w.reshape(shp)
        out["dx"] = rstd * (gw - gw.mean(axis=1, keepdims=True) - xh * (gw * xh).mean(axis=1, keepdims=True))
        red = (0,) + tuple(range(2, x.ndim))
        out["dweight"] = (g * xh).sum(axis=red)
        out["dbias"] = g.sum(axis=red)
    return out
