"""oracle/ss2d_traversal.py — TEST INFRASTRUCTURE ONLY.

The SS2D core `forward_corev2` (cross2d, vmamba.py:656-684) restated WITHOUT the four materialised traversals — the data flow
a traversal-aware scan kernel would use — next to the reference's op sequence
(`cross_scan -> x_proj -> dt_proj -> selective_scan -> cross_merge`, csm_triton.py:22-85 for the traversals).

The four directions of `cross_scan` are (csm_triton.py:26-29): k0 the row-major flattening of (H, W), k1 the flattening of the
transposed image (W, H), k2 = flip(k0), k3 = flip(k1). Hence:

* a 1x1 projection commutes with a pixel permutation, so x_proj / dt_proj of direction k can be evaluated on the image once
  (row-major for k0 / k2, transposed for k1 / k3) instead of on four permuted copies;
* scanning a flipped sequence forward is scanning the sequence itself BACKWARD (h_t = a_t h_{t+1} + b_t u_t), which leaves the
  result at the un-flipped positions — exactly where `cross_merge` (csm_triton.py:60-62: y0 + flip(y2) + T(y1 + flip(y3)))
  wants it. So only two layouts exist, the image and its transpose; no tensor is ever flipped and the merge is
  `(y0 + y2) + transpose(y1 + y3)` with the reference's association.

`scan_dir` runs the fp64 recurrence of `selective_scan_ref` (test_selective_scan.py:168-234, N states, softplus on delta + bias)
in either direction. `tests/test_oracle_golden.py::test_traversal_aware_ss2d_equals_the_cross_scan_form` holds the two
formulations against each other.
"""
from __future__ import annotations

import numpy as np


def _softplus(v):
    return np.where(v > 20.0, v, np.log1p(np.exp(np.minimum(v, 20.0))))   # F.softplus threshold (test_selective_scan.py:189)


def scan_dir(u, delta, A, Bm, Cm, Dv, bias, reverse):
    """u, delta: (D, L); A: (D, N); Bm, Cm: (N, L); Dv, bias: (D,). y_t = sum_n C[n,t] h[n,t] + D u_t, with
    h[n,t] = exp(delta_t A[n]) h[n,t-1] + delta_t B[n,t] u_t walked forward, or from the end when `reverse`."""
    Dm, L = u.shape
    dl = _softplus(delta + bias[:, None])
    y = np.empty((Dm, L))
    h = np.zeros((Dm, A.shape[1]))
    order = range(L - 1, -1, -1) if reverse else range(L)
    for t in order:
        h = np.exp(dl[:, t, None] * A) * h + (dl[:, t] * u[:, t])[:, None] * Bm[None, :, t]
        y[:, t] = (h * Cm[None, :, t]).sum(axis=1) + Dv * u[:, t]
    return y


def ss2d_cross_scan_form(x, x_proj_w, dt_w, dt_b, A, Dv):
    """the reference's sequence. x: (D, H, W); x_proj_w: (K, R + 2N, D); dt_w: (K, D, R); dt_b, Dv: (K, D); A: (K, D, N)"""
    Dm, H, W = x.shape
    L = H * W
    K, _, R = dt_w.shape
    N = A.shape[2]
    xs = np.stack([x.reshape(Dm, L), x.transpose(0, 2, 1).reshape(Dm, L)])
    xs = np.concatenate([xs, xs[:, :, ::-1]])                                  # (4, D, L)  csm_triton.py:26-29
    ys = np.empty((K, Dm, L))
    for k in range(K):
        x_dbl = x_proj_w[k] @ xs[k]                                            # (R + 2N, L)   vmamba.py:659
        dts = dt_w[k] @ x_dbl[:R]                                              # (D, L)        vmamba.py:661
        ys[k] = scan_dir(xs[k], dts, A[k], x_dbl[R:R + N], x_dbl[R + N:], Dv[k], dt_b[k], reverse=False)
    y = ys[0] + ys[2][:, ::-1]                                                 # csm_triton.py:60-62
    yt = ys[1] + ys[3][:, ::-1]
    return (y.reshape(Dm, H, W) + yt.reshape(Dm, W, H).transpose(0, 2, 1))


def ss2d_traversal_aware(x, x_proj_w, dt_w, dt_b, A, Dv):
    """same function from the image and its transpose only: directions 0 / 2 walk `x` forward / backward, directions 1 / 3
    walk `xT` forward / backward; every result is already at its pixel's place in that layout"""
    Dm, H, W = x.shape
    L = H * W
    K, _, R = dt_w.shape
    N = A.shape[2]
    layouts = (x.reshape(Dm, L), np.ascontiguousarray(x.transpose(0, 2, 1)).reshape(Dm, L))   # the only two copies of x
    ys = []
    for k in range(K):
        src = layouts[k & 1]
        x_dbl = x_proj_w[k] @ src                                              # projected in place, no permuted copy
        dts = dt_w[k] @ x_dbl[:R]
        ys.append(scan_dir(src, dts, A[k], x_dbl[R:R + N], x_dbl[R + N:], Dv[k], dt_b[k], reverse=k >= 2))
    return (ys[0] + ys[2]).reshape(Dm, H, W) + (ys[1] + ys[3]).reshape(Dm, W, H).transpose(0, 2, 1)
