"""oracle/gdmlp_tiled.py — TEST INFRASTRUCTURE ONLY.

Two CPU restatements of the feed-forward half of a VSSBlock, `x + gdMlp(norm2(x))` (vmamba.py:1331-1333 with
`gdMlp.forward`, vmamba.py:128-133: project_in -> depthwise 3x3 -> chunk -> gelu(x1) * x2 -> project_out):

* `gdmlp_plain`  : the reference's op sequence, whole image at a time (float64 numpy);
* `gdmlp_tiled`  : the SAME arithmetic in the order a fused single-pass kernel would run it:
  spatial tiles with a one-pixel halo, the hidden channels walked in groups of gate pairs, project_out accumulated over
  the groups (a K split), skip connection added last. It pins down the two things such a kernel must get right and that
  the three separate kernels of today get for free:
    - the zero padding of the depthwise conv applies to project_in's OUTPUT: a halo pixel outside the image contributes 0,
      not project_in(LayerNorm(0)) (which is bias + beta . W);
    - gate pair g is hidden channel g (GELU side) and hidden channel g + hidden (linear side) (`chunk(2, dim=1)`).

`tests/test_oracle_golden.py::test_tiled_gdmlp_schedule_equals_the_plain_one` holds the two against each other on ragged
shapes; `gdmlp_plain` itself is held against torch's ops there.
"""
from __future__ import annotations

import math

import numpy as np

_erf = np.vectorize(math.erf, otypes=[np.float64])


def _gelu(v):
    return 0.5 * v * (1.0 + _erf(v / math.sqrt(2.0)))       # exact GELU (act_layer=nn.GELU, vmamba.py:117,126)


def _ln_pixels(x, gamma, beta, eps):
    """LayerNorm over the channel axis of every pixel (LayerNorm2d, vmamba.py:58-63); x: (C, ...)"""
    mean = x.mean(axis=0, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=0, keepdims=True)
    shape = (-1,) + (1,) * (x.ndim - 1)
    return (x - mean) / np.sqrt(var + eps) * gamma.reshape(shape) + beta.reshape(shape)


def gdmlp_plain(x, gamma, beta, eps, w1, b1, wd, bd, w2, b2):
    """x: (C, H, W); w1: (2*hidden, C); wd: (2*hidden, 3, 3); w2: (Cout, hidden). Returns x + gdMlp(LN(x)) (Cout == C)."""
    x = np.asarray(x, dtype=np.float64)
    C, H, W = x.shape
    hidden = w2.shape[1]
    t = np.einsum("oc,chw->ohw", w1, _ln_pixels(x, gamma, beta, eps)) + b1[:, None, None]
    tp = np.zeros((2 * hidden, H + 2, W + 2))
    tp[:, 1:-1, 1:-1] = t                                     # zero padding of the conv's input
    d = np.zeros_like(t)
    for kh in range(3):
        for kw in range(3):
            d += wd[:, kh, kw][:, None, None] * tp[:, kh:kh + H, kw:kw + W]
    d += bd[:, None, None]
    gated = _gelu(d[:hidden]) * d[hidden:]
    return x + np.einsum("oc,chw->ohw", w2, gated) + b2[:, None, None]


def gdmlp_tiled(x, gamma, beta, eps, w1, b1, wd, bd, w2, b2, tile_h=4, tile_w=128, pairs=16):
    """Same result, computed tile by tile and gate-pair group by group; returns (out, stats) with the recompute factor of
    project_in (halo pixels are evaluated by every tile that needs them) and the peak intermediate size per tile."""
    x = np.asarray(x, dtype=np.float64)
    C, H, W = x.shape
    hidden = w2.shape[1]
    out = np.empty((w2.shape[0], H, W))
    fc1_pixels = 0
    peak = 0
    for h0 in range(0, H, tile_h):
        th = min(tile_h, H - h0)
        for w0 in range(0, W, tile_w):
            tw = min(tile_w, W - w0)
            # halo window clipped to the image; the part outside stays zero in `t` below
            ha, hb = max(h0 - 1, 0), min(h0 + th + 1, H)
            wa, wb = max(w0 - 1, 0), min(w0 + tw + 1, W)
            xn = _ln_pixels(x[:, ha:hb, wa:wb], gamma, beta, eps)          # LayerNorm of the tile + halo, once per tile
            fc1_pixels += (hb - ha) * (wb - wa)
            acc = np.zeros((w2.shape[0], th, tw))                           # project_out accumulators of the tile
            for g0 in range(0, hidden, pairs):
                g1 = min(g0 + pairs, hidden)
                ch = np.r_[g0:g1, hidden + g0:hidden + g1]                  # GELU side | linear side of the gate pairs
                t = np.zeros((len(ch), th + 2, tw + 2))
                t[:, ha - (h0 - 1):hb - (h0 - 1), wa - (w0 - 1):wb - (w0 - 1)] = (
                    np.einsum("oc,chw->ohw", w1[ch], xn) + b1[ch][:, None, None])
                peak = max(peak, t.size)
                d = np.zeros((len(ch), th, tw))
                for kh in range(3):
                    for kw in range(3):
                        d += wd[ch][:, kh, kw][:, None, None] * t[:, kh:kh + th, kw:kw + tw]
                d += bd[ch][:, None, None]
                n = g1 - g0
                gated = _gelu(d[:n]) * d[n:]
                acc += np.einsum("oc,chw->ohw", w2[:, g0:g1], gated)        # K split of project_out over the groups
            out[:, h0:h0 + th, w0:w0 + tw] = x[:, h0:h0 + th, w0:w0 + tw] + acc + b2[:, None, None]
    return out, {"fc1_recompute": fc1_pixels / float(H * W), "peak_intermediate_elems": peak}
