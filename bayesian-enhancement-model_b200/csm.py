"""CrossScan / CrossMerge — drop-in for ``cross_scan_fn`` / ``cross_merge_fn``
(basicsr/vmamba/models/csm_triton.py:491-505) and their autograd pairs (:182-273, 393-487).

Same arguments, defaults, shapes and autograd pairing (d cross_scan = cross_merge and vice versa); the work is done by
bem_cross_scan / bem_cross_merge (csrc/csm.cu). ``force_torch`` is accepted for signature compatibility: the reference
uses it to pick its PyTorch implementation, which computes the same thing bit for bit.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import lib


def _launch(fn, what, src, dst, B, Cc, H, W, img_cf, seq_cf, one_by_one, scans):
    if scans not in (0, 1, 2):
        raise RuntimeError(f"cross scan/merge: scans must be 0 (cross2d), 1 (unidi) or 2 (bidi), got {scans}")
    p = _lib.BemCsmParams(B=B, C=Cc, H=H, W=W, dtype=_lib.dtype_code(src.dtype), img_channel_first=int(img_cf),
                          seq_channel_first=int(seq_cf), one_by_one=int(one_by_one), scans=int(scans),
                          src=_lib.ptr(src), dst=_lib.ptr(dst))
    _lib.launch(what, fn, p, src.device, key=(B, Cc, H, W), nbytes=src.numel() * src.element_size() + dst.numel() * dst.element_size())


def _img_dims(x, channel_first, one_by_one):
    if one_by_one:
        if channel_first:
            B, K, Cc, H, W = x.shape
        else:
            B, H, W, K, Cc = x.shape
        if K != 4:
            raise RuntimeError("one_by_one cross scan/merge expects 4 directions")
    else:
        if channel_first:
            B, Cc, H, W = x.shape
        else:
            B, H, W, Cc = x.shape
    return int(B), int(Cc), int(H), int(W)


def _scan(x, in_cf, out_cf, one_by_one, scans, dims):
    """image-side x -> sequences y: (B,4,C,L) | (B,L,4,C)"""
    _lib.require_cuda(x)
    B, Cc, H, W = dims
    x = x.contiguous()
    y = x.new_empty((B, 4, Cc, H * W)) if out_cf else x.new_empty((B, H * W, 4, Cc))
    _launch(lib.bem_cross_scan, "cross_scan", x, y, B, Cc, H, W, in_cf, out_cf, one_by_one, scans)
    return y


def _merge(y, in_cf, out_cf, one_by_one, scans, dims):
    """sequences y ((B,4,C,H,W)-like | (B,H,W,4,C)-like, any view of that memory) -> image-side x"""
    _lib.require_cuda(y)
    B, Cc, H, W = dims
    y = y.contiguous()
    if one_by_one:
        x = y.new_empty((B, 4, Cc, H * W)) if in_cf else y.new_empty((B, H * W, 4, Cc))
    else:
        x = y.new_empty((B, Cc, H * W)) if in_cf else y.new_empty((B, H * W, Cc))
    _launch(lib.bem_cross_merge, "cross_merge", y, x, B, Cc, H, W, in_cf, out_cf, one_by_one, scans)
    return x


class CrossScanF(torch.autograd.Function):
    """csm_triton.CrossScanTritonF (csm_triton.py:393-443)"""

    @staticmethod
    def forward(ctx, x, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
        dims = _img_dims(x, in_channel_first, one_by_one)
        ctx.cfg = (in_channel_first, out_channel_first, one_by_one, scans, dims)
        return _scan(x, in_channel_first, out_channel_first, one_by_one, scans, dims)

    @staticmethod
    def backward(ctx, y):
        in_cf, out_cf, obo, scans, dims = ctx.cfg
        B, Cc, H, W = dims
        x = _merge(y, in_cf, out_cf, obo, scans, dims)
        if obo:
            x = x.view(B, 4, Cc, H, W) if in_cf else x.view(B, H, W, 4, Cc)
        else:
            x = x.view(B, Cc, H, W) if in_cf else x.view(B, H, W, Cc)
        return x, None, None, None, None


class CrossMergeF(torch.autograd.Function):
    """csm_triton.CrossMergeTritonF (csm_triton.py:446-487)"""

    @staticmethod
    def forward(ctx, y, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0):
        if out_channel_first:
            B, K, Cc, H, W = y.shape
        else:
            B, H, W, K, Cc = y.shape
        dims = (int(B), int(Cc), int(H), int(W))
        ctx.cfg = (in_channel_first, out_channel_first, one_by_one, scans, dims)
        return _merge(y, in_channel_first, out_channel_first, one_by_one, scans, dims)

    @staticmethod
    def backward(ctx, x):
        in_cf, out_cf, obo, scans, dims = ctx.cfg
        B, Cc, H, W = dims
        y = _scan(x, in_cf, out_cf, obo, scans, dims)
        y = y.view(B, 4, Cc, H, W) if out_cf else y.view(B, H, W, 4, Cc)
        return y, None, None, None, None


def cross_scan_fn(x: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0,
                  force_torch=False):
    """x: (B, C, H, W) | (B, H, W, C) | (B, 4, C, H, W) | (B, H, W, 4, C) -> y: (B, 4, C, L) | (B, L, 4, C).
    scans: 0 cross scan, 1 unidirectional, 2 bidirectional (csm_triton.py:491-496)."""
    return CrossScanF.apply(x, in_channel_first, out_channel_first, one_by_one, scans)


def cross_merge_fn(y: torch.Tensor, in_channel_first=True, out_channel_first=True, one_by_one=False, scans=0,
                   force_torch=False):
    """y: (B, 4, C, H, W) | (B, H, W, 4, C) -> x: (B, C, L) | (B, L, C) | (B, 4, C, L) | (B, L, 4, C)
    (csm_triton.py:500-505)."""
    return CrossMergeF.apply(y, in_channel_first, out_channel_first, one_by_one, scans)
