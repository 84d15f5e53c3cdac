"""Device-resident no-reference scorer: NIQE ("Making a 'Completely Blind' Image Quality Analyzer"), batched over the
Monte-Carlo predictions of one image, so that best-sample selection (Enhancement/eval.py:249-250, 272-275: `one_niqe_list.append(
calculate_niqe(pred*255, crop_border=0))`, then `index(min(...))`) never leaves the GPU.

Follows basicsr/metrics/niqe.py step by step (the numbers in brackets are its lines):
  * [183-193] the (H, W, 3) prediction scaled to [0, 255] goes through `to_y_channel`, i.e. the BT.601 luma of
    basicsr/utils/matlab_functions.bgr2ycbcr with the channels taken in the order they come (eval.py hands over RGB where the
    function expects BGR — reproduced, not corrected), then `round()`;
  * [97-101] crop to whole 96 x 96 blocks; [104-108] local mean / deviation with the 7 x 7 Gaussian window of the parameter file
    (`scipy.ndimage.convolve`, mode 'nearest') and the normalised image (MSCN coefficients);
  * [110-116, 41-60, 13-38] 18 features per block: AGGD fits of the block and of its products with four circular shifts
    (moment matching against a table of 9801 gamma values);
  * [120-122] second scale: MATLAB-style antialiased bicubic `imresize(img / 255, 0.5) * 255` (matlab_functions.py:16-178),
    applied here as two matrix products with the weight matrices of calculate_weights_indices (symmetric padding folded in);
  * [126-139] multivariate-Gaussian fit of the 36-d block features (nanmean, covariance of the NaN-free rows), distance to the
    pristine model: sqrt(d^T pinv((cov_p + cov_d) / 2) d).
All arithmetic after the luma conversion is float64 on the device (the reference is float64 numpy), batched over images;
nothing synchronises with the host. The per-pixel work (local statistics, normalisation, the block moments of the five maps)
runs in two hand-written kernels (csrc/niqe.cu); what remains per image is a few kilobytes of moment algebra in torch. The pristine parameters are the reference's own file (basicsr/metrics/niqe_pris_params.npz),
read from a path the caller gives — they are not part of this repository.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch


def _cubic(x):
    ax = np.abs(x)
    ax2, ax3 = ax ** 2, ax ** 3
    return (1.5 * ax3 - 2.5 * ax2 + 1) * (ax <= 1) + (-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2) * ((ax > 1) & (ax <= 2))


def resize_matrix(in_len: int, scale: float) -> np.ndarray:
    """(out_len, in_len) float32 matrix of MATLAB's antialiased bicubic resize along one axis:
    calculate_weights_indices (matlab_functions.py:16-86, evaluated in float32 like the reference's torch code) with the
    symmetric border copies of imresize (:125-139) folded into the columns."""
    out_len = math.ceil(in_len * scale)
    kw = 4.0
    if scale < 1:
        kw = kw / scale
    x = np.linspace(1, out_len, out_len, dtype=np.float32)
    u = (x / np.float32(scale) + np.float32(0.5 * (1 - 1 / scale))).astype(np.float32)
    left = np.floor(u - np.float32(kw / 2))
    p = math.ceil(kw) + 2
    idx = left[:, None] + np.arange(p, dtype=np.float32)[None, :]
    dist = (u[:, None] - idx).astype(np.float32)
    w = (np.float32(scale) * _cubic(dist * np.float32(scale))).astype(np.float32) if scale < 1 else _cubic(dist).astype(np.float32)
    w = (w / w.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    zero = (w == 0).sum(axis=0)
    if zero[0] != 0:
        idx, w = idx[:, 1:p - 1], w[:, 1:p - 1]       # narrow(1, 1, p - 2)
    if zero[-1] != 0:
        idx, w = idx[:, :p - 2], w[:, :p - 2]         # narrow(1, 0, p - 2): a no-op when the first narrowing already happened
    sym_s = int(-idx.min() + 1)
    idx = (idx + sym_s - 1).astype(np.int64)          # positions in the symmetrically extended signal
    M = np.zeros((out_len, in_len), np.float64)
    for i in range(out_len):
        for a, wt in zip(idx[i], w[i]):
            if a < sym_s:
                o = sym_s - 1 - a
            elif a < sym_s + in_len:
                o = a - sym_s
            else:
                o = in_len - 1 - (a - sym_s - in_len)
            M[i, o] += float(wt)
    return M


class NiqeScorer:
    """scores = NiqeScorer(params_path)(pred): pred (S, 3, H, W) in [0, 1] on a CUDA device -> (S,) float32 NIQE (lower is
    better; select with `take_min=True`). `params_path`: the reference's basicsr/metrics/niqe_pris_params.npz."""

    BLOCK = 96

    def __init__(self, params_path: str | None = None, device=None):
        path = params_path or os.environ.get("BEM_NIQE_PARAMS")
        if not path or not os.path.exists(path):
            raise RuntimeError("NiqeScorer needs the reference's parameter file basicsr/metrics/niqe_pris_params.npz "
                               "(pass its path or set BEM_NIQE_PARAMS)")
        z = np.load(path)
        self._mu = np.asarray(z["mu_pris_param"], np.float64).reshape(-1)
        self._cov = np.asarray(z["cov_pris_param"], np.float64)
        self._win = np.asarray(z["gaussian_window"], np.float64)
        gam = np.arange(0.2, 10.001, 0.001)                                     # niqe.py:24
        self._gam = gam
        # everything the fits need of the gamma function, tabulated once on the host over the 9801 candidate shapes (alpha is
        # always one of them): r(gamma) of niqe.py:26, sqrt(G(1/a) / G(3/a)) of :36-37 and G(2/a) / G(1/a) of :58
        lg = np.vectorize(math.lgamma)
        g1, g2, g3 = lg(1.0 / gam), lg(2.0 / gam), lg(3.0 / gam)
        self._r_gam = np.exp(2 * g2 - g1 - g3)
        self._beta_scale = np.exp(0.5 * (g1 - g3))
        self._mean_scale = np.exp(g2 - g1)
        self._dev = {}
        self._resize = {}

    def _consts(self, device):
        c = self._dev.get(device)
        if c is None:
            t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=device)
            c = dict(mu=t(self._mu), cov=t(self._cov), win=t(self._win).contiguous(), gam=t(self._gam), r_gam=t(self._r_gam).contiguous(),
                     beta_scale=t(self._beta_scale), mean_scale=t(self._mean_scale))
            self._dev[device] = c
        return c

    def _resize_mats(self, h, w, device):
        key = (h, w, device)
        m = self._resize.get(key)
        if m is None:
            m = (torch.as_tensor(resize_matrix(h, 0.5), dtype=torch.float64, device=device),
                 torch.as_tensor(resize_matrix(w, 0.5), dtype=torch.float64, device=device))
            self._resize[key] = m
        return m

    @staticmethod
    def _aggd_from_moments(m, n, c):
        """m: (..., 6) = [sum_{v<0} v^2, #{v<0}, sum_{v>0} v^2, #{v>0}, sum |v|, sum v^2] over n values -> alpha, beta_l, beta_r
        (niqe.py:13-38). The table of r(gamma) is strictly increasing, so `argmin((r_gam - rhatnorm)**2)` is a binary search plus
        one comparison of the two neighbours (ties -> the lower index, as argmin); a NaN statistic selects index 0 like numpy."""
        left = torch.sqrt(m[..., 0] / m[..., 1])              # mean over an empty set -> nan, like numpy
        right = torch.sqrt(m[..., 2] / m[..., 3])
        gh = left / right
        rhat = (m[..., 4] / n) ** 2 / (m[..., 5] / n)
        rn = rhat * (gh ** 3 + 1) * (gh + 1) / (gh ** 2 + 1) ** 2
        rg = c["r_gam"]
        hi = torch.searchsorted(rg, rn.contiguous()).clamp(1, rg.numel() - 1)
        lo = hi - 1
        pick_hi = (rg[hi] - rn) ** 2 < (rg[lo] - rn) ** 2
        idx = torch.where(pick_hi, hi, lo)
        idx = torch.where(torch.isnan(rn), torch.zeros_like(idx), idx)
        s = c["beta_scale"][idx]
        return c["gam"][idx], left * s, right * s, c["mean_scale"][idx]

    def _features(self, img, c, bs):
        """img: (S, H, W) float32, H / W multiples of bs -> (S, blocks, 18) float64   (niqe.py:104-116, 41-60): the per-pixel
        work runs in csrc/niqe.cu (bem_niqe_mscn, bem_niqe_block_stats), the 18 features per block come from the moments"""
        from . import _lib
        from ._lib import lib
        S, H, W = img.shape
        img = img.contiguous()
        dev = img.device
        nrm = torch.empty((S, H, W), dtype=torch.float64, device=dev)
        nb = (H // bs) * (W // bs)
        mom = torch.empty((S, nb, 5, 6), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(lib.bem_niqe_mscn(_lib.ptr(img), _lib.ptr(c["win"]), _lib.ptr(nrm), S, H, W, st), "niqe_mscn")
            _lib.check(lib.bem_niqe_block_stats(_lib.ptr(nrm), _lib.ptr(mom), S, H, W, bs, st), "niqe_block_stats")
        _lib.profile.launches += 2
        a, bl, br, g21 = self._aggd_from_moments(mom, float(bs * bs), c)     # (S, nb, 5) each
        feats = [a[..., 0], (bl[..., 0] + br[..., 0]) / 2]
        for k in range(1, 5):
            feats += [a[..., k], (br[..., k] - bl[..., k]) * g21[..., k], bl[..., k], br[..., k]]
        return torch.stack(feats, dim=-1)

    @torch.no_grad()
    def __call__(self, pred: torch.Tensor) -> torch.Tensor:
        from . import _lib
        _lib.require_cuda(pred)
        if pred.dim() != 4 or pred.shape[1] != 3:
            raise RuntimeError("NiqeScorer expects (S, 3, H, W) predictions in [0, 1]")
        c = self._consts(pred.device)
        p = pred.float() * 255.0                                           # eval.py:250 `pred*255`, float32
        p = p / 255.0                                                      # to_y_channel (metric_util.py:45)
        y = (p[:, 0].double() * 24.966 + p[:, 1].double() * 128.553 + p[:, 2].double() * 65.481 + 16.0)   # bgr2ycbcr y_only
        y = ((y / 255.0).float() * 255.0)                                  # float32 round trips of the reference (:52, _convert_output_type_range)
        img = torch.round(y)                                               # float32, integer-valued
        bs = self.BLOCK
        S, H, W = img.shape
        nh, nw = H // bs, W // bs
        if nh == 0 or nw == 0:
            raise RuntimeError(f"NiqeScorer needs images of at least {bs} x {bs} pixels")
        img = img[:, :nh * bs, :nw * bs]
        f1 = self._features(img, c, bs)
        Mh, Mw = self._resize_mats(nh * bs, nw * bs, pred.device)
        small = (Mh @ (img / 255.0).double() @ Mw.t()).float() * 255.0     # imresize works in float32
        f2 = self._features(small, c, bs // 2)
        dist = torch.cat([f1, f2], dim=-1)                                 # (S, blocks, 36)
        mu_d = torch.nanmean(dist, dim=1)
        ok = ~torch.isnan(dist).any(dim=-1)                                # rows np.cov sees
        n = ok.sum(dim=1, keepdim=True).double()
        z = torch.where(ok.unsqueeze(-1), dist, torch.zeros_like(dist))
        m = z.sum(dim=1, keepdim=True) / n.unsqueeze(-1)
        zc = torch.where(ok.unsqueeze(-1), dist - m, torch.zeros_like(dist))
        cov_d = zc.transpose(1, 2) @ zc / (n.unsqueeze(-1) - 1)
        # niqe.py:132-134: d pinv(M) d^T with M = (cov_p + cov_d) / 2. M is symmetric positive definite (the pristine
        # covariance has full rank), where the pseudo-inverse is the inverse: one batched LU solve instead of S SVDs (the SVD
        # path of torch.linalg.pinv costs ~1.4 ms per image). solve_ex neither raises nor synchronises; a singular M (only
        # possible with NaN features) yields NaN, which selection treats like the reference does.
        d = (c["mu"] - mu_d).unsqueeze(-1)                                 # (S, 36, 1)
        sol = torch.linalg.solve_ex((c["cov"] + cov_d) / 2, d)[0]
        q = torch.sqrt((d * sol).sum(dim=(1, 2)))
        return q.float()
