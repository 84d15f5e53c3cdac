"""oracle/network.py — TEST INFRASTRUCTURE ONLY.

CPU restatement of the stage-1 condition generator forward — `Network.forward` (basicsr/archs/UNet_arch.py:443-474) with
`SubNetwork.forward` (:340-362), `BasicBlock` / `VSSBlock._forwardv01` (vmamba.py:1319-1334), `SS2D.forwardv2` +
`forward_corev2` cross2d (vmamba.py:656-716), `gdMlp.forward` (vmamba.py:128-133), `PatchMerging` (UNet_arch.py:70-83),
`DualUpSample` (UNet_arch.py:147-156) and the Bayesian layers in stochastic / deterministic mode (basicsr/bayesian).

It is a functional walk over a reference-format state_dict (plain or after convert2bnn: `*.mu_weight / *.rho_weight`),
so it needs no module classes. Dense arithmetic uses torch CPU ops (what the reference itself executes on a CPU); the
selective scan uses the C restatement in scan_oracle.c (the reference's CPU scan is a Python loop over L,
csms6s.py:61-67 — far slower; using the C loop makes this baseline conservative), the traversal uses the numpy
restatement. Pinned against tests/golden/models.npz (outputs of the real reference network).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

import oracle


def _ln2d(x, w, b, eps=1e-5):
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps).permute(0, 3, 1, 2)


class _Weights:
    """resolves `<name>.weight` / `<name>.bias` either from plain keys or from mu/rho (+ eps) of a Bayesian layer"""

    def __init__(self, sd, eps=None, deterministic=False, generator=None):
        self.sd = sd
        self.eps = eps or {}
        self.det = deterministic
        self.gen = generator
        self.used_eps = {}

    def _draw(self, key, like):
        if key in self.eps:
            return torch.as_tensor(self.eps[key]).reshape(like.shape)
        e = torch.randn(like.shape, generator=self.gen)
        self.used_eps[key] = e
        return e

    def get(self, name, which="weight"):
        sd = self.sd
        if f"{name}.{which}" in sd:
            return sd[f"{name}.{which}"]
        mu_k = f"{name}.mu_{which}"
        if mu_k not in sd:
            return None
        mu = sd[mu_k]
        if self.det:
            return mu
        rho = sd[f"{name}.rho_{which}"]
        eps = self._draw(f"{name}.eps_{which}", mu)
        return mu + torch.log1p(torch.exp(rho)) * eps        # conv.py:106-107


def _conv(x, W, name, stride=1, padding=0, groups=1):
    w = W.get(name, "weight")
    b = W.get(name, "bias")
    if w.dim() == 2:
        w = w[:, :, None, None]                                # Linear2d (vmamba.py:49-51, linear.py:90)
    return F.conv2d(x, w, b, stride, padding, 1, groups)


def ss2d_core(x, sd, p):
    """forward_corev2, scan_mode cross2d, no_einsum (vmamba.py:656-698)"""
    B, D, H, W = x.shape
    L = H * W
    xw, dtw, dtb = sd[f"{p}.x_proj_weight"], sd[f"{p}.dt_projs_weight"], sd[f"{p}.dt_projs_bias"]
    A_logs, Ds = sd[f"{p}.A_logs"], sd[f"{p}.Ds"]
    K, _, R = dtw.shape
    N = A_logs.shape[1]
    xs = torch.from_numpy(oracle.cross_scan_oracle(x.numpy()))                          # (B,4,D,L)
    x_dbl = F.conv1d(xs.view(B, -1, L), xw.view(-1, D, 1), None, groups=K).view(B, K, -1, L)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = F.conv1d(dts.contiguous().view(B, -1, L), dtw.view(K * D, -1, 1), groups=K)
    As = -A_logs.float().exp()
    ys = oracle.selective_scan_oracle(xs.view(B, -1, L), dts.contiguous().view(B, -1, L), As, Bs.contiguous(),
                                      Cs.contiguous(), Ds.float(), None, dtb.view(-1).float(), True)   # fp32, csms6s.py:29-72
    y = oracle.cross_merge_oracle(ys.reshape(B, K, D, L), H, W)                           # (B,D,L)
    y = torch.from_numpy(y).view(B, D, H, W)
    return _ln2d(y, sd[f"{p}.out_norm.weight"], sd[f"{p}.out_norm.bias"])


def vss_block(x, W, p):
    sd = W.sd
    h = _ln2d(x, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"])
    h = _conv(h, W, f"{p}.op.in_proj")
    D = h.shape[1]
    h = _conv(h, W, f"{p}.op.conv2d", padding=1, groups=D)
    h = F.silu(h)
    h = ss2d_core(h, sd, f"{p}.op")
    x = x + _conv(h, W, f"{p}.op.out_proj")
    h = _ln2d(x, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"])
    h = _conv(h, W, f"{p}.mlp.project_in")
    h = _conv(h, W, f"{p}.mlp.dwconv", padding=1, groups=h.shape[1])
    x1, x2 = h.chunk(2, dim=1)
    h = F.gelu(x1) * x2
    return x + _conv(h, W, f"{p}.mlp.project_out")


def _basic_block(x, W, p):
    i = 0
    while f"{p}.blocks.{i}.norm.weight" in W.sd:
        x = vss_block(x, W, f"{p}.blocks.{i}")
        i += 1
    return x


def _patch_merging(x, W, p):
    x = torch.cat([x[:, :, 0::2, 0::2], x[:, :, 1::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 1::2]], 1)
    return _conv(_ln2d(x, W.sd[f"{p}.norm.weight"], W.sd[f"{p}.norm.bias"]), W, f"{p}.reduction")


def _dual_upsample(x, W, p):
    sd = W.sd
    xp = _conv(x, W, f"{p}.up_p.0")
    xp = F.prelu(xp, sd[f"{p}.up_p.1.weight"])
    xp = F.pixel_shuffle(xp, 2)
    xp = _conv(xp, W, f"{p}.up_p.3")
    xb = _conv(x, W, f"{p}.up_b.0")
    xb = F.prelu(xb, sd[f"{p}.up_b.1.weight"])
    xb = F.interpolate(xb, scale_factor=2, mode="bilinear", align_corners=False)
    xb = _conv(xb, W, f"{p}.up_b.3")
    return _conv(torch.cat([xp, xb], dim=1), W, f"{p}.conv")


@torch.no_grad()
def network_forward(sd, x, eps=None, deterministic=False, generator=None, return_eps=False):
    """sd: reference-format state_dict of `Network` (values: CPU tensors / numpy); x: (B,3,H,W) CPU float tensor.
    Returns the last element of the network's output list (UNet_arch.py:469-474), optionally with the eps it drew."""
    sd = {k: torch.as_tensor(v).float() for k, v in sd.items()}
    W = _Weights(sd, eps, deterministic, generator)
    x = torch.as_tensor(x).float()
    fea = _conv(x, W, "first_conv", padding=1)
    s = 0
    out = None
    while f"subnets.{s}.bottleneck.blocks.0.norm.weight" in sd:
        p = f"subnets.{s}"
        level = 0
        while f"{p}.encoder_layers.{level}.0.blocks.0.norm.weight" in sd:
            level += 1
        inp = fea
        skips = []
        for i in range(level):
            fea = _basic_block(fea, W, f"{p}.encoder_layers.{i}.0")
            skips.append(fea)
            fea = _patch_merging(fea, W, f"{p}.encoder_layers.{i}.1")
        fea = _basic_block(fea, W, f"{p}.bottleneck")
        for i in range(level):
            fea = _dual_upsample(fea, W, f"{p}.decoder_layers.{i}.0")
            fea = _conv(torch.cat([fea, skips[level - 1 - i]], dim=1), W, f"{p}.decoder_layers.{i}.1")
            fea = _basic_block(fea, W, f"{p}.decoder_layers.{i}.2")
        fea = inp + fea
        out = _conv(fea, W, "proj", padding=1)
        s += 1
    return (out, W.used_eps) if return_eps else out
