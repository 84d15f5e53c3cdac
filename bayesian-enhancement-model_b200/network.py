"""Host-side mirror of the stage-1 condition generator: `Network` of basicsr/archs/UNet_arch.py:365-474 and the blocks it
is built from (VSSBlock vmamba.py:1241-1334, gdMlp vmamba.py:116-133, BasicBlock / SubNetwork / PatchMerging /
DualUpSample UNet_arch.py:57-362).

Why it exists: BASELINE configs 2/3 (MC-sample images/s of the stage-1 Bayesian UNet) must run on a box that has no copy
of the reference. Module / parameter names and shapes equal the reference's, so `state_dict`s (including released
`{'params': ...}` checkpoints and their `mu_*/rho_*` keys after `convert2bnn_selective`) load unchanged; tests/ pins the
forward against vectors recorded from the reference model. Everything on the hot path (scan, traversal, Bayesian layers)
runs on this package's kernels; the deterministic glue around it (LayerNorm, activations, resampling, the two 3x3 stem
convolutions) is ordinary PyTorch and is outside the scope table.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from . import _lib, bayesian
from .bayesian import functional as BF
from .ss2d import SS2D, LayerNorm2d, apply_1x1, apply_residual, fuses_act


class Conv2d(_lib.InvalidatesCaches, nn.Conv2d):
    """nn.Conv2d (same name, parameters and state_dict) whose small-channel 3x3 stems run on bem_conv3x3 and whose deterministic 1x1 case — PatchMerging.reduction, the
    DualUpSample projections, the decoder fusion convs (UNet_arch.py:88-135, 161-163) — runs on the same tcgen05 pointwise
    kernel as the Bayesian 1x1 layers (one weight set, no sampling) when no gradient is needed; everything else is the
    library convolution."""

    def _is_1x1(self):
        return self.kernel_size == (1, 1) and self.stride == (1, 1) and self.padding == (0, 0) and self.groups == 1

    def _fuses_norm(self):
        return self._is_1x1()

    def forward(self, x, pre_norm=None, post_prelu=None):
        """pre_norm: a LayerNorm2d that precedes the conv (PatchMerging, UNet_arch.py:80-82); post_prelu: an nn.PReLU that
        follows it (DualUpSample); both fused when the 1x1 kernel runs, applied separately otherwise"""
        if (self._is_1x1() and x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32
                and not (torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad))):
            w = self.weight.view(1, self.out_channels, self.in_channels)
            ln = None if pre_norm is None else (pre_norm.weight, pre_norm.bias, pre_norm.eps)
            cache = self.__dict__.setdefault("_pack_cache", {})   # constant weights: packed once, reused by every call
            return BF.pointwise_conv(x, w, None if self.bias is None else self.bias.view(1, -1), 1, ln=ln, pack_cache=cache,
                                     prelu=None if post_prelu is None else post_prelu.weight)
        if (pre_norm is None and self.kernel_size == (3, 3) and self.stride == (1, 1) and self.padding == (1, 1)
                and self.dilation == (1, 1) and self.groups == 1 and self.padding_mode == "zeros"
                and min(self.in_channels, self.out_channels) <= 8 and self.in_channels * 9 * 8 * 4 <= 48 * 1024
                and x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32
                and not (torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad))):
            y = BF.conv3x3_direct(x, self.weight, self.bias)     # the full-resolution stems (first_conv, proj)
            return y if post_prelu is None else post_prelu(y)
        y = super().forward(x if pre_norm is None else pre_norm(x))
        return y if post_prelu is None else post_prelu(y)


def conv1x1_of_cat(conv, a, b):
    """conv(torch.cat([a, b], dim=1)) for a bias-free 1x1 `conv`: W_a a + W_b b as two 1x1 launches, the second adding onto
    the first in its epilogue — the concatenation is never materialised. Inference fast path; otherwise the plain form."""
    fast = (isinstance(conv, Conv2d) and conv._is_1x1() and conv.bias is None and a.is_cuda and a.dtype == torch.float32
            and b.dtype == torch.float32 and not torch.is_grad_enabled())
    if not fast:
        return conv(torch.cat([a, b], dim=1))
    ca = a.shape[1]
    cache = conv.__dict__.setdefault("_split_cache", {})
    key = (conv.weight.data_ptr(), conv.weight._version, ca, _lib.cache_generation())
    if cache.get("key") != key:
        w = conv.weight.detach().view(conv.out_channels, -1)
        cache.update(key=key, wa=w[:, :ca].contiguous().unsqueeze(0), wb=w[:, ca:].contiguous().unsqueeze(0), pa={}, pb={})
    y = BF.pointwise_conv(a, cache["wa"], None, 1, pack_cache=cache["pa"])
    return BF.pointwise_conv(b, cache["wb"], None, 1, residual=y, pack_cache=cache["pb"])


class gdMlp(nn.Module):
    """Gated-Dconv MLP (vmamba.py:116-133)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, channels_first=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.project_in = Conv2d(in_features, hidden_features * 2, kernel_size=1)
        self.dwconv = Conv2d(hidden_features * 2, hidden_features * 2, kernel_size=3, stride=1, padding=1,
                                groups=hidden_features * 2)
        self.project_out = Conv2d(hidden_features, out_features, kernel_size=1)
        self.act = act_layer()

    def forward(self, x, pre_norm=None, residual=None):
        """vmamba.py:127-133; `residual`: the block's skip connection, added by project_out when it can fuse it."""
        x = apply_1x1(self.project_in, x, pre_norm)
        if fuses_act(self.dwconv) and isinstance(self.act, nn.GELU) and self.act.approximate == "none":
            x = self.dwconv(x, post_act="gelu_gate")         # chunk -> gelu(x1) * x2 inside the depthwise kernel
        else:
            x1, x2 = self.dwconv(x).chunk(2, dim=1)
            x = self.act(x1) * x2
        return apply_residual(self.project_out, x, residual)


class VSSBlock(nn.Module):
    """VSSBlock._forwardv01 with post_norm=False, drop_path=0 (vmamba.py:1319-1334)."""

    def __init__(self, hidden_dim=0, drop_path=0, norm_layer=LayerNorm2d, channel_first=True, ssm_d_state=16, ssm_ratio=2.0,
                 ssm_dt_rank="auto", ssm_act_layer=nn.SiLU, ssm_conv=3, ssm_conv_bias=True, ssm_drop_rate=0, ssm_init="v0",
                 forward_type="v05_noz", mlp_ratio=4.0, mlp_act_layer=nn.GELU, mlp_drop_rate=0.0, mlp_type="gdmlp",
                 use_checkpoint=False, post_norm=False, **kwargs):
        super().__init__()
        if drop_path != 0 or post_norm or use_checkpoint or mlp_type != "gdmlp":
            raise NotImplementedError("bem_b200.VSSBlock mirrors the BEM configuration only (UNet_arch.py:205-228)")
        self.norm = norm_layer(hidden_dim)
        self.op = SS2D(d_model=hidden_dim, d_state=ssm_d_state, ssm_ratio=ssm_ratio, dt_rank=ssm_dt_rank,
                       act_layer=ssm_act_layer, d_conv=ssm_conv, conv_bias=ssm_conv_bias, dropout=ssm_drop_rate,
                       initialize=ssm_init, forward_type=forward_type, channel_first=channel_first)
        self.norm2 = norm_layer(hidden_dim)
        self.mlp = gdMlp(in_features=hidden_dim, hidden_features=int(hidden_dim * mlp_ratio), act_layer=mlp_act_layer,
                         drop=mlp_drop_rate, channels_first=channel_first)

    def forward(self, x):
        x = self.op(x, pre_norm=self.norm, residual=x)
        return self.mlp(x, pre_norm=self.norm2, residual=x)


class PatchMerging(nn.Module):
    """UNet_arch.PatchMerging (UNet_arch.py:57-83)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.norm = LayerNorm2d(4 * dim)
        self.reduction = Conv2d(4 * dim, 2 * dim, 1, 1, 0, bias=False)

    def forward(self, x):
        x = torch.cat([x[:, :, 0::2, 0::2], x[:, :, 1::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 1::2]], 1)
        return apply_1x1(self.reduction, x, self.norm)


class DualUpSample(nn.Module):
    """UNet_arch.DualUpSample (UNet_arch.py:97-156), scale factors 2 and 4."""

    def __init__(self, in_channels, scale_factor):
        super().__init__()
        self.factor = scale_factor
        c = in_channels
        if scale_factor == 2:
            self.conv = Conv2d(c, c // 2, 1, 1, 0, bias=False)
            self.up_p = nn.Sequential(Conv2d(c, 2 * c, 1, 1, 0, bias=False), nn.PReLU(), nn.PixelShuffle(2),
                                      Conv2d(c // 2, c // 2, 1, stride=1, padding=0, bias=False))
            self.up_b = nn.Sequential(Conv2d(c, c, 1, 1, 0), nn.PReLU(),
                                      nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False),
                                      Conv2d(c, c // 2, 1, stride=1, padding=0, bias=False))
        elif scale_factor == 4:
            self.conv = Conv2d(2 * c, c, 1, 1, 0, bias=False)
            self.up_p = nn.Sequential(Conv2d(c, 16 * c, 1, 1, 0, bias=False), nn.PReLU(), nn.PixelShuffle(4),
                                      Conv2d(c, c, 1, stride=1, padding=0, bias=False))
            self.up_b = nn.Sequential(Conv2d(c, c, 1, 1, 0), nn.PReLU(),
                                      nn.Upsample(scale_factor=4, mode="bilinear", align_corners=False),
                                      Conv2d(c, c, 1, stride=1, padding=0, bias=False))
        else:
            raise NotImplementedError(scale_factor)

    def forward(self, x):
        """UNet_arch.py:150-156: conv(cat([up_p(x), up_b(x)])). At inference two exact rewrites of the linear pieces save the
        full-resolution intermediates: the bias-free 1x1 conv that follows the bilinear upsampling is applied BEFORE it (both
        are linear, the interpolation weights sum to one, so they commute: half the channels to upsample, a quarter of the
        pixels to convolve), and conv(cat([p, b])) = W_p p + W_b b is evaluated as two 1x1 launches, the second adding onto
        the first in its epilogue, instead of materialising the concatenation."""
        fast = (x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and self.up_b[3].bias is None
                and isinstance(self.up_b[2], nn.Upsample) and self.up_b[2].mode == "bilinear")
        if not fast:
            return self.conv(torch.cat([self.up_p(x), self.up_b(x)], dim=1))
        prelu_ok = isinstance(self.up_p[1], nn.PReLU) and isinstance(self.up_b[1], nn.PReLU) and isinstance(self.up_p[0], Conv2d)
        if prelu_ok:   # conv -> PReLU in one launch
            p = self.up_p[3](self.up_p[2](self.up_p[0](x, post_prelu=self.up_p[1])))
            b = self.up_b[2](self.up_b[3](self.up_b[0](x, post_prelu=self.up_b[1])))
        else:
            p = self.up_p(x)
            b = self.up_b[2](self.up_b[3](self.up_b[1](self.up_b[0](x))))
        return conv1x1_of_cat(self.conv, p, b)


class BasicBlock(nn.Module):
    """UNet_arch.BasicBlock (UNet_arch.py:179-243); `.bayesian` marks the region convert2bnn_selective converts."""

    def __init__(self, dim, num_blocks=2, d_state=1, ssm_ratio=1, mlp_ratio=4, mlp_type="gdmlp", sam=False, condition=False,
                 bayesian=False):
        super().__init__()
        if sam or condition:
            raise NotImplementedError("SAM / condition blocks are not part of the stage-1 model (UNet_arch.py:477-487)")
        self.bayesian = bayesian
        self.sam = sam
        self.condition = condition
        self.blocks = nn.ModuleList([
            VSSBlock(hidden_dim=dim, drop_path=0, norm_layer=LayerNorm2d, channel_first=True, ssm_d_state=d_state,
                     ssm_ratio=ssm_ratio, ssm_dt_rank="auto", ssm_act_layer=nn.SiLU, ssm_conv=3, ssm_conv_bias=False,
                     ssm_drop_rate=0, ssm_init="v0", forward_type="v05_noz", mlp_ratio=mlp_ratio, mlp_act_layer=nn.GELU,
                     mlp_drop_rate=0.0, mlp_type=mlp_type, use_checkpoint=False, post_norm=False)
            for _ in range(num_blocks)])

    def forward(self, x):
        for block in self.blocks:
            x = block(x)
        return x


class SubNetwork(nn.Module):
    """UNet_arch.SubNetwork (UNet_arch.py:246-362) with use_pixelshuffle=True (PatchMerging / DualUpSample)."""

    def __init__(self, dim=31, num_blocks=(2, 4, 4), d_state=(1, 1, 1), ssm_ratio=1, mlp_ratio=4, mlp_type="gdmlp",
                 use_pixelshuffle=True, drop_path=0.0, sam=False):
        super().__init__()
        if not use_pixelshuffle or drop_path > 0 or sam:
            raise NotImplementedError("bem_b200.SubNetwork mirrors build_model()'s configuration (UNet_arch.py:477-487)")
        self.dim = dim
        level = len(num_blocks) - 1
        self.level = level
        self.encoder_layers = nn.ModuleList([])
        self.drop_path = nn.Identity()
        curr = dim
        for i in range(level):
            self.encoder_layers.append(nn.ModuleList([
                BasicBlock(dim=curr, num_blocks=num_blocks[i], d_state=d_state[i], ssm_ratio=ssm_ratio, mlp_ratio=mlp_ratio,
                           mlp_type=mlp_type, bayesian=True),
                PatchMerging(curr)]))
            curr *= 2
        self.bottleneck = BasicBlock(dim=curr, num_blocks=num_blocks[-1], d_state=d_state[level], ssm_ratio=ssm_ratio,
                                     mlp_ratio=mlp_ratio, bayesian=True)
        self.decoder_layers = nn.ModuleList([])
        for i in range(level):
            self.decoder_layers.append(nn.ModuleList([
                DualUpSample(curr, scale_factor=2),
                Conv2d(curr, curr // 2, 1, 1, bias=False),
                BasicBlock(dim=curr // 2, num_blocks=num_blocks[level - 1 - i], d_state=d_state[level - 1 - i],
                           ssm_ratio=ssm_ratio, mlp_ratio=mlp_ratio, bayesian=True)]))
            curr //= 2
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        """UNet_arch.py:331-338: every nn.Linear (hence Linear2d in_proj / out_proj) and nn.LayerNorm is re-initialised."""
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.weight, 1.0)
            nn.init.constant_(m.bias, 0)

    def forward(self, x):
        fea = x
        skips = []
        for en_block, down in self.encoder_layers:
            fea = en_block(fea)
            skips.append(fea)
            fea = down(fea)
        fea = self.bottleneck(fea)
        for i, (up, fusion, de_block) in enumerate(self.decoder_layers):
            fea = up(fea)
            fea = conv1x1_of_cat(fusion, fea, skips[self.level - 1 - i])
            fea = de_block(fea)
        return x + self.drop_path(fea)


class Network(nn.Module):
    """UNet_arch.Network (UNet_arch.py:365-474): first_conv -> `stage` SubNetworks -> proj; returns [x, out_1, ...]."""

    def __init__(self, in_channels=3, out_channels=3, n_feat=40, stage=1, num_blocks=(1, 1, 1), d_state=1, ssm_ratio=1,
                 mlp_ratio=4, mlp_type="gdmlp", use_pixelshuffle=False, drop_path=0.0, use_illu=False, sam=False,
                 last_act=None):
        super().__init__()
        self.stage = stage
        self.mask_token = nn.Parameter(torch.zeros(1, n_feat, 1, 1))
        nn.init.trunc_normal_(self.mask_token, mean=0.0, std=0.02)
        self.first_conv = Conv2d(in_channels, n_feat, 3, 1, 1, bias=True)
        nn.init.kaiming_normal_(self.first_conv.weight, mode="fan_out", nonlinearity="linear")
        nn.init.zeros_(self.first_conv.bias)
        self.subnets = nn.ModuleList([])
        self.proj = Conv2d(n_feat, out_channels, 3, 1, 1, bias=True)
        nn.init.zeros_(self.proj.bias)
        if last_act is None:
            self.last_act = nn.Identity()
        elif last_act == "relu":
            self.last_act = nn.ReLU()
        elif last_act == "softmax":
            self.last_act = nn.Softmax(dim=1)
        else:
            raise NotImplementedError
        if isinstance(d_state, int):
            d_state = [d_state] * len(num_blocks)
        for _ in range(stage):
            self.subnets.append(SubNetwork(dim=n_feat, num_blocks=list(num_blocks), d_state=list(d_state),
                                           ssm_ratio=ssm_ratio, mlp_ratio=mlp_ratio, mlp_type=mlp_type,
                                           use_pixelshuffle=use_pixelshuffle, drop_path=drop_path, sam=sam))

    def forward(self, x, mask=None):
        out_list = [x]
        fea = self.first_conv(x)
        B, C, H, W = fea.size()
        if self.training and mask is not None:
            mask_tokens = self.mask_token.expand(B, -1, H, W)
            w = mask.unsqueeze(1).type_as(mask_tokens)
            fea = fea * (1.0 - w) + mask_tokens * w
        for subnet in self.subnets:
            fea = subnet(fea)
            out_list.append(self.last_act(self.proj(fea)))
        return out_list


def build_model():
    """UNet_arch.build_model (UNet_arch.py:477-487): the stage-1 configuration."""
    return Network(stage=1, n_feat=40, num_blocks=[2, 2, 2], d_state=[1, 1, 1], ssm_ratio=1, mlp_ratio=4, mlp_type="gdmlp",
                   use_pixelshuffle=True)


def build_bayesian_model(sigma_init=0.05, decay=0.998, pretrain=False, selective=True):
    """stage-1 Bayesian condition generator: build_model() + convert2bnn exactly as ConditionGenerator.__init__ does
    (basicsr/models/condition_generator_model.py:50-59)."""
    net = build_model()
    cfg = {"sigma_init": sigma_init, "decay": decay, "pretrain": pretrain}
    if selective:
        bayesian.convert2bnn_selective(net, cfg)
    else:
        bayesian.convert2bnn(net, cfg)
    bayesian.set_mc_config(net)   # assigns layer ids
    return net
