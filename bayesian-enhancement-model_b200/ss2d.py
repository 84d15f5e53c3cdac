"""SS2D core — the caller of the hot path: SS2Dv2.forward_corev2, scan_mode cross2d / unidi / bidi, `no_einsum=True`
(basicsr/vmamba/models/vmamba.py:547-577, 656-698), as used by every BEM arch through forward_type "v05_noz"
(basicsr/archs/UNet_arch.py:205-228).

``ss2d_core`` is the function form (parameters passed explicitly) so it can be patched into the reference's SS2D
(patch.py) and used by the mirror modules in network.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .bayesian import functional as BF
from .csm import cross_merge_fn, cross_scan_fn
from .selective_scan import fused_dt_rank_ok, selective_scan_fn

_SCAN_MODES = dict(cross2d=0, unidi=1, bidi=2)


def ss2d_fwd(x, z, dt_weight, A, Dskip, delta_bias, dstate=1, delta_softplus=True):
    """bem_ss2d_fwd: x (B, D, H, W), z = x_proj output in image order (B, 4*(R+2N), H*W) -> y (B, D, H*W), fp32."""
    from ._lib import lib
    _lib.require_cuda(x, z, dt_weight, A)
    B, D, H, W = x.shape
    R = dt_weight.shape[1]
    x = x.contiguous()
    z = z.contiguous()
    dt_weight = dt_weight.to(torch.float32).contiguous()
    A = A.to(torch.float32).contiguous()
    y = torch.empty((B, D, H * W), dtype=torch.float32, device=x.device)
    ws = _lib.workspace(x.device, lib.bem_ss2d_workspace_bytes(B, D, H, W, dstate, R), kind="ss2d")
    p = _lib.BemSs2dFwdParams(batch=B, d_inner=D, H=H, W=W, dstate=dstate, dt_rank=R, delta_softplus=int(delta_softplus),
                              x=_lib.ptr(x), xdbl=_lib.ptr(z), dt_weight=_lib.ptr(dt_weight), A=_lib.ptr(A),
                              Dskip=_lib.ptr(Dskip), delta_bias=_lib.ptr(delta_bias), y=_lib.ptr(y), workspace=_lib.ptr(ws),
                              workspace_bytes=ws.numel())
    L = H * W
    fused = lib.bem_ss2d_supported(dstate, R) == 2
    # algorithmic bytes, SURVEY 8(d) fused-SS2D form with x_proj outside: x in, xdbl in, y out (the traversal-aware kernels
    # read x / xdbl twice; the composed form moves ~18 D*L units through four launches)
    nbytes = 4 * B * L * (D + 4 * (R + 2 * dstate) + D)
    _lib.launch("ss2d_fwd", lib.bem_ss2d_fwd, p, x.device, key=(B, D, H, W, R), nbytes=nbytes, kernels=3 if fused else 4)
    return y


def ss2d_scan(x, dts, As, Bs, Cs, Ds, delta_bias, H, W, delta_softplus=True, ssoflex=True, scans=0):
    """Fused operator of the SS2D core: cross_scan -> selective scan -> cross_merge for already projected
    inputs, i.e. vmamba.py:657 + :672-684 in one call.
      x          : (B, D, H, W) (or (B, D, H*W)) image-order input of the scan
      dts        : (B, K*D, L) delta before bias / softplus, in TRAVERSAL order (as forward_corev2 computes it from xs)
      As         : (K*D, N);  Bs, Cs : (B, K, N, L) traversal order;  Ds, delta_bias : (K*D) or None
    -> y : (B, D, L), the four directions merged back to image order (before out_norm). Differentiable: it composes the
    autograd functions of cross_scan_fn / selective_scan_fn / cross_merge_fn."""
    B, D = x.shape[0], x.shape[1]
    x = x.reshape(B, D, H, W)
    K = Bs.shape[1]
    xs = cross_scan_fn(x, in_channel_first=True, out_channel_first=True, scans=scans)            # (B, K, D, L)
    ys = selective_scan_fn(xs.view(B, -1, H * W), dts, As, Bs, Cs, Ds, delta_bias, delta_softplus, ssoflex)
    return cross_merge_fn(ys.view(B, K, -1, H, W), in_channel_first=True, out_channel_first=True, scans=scans)


def forward_corev2_patched(self, x=None, force_fp32=False, ssoflex=True, no_einsum=False, selective_scan_backend=None,
                           scan_mode="cross2d", scan_force_torch=False, **kwargs):
    """Replacement for SS2Dv2.forward_corev2 (vmamba.py:547-698) installed by bem_b200.patch: same signature, same result,
    the whole core on this package's kernels (ss2d_core). Modes this package does not build raise instead of falling
    back; channel_last SS2D (unused by the BEM archs) permutes around the channel-first core."""
    if selective_scan_backend not in (None, "oflex"):
        raise RuntimeError(f"bem_b200: selective_scan_backend {selective_scan_backend!r} is not built (oflex only)")
    if scan_mode == "cascade2d":
        raise RuntimeError("bem_b200: scan_mode 'cascade2d' is not built (cross2d / unidi / bidi)")
    y = ss2d_core(x, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds,
                  x_proj_bias=getattr(self, "x_proj_bias", None), out_norm=None, scan_mode=scan_mode, force_fp32=force_fp32,
                  ssoflex=ssoflex, pack_cache=self.__dict__.setdefault("_pack_cache", {}))
    if not self.channel_first:
        y = y.permute(0, 2, 3, 1).contiguous()
    return self.out_norm(y).to(x.dtype)


def ss2d_core(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, x_proj_bias=None, out_norm=None,
              scan_mode="cross2d", force_fp32=False, ssoflex=True, delta_softplus=True, pack_cache=None):
    """x: (B, D, H, W) -> y: (B, D, H, W) (after `out_norm` when given), following vmamba.py:656-698 line by line:
    cross_scan -> x_proj (grouped 1x1) -> split dt/B/C -> dt_proj (grouped 1x1) -> selective scan -> cross_merge."""
    if scan_mode not in _SCAN_MODES:
        raise RuntimeError(f"ss2d_core: scan_mode {scan_mode!r} is not built (cross2d / unidi / bidi)")
    scans = _SCAN_MODES[scan_mode]
    B, D, H, W = x.shape
    N = A_logs.shape[1]
    K, _, R = dt_projs_weight.shape
    L = H * W
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or x_proj_weight.requires_grad or dt_projs_weight.requires_grad)
    if needs_grad or x.dtype != torch.float32:
        xs = cross_scan_fn(x, in_channel_first=True, out_channel_first=True, scans=scans)            # (B, 4, D, L)
        x_dbl = F.conv1d(xs.view(B, -1, L), x_proj_weight.view(-1, D, 1),
                         bias=(x_proj_bias.view(-1) if x_proj_bias is not None else None), groups=K)  # vmamba.py:659
        x_dbl = x_dbl.view(B, K, -1, L)
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)                                             # vmamba.py:660
        dts = F.conv1d(dts.contiguous().view(B, -1, L), dt_projs_weight.view(K * D, -1, 1), groups=K)  # vmamba.py:661
    else:
        # x_proj: a 1x1 convolution commutes with the pixel permutation of a traversal, so the K projections are one
        # 1x1 conv of the UN-scanned x (D -> K*(R+2N) channels, x read once instead of the four scanned copies), whose
        # result each direction then traverses on its own channel block (cross_scan, one_by_one) — vmamba.py:659
        Cx = x_proj_weight.shape[1]
        z = BF.pointwise_conv(x.reshape(B, D, L), x_proj_weight.reshape(1, K * Cx, D),
                              None if x_proj_bias is None else x_proj_bias.reshape(1, K * Cx), 1, pack_cache=pack_cache)
        if scans == 0 and x.dtype == torch.float32 and _lib.lib.bem_ss2d_supported(N, R) > 0 and not force_fp32:
            # everything after x_proj in ONE C-ABI call (bem_ss2d_fwd): cross_scan(x), per-direction traversal of z,
            # selective scan with dt_proj fused (the (B, K*D, L) delta tensor is neither written nor read), cross_merge
            As = None if pack_cache is None else pack_cache.get("As")
            akey = (A_logs.data_ptr(), A_logs._version, _lib.cache_generation())
            if As is None or pack_cache.get("As_key") != akey:     # -exp(A_logs) once per parameter version, not per call
                As = -A_logs.detach().to(torch.float).exp()
                if pack_cache is not None:
                    pack_cache["As"], pack_cache["As_key"] = As, akey
            y = ss2d_fwd(x, z, dt_projs_weight.reshape(K * D, R), As, Ds.to(torch.float),
                         dt_projs_bias.reshape(-1).to(torch.float), N, bool(delta_softplus)).view(B, -1, H, W)
            if out_norm is not None:
                y = out_norm(y)
            return y.to(x.dtype)
        xs = cross_scan_fn(x, in_channel_first=True, out_channel_first=True, scans=scans)
        x_dbl = cross_scan_fn(z.view(B, K, Cx, H, W), in_channel_first=True, out_channel_first=True, one_by_one=True, scans=scans)
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
        # dt_proj: the K directions as the weight sets of the tcgen05 pointwise kernel; dts is read as a strided channel
        # slice of x_dbl (no .contiguous() copy) — vmamba.py:661
        dts = BF.grouped_pointwise(dts, dt_projs_weight, pack_cache=None if pack_cache is None else pack_cache.setdefault("dt_pack", {}))
    xs = xs.view(B, -1, L)
    dts = dts.contiguous().view(B, -1, L)
    As = -A_logs.to(torch.float).exp()                  # (K * D, N)
    Dsf = Ds.to(torch.float)
    delta_bias = dt_projs_bias.view(-1).to(torch.float)
    # Bs / Cs stay strided views of x_dbl (last-dim stride 1): the scan kernel takes 64-bit strides, the reference
    # materialises them with .contiguous() (vmamba.py:666-667)
    if force_fp32:
        xs, dts, Bs, Cs = (t.to(torch.float32) for t in (xs, dts, Bs, Cs))
    ys = selective_scan_fn(xs, dts, As, Bs, Cs, Dsf, delta_bias, delta_softplus, ssoflex).view(B, K, -1, H, W)
    y = cross_merge_fn(ys, in_channel_first=True, out_channel_first=True, scans=scans)             # (B, D, L)
    y = y.view(B, -1, H, W)
    if out_norm is not None:
        y = out_norm(y)
    return y.to(x.dtype)


# ---------------------------------------------------------------------------------------------------------------------
# mirror modules (same parameter names / shapes / init as the reference, so state_dicts interchange)
# ---------------------------------------------------------------------------------------------------------------------
class Linear2d(nn.Linear):
    """vmamba.Linear2d (vmamba.py:42-56): a Linear applied as a 1x1 convolution on (B, C, H, W)."""

    def forward(self, x: torch.Tensor):
        return F.conv2d(x, self.weight[:, :, None, None], self.bias)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        state_dict[prefix + "weight"] = state_dict[prefix + "weight"].view(self.weight.shape)
        return super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class LayerNorm2d(nn.LayerNorm):
    """vmamba.LayerNorm2d (vmamba.py:59-64)."""

    def forward(self, x: torch.Tensor):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


def _dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4):
    """mamba_init.dt_init (vmamba.py:224-249)"""
    dt_proj = nn.Linear(dt_rank, d_inner, bias=True)
    std = dt_rank ** -0.5 * dt_scale
    if dt_init == "constant":
        nn.init.constant_(dt_proj.weight, std)
    elif dt_init == "random":
        nn.init.uniform_(dt_proj.weight, -std, std)
    else:
        raise NotImplementedError
    dt = torch.exp(torch.rand(d_inner) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
    inv_dt = dt + torch.log(-torch.expm1(-dt))
    with torch.no_grad():
        dt_proj.bias.copy_(inv_dt)
    return dt_proj


class SS2D(_lib.InvalidatesCaches, nn.Module):
    """The SS2D configuration every BEM arch instantiates: forward_type "v05_noz", channel_first, ssm_init "v0"
    (vmamba.py:438-545 __initv2__, :700-716 forwardv2). Other forward types are out of scope and raise."""

    def __init__(self, d_model=96, d_state=16, ssm_ratio=2.0, dt_rank="auto", act_layer=nn.SiLU, d_conv=3, conv_bias=True,
                 dropout=0.0, bias=False, dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4,
                 initialize="v0", forward_type="v05_noz", channel_first=True, **kwargs):
        super().__init__()
        if forward_type != "v05_noz" or not channel_first or initialize != "v0":
            raise NotImplementedError("bem_b200.SS2D mirrors forward_type='v05_noz', channel_first=True, initialize='v0' only")
        d_inner = int(ssm_ratio * d_model)
        dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.channel_first = True
        self.with_dconv = d_conv > 1
        self.disable_z = True
        k_group = 4
        self.out_norm = LayerNorm2d(d_inner)
        self.in_proj = Linear2d(d_model, d_inner, bias=bias)
        self.act = act_layer()
        if self.with_dconv:
            self.conv2d = nn.Conv2d(d_inner, d_inner, groups=d_inner, bias=conv_bias, kernel_size=d_conv,
                                    padding=(d_conv - 1) // 2)
        x_proj = [nn.Linear(d_inner, dt_rank + d_state * 2, bias=False) for _ in range(k_group)]
        self.x_proj_weight = nn.Parameter(torch.stack([t.weight for t in x_proj], dim=0))   # (K, R + 2N, D)
        self.out_act = nn.Identity()
        self.out_proj = Linear2d(d_inner, d_model, bias=bias)
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()
        dt_projs = [_dt_init(dt_rank, d_inner, dt_scale, dt_init, dt_min, dt_max, dt_init_floor) for _ in range(k_group)]
        self.dt_projs_weight = nn.Parameter(torch.stack([t.weight for t in dt_projs], dim=0))   # (K, D, R)
        self.dt_projs_bias = nn.Parameter(torch.stack([t.bias for t in dt_projs], dim=0))       # (K, D)
        A = torch.arange(1, d_state + 1, dtype=torch.float32).view(1, -1).repeat(d_inner, 1).contiguous()
        self.A_logs = nn.Parameter(torch.log(A)[None].repeat(k_group, 1, 1).flatten(0, 1).contiguous())   # (K*D, N)
        self.Ds = nn.Parameter(torch.ones(k_group * d_inner))
        self.A_logs._no_weight_decay = True
        self.Ds._no_weight_decay = True

    def forward_core(self, x, apply_out_norm=True):
        return ss2d_core(x, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds,
                         x_proj_bias=getattr(self, "x_proj_bias", None), out_norm=self.out_norm if apply_out_norm else None,
                         pack_cache=self.__dict__.setdefault("_pack_cache", {}))

    def forward(self, x: torch.Tensor, pre_norm=None, residual=None, **kwargs):
        """forwardv2 (vmamba.py:700-716). `pre_norm`: the block's LayerNorm2d; when in_proj / out_proj are Bayesian 1x1
        layers of this package both the block norm and `out_norm` are fused into their kernels, SiLU into the depthwise
        kernel and `residual` (the block's skip connection, vmamba.py:1331) into out_proj's epilogue."""
        x = apply_1x1(self.in_proj, x, pre_norm)
        if self.with_dconv and fuses_act(self.conv2d) and isinstance(self.act, nn.SiLU):
            x = self.conv2d(x, post_act="silu")
        else:
            if self.with_dconv:
                x = self.conv2d(x)
            x = self.act(x)
        fuse_out = fuses_norm(self.out_proj) and isinstance(self.out_act, nn.Identity) and not torch.is_grad_enabled()
        y = self.forward_core(x, apply_out_norm=not fuse_out)
        y = self.out_act(y)
        if isinstance(self.dropout, nn.Identity):
            return apply_residual(self.out_proj, y, residual, self.out_norm if fuse_out else None)
        y = self.dropout(apply_1x1(self.out_proj, y, self.out_norm if fuse_out else None))
        return y if residual is None else residual + y


def fuses_norm(layer) -> bool:
    f = getattr(layer, "_fuses_norm", None)
    return bool(f and f())


def fuses_act(layer) -> bool:
    f = getattr(layer, "_fuses_act", None)
    return bool(f and f()) and not torch.is_grad_enabled()


def apply_residual(layer, x, residual, norm=None):
    """residual + layer(norm(x)), the additions / normalisation handed to the layer when it can fuse them"""
    f = getattr(layer, "_fuses_residual", None)
    if residual is not None and f and f() and not torch.is_grad_enabled():
        if norm is not None and fuses_norm(layer):
            return layer(x, pre_norm=norm, residual=residual)
        return layer(x if norm is None else norm(x), residual=residual)
    y = apply_1x1(layer, x, norm)
    return y if residual is None else residual + y


def apply_1x1(layer, x, norm):
    """norm -> 1x1 layer; the normalisation is handed to the layer when it can fuse it (bem_b200.bayesian 1x1 layers)."""
    if norm is None:
        return layer(x)
    if fuses_norm(layer):
        return layer(x, pre_norm=norm)
    return layer(norm(x))
