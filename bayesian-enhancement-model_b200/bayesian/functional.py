"""Kernel wrappers + autograd for the Bayesian layers (csrc/bayes.cu).

sample_weights      w[s] = mu + log1p(exp(rho)) * eps[s]          (basicsr/bayesian/conv.py:106-107)
pointwise_conv      S-batched 1x1 convolution with per-sample weights (conv.py:114 / linear.py:90 for 1x1 shapes)
depthwise_conv3x3   S-batched depthwise 3x3 (conv.py:114 with groups == channels)

Forward passes run the hand-written kernels. Backward passes (stage-1 training only) reuse the same kernels for the
data gradients; the tiny weight-gradient reductions use torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .._lib import lib


def _f32c(t, name):
    if t is None:
        return None
    _lib.require_cuda(t)
    if t.dtype != torch.float32:
        raise RuntimeError(f"bem_b200.bayesian: {name} must be float32 (got {t.dtype})")
    return t.contiguous()


def _sample_raw(mu, rho, eps, n_samples, seed, stream_id, sample0, want_eps):
    mu = _f32c(mu, "mu")
    rho = _f32c(rho, "rho")
    eps = _f32c(eps, "eps")
    numel = mu.numel()
    if eps is not None and eps.numel() != n_samples * numel:
        raise RuntimeError(f"eps has {eps.numel()} elements, expected n_samples * numel = {n_samples * numel}")
    w = torch.empty((n_samples,) + tuple(mu.shape), dtype=torch.float32, device=mu.device)
    eps_out = torch.empty_like(w) if (want_eps and eps is None and rho is not None) else None
    p = _lib.BemBayesSampleParams(numel=numel, n_samples=n_samples, mu=_lib.ptr(mu), rho=_lib.ptr(rho), eps=_lib.ptr(eps),
                                  w=_lib.ptr(w), eps_out=_lib.ptr(eps_out), seed=int(seed) & (2 ** 64 - 1),
                                  stream_id=int(stream_id), sample0=int(sample0))
    _lib.launch("bayes_sample", lib.bem_bayes_sample, p, mu.device, nbytes=4 * numel * (2 + n_samples))
    used = eps.view_as(w) if eps is not None else eps_out
    return w, used


class _SampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, rho, eps, n_samples, seed, stream_id, sample0):
        w, used = _sample_raw(mu, rho, eps, n_samples, seed, stream_id, sample0, want_eps=True)
        ctx.save_for_backward(rho, used)
        ctx.mark_non_differentiable(used)
        return w, used

    @staticmethod
    def backward(ctx, dw, _deps):
        rho, eps = ctx.saved_tensors
        dmu = dw.sum(0)
        drho = (dw * eps).sum(0) * torch.sigmoid(rho)   # d log1p(exp(rho)) / d rho
        return dmu, drho, None, None, None, None, None


def sample_weights(mu, rho, eps=None, n_samples=1, seed=0, stream_id=0, sample0=0):
    """-> (w, eps_used); w: (S, *mu.shape). eps=None draws the counter-based Philox stream inside the kernel
    (keyed (seed, stream_id, sample0 + s): independent of how samples are split over ranks)."""
    if torch.is_grad_enabled() and (mu.requires_grad or rho.requires_grad):
        return _SampleFn.apply(mu, rho, eps, n_samples, seed, stream_id, sample0)
    return _sample_raw(mu, rho, eps, n_samples, seed, stream_id, sample0, want_eps=True)


# ---------------------------------------------------------------------------------------------------
class PackPlan:
    """The pack steps of the 1x1 layers of one forward, batched into one launch (bem_bayes_pointwise_pack_run).

    A Monte-Carlo forward re-draws all Bayesian weights first (mc.MCArena.draw), so nothing a layer's pack kernel reads
    depends on the layers before it. recording(): one ordinary forward during which every eligible pointwise call (weights
    given as a tensor, no per-layer pack cache) gets a workspace of its own and is noted; build() turns the notes into the
    device table; run() packs them all; playing(): the same forward again, each call now passing `prepacked`. Calls are
    matched by position and checked against the recorded signature (weight / bias / LayerNorm pointers, shape, alignment
    class of x) — a call that does not match simply packs for itself as before. Only calls whose weight (and bias) live
    inside `stable` = (first, last) byte address of the buffer the draw fills are eligible: anything computed during the
    forward is not there yet when run() packs.
    The active plan is process-global state (like torch's grad mode is per thread, this is per process): one planned forward at a
    time; the lanes of an MCSampler record and capture one after the other and only REPLAY concurrently, which involves no plan."""

    def __init__(self, stable):
        self.stable = (int(stable[0]), int(stable[1]))
        self.entries = []      # (signature, params, workspace, tensors kept alive)
        self.table = None
        self.total_blocks = 0
        self.mode = None
        self.cursor = 0
        self.misses = 0

    def recording(self):
        self.entries, self.table, self.mode, self.cursor = [], None, "record", 0
        return _PlanScope(self)

    def playing(self):
        self.mode, self.cursor = "play", 0
        return _PlanScope(self)

    def build(self, device):
        n = len(self.entries)
        if n == 0:
            return False
        arr = (_lib.BemBayesPointwiseParams * n)(*[e[1] for e in self.entries])
        nbytes = lib.bem_bayes_pointwise_pack_table_bytes(n)
        host = torch.empty(nbytes, dtype=torch.uint8)
        total = C.c_int32(0)
        _lib.check(lib.bem_bayes_pointwise_pack_table(arr, n, C.c_void_p(host.data_ptr()), C.byref(total)), "bayes_pointwise_pack_table")
        self.table = host.to(device)
        self.total_blocks = int(total.value)
        return True

    def run(self):
        """pack every recorded layer from the current contents of its weight tensors (one launch, current stream)"""
        if self.table is None:
            return
        dev = self.table.device
        with torch.cuda.device(dev):
            code = lib.bem_bayes_pointwise_pack_run(_lib.ptr(self.table), len(self.entries), self.total_blocks, _lib.stream_ptr(dev))
        _lib.profile.launches += 1
        _lib.check(code, "bayes_pointwise_pack_run")

    def eligible(self, w, bias):
        lo, hi = self.stable
        return w is not None and lo <= w.data_ptr() < hi and (bias is None or lo <= bias.data_ptr() < hi)

    def _lookup(self, sig):
        if self.cursor < len(self.entries) and self.entries[self.cursor][0] == sig and self.table is not None:
            ws = self.entries[self.cursor][2]
            self.cursor += 1
            return ws
        self.misses += 1
        self.cursor = len(self.entries)      # out of step: every later call packs for itself
        return None


class _PlanScope:
    def __init__(self, plan):
        self.plan = plan

    def __enter__(self):
        global _ACTIVE_PLAN
        self.prev, _ACTIVE_PLAN = _ACTIVE_PLAN, self.plan
        return self.plan

    def __exit__(self, *exc):
        global _ACTIVE_PLAN
        _ACTIVE_PLAN = self.prev
        self.plan.mode = None
        return False


_ACTIVE_PLAN = None


def _pointwise_raw(x, w=None, bias=None, mu=None, sigma=None, eps=None, n_samples=1, ln=None, force_simt=False,
                   interleave=False, residual=None, pack_cache=None, prelu=None):
    """x: (S*Bx, Cin, *spatial) fp32 — any image stride, channels P apart; w: (S|1, Cout, Cin) or mu/sigma/eps for the
    fused sample-on-load path; ln = (gamma, beta, eps): LayerNorm over the channels of every pixel fused into the
    activation staging; interleave: image i uses weight set i % S instead of i // Bx; residual: (batch, Cout, *spatial)
    contiguous tensor added to the result in the epilogue (the block's skip connection)."""
    _lib.require_cuda(x)
    if x.dtype != torch.float32:
        raise RuntimeError(f"bem_b200.bayesian: input must be float32 (got {x.dtype})")
    w, bias, mu, sigma, eps = (_f32c(t, n) for t, n in ((w, "w"), (bias, "bias"), (mu, "mu"), (sigma, "sigma"), (eps, "eps")))
    ref = w if w is not None else mu
    cout, cin = int(ref.shape[-2]), int(ref.shape[-1])
    if x.dim() < 3 or x.shape[1] != cin:
        raise RuntimeError(f"pointwise conv: input {tuple(x.shape)} does not match weight (.., {cout}, {cin})")
    batch = x.shape[0]
    P = x[0, 0].numel()
    # accepted without a copy: pixels contiguous, channels P apart, any image stride (e.g. a channel slice of x_dbl)
    inner_ok = x[0, 0].is_contiguous() and (cin == 1 or x.stride(1) == P)
    if not inner_ok:
        x = x.contiguous()
    img_stride = x.stride(0) if batch > 1 else cin * P
    out = torch.empty((batch, cout) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    g = b = None
    ln_eps = 0.0
    if ln is not None:
        g, b, ln_eps = _f32c(ln[0], "ln weight"), _f32c(ln[1], "ln bias"), float(ln[2])
        if g.numel() != cin:
            raise RuntimeError(f"fused LayerNorm expects {cin} channels, got {g.numel()}")
    if residual is not None:
        residual = _f32c(residual, "residual")
        if tuple(residual.shape) != tuple(out.shape):
            raise RuntimeError(f"pointwise conv: residual {tuple(residual.shape)} does not match output {tuple(out.shape)}")
    prelu = _f32c(prelu, "prelu slope")
    if prelu is not None and prelu.numel() not in (1, cout):
        raise RuntimeError(f"pointwise conv: PReLU with {prelu.numel()} slopes does not match {cout} output channels")
    need = lib.bem_bayes_pointwise_workspace_bytes(n_samples, cin, cout)
    prepacked = 0
    sig = None
    if pack_cache is not None and not force_simt:
        # constant weights (deterministic layers): the packed tiles live in the layer's own buffer and are rebuilt only when
        # a participating tensor changes (in-place updates bump ._version) or x changes alignment class
        # One buffer PER key, never rewritten for another key: a captured CUDA graph has `prepacked = 1` and the buffer's
        # address baked in, so a later eager call with another alignment class (another image size) must not repack the
        # tiles the graph reads. A new cache generation (weights written through .data, load_state_dict, ...) clears the lot;
        # graphs are re-captured on the same signal (mc.MCArena.valid).
        gen = _lib.cache_generation()
        if pack_cache.get("gen") != gen:
            pack_cache.clear()
            pack_cache["gen"] = gen
        key = tuple((t.data_ptr(), t._version) for t in (w, bias, g, b) if t is not None) + (n_samples, P % 4 == 0, int(img_stride) % 4 == 0,
                                                                                              x.data_ptr() % 16 == 0, float(ln_eps), x.device.index)
        packs = pack_cache.setdefault("packs", {})
        ws = packs.get(key)
        prepacked = int(ws is not None)
        if ws is None:
            if len(packs) >= 8:      # stale tensor versions (optimizer steps between evaluations): keep the dict small
                packs.clear()
            ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            packs[key] = ws
    else:
        ws = None
        plan = _ACTIVE_PLAN
        if plan is not None and not force_simt and plan.eligible(w, bias):
            sig = (w.data_ptr(), tuple(w.shape)) + tuple(None if t is None else t.data_ptr() for t in (bias, g, b)) + (
                float(ln_eps), n_samples, batch, P, int(img_stride) % 4 == 0, x.data_ptr() % 16 == 0, bool(interleave))
            if plan.mode == "play":
                ws = plan._lookup(sig)
                prepacked = int(ws is not None)
            elif plan.mode == "record":
                ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        if ws is None:
            ws = _lib.workspace(x.device, need, kind="pointwise")
    p = _lib.BemBayesPointwiseParams(n_samples=n_samples, batch=batch, cin=cin, cout=cout, P=P, x=_lib.ptr(x),
                                     w=_lib.ptr(w), mu=_lib.ptr(mu), rho=None, eps=_lib.ptr(eps), bias=_lib.ptr(bias),
                                     out=_lib.ptr(out), sigma=_lib.ptr(sigma), ln_gamma=_lib.ptr(g), ln_beta=_lib.ptr(b),
                                     ln_eps=ln_eps, force_simt=int(bool(force_simt)), x_img_stride=int(img_stride),
                                     sample_interleave=int(bool(interleave)), workspace=_lib.ptr(ws), workspace_bytes=ws.numel(),
                                     residual=_lib.ptr(residual), prepacked=prepacked, prelu_slope=_lib.ptr(prelu),
                                     prelu_n=0 if prelu is None else prelu.numel())
    if pack_cache is None and _ACTIVE_PLAN is not None and _ACTIVE_PLAN.mode == "record" and sig is not None:
        q = _lib.BemBayesPointwiseParams.from_buffer_copy(p)
        _ACTIVE_PLAN.entries.append((sig, q, ws, (w, bias, g, b)))
    _lib.launch("bayes_pointwise", lib.bem_bayes_pointwise, p, x.device, key=(batch, cin, cout, P),
                nbytes=4 * batch * P * (cin + cout * (2 if residual is not None else 1)), kernels=1 if (force_simt or prepacked) else 2)
    return out


class _PointwiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, n_samples):
        ctx.save_for_backward(x, w)
        ctx.n_samples = n_samples
        ctx.has_bias = bias is not None
        return _pointwise_raw(x, w=w, bias=bias, n_samples=n_samples)

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        S = ctx.n_samples
        dout = dout.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _pointwise_raw(dout, w=w.transpose(-1, -2).contiguous(), n_samples=S)   # W^T dout with the same kernel
        if ctx.needs_input_grad[1] or ctx.has_bias:
            Bx = x.shape[0] // S
            xs = x.reshape(S, Bx, x.shape[1], -1)
            ds = dout.reshape(S, Bx, dout.shape[1], -1)
            if ctx.needs_input_grad[1]:
                dw = torch.einsum("sbop,sbip->soi", ds, xs).reshape(w.shape) if w.shape[0] == S else \
                    torch.einsum("sbop,sbip->oi", ds, xs).reshape(w.shape)
            if ctx.has_bias:
                db = ds.sum(dim=(1, 3))
        return dx, dw, db, None


def grouped_pointwise(x, w, bias=None, pack_cache=None):
    """Grouped 1x1 convolution as the reference's `F.conv1d(x.view(B, K*Cin, L), w.view(K*Cout, Cin, 1), groups=K)`
    (vmamba.py:659-661): x: (B, K, Cin, L) (any stride between (b, k) images), w: (K, Cout, Cin) -> (B, K, Cout, L).
    The K groups are the kernel's weight sets. Inference only (no autograd)."""
    B, K, Cin, L = x.shape
    xv = x.reshape(B * K, Cin, L) if x.is_contiguous() else x.flatten(0, 1)
    out = _pointwise_raw(xv, w=w, bias=bias, n_samples=K, interleave=True, pack_cache=pack_cache)
    return out.view(B, K, -1, L)


def pointwise_conv(x, w, bias=None, n_samples=1, ln=None, force_simt=False, residual=None, pack_cache=None, prelu=None):
    """out[img] = w[s(img)] @ LN(x[img]) + bias[s(img)] (+ residual[img]); w: (S, Cout, Cin), bias: (S, Cout) | None.
    ln = (gamma, beta, eps) fuses the preceding LayerNorm2d, residual the skip connection (both inference only)."""
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or (bias is not None and bias.requires_grad)):
        if ln is not None:
            raise RuntimeError("the fused LayerNorm path has no backward; apply the norm module separately when training")
        y = _PointwiseFn.apply(x, w, bias, n_samples)
        y = y if residual is None else residual + y
        return y if prelu is None else torch.nn.functional.prelu(y, prelu)
    if residual is not None and prelu is not None and not force_simt:   # not a combination the tensor-core epilogue is built for
        y = _pointwise_raw(x, w=w, bias=bias, n_samples=n_samples, ln=ln, residual=residual, pack_cache=pack_cache)
        return torch.nn.functional.prelu(y, prelu)
    return _pointwise_raw(x, w=w, bias=bias, n_samples=n_samples, ln=ln, force_simt=force_simt, residual=residual,
                          pack_cache=pack_cache, prelu=prelu)


def pointwise_conv_sampled(x, mu, sigma, eps, bias=None, n_samples=1, ln=None, force_simt=False, residual=None):
    """inference path: w = mu + sigma * eps (sigma = log1p(exp(rho)) precomputed once per layer) is formed while the
    weight tile is staged — the sampled weight never exists in HBM."""
    return _pointwise_raw(x, mu=mu, sigma=sigma, eps=eps, bias=bias, n_samples=n_samples, ln=ln, force_simt=force_simt,
                          residual=residual)


# ---------------------------------------------------------------------------------------------------
ACTS = {None: 0, "none": 0, "silu": 1, "gelu_gate": 2}


def _depthwise_raw(x, w, bias, n_samples, act=None):
    x = _f32c(x, "input")
    w = _f32c(w, "w")
    bias = _f32c(bias, "bias")
    batch, Cc, H, W = x.shape
    K = int(w.shape[-1])
    code = ACTS[act]
    if code == 2 and Cc % 2:
        raise RuntimeError("gated GELU needs an even channel count")
    out = torch.empty((batch, Cc // 2 if code == 2 else Cc, H, W), dtype=torch.float32, device=x.device)
    p = _lib.BemBayesDepthwiseParams(n_samples=n_samples, batch=batch, C=Cc, H=H, W=W, K=K, x=_lib.ptr(x), w=_lib.ptr(w),
                                     bias=_lib.ptr(bias), out=_lib.ptr(out), act=code)
    _lib.launch("bayes_depthwise", lib.bem_bayes_depthwise, p, x.device, key=(batch, Cc, H, W, code),
                nbytes=4 * (x.numel() + out.numel()))
    return out


def apply_act(y, act):
    """the unfused form of the depthwise kernel's `act` codes (training / non-depthwise geometries)"""
    if ACTS[act] == 1:
        return torch.nn.functional.silu(y)
    if ACTS[act] == 2:
        y1, y2 = y.chunk(2, dim=1)
        return torch.nn.functional.gelu(y1) * y2
    return y


class _DepthwiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, n_samples):
        ctx.save_for_backward(x, w)
        ctx.n_samples = n_samples
        ctx.has_bias = bias is not None
        return _depthwise_raw(x, w, bias, n_samples)

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        S = ctx.n_samples
        dout = dout.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _depthwise_raw(dout, torch.flip(w, dims=(-1, -2)).contiguous(), None, S)   # correlation with the flipped taps
        if ctx.needs_input_grad[1]:
            Bx = x.shape[0] // S
            Cc, H, W = x.shape[1:]
            xp = torch.nn.functional.pad(x, (1, 1, 1, 1)).reshape(S, Bx, Cc, H + 2, W + 2)
            ds = dout.reshape(S, Bx, Cc, H, W)
            taps = [(ds * xp[..., i:i + H, j:j + W]).sum(dim=(1, 3, 4)) for i in range(3) for j in range(3)]
            dw = torch.stack(taps, dim=-1).reshape(S, Cc, 3, 3)
            dw = dw.reshape(w.shape) if w.shape[0] == S else dw.sum(0).reshape(w.shape)
        if ctx.has_bias:
            db = dout.reshape(S, -1, dout.shape[1], dout.shape[2] * dout.shape[3]).sum(dim=(1, 3))
        return dx, dw, db, None


def depthwise_conv3x3(x, w, bias=None, n_samples=1, act=None):
    """x: (S*Bx, C, H, W); w: (S, C, 3, 3) (or (S, C, 1, 3, 3)); bias: (S, C) | None; stride 1, zero padding 1.
    act: None | "silu" | "gelu_gate" — the activation that follows the convolution, fused when no gradient is needed."""
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or (bias is not None and bias.requires_grad)):
        return apply_act(_DepthwiseFn.apply(x, w, bias, n_samples), act)
    return _depthwise_raw(x, w, bias, n_samples, act)


def sample_batched(entries, blocks, n_blocks, seed, sample0=0, sample0_dev=None):
    """one launch sampling every tensor listed in the device table `entries` (see MCArena)"""
    p = _lib.BemBayesSampleBatchedParams(entries=_lib.ptr(entries), blocks=_lib.ptr(blocks), n_blocks=int(n_blocks),
                                         seed=int(seed), sample0=int(sample0), sample0_dev=_lib.ptr(sample0_dev))
    _lib.launch("bayes_sample", lib.bem_bayes_sample_batched, p, entries.device, key=("batched", int(n_blocks)),
                nbytes=0)


def conv3x3_direct(x, w, bias=None):
    """Dense 3x3 convolution, stride 1, zero padding 1 (bem_conv3x3): the network's small-channel stems. Inference only."""
    _lib.require_cuda(x, w)
    x = _f32c(x, "input")
    w = _f32c(w, "w")
    bias = _f32c(bias, "bias")
    B, cin, H, W = x.shape
    cout = w.shape[0]
    if tuple(w.shape) != (cout, cin, 3, 3):
        raise RuntimeError(f"conv3x3_direct: weight {tuple(w.shape)} does not match input channels {cin}")
    out = torch.empty((B, cout, H, W), dtype=torch.float32, device=x.device)
    p = _lib.BemConv3x3Params(batch=B, cin=cin, cout=cout, H=H, W=W, x=_lib.ptr(x), w=_lib.ptr(w), bias=_lib.ptr(bias), out=_lib.ptr(out))
    _lib.launch("conv3x3", lib.bem_conv3x3, p, x.device, key=(B, cin, cout, H, W), nbytes=4 * (x.numel() + out.numel()))
    return out
