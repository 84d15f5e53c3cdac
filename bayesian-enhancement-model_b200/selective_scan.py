"""Selective scan — the reference's three nested boundaries on top of bem_scan_fwd / bem_scan_bwd.

1. ``selective_scan_cuda_oflex``-compatible ``fwd`` / ``bwd`` (same positional signature, same checks and return lists
   as kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_oflex.cpp:157-358, pybind at :360-363)
2. ``SelectiveScanCuda`` / ``selective_scan_fn`` — the product API of basicsr/vmamba/models/csms6s.py:75-130
3. ``build_selective_scan_fn`` / ``selective_scan_fn_test_api`` — the mamba-style API of
   kernels/selective_scan/test_selective_scan.py:18-165 (B/C 3-D or 4-D, z, return_last_state)

Every path ends in the sm_100a kernels; CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import lib

MAX_DSTATE = 256      # selective_scan_oflex.cpp:190


def chunk_len(dtype: torch.dtype) -> int:
    """positions per carry chunk of ``x`` for an element type (the reference fixes 2048, selective_scan_oflex.cpp:218)"""
    return lib.bem_scan_chunk_len(_lib.dtype_code(dtype))


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)   # what TORCH_CHECK surfaces as in Python


def _common_checks(u, delta, A, B, C, D_, delta_bias_):
    _lib.require_cuda(u, delta, A, B, C, D_, delta_bias_)
    it = u.dtype
    _check(it in (torch.float32, torch.float16, torch.bfloat16), "selective_scan: input must be float32 / float16 / bfloat16")
    _check(A.dtype == torch.float32, "selective_scan: A must be float32")
    _check(delta.dtype == it and B.dtype == it and C.dtype == it, "selective_scan: u, delta, B, C must share a dtype")
    _check(u.dim() == 3 and delta.dim() == 3, "selective_scan: u, delta must be (batch, dim, seqlen)")
    _check(u.stride(-1) == 1 or u.size(-1) == 1, "selective_scan: u must be contiguous in the last dim")
    _check(delta.stride(-1) == 1 or delta.size(-1) == 1, "selective_scan: delta must be contiguous in the last dim")
    batch, dim, seqlen = u.shape
    _check(A.dim() == 2 and B.dim() == 4 and C.dim() == 4, "selective_scan: A must be (dim, dstate), B/C (batch, groups, dstate, seqlen)")
    dstate, n_groups = A.size(1), B.size(1)
    _check(dim % n_groups == 0, "dims should be dividable by n_groups")
    _check(dstate <= MAX_DSTATE, "selective_scan only supports state dimension <= 256")
    _check(tuple(delta.shape) == (batch, dim, seqlen), "delta must have shape (batch, dim, seqlen)")
    _check(tuple(A.shape) == (dim, dstate), "A must have shape (dim, dstate)")
    _check(tuple(B.shape) == (batch, n_groups, dstate, seqlen), "B must have shape (batch, n_groups, dstate, seqlen)")
    _check(tuple(C.shape) == (batch, n_groups, dstate, seqlen), "C must have shape (batch, n_groups, dstate, seqlen)")
    _check(B.stride(-1) == 1 or B.size(-1) == 1, "B must be contiguous in the last dim")
    _check(C.stride(-1) == 1 or C.size(-1) == 1, "C must be contiguous in the last dim")
    for name, t in (("D", D_), ("delta_bias", delta_bias_)):
        if t is not None:
            _check(t.dtype == torch.float32, f"{name} must be float32")
            _check(tuple(t.shape) == (dim,), f"{name} must have shape (dim,)")
            _check(t.stride(-1) == 1 or t.size(-1) == 1, f"{name} must be contiguous")
    return batch, dim, seqlen, dstate, n_groups


def fwd(u, delta, A, B, C, D_, delta_bias_, delta_softplus, nrows=1, out_float=True, dt_weight=None):
    """``selective_scan_cuda_oflex.fwd`` (selective_scan_oflex.cpp:157-243) -> ``[out, x]``.
    ``nrows`` is accepted and ignored exactly as in the reference (:236-238).
    ``dt_weight`` (extension, inference): (dim, R) dt_proj weight; ``delta`` is then the low-rank dt (batch, n_groups, R, L)
    and the projection ``F.conv1d(dts, dt_projs_weight, groups=K)`` (vmamba.py:661) happens inside the scan kernel."""
    if dt_weight is not None:
        return _fwd_lowrank(u, delta, A, B, C, D_, delta_bias_, delta_softplus, out_float, dt_weight)
    batch, dim, seqlen, dstate, n_groups = _common_checks(u, delta, A, B, C, D_, delta_bias_)
    dev = u.device
    dt = _lib.dtype_code(u.dtype)
    CL = lib.bem_scan_chunk_len(dt)
    n_chunks = (seqlen + CL - 1) // CL
    out = torch.empty((batch, dim, seqlen), dtype=torch.float32 if out_float else u.dtype, device=dev)
    x = torch.empty((batch, dim, n_chunks, dstate * 2), dtype=torch.float32, device=dev)
    need = lib.bem_scan_workspace_bytes(batch, dim, seqlen, dstate, dt)
    ws = _lib.workspace(dev, need)
    p = _lib.BemScanFwdParams(
        batch=batch, dim=dim, seqlen=seqlen, dstate=dstate, n_groups=n_groups, dtype=dt,
        out_dtype=_lib.BEM_F32 if out_float else dt, delta_softplus=int(bool(delta_softplus)),
        u=_lib.ptr(u), delta=_lib.ptr(delta), A=_lib.ptr(A), B=_lib.ptr(B), C=_lib.ptr(C), D=_lib.ptr(D_),
        delta_bias=_lib.ptr(delta_bias_), out=_lib.ptr(out), x=_lib.ptr(x),
        u_bs=u.stride(0), u_ds=u.stride(1), delta_bs=delta.stride(0), delta_ds=delta.stride(1),
        A_ds=A.stride(0), A_ns=A.stride(1),
        B_bs=B.stride(0), B_gs=B.stride(1), B_ns=B.stride(2), C_bs=C.stride(0), C_gs=C.stride(1), C_ns=C.stride(2),
        out_bs=out.stride(0), out_ds=out.stride(1), workspace=_lib.ptr(ws), workspace_bytes=ws.numel())
    es, eo = u.element_size(), out.element_size()
    nbytes = batch * dim * seqlen * (2 * es + eo) + 2 * batch * n_groups * dstate * seqlen * es   # SURVEY 8d, boundary form
    _lib.launch("scan_fwd", lib.bem_scan_fwd, p, dev, key=(batch, dim, dstate, seqlen, str(u.dtype)), nbytes=nbytes)
    return [out, x]


def fused_dt_rank_ok(dt_rank: int, dstate: int, dtype) -> bool:
    """configurations the fused dt_proj path of the forward kernel is built for"""
    return 0 < dt_rank <= 8 and dstate == 1 and dtype == torch.float32


def _fwd_lowrank(u, dtl, A, B, C, D_, delta_bias_, delta_softplus, out_float, dt_weight):
    batch, dim, seqlen = u.shape
    _lib.require_cuda(u, dtl, A, B, C, dt_weight)
    n_groups, R = dtl.shape[1], dtl.shape[2]
    dstate = A.shape[1]
    _check(u.dtype == torch.float32 and dtl.dtype == torch.float32, "fused dt_proj: float32 only")
    _check(tuple(dtl.shape) == (batch, n_groups, R, seqlen) and dtl.stride(-1) == 1, "low-rank delta must be (batch, groups, rank, L)")
    _check(tuple(dt_weight.shape) == (dim, R), "dt_weight must be (dim, rank)")
    _check(fused_dt_rank_ok(R, dstate, u.dtype), "fused dt_proj: dstate 1, rank <= 8")
    _check(u.stride(-1) == 1 and B.stride(-1) == 1 and C.stride(-1) == 1, "u, B, C must be contiguous in the last dim")
    dev = u.device
    dt = _lib.dtype_code(u.dtype)
    CL = lib.bem_scan_chunk_len(dt)
    n_chunks = (seqlen + CL - 1) // CL
    out = torch.empty((batch, dim, seqlen), dtype=torch.float32, device=dev)
    x = torch.empty((batch, dim, n_chunks, dstate * 2), dtype=torch.float32, device=dev)
    ws = _lib.workspace(dev, lib.bem_scan_workspace_bytes(batch, dim, seqlen, dstate, dt))
    w = dt_weight.to(torch.float32).contiguous()
    A = A.to(torch.float32)
    p = _lib.BemScanFwdParams(
        batch=batch, dim=dim, seqlen=seqlen, dstate=dstate, n_groups=n_groups, dtype=dt, out_dtype=_lib.BEM_F32,
        delta_softplus=int(bool(delta_softplus)), u=_lib.ptr(u), delta=_lib.ptr(dtl), A=_lib.ptr(A), B=_lib.ptr(B), C=_lib.ptr(C),
        D=_lib.ptr(D_), delta_bias=_lib.ptr(delta_bias_), out=_lib.ptr(out), x=_lib.ptr(x),
        u_bs=u.stride(0), u_ds=u.stride(1), delta_bs=dtl.stride(0), delta_ds=dtl.stride(2), A_ds=A.stride(0), A_ns=A.stride(1),
        B_bs=B.stride(0), B_gs=B.stride(1), B_ns=B.stride(2), C_bs=C.stride(0), C_gs=C.stride(1), C_ns=C.stride(2),
        out_bs=out.stride(0), out_ds=out.stride(1), workspace=_lib.ptr(ws), workspace_bytes=ws.numel(),
        dt_rank=R, dt_weight=_lib.ptr(w), delta_gs=dtl.stride(1))
    nbytes = 4 * (2 * batch * dim * seqlen + batch * n_groups * (R + 2 * dstate) * seqlen)
    _lib.launch("scan_fwd", lib.bem_scan_fwd, p, dev, key=(batch, dim, dstate, seqlen, "fused_dt%d" % R), nbytes=nbytes)
    return [out, x]


def bwd(u, delta, A, B, C, D_, delta_bias_, dout, x_, delta_softplus, nrows=1):
    """``selective_scan_cuda_oflex.bwd`` (selective_scan_oflex.cpp:245-358) ->
    ``[du, ddelta, dA, dB, dC, dD, ddelta_bias]`` (dD / ddelta_bias are None when the input was absent)."""
    batch, dim, seqlen, dstate, n_groups = _common_checks(u, delta, A, B, C, D_, delta_bias_)
    _lib.require_cuda(dout, x_)
    _check(dout.dtype == u.dtype or dout.dtype == torch.float32, "dout must be the input dtype or float32")
    _check(tuple(dout.shape) == (batch, dim, seqlen), "dout must have shape (batch, dim, seqlen)")
    _check(dout.stride(-1) == 1 or dout.size(-1) == 1, "dout must be contiguous in the last dim")
    dev = u.device
    dt = _lib.dtype_code(u.dtype)
    CL = lib.bem_scan_chunk_len(dt)
    n_chunks = (seqlen + CL - 1) // CL
    if n_chunks > 1:
        _check(x_ is not None, "x is required when seqlen spans more than one chunk")   # selective_scan_oflex.cpp:315
    if x_ is not None:
        _check(x_.dtype == torch.float32 and x_.is_contiguous(), "x must be contiguous float32")
        _check(tuple(x_.shape) == (batch, dim, n_chunks, 2 * dstate), "x must have shape (batch, dim, n_chunks, 2*dstate)")
    du = torch.empty_like(u, memory_format=torch.contiguous_format)
    ddelta = torch.empty_like(delta, memory_format=torch.contiguous_format)
    dA = torch.zeros((dim, dstate), dtype=torch.float32, device=dev)
    dB = torch.zeros((batch, n_groups, dstate, seqlen), dtype=torch.float32, device=dev)
    dC = torch.zeros((batch, n_groups, dstate, seqlen), dtype=torch.float32, device=dev)
    dD = torch.zeros_like(D_) if D_ is not None else None
    ddelta_bias = torch.zeros_like(delta_bias_) if delta_bias_ is not None else None
    need = lib.bem_scan_workspace_bytes(batch, dim, seqlen, dstate, dt)
    ws = _lib.workspace(dev, need)
    p = _lib.BemScanBwdParams(
        batch=batch, dim=dim, seqlen=seqlen, dstate=dstate, n_groups=n_groups, dtype=dt,
        dout_dtype=_lib.dtype_code(dout.dtype), delta_softplus=int(bool(delta_softplus)),
        u=_lib.ptr(u), delta=_lib.ptr(delta), A=_lib.ptr(A), B=_lib.ptr(B), C=_lib.ptr(C), D=_lib.ptr(D_),
        delta_bias=_lib.ptr(delta_bias_), dout=_lib.ptr(dout), x=_lib.ptr(x_), du=_lib.ptr(du), ddelta=_lib.ptr(ddelta),
        dA=_lib.ptr(dA), dB=_lib.ptr(dB), dC=_lib.ptr(dC), dD=_lib.ptr(dD), ddelta_bias=_lib.ptr(ddelta_bias),
        u_bs=u.stride(0), u_ds=u.stride(1), delta_bs=delta.stride(0), delta_ds=delta.stride(1),
        A_ds=A.stride(0), A_ns=A.stride(1),
        B_bs=B.stride(0), B_gs=B.stride(1), B_ns=B.stride(2), C_bs=C.stride(0), C_gs=C.stride(1), C_ns=C.stride(2),
        dout_bs=dout.stride(0), dout_ds=dout.stride(1), du_bs=du.stride(0), du_ds=du.stride(1),
        ddelta_bs=ddelta.stride(0), ddelta_ds=ddelta.stride(1), workspace=_lib.ptr(ws), workspace_bytes=ws.numel())
    es, eo = u.element_size(), dout.element_size()
    nbytes = batch * dim * seqlen * (4 * es + eo) + 4 * batch * n_groups * dstate * seqlen * es
    _lib.launch("scan_bwd", lib.bem_scan_bwd, p, dev, key=(batch, dim, dstate, seqlen, str(u.dtype)), nbytes=nbytes)
    return [du, ddelta, dA, dB.to(B.dtype), dC.to(C.dtype), dD, ddelta_bias]   # casts as in selective_scan_oflex.cpp:356


class _OflexModule:
    """Stands in for the pybind module ``selective_scan_cuda_oflex`` (selective_scan_oflex.cpp:360-363)."""
    __name__ = "selective_scan_cuda_oflex"
    fwd = staticmethod(fwd)
    bwd = staticmethod(bwd)


selective_scan_cuda_oflex = _OflexModule()


# ---------------------------------------------------------------------------------------------------
# product API — csms6s.py:75-130
# ---------------------------------------------------------------------------------------------------
class SelectiveScanCuda(torch.autograd.Function):
    """csms6s.SelectiveScanCuda (csms6s.py:75-113) with the single backend this package has."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, oflex=True, backend=None):
        if backend not in (None, "oflex"):
            raise RuntimeError(f"bem_b200 has one scan backend (the sm_100a kernels); backend={backend!r} does not exist here")
        ctx.delta_softplus = delta_softplus
        ctx.backend = "oflex"
        out, x = fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1, oflex)
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, ddelta_bias = bwd(u, delta, A, B, C, D, delta_bias, dout, x, ctx.delta_softplus, 1)
        return du, ddelta, dA, dB, dC, dD, ddelta_bias, None, None, None


def selective_scan_fn(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, oflex=True, backend=None):
    """csms6s.selective_scan_fn (csms6s.py:116-130). u, delta: (B, K*C, L); A: (K*C, N); B, C: (B, K, N, L);
    D, delta_bias: (K*C). ``backend="torch"`` (the reference's pure-PyTorch loop) is not provided: see oracle/."""
    return SelectiveScanCuda.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, oflex, backend)


# ---------------------------------------------------------------------------------------------------
# mamba-style test API — test_selective_scan.py:18-165
# ---------------------------------------------------------------------------------------------------
def build_selective_scan_fn(selective_scan_cuda=None, mode="ssoflex", tag=None):
    """test_selective_scan.build_selective_scan_fn for mode "ssoflex" (the only mode the reference builds,
    kernels/selective_scan/setup.py:40). `selective_scan_cuda` defaults to this package's module."""
    if mode != "ssoflex":
        raise RuntimeError(f"bem_b200 implements mode 'ssoflex' only (got {mode!r})")
    ext = selective_scan_cuda or selective_scan_cuda_oflex

    class SelectiveScanFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, return_last_state=False,
                    nrows=1, backnrows=-1):
            if u.stride(-1) != 1:
                u = u.contiguous()
            if delta.stride(-1) != 1:
                delta = delta.contiguous()
            if D is not None:
                D = D.contiguous()
            if B.stride(-1) != 1:
                B = B.contiguous()
            if C.stride(-1) != 1:
                C = C.contiguous()
            ctx.squeeze_B = ctx.squeeze_C = False
            if B.dim() == 3:
                B = B.unsqueeze(1)
                ctx.squeeze_B = True
            if C.dim() == 3:
                C = C.unsqueeze(1)
                ctx.squeeze_C = True
            ctx._d_dtype = ctx._delta_bias_dtype = None
            if D is not None and D.dtype != torch.float:
                ctx._d_dtype = D.dtype
                D = D.float()
            if delta_bias is not None and delta_bias.dtype != torch.float:
                ctx._delta_bias_dtype = delta_bias.dtype
                delta_bias = delta_bias.float()
            assert u.shape[1] % (B.shape[1] * nrows) == 0
            assert nrows in [1, 2, 3, 4]
            if backnrows > 0:
                assert u.shape[1] % (B.shape[1] * backnrows) == 0
                assert backnrows in [1, 2, 3, 4]
            else:
                backnrows = nrows
            ctx.backnrows = backnrows
            out, x = ext.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, True)
            ctx.delta_softplus = delta_softplus
            ctx.has_D, ctx.has_bias = D is not None, delta_bias is not None
            last_state = x[:, :, -1, 1::2]   # (batch, dim, dstate)  test_selective_scan.py:79
            ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
            if return_last_state:
                ctx.mark_non_differentiable(last_state)
                return out, last_state
            return out

        @staticmethod
        def backward(ctx, dout, *args):
            u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
            if dout.stride(-1) != 1:
                dout = dout.contiguous()
            du, ddelta, dA, dB, dC, dD, ddelta_bias = ext.bwd(u, delta, A, B, C, D, delta_bias, dout, x,
                                                              ctx.delta_softplus, ctx.backnrows)
            dB = dB.squeeze(1) if ctx.squeeze_B else dB
            dC = dC.squeeze(1) if ctx.squeeze_C else dC
            if dD is not None and ctx._d_dtype is not None:
                dD = dD.to(ctx._d_dtype)
            if ddelta_bias is not None and ctx._delta_bias_dtype is not None:
                ddelta_bias = ddelta_bias.to(ctx._delta_bias_dtype)
            return du, ddelta, dA, dB, dC, dD, ddelta_bias, None, None, None, None

    def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                          return_last_state=False, nrows=1, backnrows=-1):
        """if return_last_state is True, returns (out, last_state); last_state has shape (batch, dim, dstate) and carries
        no gradient. `z` gates the output as in selective_scan_ref (test_selective_scan.py:233-234): out * silu(z).
        (The reference's ssoflex wrapper accepts z and silently ignores it, test_selective_scan.py:88-89.)"""
        outs = SelectiveScanFn.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, return_last_state, nrows, backnrows)
        out, last = (outs if return_last_state else (outs, None))
        if z is not None:
            out = out * torch.nn.functional.silu(z.float())
        out = out.to(u.dtype)   # test_selective_scan.py:158-159
        return (out, last) if return_last_state else out

    selective_scan_fn.__repr__ = lambda *_: f"selective_scan_fn | {mode} | {tag}"
    return selective_scan_fn


selective_scan_fn_test_api = build_selective_scan_fn()
