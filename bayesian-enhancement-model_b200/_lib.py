"""ctypes binding of libbem_b200.so (include/bem_b200.h). The library is the product: if it is missing this module
raises at import time — there is no Python / PyTorch / CPU fallback for any operator in this package."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbem_b200.so")

BEM_F32, BEM_F16, BEM_BF16 = 0, 1, 2
BEM_OK, BEM_ERR_BAD_ARG, BEM_ERR_WORKSPACE, BEM_ERR_UNSUPPORTED = 0, 10001, 10002, 10003
ABI_VERSION = 14

i32, i64, u64, vp = C.c_int32, C.c_int64, C.c_uint64, C.c_void_p


class BemScanFwdParams(C.Structure):
    _fields_ = [(n, i32) for n in ("batch", "dim", "seqlen", "dstate", "n_groups", "dtype", "out_dtype", "delta_softplus")] + \
               [(n, vp) for n in ("u", "delta", "A", "B", "C", "D", "delta_bias", "out", "x")] + \
               [(n, i64) for n in ("u_bs", "u_ds", "delta_bs", "delta_ds", "A_ds", "A_ns", "B_bs", "B_gs", "B_ns",
                                   "C_bs", "C_gs", "C_ns", "out_bs", "out_ds")] + \
               [("workspace", vp), ("workspace_bytes", i64), ("dt_rank", i32), ("dt_weight", vp), ("delta_gs", i64)]


class BemScanBwdParams(C.Structure):
    _fields_ = [(n, i32) for n in ("batch", "dim", "seqlen", "dstate", "n_groups", "dtype", "dout_dtype", "delta_softplus")] + \
               [(n, vp) for n in ("u", "delta", "A", "B", "C", "D", "delta_bias", "dout", "x", "du", "ddelta", "dA", "dB",
                                  "dC", "dD", "ddelta_bias")] + \
               [(n, i64) for n in ("u_bs", "u_ds", "delta_bs", "delta_ds", "A_ds", "A_ns", "B_bs", "B_gs", "B_ns",
                                   "C_bs", "C_gs", "C_ns", "dout_bs", "dout_ds", "du_bs", "du_ds", "ddelta_bs", "ddelta_ds")] + \
               [("workspace", vp), ("workspace_bytes", i64)]


class BemCsmParams(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "C", "H", "W", "dtype", "img_channel_first", "seq_channel_first", "one_by_one",
                                   "scans")] + [("src", vp), ("dst", vp)]


class BemSs2dFwdParams(C.Structure):
    _fields_ = [(n, i32) for n in ("batch", "d_inner", "H", "W", "dstate", "dt_rank", "delta_softplus")] + \
               [(n, vp) for n in ("x", "xdbl", "dt_weight", "A", "Dskip", "delta_bias", "y", "workspace")] + \
               [("workspace_bytes", i64)]


class BemBayesSampleParams(C.Structure):
    _fields_ = [("numel", i64), ("n_samples", i32)] + [(n, vp) for n in ("mu", "rho", "eps", "w", "eps_out")] + \
               [("seed", u64), ("stream_id", u64), ("sample0", i64)]


class BemBayesPointwiseParams(C.Structure):
    _fields_ = [(n, i32) for n in ("n_samples", "batch", "cin", "cout")] + [("P", i64)] + \
               [(n, vp) for n in ("x", "w", "mu", "rho", "eps", "bias", "out", "sigma", "ln_gamma", "ln_beta")] + \
               [("ln_eps", C.c_float), ("force_simt", i32), ("x_img_stride", i64), ("sample_interleave", i32),
                ("workspace", vp), ("workspace_bytes", i64), ("residual", vp), ("prelu_slope", vp), ("prelu_n", i32), ("prepacked", i32)]


class BemBayesDepthwiseParams(C.Structure):
    _fields_ = ([(n, i32) for n in ("n_samples", "batch", "C", "H", "W", "K")] + [(n, vp) for n in ("x", "w", "bias", "out")]
                + [("act", i32)])


class BemConv3x3Params(C.Structure):
    _fields_ = [(n, i32) for n in ("batch", "cin", "cout", "H", "W")] + [(n, vp) for n in ("x", "w", "bias", "out")]


class BemBayesSampleBatchedParams(C.Structure):
    _fields_ = [("entries", vp), ("blocks", vp), ("n_blocks", i32), ("seed", C.c_uint64), ("sample0", i64), ("sample0_dev", vp)]


# every symbol include/bem_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "bem_abi_version": (C.c_int, []),
    "bem_error_string": (C.c_char_p, [C.c_int]),
    "bem_scan_chunk_len": (C.c_int, [C.c_int]),
    "bem_scan_workspace_bytes": (i64, [C.c_int] * 5),
    "bem_scan_fwd": (C.c_int, [C.POINTER(BemScanFwdParams), vp]),
    "bem_scan_bwd": (C.c_int, [C.POINTER(BemScanBwdParams), vp]),
    "bem_cross_scan": (C.c_int, [C.POINTER(BemCsmParams), vp]),
    "bem_cross_merge": (C.c_int, [C.POINTER(BemCsmParams), vp]),
    "bem_ss2d_supported": (C.c_int, [C.c_int] * 2),
    "bem_ss2d_workspace_bytes": (i64, [C.c_int] * 6),
    "bem_ss2d_fwd": (C.c_int, [C.POINTER(BemSs2dFwdParams), vp]),
    "bem_bayes_sample": (C.c_int, [C.POINTER(BemBayesSampleParams), vp]),
    "bem_bayes_sample_batched": (C.c_int, [C.POINTER(BemBayesSampleBatchedParams), vp]),
    "bem_bayes_pointwise_workspace_bytes": (i64, [C.c_int] * 3),
    "bem_bayes_pointwise": (C.c_int, [C.POINTER(BemBayesPointwiseParams), vp]),
    "bem_bayes_pointwise_pack_table_bytes": (i64, [C.c_int]),
    "bem_bayes_pointwise_pack_table": (C.c_int, [C.POINTER(BemBayesPointwiseParams), C.c_int, vp, C.POINTER(i32)]),
    "bem_bayes_pointwise_pack_run": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "bem_bayes_depthwise": (C.c_int, [C.POINTER(BemBayesDepthwiseParams), vp]),
    "bem_conv3x3": (C.c_int, [C.POINTER(BemConv3x3Params), vp]),
    "bem_select_best": (C.c_int, [vp, i32, i32, vp, vp, vp]),
    "bem_niqe_mscn": (C.c_int, [vp, vp, vp, i32, i32, i32, vp]),
    "bem_niqe_block_stats": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "bem_layernorm2d_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, C.c_int64, C.c_float, vp]),
    "bem_layernorm2d_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, C.c_int64, vp]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python bayesian-enhancement-model_b200/build.py` "
        "(nvcc, sm_100a). bem_b200 has no fallback implementation.")
lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)      # AttributeError here = library/header mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args
if lib.bem_abi_version() != ABI_VERSION:
    raise ImportError(f"libbem_b200.so ABI {lib.bem_abi_version()} != binding ABI {ABI_VERSION}: rebuild the library")

_DT = {torch.float32: BEM_F32, torch.float16: BEM_F16, torch.bfloat16: BEM_BF16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise RuntimeError(f"bem_b200: unsupported dtype {dt} (float32 / float16 / bfloat16 only)")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(code: int, what: str):
    if code != 0:
        raise RuntimeError(f"bem_b200.{what} failed: {lib.bem_error_string(code).decode()} (code {code})")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("bem_b200 operators run on CUDA tensors only (sm_100a kernels; there is no CPU fallback)")


# ---------------------------------------------------------------------------------------------------------------------
# cache generation. The inference paths cache tensors derived from parameters (sigma = log1p(exp(rho)), packed weight tiles
# of constant-weight 1x1 layers, -exp(A_logs), split fusion weights) and CUDA graphs with those addresses baked in. Their
# keys hold (data_ptr, tensor._version) of the sources — but writes through `.data` (the reference's EMA update,
# basicsr/models/base_model.py:84 `p.data.mul_(decay).add_(...)`, init_parameters(), tools.py `.data.copy_`) do NOT bump
# `_version`. Every key therefore also holds this process-wide generation, which module.train() / .eval(),
# load_state_dict and .to() / .cuda() (nn.Module._apply) of this package's modules advance, and which callers who write
# parameters through `.data` must advance themselves: bem_b200.invalidate_caches().
# ---------------------------------------------------------------------------------------------------------------------
_cache_generation = [0]


def invalidate_caches():
    """drop every derived-weight cache and captured graph of this package at its next use (call after writing parameters
    through `.data`, e.g. an EMA update)"""
    _cache_generation[0] += 1


def cache_generation() -> int:
    return _cache_generation[0]


class InvalidatesCaches:
    """mixin for nn.Modules whose parameters feed the caches above"""

    def train(self, mode: bool = True):
        invalidate_caches()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        invalidate_caches()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        invalidate_caches()
        return super()._load_from_state_dict(*args, **kwargs)


_workspaces: dict = {}


def workspace(device: torch.device, nbytes: int, kind: str = "scan") -> torch.Tensor:
    """Scratch (scan look-back descriptors / packed weight tiles), cached per (kind, device, stream): launches on one
    stream are ordered, so they may share it; different streams get their own."""
    key = (kind, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)   # scan kernels need a zero-filled first use
        _workspaces[key] = ws
    return ws


def release_workspaces(stream_handle: int):
    """drop the scratch buffers cached for a stream that is no longer used (e.g. a warm-up side stream)"""
    for key in [k for k in _workspaces if k[2] == stream_handle]:
        del _workspaces[key]


def scan_error_word(device: torch.device) -> int:
    """Watchdog word of the last scan launch on the current stream (0 = clean). Synchronises; tests only."""
    key = ("scan", device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        return 0
    return int(ws[4:8].view(torch.int32).item())


# ---------------------------------------------------------------------------------------------------------------------
# launch accounting (bench.py): every C-ABI call goes through `launch`, which counts kernels and, when a profile is armed,
# brackets the call with CUDA events recorded on the launching stream.
# ---------------------------------------------------------------------------------------------------------------------
class _Profile:
    def __init__(self):
        self.launches = 0          # kernels of libbem_b200.so launched since the last reset
        self.armed = False
        self.records = []          # (name, key, start_event, end_event, algorithmic_bytes)

    def reset(self, armed=False):
        self.launches = 0
        self.armed = armed
        self.records = []

    def summary(self):
        """{name: dict(calls, ms, bytes)} — synchronises; call after the timed region."""
        torch.cuda.synchronize()
        out = {}
        for name, key, e0, e1, nbytes in self.records:
            d = out.setdefault(name, dict(calls=0, ms=0.0, bytes=0, by_key={}))
            ms = e0.elapsed_time(e1)
            d["calls"] += 1
            d["ms"] += ms
            d["bytes"] += nbytes
            k = d["by_key"].setdefault(str(key), dict(calls=0, ms=0.0, bytes=0))
            k["calls"] += 1
            k["ms"] += ms
            k["bytes"] += nbytes
        return out


profile = _Profile()


def launch(name, fn, params, device, key=None, nbytes=0, kernels=1):
    """run one C-ABI entry point on `device`'s current stream, check its return code, account for it"""
    with torch.cuda.device(device):
        st = stream_ptr(device)
        if profile.armed:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            code = fn(C.byref(params), st)
            e1.record()
            profile.records.append((name, key, e0, e1, nbytes))
        else:
            code = fn(C.byref(params), st)
    profile.launches += kernels
    check(code, name)
