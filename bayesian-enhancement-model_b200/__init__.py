"""bem_b200 — B200-native (sm_100a) implementation of the Bayesian-Enhancement-Model hot path.

    selective scan      bem_b200.selective_scan_fn / SelectiveScanCuda / selective_scan_cuda_oflex / build_selective_scan_fn
    traversal           bem_b200.cross_scan_fn / cross_merge_fn
    SS2D core           bem_b200.ss2d_core, bem_b200.ss2d_scan (fused operator), bem_b200.ss2d_fwd (C-ABI bem_ss2d_fwd), bem_b200.SS2D
    Bayesian layers     bem_b200.bayesian.{Conv2d,Linear2d,Linear}Reparameterization, convert2bnn*, set_prediction_type, ...
    MC inference        bem_b200.mc.{MCSampler, mc_infer, select_best}; bem_b200.NiqeScorer (device-resident no-reference score)
    stage-1 network     bem_b200.network.{Network, build_model, build_bayesian_model}
    reference patching  bem_b200.patch.install(...) / uninstall()
    training            bem_b200.GraphedTrainStep (forward + loss + backward + optimizer step as one CUDA-graph replay)
    after `.data` writes  bem_b200.invalidate_caches()  (derived-weight caches and captured graphs key on tensor versions,
                        which in-place writes through `.data` — e.g. an EMA update — do not advance)

All operators call libbem_b200.so (include/bem_b200.h) through ctypes; importing this package without the built library
raises ImportError — there is no fallback path.
"""
from . import _lib  # noqa: F401  (loads libbem_b200.so, raises if it is missing)
from ._lib import invalidate_caches  # noqa: F401
from . import bayesian, graphed, layernorm, mc, network, niqe, patch  # noqa: F401
from .layernorm import layer_norm_2d  # noqa: F401
from .graphed import GraphedTrainStep  # noqa: F401
from .niqe import NiqeScorer  # noqa: F401
from .csm import CrossMergeF, CrossScanF, cross_merge_fn, cross_scan_fn  # noqa: F401
from .selective_scan import (SelectiveScanCuda, build_selective_scan_fn, chunk_len, selective_scan_cuda_oflex,  # noqa: F401
                             selective_scan_fn, selective_scan_fn_test_api)
from .ss2d import SS2D, LayerNorm2d, Linear2d, ss2d_core, ss2d_fwd, ss2d_scan  # noqa: F401

__version__ = "0.1.0"
