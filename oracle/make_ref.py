#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the GPU box.  TEST / BENCH INFRASTRUCTURE — never imported by the product.

    python oracle/make_ref.py            # called from __graft_entry__.build() when /root/reference is present

/root/reference exists only in the build container; the GPU box gets a snapshot of this repository. Everything the
`-m gpu` parity tests and bench.py's reference columns need from the reference is therefore placed under the git-ignored
(not gpurun-ignored) directory oracle/_ref/ by this recipe — nothing from the reference enters the repository's history:

  oracle/_ref/reference/   verbatim copies of the reference's own Python for the hot path and its callers
                           (basicsr/{vmamba,bayesian,archs,metrics,utils,ops,models,losses,QD/*.py + checkpoints},
                           kernels/selective_scan, Enhancement, Options), read through oracle/ref_loader.py
  oracle/_ref/selective_scan_cuda_oflex.so
                           the reference's CUDA extension (kernels/selective_scan/csrc/selective_scan/cusoflex/*, with the
                           flags of kernels/selective_scan/setup.py:114-135) compiled unmodified for sm_100a from the sources
                           where they lie — the GPU oracle of the scan and the kernel to beat

`pip install --target baseline/_ref /root/reference` (the generic contract) is not usable: setup.py:131-150 builds three
extensions from basicsr/models/ops/*, a directory that does not exist in the repository (the sources are in basicsr/ops),
and the package would not contain kernels/selective_scan at all.
"""
from __future__ import annotations

import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PY_DST = os.path.join(DST, "reference")
EXT_NAME = "selective_scan_cuda_oflex"
EXT_SO = os.path.join(DST, EXT_NAME + ".so")

# (subtree, ignore patterns)
TREES = [
    ("basicsr", ("evaluation_results", "*.log", "__pycache__", "data", "*.pyc", "test_metrics")),
    ("kernels/selective_scan", ("__pycache__", "*.pyc", "build", "*.egg-info")),
    ("Enhancement", ("__pycache__", "*.pyc")),
    ("Options", ()),
]


def stage_python(verbose=True) -> str:
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} is not present: the staged copy can only be made in the build container")
    for sub, ignore in TREES:
        src, dst = os.path.join(REF, sub), os.path.join(PY_DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns(*ignore))
    # basicsr/data is needed only as an importable (empty) package by a few `from basicsr.data...` lines we never execute
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(PY_DST))
        print(f"staged {n} reference files under {PY_DST}")
    return PY_DST


def build_oflex(verbose=True) -> str:
    """torch.utils.cpp_extension build of the reference's oflex extension for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(REF, "kernels/selective_scan/csrc/selective_scan")
    sources = [os.path.join(src, "cusoflex", f) for f in
               ("selective_scan_oflex.cpp", "selective_scan_core_fwd.cu", "selective_scan_core_bwd.cu")]
    if os.path.exists(EXT_SO) and all(os.path.getmtime(EXT_SO) >= os.path.getmtime(s) for s in sources):
        return EXT_SO
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    build_dir = os.path.join(DST, "_build_oflex")
    os.makedirs(build_dir, exist_ok=True)
    load(name=EXT_NAME, build_directory=build_dir, extra_include_paths=[src], sources=sources, verbose=verbose,
         extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                            "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                            "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                            "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo"])
    shutil.copy2(os.path.join(build_dir, EXT_NAME + ".so"), EXT_SO)
    shutil.rmtree(build_dir, ignore_errors=True)
    return EXT_SO


def main(verbose=True):
    os.makedirs(DST, exist_ok=True)
    stage_python(verbose)
    so = build_oflex(verbose)
    if verbose:
        print("reference extension:", so)


if __name__ == "__main__":
    main(verbose="-q" not in sys.argv)
