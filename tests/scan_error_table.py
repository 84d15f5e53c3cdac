"""Diagnostic (GPU box; `python tests/scan_error_table.py`): normalised max error of the sm_100a scan vs the fp64 oracle, next to
the error of the reference-order fp32 oracle, for a list of shapes. Test infrastructure (it uses the oracle), not part of the
product; reads only this repo. Not collected by pytest (no test_ prefix)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bem_b200  # noqa: E402
import oracle  # noqa: E402
from conftest import nmax_err  # noqa: E402
from test_scan_gpu import make_inputs  # noqa: E402

CASES = [((1, 384, 16, 4, 4096), 1), ((2, 96, 16, 4, 4096), 4112), ((2, 48, 1, 2, 1024), 1025), ((8, 160, 1, 4, 4096), 1),
         ((1, 640, 1, 4, 15000), 1), ((1, 160, 1, 4, 60000), 3), ((2, 96, 4, 2, 4096), 5), ((1, 96, 16, 4, 384), 6),
         ((1, 96, 16, 4, 768), 6)]
for cfg, seed in CASES:
    inp = make_inputs(*cfg, torch.float32, seed=seed)
    leaves = {k: (v.clone().requires_grad_() if k != "dout" else v) for k, v in inp.items()}
    out = bem_b200.selective_scan_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                                     leaves["delta_bias"], True, True)
    out.backward(inp["dout"])
    o = oracle.selective_scan_oracle_f64(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["delta_bias"], True,
                                         dout=inp["dout"])
    o32 = oracle.selective_scan_oracle(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None, inp["delta_bias"], True)
    e = {"out": nmax_err(out.detach().cpu().numpy(), o["out"])}
    for gk, lk in (("du", "u"), ("ddelta", "delta"), ("dA", "A"), ("dB", "B"), ("dC", "C"), ("dD", "D"), ("ddelta_bias", "delta_bias")):
        e[gk] = nmax_err(leaves[lk].grad.cpu().numpy(), o[gk])
    print(cfg, "ref32 out %.1e |" % nmax_err(o32, o["out"]), " ".join(f"{k} {v:.1e}" for k, v in e.items()), flush=True)
