"""Timing driver (GPU box): the Bayesian 1x1 layer shapes of the stage-1 network at 600x400, tcgen05 vs CUDA-core kernel."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bem_b200.bayesian import functional as BF  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    """device time of one call (its kernels captured in a CUDA graph: no host launch gaps), cold L2"""
    fn()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    fn = g.replay
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


shapes = [(40, 40, 240000, True), (40, 320, 240000, True), (160, 40, 240000, False), (80, 80, 60000, True), (80, 640, 60000, True),
          (320, 80, 60000, False), (160, 160, 15000, True), (160, 1280, 15000, True), (640, 160, 15000, False)]
for cin, cout, P, ln in shapes:
    x = torch.randn(1, cin, P, device=dev)
    mu = torch.randn(cout, cin, device=dev) / cin ** 0.5
    sig = torch.full_like(mu, 0.05)
    eps = torch.randn(1, cout, cin, device=dev)
    b = torch.randn(1, cout, device=dev)
    lnp = (torch.ones(cin, device=dev), torch.zeros(cin, device=dev), 1e-5) if ln else None
    BF.pointwise_conv_sampled(x, mu, sig, eps, b, 1, ln=lnp)
    t_tc = timed(lambda: BF.pointwise_conv_sampled(x, mu, sig, eps, b, 1, ln=lnp))
    t_simt = timed(lambda: BF.pointwise_conv_sampled(x, mu, sig, eps, b, 1, force_simt=True))
    nbytes = 4 * P * (cin + cout)
    flops = 2.0 * P * cin * cout
    print(f"cin {cin:4d} cout {cout:5d} P {P:6d} ln {int(ln)}: tcgen05 {t_tc * 1e3:7.1f} us {nbytes / t_tc / 1e6:7.0f} GB/s {flops / t_tc / 1e9:7.1f} TFLOP/s"
          f" | simt {t_simt * 1e3:7.1f} us {nbytes / t_simt / 1e6:7.0f} GB/s", flush=True)
