"""Profiling / timing driver (GPU box): the level-0 BEM-I scan (B1 KD160 N1 L240000 fp32) and friends, launched straight
through the C-ABI with pre-allocated buffers so that host overhead does not pace the loop. Per-launch CUDA events."""
import ctypes as ct
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bem_b200  # noqa: E402
from bem_b200 import _lib  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "L0"
B, KD, N, G, L = {"L0": (1, 160, 1, 4, 240000), "L1": (1, 320, 1, 4, 60000), "L2": (1, 640, 1, 4, 15000),
                  "T0": (8, 160, 1, 4, 4096), "T1": (8, 320, 1, 4, 1024), "T2": (8, 640, 1, 4, 256),
                  "C1": (1, 384, 16, 4, 4096), "HD": (1, 640, 1, 4, 129600)}[shape]
dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
dev = torch.device("cuda")
torch.manual_seed(0)
u = torch.randn(B, KD, L, device=dev).to(dt)
delta = (0.5 * torch.rand(B, KD, L, device=dev)).to(dt)
A = -0.5 * torch.rand(KD, N, device=dev)
Bm = torch.randn(B, G, N, L, device=dev).to(dt)
Cm = torch.randn(B, G, N, L, device=dev).to(dt)
D = torch.randn(KD, device=dev)
bias = 0.5 * torch.rand(KD, device=dev)
dout = torch.randn(B, KD, L, device=dev)
ext = bem_b200.selective_scan_cuda_oflex
out, x = ext.fwd(u, delta, A, Bm, Cm, D, bias, True, 1, True)   # warm-up + allocates the workspace
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2: the inputs are evicted between launches


def timed(fn, iters=20):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


es = u.element_size()
fwd_bytes = B * KD * L * (2 * es + 4) + 2 * B * G * N * L * es
med, best = timed(lambda: ext.fwd(u, delta, A, Bm, Cm, D, bias, True, 1, True))
print(f"{shape} fwd  median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {fwd_bytes / med / 1e6:7.1f} GB/s (median)")
if "--bwd" in sys.argv:
    bwd_bytes = B * KD * L * (4 * es + 4) + 4 * B * G * N * L * es
    med, best = timed(lambda: ext.bwd(u, delta, A, Bm, Cm, D, bias, dout, x, True, 1))
    print(f"{shape} bwd  median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {bwd_bytes / med / 1e6:7.1f} GB/s (median, incl. grad zero-fill)")
