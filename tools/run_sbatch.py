"""S-batched kernels (GPU box): per-image time of the Bayesian 1x1 / depthwise / SS2D-core launches of the stage-1 network at
600x400 when one launch carries S Monte-Carlo samples (S images, S weight sets) instead of one."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bem_b200  # noqa: E402
from bem_b200 import ss2d  # noqa: E402
from bem_b200.bayesian import functional as BF  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tools"))
from run_pointwise import timed  # noqa: E402

dev = torch.device("cuda")
shapes = [(40, 40, 240000, True), (40, 320, 240000, True), (160, 40, 240000, False), (80, 80, 60000, True), (80, 640, 60000, True),
          (320, 80, 60000, False), (160, 160, 15000, True), (160, 1280, 15000, True), (640, 160, 15000, False)]
for S in (1, 2, 4):
    print(f"--- S = {S}")
    for cin, cout, P, ln in shapes:
        x = torch.randn(S, cin, P, device=dev)
        w = torch.randn(S, cout, cin, device=dev) / cin ** 0.5
        b = torch.randn(S, cout, device=dev)
        lnp = (torch.ones(cin, device=dev), torch.zeros(cin, device=dev), 1e-5) if ln else None
        t = timed(lambda: BF.pointwise_conv(x, w, b, S, ln=lnp))
        print(f"1x1 {cin:4d}->{cout:5d} P {P:6d}: {t * 1e3 / S:7.1f} us per image  ({4 * P * (cin + cout) * S / t / 1e6:6.0f} GB/s)", flush=True)
    for C, H, W, act in ((40, 400, 600, "silu"), (320, 400, 600, "gelu_gate"), (80, 200, 300, "silu"), (640, 200, 300, "gelu_gate"),
                         (160, 100, 150, "silu"), (1280, 100, 150, "gelu_gate")):
        x = torch.randn(S, C, H, W, device=dev)
        w = torch.randn(S, C, 3, 3, device=dev)
        t = timed(lambda: BF.depthwise_conv3x3(x, w, None, S, act=act))
        print(f"dw  {C:4d} {H}x{W} {act}: {t * 1e3 / S:7.1f} us per image", flush=True)
    for (D, H, W, R) in ((40, 400, 600, 3), (80, 200, 300, 5), (160, 100, 150, 10)):
        x = torch.randn(S, D, H, W, device=dev)
        z = torch.randn(S, 4 * (R + 2), H * W, device=dev) * 0.5
        dtw = torch.randn(4 * D, R, device=dev) * 0.5
        A = -torch.rand(4 * D, 1, device=dev) - 0.5
        Ds = torch.randn(4 * D, device=dev)
        bias = torch.randn(4 * D, device=dev) * 0.5
        t = timed(lambda: ss2d.ss2d_fwd(x, z, dtw, A, Ds, bias))
        print(f"ss2d D={D} {H}x{W}: {t * 1e3 / S:7.1f} us per image", flush=True)
