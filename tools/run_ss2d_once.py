"""SS2D core (x_proj output -> y) at the three BEM levels of the 600x400 workload: bem_ss2d_fwd timed with CUDA events over
graph replays on rotating inputs, plus parity of the traversal-aware kernels against the explicit operator chain
(cross_scan -> dt_proj -> selective_scan -> cross_merge, each tested on its own against the oracle).
    python tools/run_ss2d_once.py            # BEM_SS2D_COMPOSED=1 for the composed form"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bem_b200  # noqa: E402
from bem_b200 import ss2d  # noqa: E402
from bem_b200.bayesian import functional as BF  # noqa: E402

dev = torch.device("cuda")
LEVELS = [(40, 400, 600, 3), (80, 200, 300, 5)] + ([] if os.environ.get("BEM_SS2D_COMPOSED") else [(160, 100, 150, 10)])


def chain(x, z, dtw, A, Ds, bias, R):
    B, D, H, W = x.shape
    L = H * W
    xs = bem_b200.cross_scan_fn(x, True, True, False, 0).view(B, -1, L)
    xd = bem_b200.cross_scan_fn(z.view(B, 4, R + 2, H, W), True, True, True, 0)
    dts, Bs, Cs = torch.split(xd, [R, 1, 1], dim=2)
    dts = BF.grouped_pointwise(dts, dtw.view(4, D, R)).reshape(B, -1, L)
    ys = bem_b200.selective_scan_fn(xs, dts, A, Bs, Cs, Ds, bias, True, True)
    return bem_b200.cross_merge_fn(ys.view(B, 4, D, H, W), True, True, False, 0)


for (D, H, W, R) in LEVELS:
    torch.manual_seed(D)
    sets = []
    n_sets = max(2, int(400e6 // (4 * H * W * (2 * D + 4 * (R + 2)))) + 1)
    for _ in range(n_sets):
        x = torch.randn(1, D, H, W, device=dev)
        z = torch.randn(1, 4 * (R + 2), H * W, device=dev) * 0.5
        sets.append((x, z))
    dtw = torch.randn(4 * D, R, device=dev) * 0.5
    A = -torch.rand(4 * D, 1, device=dev) - 0.5
    Ds = torch.randn(4 * D, device=dev)
    bias = torch.randn(4 * D, device=dev) * 0.5
    y = ss2d.ss2d_fwd(sets[0][0], sets[0][1], dtw, A, Ds, bias)
    ref = chain(sets[0][0], sets[0][1], dtw, A, Ds, bias, R)
    err = float((y - ref).abs().max() / ref.abs().max())
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for (x, z) in sets:
            ss2d.ss2d_fwd(x, z, dtw, A, Ds, bias)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for (x, z) in sets:
            ss2d.ss2d_fwd(x, z, dtw, A, Ds, bias)
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / len(sets))
    ts = sorted(ts[1:])
    nb = 4 * H * W * (2 * D + 4 * (R + 2))
    print(f"D={D} {H}x{W} R={R}: ss2d_fwd {1e3 * ts[len(ts) // 2]:.1f} us  ({nb / ts[len(ts) // 2] / 1e6:.0f} GB/s algorithmic, "
          f"supported={bem_b200._lib.lib.bem_ss2d_supported(1, R)})  nmax err vs operator chain {err:.2e}")
