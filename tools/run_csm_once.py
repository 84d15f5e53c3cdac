"""Device time of the tiled cross scan / merge kernels at the BEM level-0/1 shapes (CUDA graph of 8 calls, CUDA events / 8)."""
import os, sys, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bem_b200

dev = torch.device("cuda")


def timed(fn, reps=8, iters=10):
    fn()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return statistics.median(ts)


for C, H, W in ((40, 400, 600), (80, 200, 300)):
    x = torch.randn(1, C, H, W, device=dev)
    ys = torch.randn(1, 4, C, H * W, device=dev)
    nbytes = 4 * C * H * W * 5
    t = timed(lambda: bem_b200.cross_scan_fn(x, in_channel_first=True, out_channel_first=True, scans=0))
    print(f"cross_scan  C{C} {H}x{W}: {t * 1e3:6.1f} us  {nbytes / t / 1e6:6.0f} GB/s  {nbytes / t / 1e6 / 6545:.2f} of peak")
    t = timed(lambda: bem_b200.cross_merge_fn(ys.view(1, 4, C, H, W), in_channel_first=True, out_channel_first=True, scans=0))
    print(f"cross_merge C{C} {H}x{W}: {t * 1e3:6.1f} us  {nbytes / t / 1e6:6.0f} GB/s  {nbytes / t / 1e6 / 6545:.2f} of peak")
