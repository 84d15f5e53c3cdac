"""Row-sequential scan kernels (scan_rows.cu) against the look-back kernels at the multi-state shapes of BASELINE configs[0] / [3]:
CUDA events over graph replays of rotating input sets (working set > L2). `python tools/run_scan_rows.py`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bem_b200  # noqa: E402

ext = bem_b200.selective_scan_cuda_oflex
dev = torch.device("cuda")


def bench(fn, n_sets, reps=5):
    for i in range(n_sets):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(n_sets):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for i in range(n_sets):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n_sets)
    return sorted(ts)[len(ts) // 2]


def case(name, B, KD, N, G, L, dtype, bwd=True):
    per_set = (3 * B * KD * L + 2 * B * G * N * L) * torch.finfo(dtype).bits // 8
    n_sets = max(2, min(16, int(300e6 // per_set) + 1))
    sets = []
    for i in range(n_sets):
        torch.manual_seed(i)
        sets.append(dict(u=torch.randn(B, KD, L, device=dev, dtype=dtype), delta=(0.5 * torch.rand(B, KD, L, device=dev)).to(dtype),
                         A=-0.5 * torch.rand(KD, N, device=dev), B=torch.randn(B, G, N, L, device=dev, dtype=dtype),
                         C=torch.randn(B, G, N, L, device=dev, dtype=dtype), D=torch.randn(KD, device=dev),
                         bias=0.5 * torch.rand(KD, device=dev), dout=torch.randn(B, KD, L, device=dev, dtype=dtype)))
    out = {}
    for rows in ("0", "1"):
        os.environ["BEM_SCAN_ROWS"] = rows
        xs = [None] * n_sets

        def f(i):
            s = sets[i]
            o, x = ext.fwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], True, 1, True)
            xs[i] = x

        def b(i):
            s = sets[i]
            ext.bwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], s["dout"], xs[i], True, 1)

        tf = bench(f, n_sets)
        tb = bench(b, n_sets) if bwd else float("nan")
        out[rows] = (tf, tb)
    print(f"{name}: look-back fwd {out['0'][0]:.1f} us bwd {out['0'][1]:.1f} us | rows fwd {out['1'][0]:.1f} us bwd {out['1'][1]:.1f} us", flush=True)


case("c1  B1 KD384 N16 L4096 fp32", 1, 384, 16, 4, 4096, torch.float32)
case("vmamba-t stage1 B8 KD384 N16 L3136 fp32", 8, 384, 16, 4, 3136, torch.float32)
case("N4  B1 KD384 N4 L4096 fp32", 1, 384, 4, 4, 4096, torch.float32)
case("hd  B1 KD384 N16 L129600 bf16", 1, 384, 16, 4, 129600, torch.bfloat16, bwd=True)
