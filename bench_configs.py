"""bench_configs.py — the measurement legs bench.py dispatches to beside the headline (BASELINE.json configs[1]):

  reference arms   the UNMODIFIED reference staged under oracle/_ref (oracle/ref_loader.py): its pure-PyTorch CPU path for
                   `--impl reference`, and the reference's own GPU path (CUDA extension recompiled for sm_100a + Triton cross
                   scan / merge + eager Bayesian layers) as the `reference_gpu` column beside our numbers
  --config c1      BASELINE configs[0]: scan fwd+bwd at B1 K4 D96 N16 L64x64 fp32
  --config hd      BASELINE configs[3]: long-sequence scans of a 1920x1080 image (L = 129600 per direction), bf16 in, fp32 out
  --config train   BASELINE configs[4]: fwd+bwd+optimizer step of the 18 VSSBlocks of DecompDualBranch2DDWavelet_4 on 8x128x128
                   patches per rank, DDP over the ranks

This file is measurement harness: it may execute oracle/ (reference arms and cpu_baseline legs only), the product never does.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ref():
    from oracle import ref_loader as R
    return R


# ---------------------------------------------------------------------------------------------------------------------
# timing helpers
# ---------------------------------------------------------------------------------------------------------------------
class L2Flush:
    """a buffer larger than the 126 MB L2, rewritten between timed launches whose working set would otherwise stay cached"""

    def __init__(self, dev, mb=384):
        import torch
        self.buf = torch.empty(mb << 20, dtype=torch.uint8, device=dev)

    def __call__(self):
        self.buf.zero_()


def time_calls(fn, dev, reps=20, warmup=5, flush=None):
    """mean / min ms of `fn()` over `reps` calls, CUDA events on the current stream around each call, optional L2 flush before
    each (outside the events)"""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return sum(ts) / len(ts), ts[0], ts[len(ts) // 2]


def time_graph_rotating(make_fn, n_sets, dev, per_graph=None, reps=6):
    """`make_fn(i)` returns the launch closure on input set i; the n_sets sets together exceed the L2, so each launch inside
    the captured graph starts with none of its inputs cached. Returns mean ms per launch over the replays."""
    import torch
    fns = [make_fn(i) for i in range(n_sets)]
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    ts = []
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1) / len(fns))
    ts = sorted(ts[1:])
    return sum(ts) / len(ts), ts[len(ts) // 2]


def scan_inputs(Bn, KD, G, N, L, dtype, dev, seed=1, n_sets=1):
    """synthetic inputs of test_selective_scan.py:406-441 (A = -0.5 U, B/C/u ~ N, delta = 0.5 U, bias = 0.5 U, D ~ N)"""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    sets = []
    for _ in range(n_sets):
        u = torch.randn(Bn, KD, L, generator=g).to(dev, dtype)
        delta = (0.5 * torch.rand(Bn, KD, L, generator=g)).to(dev, dtype)
        A = (-0.5 * torch.rand(KD, N, generator=g)).to(dev)
        Bm = torch.randn(Bn, G, N, L, generator=g).to(dev, dtype)
        Cm = torch.randn(Bn, G, N, L, generator=g).to(dev, dtype)
        D = torch.randn(KD, generator=g).to(dev)
        bias = (0.5 * torch.rand(KD, generator=g)).to(dev)
        dout = torch.randn(Bn, KD, L, generator=g).to(dev)
        sets.append(dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=D, bias=bias, dout=dout))
    return sets


def scan_bytes(Bn, KD, G, N, L, s, s_o=4):
    """SURVEY 8(d), boundary form"""
    fwd = Bn * KD * L * (2 * s + s_o) + 2 * Bn * G * N * L * s
    bwd = Bn * KD * L * (4 * s + s_o) + 4 * Bn * G * N * L * s
    return fwd, bwd


# ---------------------------------------------------------------------------------------------------------------------
# the reference on the GPU: the kernel(s) to beat (column 2)
# ---------------------------------------------------------------------------------------------------------------------
def reference_gpu_kernels(dev, H=400, W=600):
    """Reference GPU path per kernel at the level-0 shapes of the 600x400 workload: oflex extension fwd / bwd (B1 KD160 N1
    L240000 fp32), Triton cross_scan / cross_merge (1, 40, 400, 600), eager Bayesian 1x1 (40 -> 320) = sample + F.conv2d."""
    import torch
    R = _ref()
    if not R.available():
        return {"unavailable": R.why_unavailable()}
    out = {}
    ext = R.oflex_ext()
    L = H * W
    s = scan_inputs(1, 160, 4, 1, L, torch.float32, dev)[0]
    fwd = lambda: ext.fwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], True, 1, True)
    o, x = fwd()[:2]
    bwd = lambda: ext.bwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], s["dout"], x, True, 1)
    fb, bb = scan_bytes(1, 160, 4, 1, L, 4)
    for name, fn, nb in (("scan_fwd_L0", fwd, fb), ("scan_bwd_L0", bwd, bb)):
        ms = time_calls(fn, dev, reps=10, warmup=3)[2]
        out[name] = {"ms": ms, "GBps": nb / ms / 1e6, "what": "selective_scan_cuda_oflex (reference, unmodified, sm_100a)"}
    csm = R.csm_triton()
    x4 = torch.randn(1, 40, H, W, device=dev)
    ys = torch.randn(1, 4, 40, H, W, device=dev)
    ms = time_calls(lambda: csm.cross_scan_fn(x4, True, True, False, 0), dev, reps=10, warmup=3)[2]
    out["cross_scan_L0"] = {"ms": ms, "GBps": 5 * x4.numel() * 4 / ms / 1e6, "what": "csm_triton.cross_scan_fn (Triton)"}
    ms = time_calls(lambda: csm.cross_merge_fn(ys, True, True, False, 0), dev, reps=10, warmup=3)[2]
    out["cross_merge_L0"] = {"ms": ms, "GBps": 5 * x4.numel() * 4 / ms / 1e6, "what": "csm_triton.cross_merge_fn (Triton)"}
    refb = R.bayesian()
    layer = refb.Conv2dReparameterization(40, 320, 1, bias=True).to(dev).eval()
    ln = torch.nn.LayerNorm(40).to(dev)
    xp = torch.randn(1, 40, H, W, device=dev)
    with torch.no_grad():
        def eager():
            y = torch.nn.functional.layer_norm(xp.permute(0, 2, 3, 1), (40,), ln.weight, ln.bias, 1e-5).permute(0, 3, 1, 2)
            return layer(y)
        ms = time_calls(eager, dev, reps=10, warmup=3)[2]
    out["bayes_1x1_40_320_ln_L0"] = {"ms": ms, "GBps": 4 * L * 360 / ms / 1e6,
                                     "what": "LayerNorm2d + bayesian.Conv2dReparameterization (eager: normal_, softplus, F.conv2d)"}
    return out


def reference_gpu_network(dev, steps=5, H=400, W=600):
    """the reference's stage-1 Bayesian `Network`, unpatched, on the GPU: reference CUDA extension + Triton cross scan / merge +
    eager Bayesian layers, one MC sample per step (what Enhancement/eval.py:199-211 runs per sample)"""
    import torch
    R = _ref()
    if not R.available():
        return {"unavailable": R.why_unavailable()}
    unet = R.unet_arch(True)
    refb = R.bayesian()
    torch.manual_seed(0)
    net = unet.build_model()
    refb.convert2bnn_selective(net, {"sigma_init": 0.05, "decay": 0.998, "pretrain": False})
    net = net.to(dev).eval()
    refb.set_prediction_type(net, deterministic=False)
    img = torch.rand(1, 3, H, W, device=dev)
    with torch.no_grad():
        for _ in range(2):
            net(img)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            y = torch.clamp(net(img)[-1], 0, 1)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    del net
    return {"images_per_s": 1e3 / ms, "ms_per_step": ms, "steps": steps,
            "what": "reference UNet_arch.Network + convert2bnn_selective, unpatched, eager: oflex extension (sm_100a) + Triton csm + F.conv2d"}


# ---------------------------------------------------------------------------------------------------------------------
# the reference on the CPU (--impl reference, cpu_baseline kind "reference")
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_network(steps, warmup, budget_s, H=400, W=600):
    """The reference's own pure-PyTorch CPU path for the headline workload: UNet_arch.Network + convert2bnn_selective, scans by
    csms6s.selective_scan_torch (the Python loop over L, csms6s.py:29-72), traversal by the torch cross_scan / cross_merge,
    Bayesian layers eager. Each step = one MC sample on a top crop of the 600x400 image sized so that warmup + steps fit the
    budget (scan cost is linear in the pixel count); throughput is scaled to whole images by the pixel fraction."""
    import torch
    R = _ref()
    if not R.available():
        raise RuntimeError(R.why_unavailable())
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet = R.unet_arch(False)          # no extension registered: csms6s falls to selective_scan_torch
    vm = R.vmamba(False)
    import importlib
    cs = importlib.import_module("basicsr.vmamba.models.csms6s")
    assert not (cs.WITH_SELECTIVESCAN_OFLEX or cs.WITH_SELECTIVESCAN_CORE or cs.WITH_SELECTIVESCAN_MAMBA), "CPU arm must run the torch scan"
    refb = R.bayesian()
    torch.manual_seed(0)
    net = unet.build_model()
    refb.convert2bnn_selective(net, {"sigma_init": 0.05, "decay": 0.998, "pretrain": False})
    net.eval()
    refb.set_prediction_type(net, deterministic=False)
    img = torch.rand(1, 3, H, W)
    with torch.no_grad():
        t0 = time.perf_counter()
        net(img[:, :, :16])
        per_row = (time.perf_counter() - t0) / 16
        n = max(1, steps + warmup)
        rows = int(min(H, max(16, (budget_s / n) / per_row)) // 16 * 16)
        crop = img[:, :, :rows]
        for _ in range(warmup):
            net(crop)
        t0 = time.perf_counter()
        for _ in range(steps):
            torch.clamp(net(crop)[-1], 0, 1)
        dt = time.perf_counter() - t0
    frac = rows / H
    return dict(value=steps * frac / dt, unit="images/s", cores=cores, kind="reference",
                sample=f"{steps} MC samples of the top {rows}x{W} crop ({frac:.2f} image each) through the unmodified reference "
                       f"Network (pure PyTorch: selective_scan_torch loop, torch cross scan/merge, eager Bayesian layers), {cores} threads",
                ms_per_step=1e3 * dt / max(steps, 1))


def cpu_reference_scan(Bn, KD, G, N, L, dtype_name, with_bwd, budget_s=20.0):
    """selective_scan_ref (kernels/selective_scan/test_selective_scan.py:168-234) on the host cores, on a bounded prefix of the
    sequence: forward cost is linear in L; the autograd backward through the stacked per-step outputs is quadratic in memory
    traffic, so fwd+bwd is timed at a shorter prefix and both are scaled to the full length."""
    import torch
    R = _ref()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sref = R.selective_scan_ref()
    dt = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype_name]
    Lf = min(L, 2048)
    s = scan_inputs(Bn, KD, G, N, Lf, dt, "cpu")[0]
    t0 = time.perf_counter()
    sref(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], None, s["bias"], True)
    t_fwd = (time.perf_counter() - t0) * (L / Lf)
    t_bwd = None
    sample = f"selective_scan_ref forward on the first {Lf} of {L} positions, scaled linearly"
    if with_bwd:
        Lb = min(L, 512)
        s = scan_inputs(Bn, KD, G, N, Lb, dt, "cpu")[0]
        leaves = [s[k].clone().requires_grad_() for k in ("u", "delta", "A", "B", "C", "D", "bias")]
        t0 = time.perf_counter()
        o = sref(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], None, leaves[6], True)
        t1 = time.perf_counter()
        o.backward(s["dout"].to(o.dtype))
        t_b = time.perf_counter() - t1
        t_bwd = t_b * (L / Lb) ** 2
        sample += f"; autograd backward on the first {Lb} positions, scaled quadratically (BASELINE.md section 3: 254 s measured at full length on 8 cores)"
    total = t_fwd + (t_bwd or 0.0)
    return dict(value=1.0 / total, unit="scans/s", cores=cores, kind="reference", sample=sample,
                fwd_s=t_fwd, bwd_s=t_bwd, ms_per_step=1e3 * total)


# ---------------------------------------------------------------------------------------------------------------------
# --config c1 / hd : scan-only configurations
# ---------------------------------------------------------------------------------------------------------------------
SCAN_CONFIGS = {
    # name: (workload text, [(tag, Bn, KD, G, N, L, dtype)], with_bwd)
    "c1": ("BASELINE configs[0]: SS2D selective scan fwd+bwd, B=1, K=4 directions, D=96, N=16, L=64x64, fp32",
           [("c1", 1, 384, 4, 16, 4096, "f32")], True),
    "hd": ("BASELINE configs[3]: SS2D long-sequence scans of a 1920x1080 image at the 270x480 level (L=129600 per direction), bf16 in / fp32 out",
           [("hd_bem_kd640_n1", 1, 640, 4, 1, 129600, "bf16"), ("hd_kd384_n16", 1, 384, 4, 16, 129600, "bf16")], False),
}


def _ours_scan_closures(s, with_bwd):
    from bem_b200 import selective_scan as ss
    fwd = lambda: ss.fwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], True, 1, True)
    x = fwd()[1]
    bwd = (lambda: ss.bwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], s["dout"], x, True, 1)) if with_bwd else None
    return fwd, bwd


def _ref_scan_closures(s, with_bwd):
    ext = _ref().oflex_ext()
    fwd = lambda: ext.fwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], True, 1, True)
    x = fwd()[1]
    bwd = (lambda: ext.bwd(s["u"], s["delta"], s["A"], s["B"], s["C"], s["D"], s["bias"], s["dout"], x, True, 1)) if with_bwd else None
    return fwd, bwd


def run_scan_config(name, args, rank, world, dev):
    """one scan configuration: a step = fwd (+ bwd) of every shape of the configuration on this rank's replica (the scan does
    not shard: replicas only). value = steps/s summed over ranks."""
    import torch
    import torch.distributed as dist
    workload, shapes, with_bwd = SCAN_CONFIGS[name]
    peak, peak_src = _peaks()
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}
    # rotating input sets so that consecutive launches never find their inputs in the 126 MB L2
    per_shape = []
    for (tag, Bn, KD, G, N, L, dn) in shapes:
        fb, bb = scan_bytes(Bn, KD, G, N, L, 4 if dn == "f32" else 2)
        n_sets = max(2, min(24, int(400e6 // max(fb, 1)) + 1))
        sets = scan_inputs(Bn, KD, G, N, L, tdt[dn], dev, n_sets=n_sets)
        per_shape.append((tag, (Bn, KD, G, N, L, dn), fb, bb, sets))

    def one_step(i):
        for tag, shp, fb, bb, sets in per_shape:
            s = sets[i % len(sets)]
            f, b = closures[(tag, i % len(sets))]
            f()
            if b is not None:
                b()

    closures = {}
    for tag, shp, fb, bb, sets in per_shape:
        for j, s in enumerate(sets):
            closures[(tag, j)] = _ours_scan_closures(s, with_bwd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        one_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # e2e: the same through the public operator with HOST buffers (pinned): H2D of u/delta/B/C (+dout), op, D2H of out (+grads)
    tag, shp, fb, bb, sets = per_shape[0]
    s0 = sets[0]
    host = {k: v.cpu().pin_memory() for k, v in s0.items()}
    from bem_b200 import selective_scan as ss

    def e2e_step():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out, x = ss.fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["bias"], True, 1, True)
        res = [out]
        if with_bwd:
            res += [g for g in ss.bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["bias"], d["dout"], x, True, 1) if g is not None]
        outs = [r.to("cpu", non_blocking=True) for r in res]
        torch.cuda.synchronize(dev)
        return outs
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        outs = e2e_step()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / n_e2e
    h2d = sum(v.numel() * v.element_size() for k, v in host.items() if with_bwd or k != "dout")
    d2h = sum(o.numel() * o.element_size() for o in outs)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return None
    # per-kernel: ours and the reference extension, per call (L2 flushed) and graph-batched over rotating inputs
    flush = L2Flush(dev)
    kern = {}
    R = _ref()
    for tag, (Bn, KD, G, N, L, dn), fb, bb, sets in per_shape:
        rec = {"shape": dict(B=Bn, KD=KD, G=G, N=N, L=L, dtype=dn), "fwd_bytes": fb, "bwd_bytes": bb if with_bwd else None}
        f, b = closures[(tag, 0)]
        for nm, fn, nb in (("fwd", f, fb), ("bwd", b, bb)):
            if fn is None:
                continue
            per_call = time_calls(fn, dev, reps=20, warmup=3, flush=flush)[2]
            which = 0 if nm == "fwd" else 1
            batched = time_graph_rotating(lambda i: closures[(tag, i)][which], len(sets), dev)[1]
            rec[nm] = {"per_call_ms": per_call, "graph_batched_ms": batched, "GBps": nb / batched / 1e6, "frac": nb / batched / 1e6 / peak}
        if R.available():
            rc = {j: _ref_scan_closures(s, with_bwd) for j, s in enumerate(sets)}
            for nm, which, nb in (("fwd", 0, fb), ("bwd", 1, bb)):
                if rc[0][which] is None:
                    continue
                per_call = time_calls(rc[0][which], dev, reps=10, warmup=3, flush=flush)[2]
                batched = time_graph_rotating(lambda i: rc[i][which], len(sets), dev)[1]
                rec["reference_" + nm] = {"per_call_ms": per_call, "graph_batched_ms": batched, "GBps": nb / batched / 1e6,
                                          "frac": nb / batched / 1e6 / peak}
        kern[tag] = rec
    tag0 = per_shape[0][0]
    dom = "bwd" if with_bwd else "fwd"
    nb0 = per_shape[0][3] if with_bwd else per_shape[0][2]
    r0 = kern[tag0][dom]
    roof = {"bound": "hbm", "achieved": r0["GBps"], "peak": peak, "unit": "GB/s", "frac": r0["frac"], "traffic": None,
            "kernel": f"scan {dom} {kern[tag0]['shape']}", "peak_source": peak_src, "bytes_per_launch": float(nb0),
            "ms_per_launch": r0["graph_batched_ms"],
            "timing": "CUDA graph of launches over rotating input sets (> 126 MB L2 in total), CUDA events around the replay / launches"}
    line = {"metric": f"ss2d_scan_{'fwd_bwd' if with_bwd else 'fwd'}_steps_per_sec_{name}", "value": world * args.steps / (ms * 1e-3),
            "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": shapes[0][6], "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs rotate over sets that together exceed the 126 MB L2",
                       "sharding": "replicas only (the scan does not shard across GPUs)"},
            "e2e": {"value": world * 1e3 / ms_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": f"first shape only ({tag0}): pinned host tensors -> device, fwd{'+bwd' if with_bwd else ''}, results -> host"},
            "gpu_launches": args.steps * len(shapes) * (2 if with_bwd else 1), "roofline": roof, "kernels": kern}
    if R.available():
        line["reference_gpu"] = {tag: {k: v for k, v in rec.items() if k.startswith("reference_")} for tag, rec in kern.items()}
    if not args.no_cpu_baseline and world == 1:
        try:
            tag, Bn, KD, G, N, L, dn = shapes[0]
            r = cpu_reference_scan(Bn, KD, G, N, L, dn, with_bwd)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    return line


def run_scan_config_reference(name, args):
    workload, shapes, with_bwd = SCAN_CONFIGS[name]
    tag, Bn, KD, G, N, L, dn = shapes[0]
    r = cpu_reference_scan(Bn, KD, G, N, L, dn, with_bwd, budget_s=60.0)
    return {"impl": "reference", "metric": f"ss2d_scan_{'fwd_bwd' if with_bwd else 'fwd'}_steps_per_sec_{name}", "value": r["value"],
            "unit": "steps/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dn, "data": "synthetic", "config": {"workload": workload},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ---------------------------------------------------------------------------------------------------------------------
# --config train : BASELINE configs[4]
# ---------------------------------------------------------------------------------------------------------------------
TRAIN_LEVELS = ((40, 64, 8), (80, 32, 8), (160, 16, 2))   # (hidden_dim, H = W, VSSBlocks): the 18 blocks of DecompDualBranchDDWavelet


def build_train_stack(dev):
    """The scan-carrying part of DecompDualBranchDDWavelet (basicsr/archs/DecompDualBranchDDWavelet_arch.py:147: two U-shaped
    branches of VSSBlocks, d_state 1, on the wavelet half-resolution of 128x128 patches): 8 blocks at (B, 40, 64, 64), 8 at
    (B, 80, 32, 32), 2 at (B, 160, 16, 16) — scans 8x(B,160,4096), 8x(B,320,1024), 2x(B,640,256) as in SURVEY 8(d) config 5.
    The quaternion / wavelet glue around them (a frozen pretrained decomposition net, conv stems) is outside SURVEY 8."""
    import torch.nn as nn
    from bem_b200 import network

    class Stack(nn.Module):
        def __init__(self):
            super().__init__()
            self.levels = nn.ModuleList([nn.ModuleList([network.VSSBlock(hidden_dim=c, ssm_d_state=1, ssm_ratio=1, ssm_conv_bias=False,
                                                                         mlp_ratio=4) for _ in range(n)]) for c, _, n in TRAIN_LEVELS])

        def forward(self, xs):
            loss = 0.0
            for blocks, x in zip(self.levels, xs):
                for b in blocks:
                    x = b(x)
                loss = loss + x.float().pow(2).mean()
            return loss
    return Stack().to(dev)


def build_reference_train_model(dev, patched=True):
    """The REAL configs[4] network: the reference's DecompDualBranchDDWavelet (model code from oracle/_ref, built exactly as
    Options/DecompDualBranch2DDWavelet_4.yml:54-68) AFTER bem_b200.patch.install() — the reference's model and trainer-side code
    unmodified, its selective scan / traversal operators replaced by this package's kernels (what INTEGRATION.md section 1 gives a user)."""
    import bem_b200
    R = _ref()
    R.arch("DecompDualBranchDDWavelet_arch", True)       # import the reference's modules first: install() patches what is loaded
    if patched:
        done = bem_b200.patch.install()
        assert any(n.endswith("vmamba") for n in done), done
    return R.train_model(True, device=str(dev))


def run_train_config(args, rank, world, dev):
    import torch
    import torch.distributed as dist
    from bem_b200 import _lib
    torch.manual_seed(0)
    Bp = 8
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    real = _ref().available()
    if real:
        net = build_reference_train_model(dev).train()
        host = [torch.rand(Bp, 6, 128, 128, generator=g).pin_memory(), torch.rand(Bp, 3, 128, 128, generator=g).pin_memory()]

        class Wrap(torch.nn.Module):       # loss inside the DDP-wrapped module's graph, as basicsr's trainer computes it after net_g(lq)
            def __init__(self, m):
                super().__init__()
                self.m = m

            def forward(self, xs):
                return torch.nn.functional.l1_loss(self.m(xs[0])[-1], xs[1])    # train.pixel_opt: L1Loss (yml:100-103)
        core = Wrap(net)
        opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=2e-4, weight_decay=1e-4, betas=(0.9, 0.999),
                                capturable=True)   # yml:88-92
        workload = ("BASELINE configs[4]: fwd + L1 loss + bwd + AdamW step of the reference's DecompDualBranchDDWavelet "
                    "(Options/DecompDualBranch2DDWavelet_4.yml; model code unmodified from oracle/_ref, bem_b200.patch.install() applied: scan fwd/bwd and "
                    "cross scan / merge on libbem_b200), 8 x 6 x 128 x 128 per rank, DDP gradient all-reduce over the ranks")
    else:
        net = build_train_stack(dev).train()
        host = [torch.randn(Bp, c, h, h, generator=g).pin_memory() for c, h, _ in TRAIN_LEVELS]
        core = net
        opt = torch.optim.Adam(net.parameters(), lr=2e-4, capturable=True)
        workload = ("BASELINE configs[4] (scan-carrying part; the reference model is not staged): fwd+bwd+Adam step of the 18 VSSBlocks of "
                    "DecompDualBranch2DDWavelet_4 (8 @ 40ch 64x64, 8 @ 80ch 32x32, 2 @ 160ch 16x16; d_state 1), 8 patches of 128x128 per rank, DDP")
    side = torch.cuda.Stream(device=dev)      # DDP is built on the stream the step is later captured on (bem_b200/graphed.py)
    if world > 1:
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            model = torch.nn.parallel.DistributedDataParallel(core, device_ids=[dev.index])
        torch.cuda.current_stream(dev).wait_stream(side)
    else:
        model = core
    xs = [t.to(dev) for t in host]

    def step(inputs):
        opt.zero_grad(set_to_none=True)
        loss = model(inputs)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(xs)
    barrier()
    _lib.profile.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(xs)
    e1.record()
    barrier()
    ms_eager = e0.elapsed_time(e1)
    launches = _lib.profile.launches
    # the product's execution strategy for a training step: the whole step (forward, loss, backward, gradient all-reduce under DDP,
    # optimizer) captured as one CUDA graph and replayed (bem_b200.GraphedTrainStep); the eager figure is kept beside it
    graphed = None
    graph_error = "capture under DDP is opt-in, --graph-ddp"
    gs = None
    if world == 1 or getattr(args, "graph_ddp", False):
        try:
            import bem_b200
            gs = bem_b200.GraphedTrainStep(model, lambda m, *ins: m(list(ins)), opt, xs, warmup=11 if world > 1 else 3, stream=side)
        except Exception as ex:
            gs = None
            graph_error = f"{type(ex).__name__}: {ex}"
            if os.environ.get("BEM_BENCH_TRACE"):
                import traceback
                traceback.print_exc()
            torch.cuda.synchronize()
        if world > 1:      # every rank replays the graph or none does: a rank that fell back would issue different collectives
            ok = torch.tensor([1 if gs is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if gs is not None:
                    graph_error = "capture failed on another rank"
                gs = None
    if gs is not None:
        for _ in range(3):
            gs(*xs)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            gs(*xs)
        g1.record()
        barrier()
        graphed = g0.elapsed_time(g1)
    ms = graphed if graphed is not None else ms_eager
    run = (lambda ins: gs(*ins)) if graphed is not None else step
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        loss = run([t.to(dev, non_blocking=True) for t in host])
        lv = float(loss.item())
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    # every rank trains on its own patches: after the timed steps the replicas hold the same parameters only if the gradient
    # all-reduce really ran inside the (captured) step
    in_sync = None
    if world > 1:
        chk = torch.stack([p.detach().double().sum() for p in core.parameters() if p.requires_grad]).sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        in_sync = bool(all(torch.equal(allc[0], c) for c in allc))
    if rank != 0:
        return None
    peak, peak_src = _peaks()
    # dominant scan of the step: level-0 backward (B8 KD160 N1 L4096 fp32), graph-batched over rotating inputs
    fb, bb = scan_bytes(Bp, 160, 4, 1, 4096, 4)
    sets = scan_inputs(Bp, 160, 4, 1, 4096, torch.float32, dev, n_sets=6)
    cl = {j: _ours_scan_closures(s, True) for j, s in enumerate(sets)}
    fwd_ms = time_graph_rotating(lambda i: cl[i][0], len(sets), dev)[1]
    bwd_ms = time_graph_rotating(lambda i: cl[i][1], len(sets), dev)[1]
    roof = {"bound": "hbm", "achieved": bb / bwd_ms / 1e6, "peak": peak, "unit": "GB/s", "frac": bb / bwd_ms / 1e6 / peak, "traffic": None,
            "kernel": "scan bwd (B8 KD160 N1 L4096 fp32), level-0 scan of the train step", "peak_source": peak_src,
            "bytes_per_launch": float(bb), "ms_per_launch": bwd_ms, "fwd": {"ms": fwd_ms, "GBps": fb / fwd_ms / 1e6, "frac": fb / fwd_ms / 1e6 / peak}}
    patches = world * Bp * args.steps
    line = {"metric": "train_patches_per_sec_128x128", "value": patches / (ms * 1e-3), "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload,
                       "l2": "activations of a step (8 x 40 x 4096 x 4 B x ~30 tensors per block) exceed the L2 only at level 0; no flush inside a step",
                       "parallelism": f"ddp{world}", "loss_last": lv,
                       "execution": ("whole step replayed as one CUDA graph (bem_b200.GraphedTrainStep)" if graphed is not None
                                     else "eager launches (no graph: " + graph_error + ")"),
                       "eager_ms_per_step": ms_eager / args.steps, "ddp_replicas_in_sync_after_run": in_sync},
            "e2e": {"value": world * Bp * 1e3 / ms_e2e, "unit": "patches/s", "h2d_bytes_per_step": sum(t.numel() * 4 for t in host),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "roofline": roof}
    R = _ref()
    if R.available():
        rc = {j: _ref_scan_closures(s, True) for j, s in enumerate(sets)}
        line["reference_gpu"] = {"scan_fwd_B8_KD160_L4096_ms": time_graph_rotating(lambda i: rc[i][0], len(sets), dev)[1],
                                 "scan_bwd_B8_KD160_L4096_ms": time_graph_rotating(lambda i: rc[i][1], len(sets), dev)[1],
                                 "what": "selective_scan_cuda_oflex (reference, unmodified, sm_100a), same shape and protocol"}
        if real:
            try:   # the same training step on the UNPATCHED reference model (its CUDA extension + Triton traversal), this GPU, no DDP
                import bem_b200
                bem_b200.patch.uninstall()
                torch.manual_seed(0)
                ref_net = R.train_model(True, device=str(dev)).train()
                ref_opt = torch.optim.AdamW([p for p in ref_net.parameters() if p.requires_grad], lr=2e-4, weight_decay=1e-4)

                def ref_step():
                    ref_opt.zero_grad(set_to_none=True)
                    torch.nn.functional.l1_loss(ref_net(xs[0])[-1], xs[1]).backward()
                    ref_opt.step()
                for _ in range(3):
                    ref_step()
                torch.cuda.synchronize(dev)
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                nref = max(3, min(args.steps, 10))
                for _ in range(nref):
                    ref_step()
                r1.record()
                torch.cuda.synchronize(dev)
                rms = r0.elapsed_time(r1) / nref
                line["reference_gpu"]["train_step_ms"] = rms
                line["reference_gpu"]["train_patches_per_s_one_gpu"] = Bp * 1e3 / rms
                line["reference_gpu"]["train_what"] = "the same model, unpatched: reference CUDA extension (sm_100a) + Triton cross scan / merge, one GPU"
            except Exception as ex:
                line["reference_gpu"]["train_step_ms"] = f"failed: {type(ex).__name__}: {ex}"
    if not args.no_cpu_baseline and world == 1:
        try:
            r = cpu_reference_scan(Bp, 160, 4, 1, 4096, "f32", True)
            line["cpu_baseline"] = {"value": r["value"], "unit": "level-0 scans (fwd+bwd)/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "scans/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    return line
