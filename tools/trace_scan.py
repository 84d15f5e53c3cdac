"""Timeline of CTA 0 of the scan forward kernel (env BEM_SCAN_TRACE=1): producer and two consumer warps."""
import ctypes as C, os, sys, collections
os.environ.setdefault("BEM_SCAN_TRACE", "2")   # 1: classic schedule (scan_fwd.cu), 2: deferred-finish schedule
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bem_b200
from bem_b200 import _lib
dev = torch.device("cuda"); B, KD, N, G, L = 1, 160, 1, 4, 240000
torch.manual_seed(0)
u = torch.randn(B, KD, L, device=dev); delta = 0.5 * torch.rand(B, KD, L, device=dev); A = -0.5 * torch.rand(KD, N, device=dev)
Bm = torch.randn(B, G, N, L, device=dev); Cm = torch.randn(B, G, N, L, device=dev); D = torch.randn(KD, device=dev); bias = 0.5 * torch.rand(KD, device=dev)
ext = bem_b200.selective_scan_cuda_oflex
fn = _lib.lib.bem_dbg_scan_trace if os.environ["BEM_SCAN_TRACE"] == "1" else _lib.lib.bem_dbg_scan_deferred_trace; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_uint32 * (4 * 8192))()
for _ in range(3):
    ext.fwd(u, delta, A, Bm, Cm, D, bias, True, 1, True)
    n = fn(buf, 8192)
rec = sorted(((buf[4 * i + 2], i // 2048, buf[4 * i], buf[4 * i + 1]) for i in range(n) if buf[4 * i + 3]))
t0 = rec[0][0]
names = {4: "P1 wait hdr", 5: "P1 hdr", 6: "P1 issued", 1: "P start", 2: "P slot free", 3: "P issued", 6: "P hdr stored", 7: "P pass0 done", 10: "C wait", 11: "C full", 12: "C scanned", 13: "C lookback done", 14: "C released"}
print("records", len(rec), "span clk", rec[-1][0] - t0)
skip, lim = int(os.environ.get("SKIP", 200)), int(os.environ.get("LINES", 90))
for t, role, tag, arg in rec[skip:skip + lim]:
    print(f"{t - t0:8d}  role {role}  {names.get(tag, tag):16s} {arg}")
# phase durations per consumer role
for role in (1, 2):
    r = [x for x in rec if x[1] == role]
    dur = collections.defaultdict(list)
    for a, b in zip(r, r[1:]):
        dur[(a[2], b[2])].append(b[0] - a[0])
    for k, v in sorted(dur.items()):
        v2 = sorted(v); print(f"role {role} {names.get(k[0])} -> {names.get(k[1])}: n {len(v)} median {v2[len(v2)//2]} mean {sum(v)/len(v):.0f}")
r = [x for x in rec if x[1] == 0]
dur = collections.defaultdict(list)
for a, b in zip(r, r[1:]):
    dur[(a[2], b[2])].append(b[0] - a[0])
for k, v in sorted(dur.items()):
    v2 = sorted(v); print(f"producer {names.get(k[0])} -> {names.get(k[1])}: n {len(v)} median {v2[len(v2)//2]} mean {sum(v)/len(v):.0f}")

r = [x for x in rec if x[1] == 3]
dur = collections.defaultdict(list)
for a, b in zip(r, r[1:]):
    dur[(a[2], b[2])].append(b[0] - a[0])
for k, v in sorted(dur.items()):
    v2 = sorted(v); print(f"producer1 {names.get(k[0])} -> {names.get(k[1])}: n {len(v)} median {v2[len(v2)//2]} mean {sum(v)/len(v):.0f}")
