"""Timeline of CTA 0 of the persistent pointwise kernel (env BEM_PW_TRACE=1): prints per-role event gaps."""
import ctypes as C, os, sys, collections
os.environ["BEM_PW_TRACE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bem_b200 import _lib
from bem_b200.bayesian import functional as BF
dev = torch.device("cuda"); cin, cout, P = int(os.environ.get("CIN", 160)), int(os.environ.get("COUT", 40)), int(os.environ.get("NPIX", 240000))
x = torch.randn(1, cin, P, device=dev); w = torch.randn(1, cout, cin, device=dev)
lnp = (torch.ones(cin, device=dev), torch.zeros(cin, device=dev), 1e-5) if int(os.environ.get("LN", 0)) else None
fn = _lib.lib.bem_dbg_pointwise_trace; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_uint32 * (4 * 16384))()
for _ in range(3):
    BF.pointwise_conv(x, w, None, 1, ln=lnp)
    n = fn(buf, 16384)
rec = sorted(((buf[4 * i + 2], buf[4 * i], buf[4 * i + 1]) for i in range(n) if buf[4 * i + 3]))
n = len(rec)
t0 = rec[0][0]
names = {2: "prod slot free", 3: "prod landed", 1: "prod issue", 10: "xf0 raw_full", 11: "xf1 raw_full", 12: "xf0 a_empty", 13: "xf1 a_empty", 14: "xf0 arrive", 15: "xf1 arrive",
         40: "cta entry", 41: "prologue done", 42: "pdl wait done", 43: "cta exit", 20: "mma acc_empty", 21: "mma a_full", 22: "mma2 acc_empty", 23: "mma2 a_full", 30: "epi acc_full", 31: "epi done"}
print("records", n, "span clk", rec[-1][0] - t0)
lim = int(os.environ.get("LINES", 150)); skip = int(os.environ.get("SKIP", 400))
only = os.environ.get("ONLY")
if only:
    rec_p = [r for r in rec if str(r[1]) in only.split(",")]
else:
    rec_p = rec
for t, tag, arg in rec_p[skip:skip + lim]:
    print(f"{t - t0:8d}  {names.get(tag, tag):14s} {arg}")
last = {}; gaps = collections.defaultdict(list)
for t, tag, arg in rec:
    if tag in last: gaps[tag].append(t - last[tag])
    last[tag] = t
for tag, g in sorted(gaps.items()):
    g2 = sorted(g); print(f"{names.get(tag, tag):14s} n {len(g):5d} median gap {g2[len(g2)//2]:6d} mean {sum(g)/len(g):8.1f} max {g2[-1]}")
