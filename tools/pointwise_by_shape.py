"""Per-shape table of the 1x1 launches of one MC sample (eager, CUDA events per C-ABI call)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bem_b200 import _lib, mc, network
torch.manual_seed(0)
dev = torch.device("cuda")
net = network.build_bayesian_model().to(dev).eval()
s = mc.MCSampler(net, seed=1, arena=True, graph=False)
img = torch.rand(1, 3, 400, 600, device=dev)
s.sample(img, [0]); s.sample(img, [1])
_lib.profile.reset(armed=True)
for i in range(3):
    s.sample(img, [2 + i])
prof = _lib.profile.summary()
for name in ("bayes_pointwise", "bayes_depthwise", "scan_fwd", "cross_scan", "cross_merge"):
    rec = prof[name]
    print(f"== {name}: {rec['ms']/3:.3f} ms per sample")
    for k, v in sorted(rec["by_key"].items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {k:40s} calls/sample {v['calls']/3:5.1f}  us/call {1e3*v['ms']/v['calls']:7.1f}  GB/s {v['bytes']/v['ms']/1e6:7.0f}  share {100*v['ms']/rec['ms']:5.1f}%")
