import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bem_b200.bayesian import functional as BF
dev = torch.device("cuda"); C, H, W = int(os.environ.get("C", 320)), 400, 600
act = os.environ.get("ACT", "gelu_gate"); act = None if act == "none" else act
x = torch.randn(1, C, H, W, device=dev); w = torch.randn(1, C, 3, 3, device=dev); b = torch.randn(1, C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = BF.depthwise_conv3x3(x, w, b, 1, act=act); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[len(ts) // 2]
print(f"depthwise C={C} act={act}: {t*1e3:.1f} us  {(x.numel()+y.numel())*4/t/1e6:.0f} GB/s")
