"""One forward + backward of the row-sequential scan at configs[0] (B1 KD384 N16 L4096 fp32), twice (ncu: read the second)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BEM_SCAN_ROWS"] = "1"
import bem_b200  # noqa: E402

ext = bem_b200.selective_scan_cuda_oflex
dev = torch.device("cuda")
B, KD, N, G, L = 1, 384, 16, 4, int(os.environ.get("ROWS_L", "4096"))
dt = torch.bfloat16 if os.environ.get("ROWS_BF16") else torch.float32
torch.manual_seed(0)
u = torch.randn(B, KD, L, device=dev, dtype=dt)
delta = (0.5 * torch.rand(B, KD, L, device=dev)).to(dt)
A = -0.5 * torch.rand(KD, N, device=dev)
Bm = torch.randn(B, G, N, L, device=dev, dtype=dt)
Cm = torch.randn(B, G, N, L, device=dev, dtype=dt)
D = torch.randn(KD, device=dev)
bias = 0.5 * torch.rand(KD, device=dev)
dout = torch.randn(B, KD, L, device=dev, dtype=dt)
for _ in range(2):
    out, x = ext.fwd(u, delta, A, Bm, Cm, D, bias, True, 1, True)
    ext.bwd(u, delta, A, Bm, Cm, D, bias, dout, x, True, 1)
torch.cuda.synchronize()
print("done")
