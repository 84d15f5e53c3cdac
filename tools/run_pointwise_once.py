import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bem_b200.bayesian import functional as BF
dev = torch.device("cuda"); cin, cout, P = int(os.environ.get("CIN", 40)), int(os.environ.get("COUT", 320)), int(os.environ.get("NPIX", 240000))
x = torch.randn(1, cin, P, device=dev); mu = torch.randn(cout, cin, device=dev) / cin ** 0.5
sig = torch.full_like(mu, 0.05); eps = torch.randn(1, cout, cin, device=dev); b = torch.randn(1, cout, device=dev)
lnp = (torch.ones(cin, device=dev), torch.zeros(cin, device=dev), 1e-5) if int(os.environ.get("LN", 1)) else None
for _ in range(5):
    BF.pointwise_conv_sampled(x, mu, sig, eps, b, 1, ln=lnp)
torch.cuda.synchronize(); print("done")
