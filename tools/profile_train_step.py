"""Top kernels of one training step of BASELINE configs[4] on the reference's model after patch.install() (torch.profiler,
device time): `python tools/profile_train_step.py [--unpatched]`."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_configs as BC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--unpatched", action="store_true")
ap.add_argument("--rows", type=int, default=40)
a = ap.parse_args()
dev = torch.device("cuda")
net = BC.build_reference_train_model(dev, patched=not a.unpatched).train()
opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=2e-4, weight_decay=1e-4)
x = torch.rand(8, 6, 128, 128, device=dev)
gt = torch.rand(8, 3, 128, 128, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.l1_loss(net(x)[-1], gt)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=a.rows, max_name_column_width=90))
