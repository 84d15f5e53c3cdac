"""Time the direct 3x3 stem kernels at the level-0 shapes (CUDA events, median of 20): `python tools/run_conv3x3_once.py`."""
import importlib, os, sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BF = importlib.import_module("bayesian-enhancement-model_b200.bayesian.functional")


def main():
    dev = torch.device("cuda:0")
    H, W = 400, 600
    torch.manual_seed(0)
    for cin, cout in ((3, 40), (40, 3)):
        x = torch.randn(1, cin, H, W, device=dev)
        w = torch.randn(cout, cin, 3, 3, device=dev) * 0.1
        b = torch.randn(cout, device=dev)
        ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
        got = BF.conv3x3_direct(x, w, b)
        err = ((got.double() - ref).abs().max() / ref.abs().max()).item()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); BF.conv3x3_direct(x, w, b); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"conv3x3 {cin}->{cout}: {ts[len(ts)//2]:.1f} us  rel err {err:.2e}")


if __name__ == "__main__":
    main()
