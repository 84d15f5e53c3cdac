"""Channel-first LayerNorm kernels (csrc/ln2d.cu) against torch.nn.functional.layer_norm, the op LayerNorm2d wraps
(basicsr/vmamba/models/vmamba.py:58-63), forward and backward; tolerance 1e-5 normalised max error (fp32 tier)."""
import pytest
import torch

import oracle
from conftest import nmax_err

pytestmark = pytest.mark.gpu


def _ref(x, w, b, eps):
    y = torch.nn.functional.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps)
    return y.permute(0, 3, 1, 2)


@pytest.mark.parametrize("shape", [(8, 40, 64, 64), (2, 80, 32, 32), (1, 160, 16, 16), (3, 7, 5, 9), (1, 3, 4, 4), (2, 40, 100, 150), (2, 200, 9, 11), (1, 161, 33, 3)])
@pytest.mark.parametrize("affine", ["wb", "w", "none"])
def test_layernorm2d_matches_torch_forward_and_backward(shape, affine):
    import bem_b200
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g, dtype=torch.float64) * 1.5 + 0.3).cuda()
    C = shape[1]
    w = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).cuda() if affine != "none" else None
    b = torch.randn(C, generator=g, dtype=torch.float64).cuda() if affine == "wb" else None
    dy = torch.randn(*shape, generator=g, dtype=torch.float64).cuda()
    leaves64 = [t.clone().requires_grad_() if t is not None else None for t in (x, w, b)]
    y64 = _ref(*leaves64, 1e-5)
    y64.backward(dy)
    leaves = [t.float().clone().requires_grad_() if t is not None else None for t in (x, w, b)]
    y = bem_b200.layer_norm_2d(*leaves, eps=1e-5)
    y.backward(dy.float())
    assert nmax_err(y.detach().cpu().numpy(), y64.detach().cpu().numpy()) < 1e-5
    for got, want, name in zip(leaves, leaves64, ("dx", "dweight", "dbias")):
        if got is not None:
            assert nmax_err(got.grad.cpu().numpy(), want.grad.cpu().numpy()) < 1e-5, name
    # ... and against the numpy restatement of the op (oracle.layernorm2d_oracle, itself pinned to torch on the CPU)
    o = oracle.layernorm2d_oracle(x.cpu().numpy(), None if w is None else w.cpu().numpy(), None if b is None else b.cpu().numpy(), 1e-5,
                                  dy.cpu().numpy())
    assert nmax_err(y.detach().cpu().numpy(), o["y"]) < 1e-5
    assert nmax_err(leaves[0].grad.cpu().numpy(), o["dx"]) < 1e-5


def test_layernorm2d_inference_forward_saves_nothing_and_rejects_what_it_cannot_run():
    import bem_b200
    x = torch.randn(2, 40, 8, 8, device="cuda")
    w, b = torch.ones(40, device="cuda"), torch.zeros(40, device="cuda")
    with torch.no_grad():
        y = bem_b200.layer_norm_2d(x, w, b)
    assert nmax_err(y.cpu().numpy(), _ref(x, w, b, 1e-5).cpu().numpy()) < 1e-5
    with pytest.raises(RuntimeError):
        bem_b200.layer_norm_2d(x.half(), w, b)
    with pytest.raises(RuntimeError):
        bem_b200.layer_norm_2d(x.cpu(), w.cpu(), b.cpu())


def test_graphed_train_step_replays_the_eager_step():
    """GraphedTrainStep: the captured forward + loss + backward + AdamW step gives the same losses and parameters as eager steps
    on the same inputs (atomics in the backward kernels reorder sums: tolerance instead of bit equality). The model runs this
    package's autograd operators (scan, channel-first LayerNorm) and indexes a tensor with a Python list, the pattern that stops
    a plain capture of the reference's decomposition front end (DecompDualBranchDDWavelet_arch.py:129-130)."""
    import copy

    import bem_b200

    torch.manual_seed(0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = torch.nn.Conv2d(4, 16, 3, padding=1)
            self.norm = torch.nn.LayerNorm(16)
            self.proj = torch.nn.Conv2d(16, 18, 1)
            self.A_log = torch.nn.Parameter(torch.zeros(16, 1))
            self.D = torch.nn.Parameter(torch.ones(16))
            self.dt_bias = torch.nn.Parameter(torch.full((16,), -2.0))
            self.head = torch.nn.Conv2d(16, 3, 1)

        def forward(self, x):
            x = x[:, [0, 2, 1, 3]]                     # host-built index tensor on every call
            f = bem_b200.layer_norm_2d(self.stem(x), self.norm.weight, self.norm.bias, self.norm.eps)
            B, C, H, W = f.shape
            z = self.proj(f).flatten(2)
            dt, Bm, Cm = z[:, :16].contiguous(), z[:, 16:17].reshape(B, 1, 1, H * W), z[:, 17:18].reshape(B, 1, 1, H * W)
            y = bem_b200.selective_scan_fn(f.flatten(2), dt, -torch.exp(self.A_log), Bm.contiguous(), Cm.contiguous(), self.D,
                                           self.dt_bias, True)
            return self.head(y.view(B, C, H, W))

    net = Net().cuda()
    ref = copy.deepcopy(net)
    loss_fn = lambda m, x, t: torch.nn.functional.l1_loss(m(x), t)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, capturable=True)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3, capturable=True)
    x0, t0 = torch.rand(2, 4, 32, 32, device="cuda"), torch.rand(2, 3, 32, 32, device="cuda")
    step = bem_b200.GraphedTrainStep(net, loss_fn, opt, [x0, t0], warmup=2)   # the warm-up iterations are real optimizer steps
    for _ in range(2):
        opt_ref.zero_grad(set_to_none=False)
        loss_fn(ref, x0, t0).backward()
        opt_ref.step()
    for i in range(3):
        xi, ti = torch.rand(2, 4, 32, 32, device="cuda"), torch.rand(2, 3, 32, 32, device="cuda")
        l_graph = float(step(xi, ti))
        opt_ref.zero_grad(set_to_none=False)
        l_ref = loss_fn(ref, xi, ti)
        l_ref.backward()
        opt_ref.step()
        assert abs(l_graph - float(l_ref)) < 1e-4 * max(1.0, abs(float(l_ref))), i
    for (n, p), q in zip(net.named_parameters(), ref.parameters()):
        # Adam divides by sqrt(v): rounding-level differences of a near-zero gradient move such a parameter by up to lr per step
        assert nmax_err(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 1e-3, n
