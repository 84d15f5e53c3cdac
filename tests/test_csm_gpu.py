"""GPU parity of CrossScan / CrossMerge: pure data movement and fixed-association adds -> bit-exact."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _layouts():
    for scans in (0, 1, 2):
        for icf in (1, 0):
            for ocf in (1, 0):
                for obo in (0, 1):
                    yield scans, icf, ocf, obo


def _img(x, icf, obo):
    if icf:
        return x
    return np.transpose(x, (0, 3, 4, 1, 2)) if obo else np.transpose(x, (0, 2, 3, 1))


@pytest.mark.parametrize("scans,icf,ocf,obo", list(_layouts()))
def test_golden_reference_vectors(golden_csm, scans, icf, ocf, obo):
    """outputs recorded from the reference's torch path (csm_triton.py:22-179); broken reference combos are absent/skipped"""
    import bem_b200
    tag = f"s{scans}_i{icf}_o{ocf}_b{obo}"
    x = golden_csm["x4"] if obo else golden_csm["x"]
    H, W = x.shape[-2:]
    if f"scan/{tag}" in golden_csm and not (obo and scans == 2 and not icf) and not (obo and scans == 1 and icf and not ocf):
        src = torch.tensor(np.ascontiguousarray(_img(x, icf, obo)), device="cuda")
        y = bem_b200.cross_scan_fn(src, bool(icf), bool(ocf), bool(obo), scans)
        ref = golden_csm[f"scan/{tag}"]
        np.testing.assert_array_equal(y.cpu().numpy().reshape(-1), ref.reshape(-1))
        assert y.shape == ((x.shape[0], 4, x.shape[-3], H * W) if ocf else (x.shape[0], H * W, 4, x.shape[-3]))
    if f"merge/{tag}" in golden_csm:
        ys = torch.tensor(golden_csm[f"merge_in/{tag}"], device="cuda")
        m = bem_b200.cross_merge_fn(ys, bool(icf), bool(ocf), bool(obo), scans)
        ref = golden_csm[f"merge/{tag}"]
        assert tuple(m.shape) == ref.shape
        np.testing.assert_allclose(m.cpu().numpy(), ref, rtol=0, atol=1e-6)


@pytest.mark.parametrize("scans,icf,ocf,obo", list(_layouts()))
@pytest.mark.parametrize("shape", [(2, 5, 56, 57), (1, 3, 1, 9), (1, 2, 33, 31), (2, 3, 56, 60), (1, 2, 132, 72), (1, 2, 64, 128)])
def test_against_oracle_all_layouts(shape, scans, icf, ocf, obo):
    """non-square, non-multiple-of-32 shapes like the reference's own check (csm_triton.py:514: 56 x 57); the shapes with H and W
    multiples of 4 take the 16-byte tiled kernels (partial, multiple and exact 64 x 64 tiles)"""
    import bem_b200
    Bt, Cc, H, W = shape
    rng = np.random.RandomState(0)
    x = rng.randn(Bt, 4, Cc, H, W).astype(np.float32) if obo else rng.randn(Bt, Cc, H, W).astype(np.float32)
    src = np.ascontiguousarray(_img(x, icf, obo))
    y = bem_b200.cross_scan_fn(torch.tensor(src, device="cuda"), bool(icf), bool(ocf), bool(obo), scans)
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.cross_scan_oracle(src, bool(icf), bool(ocf), bool(obo), scans))
    ys = rng.randn(Bt, 4, Cc, H * W).astype(np.float32) if ocf else rng.randn(Bt, H * W, 4, Cc).astype(np.float32)
    yin = torch.tensor(ys, device="cuda")
    yin = yin.view(Bt, 4, Cc, H, W) if ocf else yin.view(Bt, H, W, 4, Cc)
    m = bem_b200.cross_merge_fn(yin, bool(icf), bool(ocf), bool(obo), scans)
    ref = oracle.cross_merge_oracle(ys, H, W, bool(icf), bool(ocf), bool(obo), scans)
    if scans == 1 and not obo:
        np.testing.assert_allclose(m.cpu().numpy(), ref, rtol=0, atol=1e-6)   # sum order of torch.sum is unspecified
    else:
        np.testing.assert_array_equal(m.cpu().numpy(), ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_low_precision_is_bit_exact(dtype):
    import bem_b200
    x = torch.randn(2, 6, 40, 37, device="cuda").to(dtype)
    y = bem_b200.cross_scan_fn(x)
    ref = oracle.cross_scan_oracle(x.float().cpu().numpy())
    np.testing.assert_array_equal(y.float().cpu().numpy(), ref)
    ys = torch.randn(2, 4, 6, 40, 37, device="cuda").to(dtype)
    m = bem_b200.cross_merge_fn(ys)
    # torch semantics: every add rounds to the tensor dtype (csm_triton.py:60-62)
    v = ys.view(2, 4, 6, -1)
    a = v[:, 0:2] + v[:, 2:4].flip(dims=[-1])
    ref = a[:, 0] + a[:, 1].view(2, 6, 37, 40).transpose(2, 3).contiguous().view(2, 6, -1)
    assert torch.equal(m, ref)


def test_autograd_pairing(golden_csm):
    """d cross_scan / dx = cross_merge and d cross_merge / dy = cross_scan (csm_triton.py:207-225, 248-273)"""
    import bem_b200
    x = torch.tensor(golden_csm["x"], device="cuda", requires_grad=True)
    gy = torch.tensor(golden_csm["scan_bwd/gy"], device="cuda")
    y = bem_b200.cross_scan_fn(x)
    y.backward(gy)
    np.testing.assert_allclose(x.grad.cpu().numpy(), golden_csm["scan_bwd/gx"], rtol=0, atol=1e-6)
    ys = torch.randn(2, 4, 3, 5, 7, device="cuda", requires_grad=True)
    m = bem_b200.cross_merge_fn(ys)
    gm = torch.randn_like(m)
    m.backward(gm)
    ref = oracle.cross_scan_oracle(gm.view(2, 3, 5, 7).cpu().numpy()).reshape(2, 4, 3, 5, 7)
    np.testing.assert_array_equal(ys.grad.cpu().numpy(), ref)


def test_full_size_round_trip():
    """600x400 level-0 shape: merge(scan(x)) == ((x + x) + (x + x)) exactly — a size-independent property"""
    import bem_b200
    x = torch.randn(1, 40, 400, 600, device="cuda")
    xs = bem_b200.cross_scan_fn(x)
    assert torch.equal(xs[:, 0].reshape(1, 40, 400, 600), x)
    assert torch.equal(xs[:, 1].reshape(1, 40, 600, 400), x.transpose(2, 3))
    assert torch.equal(xs[:, 2], xs[:, 0].flip(-1)) and torch.equal(xs[:, 3], xs[:, 1].flip(-1))
    m = bem_b200.cross_merge_fn(xs.view(1, 4, 40, 400, 600)).view(1, 40, 400, 600)
    assert torch.equal(m, (x + x) + (x + x))


def test_cpu_tensor_is_rejected():
    import bem_b200
    with pytest.raises(RuntimeError):
        bem_b200.cross_scan_fn(torch.randn(1, 2, 4, 4))
