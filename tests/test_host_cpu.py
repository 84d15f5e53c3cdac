"""CPU-side checks (`-m "not gpu"`): the C-ABI library loads and exports every symbol include/bem_b200.h declares, the
host logic (sample sharding, gather / select / broadcast under gloo world_size 2, module contracts) behaves."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bem():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
    import bem_b200
    return bem_b200


def test_library_exports_every_declared_symbol(bem):
    hdr = open(os.path.join(ROOT, "include", "bem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bem_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    out = subprocess.check_output(["nm", "-D", "--defined-only", bem._lib.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert declared <= exported, declared - exported
    assert declared == set(bem._lib.SYMBOLS.keys())
    assert bem._lib.lib.bem_abi_version() == bem._lib.ABI_VERSION == 14
    assert b"workspace" in bem._lib.lib.bem_error_string(10002)


def test_size_queries_without_gpu(bem):
    lib = bem._lib.lib
    assert lib.bem_scan_chunk_len(0) == 384 and lib.bem_scan_chunk_len(1) == lib.bem_scan_chunk_len(2) == 512
    assert lib.bem_scan_chunk_len(7) == 0
    # B1 KD160 L240000 N1 fp32: 625 chunks x 160 rows x 16 B + header
    assert lib.bem_scan_workspace_bytes(1, 160, 240000, 1, 0) == 128 + 2 * 160 * 625 * 16
    assert lib.bem_scan_workspace_bytes(0, 160, 100, 1, 0) == 0


def test_pack_table_is_built_on_the_host(bem):
    """bem_bayes_pointwise_pack_table (no GPU work): tiling, block ranges and workspace carving of every entry"""
    import ctypes as C
    L = bem._lib
    n = 3
    shapes = [(40, 320, 240000), (160, 40, 240000), (640, 160, 15000)]      # (cin, cout, P)
    arr = (L.BemBayesPointwiseParams * n)()
    need = []
    for e, (cin, cout, P) in zip(arr, shapes):
        nb = L.lib.bem_bayes_pointwise_workspace_bytes(1, cin, cout)
        need.append(nb)
        e.n_samples, e.batch, e.cin, e.cout, e.P = 1, 1, cin, cout, P
        e.x, e.w, e.out = 0x10000, 0x20000, 0x30000           # never dereferenced on the host
        e.workspace, e.workspace_bytes = 0x7000000 + 0x1000000 * len(need), nb
    nbytes = L.lib.bem_bayes_pointwise_pack_table_bytes(n)
    assert nbytes > 0 and nbytes % n == 0

    class Entry(C.Structure):
        _fields_ = [("p", L.BemBayesPointwiseParams), ("pack", C.c_void_p), ("vec", C.c_void_p)] + [
            (k, C.c_int32) for k in ("NT", "ntiles", "nk", "fold_ln", "block0", "nblocks")]
    assert C.sizeof(Entry) == nbytes // n
    tab = (Entry * n)()
    total = C.c_int32(0)
    assert L.lib.bem_bayes_pointwise_pack_table(arr, n, C.cast(tab, C.c_void_p), C.byref(total)) == 0
    blocks = 0
    for t, (cin, cout, P), nb in zip(tab, shapes, need):
        assert t.block0 == blocks and t.nblocks > 0
        assert t.nk == (cin + 15) // 16 and t.NT % 16 == 0 and t.NT * t.ntiles >= cout and t.NT <= 192
        assert t.fold_ln == 1                                                # aligned input -> persistent kernel layout
        assert t.nblocks == t.ntiles * t.nk + (cout + 7) // 8
        assert t.pack == t.p.workspace and t.pack < t.vec <= t.p.workspace + nb - 8 * t.NT * t.ntiles
        blocks += t.nblocks
    assert total.value == blocks
    arr[1].workspace_bytes = 16                                              # too small -> workspace error, nothing written past
    assert L.lib.bem_bayes_pointwise_pack_table(arr, n, C.cast(tab, C.c_void_p), C.byref(total)) == 10002
    assert L.lib.bem_bayes_pointwise_pack_table(None, n, C.cast(tab, C.c_void_p), C.byref(total)) == 10001
    assert L.lib.bem_bayes_pointwise_pack_run(None, 1, 1, None) == 10001


def test_pack_plan_matches_calls_by_position_and_signature(bem):
    """functional.PackPlan host logic: only tensors inside the drawn buffer are eligible; a call out of step (or with another
    signature) makes this and every later call pack for itself"""
    BF = bem.bayesian.functional
    buf = torch.zeros(64)
    lo = buf.data_ptr()
    plan = BF.PackPlan((lo, lo + 4 * buf.numel()))
    inside, outside = buf[8:24], torch.zeros(16)
    assert plan.eligible(inside, None) and plan.eligible(inside, buf[30:32])
    assert not plan.eligible(outside, None) and not plan.eligible(inside, outside) and not plan.eligible(None, None)
    with plan.recording():
        assert BF._ACTIVE_PLAN is plan and plan.mode == "record"
        plan.entries.append((("a",), None, "ws_a", ()))
        plan.entries.append((("b",), None, "ws_b", ()))
    assert BF._ACTIVE_PLAN is None and plan.mode is None
    with plan.playing():
        assert plan._lookup(("a",)) is None            # no device table yet: nothing is prepacked
    plan.table = object()
    with plan.playing():
        assert plan._lookup(("a",)) == "ws_a" and plan._lookup(("b",)) == "ws_b" and plan._lookup(("c",)) is None
    misses = plan.misses
    with plan.playing():
        assert plan._lookup(("b",)) is None            # out of step
        assert plan._lookup(("a",)) is None            # ... stays out of step for the rest of the forward
    assert plan.misses == misses + 2
    with plan.playing():                               # the next forward starts over
        assert plan._lookup(("a",)) == "ws_a"


def test_training_helpers_refuse_the_cpu(bem):
    """layer_norm_2d has no CPU path (RuntimeError like the other operators); the list-index mode of GraphedTrainStep leaves CPU
    indexing alone"""
    with pytest.raises(RuntimeError):
        bem.layer_norm_2d(torch.randn(1, 4, 3, 3), torch.ones(4), torch.zeros(4))
    assert not bem.layernorm.supported(torch.randn(1, 4, 3, 3))
    from bem_b200.graphed import _DeviceIndexMode
    t = torch.arange(24.0).view(2, 4, 3)
    with _DeviceIndexMode() as mode:
        got = t[:, [0, 2]]
    assert torch.equal(got, t[:, [0, 2]]) and not mode.cache


def test_no_cpu_fallback(bem):
    u = torch.randn(1, 4, 16)
    with pytest.raises(RuntimeError):
        bem.selective_scan_fn(u, u, torch.randn(4, 1), torch.randn(1, 1, 1, 16), torch.randn(1, 1, 1, 16))
    with pytest.raises(RuntimeError):
        bem.cross_scan_fn(torch.randn(1, 2, 4, 4))
    with pytest.raises(RuntimeError):
        bem.bayesian.Conv2dReparameterization(4, 4, 1)(torch.randn(1, 4, 3, 3))
    with pytest.raises(RuntimeError):
        bem.mc.select_best(torch.randn(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bayesian-enhancement-model_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_sharding_helpers(bem):
    mc = bem.mc
    shards = [mc.shard_samples(100, r, 8) for r in range(8)]
    assert sorted(len(s) for s in shards) == [12] * 4 + [13] * 4          # 13/12 split
    assert sorted(i for s in shards for i in s) == list(range(100))
    assert all(mc.owner_of(i, 8) == r for r, s in enumerate(shards) for i in s)
    assert mc.shard_samples(3, 5, 8) == []


def test_module_contracts_on_cpu(bem):
    net = bem.network.build_bayesian_model()
    assert sum(p.numel() for p in net.parameters()) == 2768887            # SURVEY Appendix A probe 6
    layers = bem.bayesian.bayesian_layers(net)
    assert len(layers) == 60 and [l.layer_id for l in layers] == list(range(60))
    keys = net.state_dict().keys()
    assert not any("eps_" in k or "prior_" in k for k in keys)
    assert any(k.endswith("op.in_proj.mu_weight") for k in keys) and any(k.endswith("mlp.dwconv.rho_bias") for k in keys)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bem_b200
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class StubSampler:   # deterministic stand-in for MCSampler: prediction i is a constant image of value f(i)
        def sample(self, x, ids):
            vals = torch.tensor([((i * 37) % 11) / 10.0 for i in ids], dtype=x.dtype)
            return vals.view(-1, 1, 1, 1) * torch.ones(len(ids), *x.shape[1:], dtype=x.dtype)

    x = torch.zeros(1, 3, 4, 5)
    score = lambda p: p.mean(dim=(1, 2, 3))

    def select_stub(scores, take_min):   # the product's selection is a CUDA kernel; the host logic under test is the exchange
        lst = scores.tolist()
        i = lst.index(min(lst) if take_min else max(lst))
        return i, lst[i]
    with pytest.raises(RuntimeError):    # no CPU selection in the product
        bem_b200.mc.mc_infer(StubSampler(), x, 7, score_fn=score)
    res = bem_b200.mc.mc_infer(StubSampler(), x, 7, score_fn=score, monte_carlo_mean=True, select_fn=select_stub)
    q.put((rank, res["index"], float(res["best"].mean()), res["scores"].tolist(), float(res["mean"].mean())))
    dist.destroy_process_group()


def test_mc_infer_gloo_world_size_2(bem):
    """ragged shards (4 + 3 samples), all_gather of scores in global order, first-max selection, winner broadcast,
    Monte-Carlo mean all_reduce — on the gloo backend"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vals = [((i * 37) % 11) / 10.0 for i in range(7)]
    want = vals.index(max(vals))
    for rank, index, best, scores, mean in got:
        assert index == want
        assert abs(best - vals[want]) < 1e-6
        assert all(abs(a - b) < 1e-6 for a, b in zip(scores, vals))
        assert abs(mean - sum(vals) / 7) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# round 2: patch install / uninstall on the staged reference, DDP buffer behaviour, cache generation, resume side-state
# ---------------------------------------------------------------------------------------------------------------------
def test_patch_install_and_uninstall_on_the_staged_reference(bem):
    """bem_b200.patch.install() against the REAL reference modules (oracle/_ref, staged by oracle/make_ref.py): every name
    SURVEY 8(b) lists is replaced in both import spellings of vmamba, and uninstall() restores the reference"""
    from oracle import ref_loader as R
    if not R.available():
        pytest.skip(R.why_unavailable())
    vm = R.vmamba(False)
    refb = R.bayesian()
    cs = sys.modules["basicsr.vmamba.models.csms6s"]
    ct = sys.modules["basicsr.vmamba.models.csm_triton"]
    before = (vm.selective_scan_fn, vm.cross_scan_fn, vm.cross_merge_fn, vm.SS2D.forward_corev2, cs.selective_scan_fn,
              cs.SelectiveScanCuda, ct.cross_scan_fn, sys.modules["bayesian"], vm.LayerNorm2d.forward)
    names = bem.patch.install()
    try:
        assert {"basicsr.vmamba.models.vmamba", "basicsr.vmamba.models.csms6s", "basicsr.vmamba.models.csm_triton", "bayesian"} <= set(names)
        assert vm.selective_scan_fn is bem.selective_scan_fn and cs.selective_scan_fn is bem.selective_scan_fn
        assert vm.cross_scan_fn is bem.cross_scan_fn and ct.cross_merge_fn is bem.cross_merge_fn
        assert cs.SelectiveScanCuda is bem.SelectiveScanCuda and cs.WITH_SELECTIVESCAN_OFLEX is True
        assert vm.SS2D.forward_corev2 is bem.ss2d.forward_corev2_patched
        assert vm.LayerNorm2d.forward is bem.layernorm.layernorm2d_forward_patched
        # tensors the kernels do not take (here: CPU) go through the reference's own permute / F.layer_norm / permute
        ln = vm.LayerNorm2d(6)
        xin = torch.randn(2, 6, 5, 4)
        want = torch.nn.functional.layer_norm(xin.permute(0, 2, 3, 1), (6,), ln.weight, ln.bias, ln.eps).permute(0, 3, 1, 2)
        assert torch.equal(ln(xin), want)
        assert sys.modules["bayesian"] is bem.bayesian
        # the reference's own model code, built AFTER the patch, is converted by this package's layers
        unet = R.unet_arch(False)
        net = unet.Network(stage=1, n_feat=8, num_blocks=[1, 1, 1], d_state=[1, 1, 1], ssm_ratio=1, mlp_ratio=4, mlp_type="gdmlp",
                           use_pixelshuffle=True)
        sys.modules["bayesian"].convert2bnn_selective(net, {"sigma_init": 0.05, "decay": 0.998, "pretrain": False})
        kinds = {type(m).__module__.split(".")[0] for m in net.modules() if hasattr(m, "deterministic")}
        assert kinds == {"bem_b200"}
        with pytest.raises(RuntimeError):      # ... and runs on the CUDA kernels only
            net(torch.rand(1, 3, 16, 16))
    finally:
        bem.patch.uninstall()
    after = (vm.selective_scan_fn, vm.cross_scan_fn, vm.cross_merge_fn, vm.SS2D.forward_corev2, cs.selective_scan_fn,
             cs.SelectiveScanCuda, ct.cross_scan_fn, sys.modules["bayesian"], vm.LayerNorm2d.forward)
    assert all(a is b for a, b in zip(before, after)) and sys.modules["bayesian"] is refb


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.nn as nn
    sys.path.insert(0, ROOT)
    import bem_b200
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)            # different init per rank: DDP must make rank 0's state the common one

    class Probe(nn.Module):                  # forward touches parameters and buffers without launching a CUDA kernel
        def __init__(self):
            super().__init__()
            self.conv = bem_b200.bayesian.Conv2dReparameterization(3, 4, 1, bias=True)
            self.lin = bem_b200.bayesian.Linear2dReparameterization(4, 4, bias=False)

        def forward(self, x):
            c, l = self.conv, self.lin
            return x * (c.mu_weight.sum() + c.rho_bias.sum() + l.mu_weight.sum()) + c.prior_mu_weight.sum() + l.prior_rho_weight.sum()

    m = Probe()
    names = sorted(n for n, _ in m.named_buffers())
    ddp = nn.parallel.DistributedDataParallel(m)      # base_model.py:97-100: find_unused_parameters as configured, buffers broadcast
    mu0 = m.conv.mu_weight.detach().clone()           # construction broadcasts rank 0's parameters AND buffers
    prior0 = m.conv.prior_mu_weight.detach().clone()
    with torch.no_grad():                              # rank-local drift of a prior buffer and of eps before the next forward
        m.conv.prior_mu_weight.add_(float(rank + 1))
        m.conv.prior_rho_weight = m.conv.prior_rho_weight + float(rank)   # re-ASSIGNED like conv.py:96-97: must stay a registered buffer
    still_buffer = "conv.prior_rho_weight" in dict(m.named_buffers())
    out = ddp(torch.ones(1))                            # broadcast_buffers=True (default): rank 0's buffers before every forward
    out.sum().backward()
    g = m.conv.mu_weight.grad.detach().clone()
    q.put((rank, names, mu0.tolist(), prior0.tolist(), m.conv.prior_mu_weight.tolist(), m.conv.prior_rho_weight.tolist(),
           still_buffer, g.tolist(), sorted(k for k in m.state_dict().keys())))
    dist.destroy_process_group()


def test_ddp_wraps_the_bayesian_layers_like_the_reference_gloo_world_size_2(bem):
    """stock DistributedDataParallel around this package's layers — parameters and the non-persistent
    `prior_*` / `eps_*` buffers are broadcast from rank 0 at construction and (broadcast_buffers default) before every forward,
    also after a buffer was re-assigned (conv.py:96-97 does `self.prior_mu_weight = ...` each step); gradients are averaged;
    the checkpoint keys hold none of the buffers."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29811 + os.getpid() % 150
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, names0, mu_a, prior_a, pm_a, pr_a, sb_a, g_a, keys_a), (r1, names1, mu_b, prior_b, pm_b, pr_b, sb_b, g_b, keys_b) = got
    want = ["conv.eps_bias", "conv.eps_weight", "conv.prior_mu_bias", "conv.prior_mu_weight", "conv.prior_rho_bias",
            "conv.prior_rho_weight", "lin.eps_weight", "lin.prior_mu_weight", "lin.prior_rho_weight"]
    assert names0 == names1 == want                              # the buffer names DDP broadcasts (conv.py:39-52, linear.py:26-39)
    assert mu_a == mu_b and prior_a == prior_b                   # rank 0's init everywhere after construction
    assert sb_a and sb_b                                         # re-assignment keeps the tensor in _buffers
    assert pm_a == pm_b and pr_a == pr_b                         # ... and the next forward re-broadcasts rank 0's values
    assert g_a == g_b                                            # averaged gradient
    assert keys_a == keys_b == ["conv.mu_bias", "conv.mu_weight", "conv.rho_bias", "conv.rho_weight", "lin.mu_weight", "lin.rho_weight"]


def test_cache_generation_covers_data_writes(bem):
    """`.data` writes (the reference's EMA update, base_model.py:84) do not bump tensor._version; the derived-weight
    caches therefore also key on the package's cache generation, advanced by train() / load_state_dict / _apply and by
    bem_b200.invalidate_caches()"""
    L = bem._lib
    layer = bem.bayesian.Conv2dReparameterization(4, 4, 1, bias=True)
    s0 = layer._sigma_cached().clone()
    v = layer.rho_weight._version
    layer.rho_weight.data.add_(1.0)                    # EMA-style write: version unchanged
    assert layer.rho_weight._version == v
    assert torch.equal(layer._sigma_cached(), s0)      # ... so the cache cannot see it by itself
    g = L.cache_generation()
    bem.invalidate_caches()
    assert L.cache_generation() == g + 1
    s1 = layer._sigma_cached()
    assert torch.allclose(s1, torch.log1p(torch.exp(layer.rho_weight.detach()))) and not torch.equal(s1, s0)
    for fn in (lambda: layer.train(), lambda: layer.eval(), lambda: layer.load_state_dict(layer.state_dict()), lambda: layer.float()):
        g = L.cache_generation()
        fn()
        assert L.cache_generation() > g
    net = bem.network.build_model()
    g = L.cache_generation()
    net.load_state_dict(net.state_dict())
    assert L.cache_generation() > g


def test_mc_config_scope_restores_the_layers(bem):
    """a sampler's Philox configuration must not leak into later direct forwards of the shared layers"""
    net = bem.network.build_bayesian_model()
    layers = bem.bayesian.bayesian_layers(net)
    layers[3].eps_source, layers[3].mc_seed = "torch", 5
    with bem.mc._McConfigScope(net):
        bem.bayesian.set_mc_config(net, mc_samples=4, eps_source="philox", seed=9, sample0=17)
        assert layers[0].eps_source == "philox" and layers[3].mc_seed == 9 and layers[0].mc_sample0 == 17
    assert layers[0].eps_source == "torch" and "eps_source" not in layers[0].__dict__      # class default again
    assert layers[3].eps_source == "torch" and layers[3].mc_seed == 5
    assert all(l.mc_samples == 1 and l.mc_sample0 == 0 for l in layers)


def test_prior_state_side_dict_closes_the_resume_gap(bem):
    """priors / step are not in the checkpoint (reference behaviour, keys unchanged); the side dict restores them"""
    B = bem.bayesian
    a = B.Conv2dReparameterization(3, 3, 1, bias=True)
    with torch.no_grad():
        a.prior_mu_weight.add_(0.5)
        a.prior_rho_bias.add_(-0.25)
    a.step = 41
    side = B.prior_state_dict(torch.nn.Sequential(a))
    assert set(side) == {"0.prior_mu_weight", "0.prior_rho_weight", "0.prior_mu_bias", "0.prior_rho_bias", "0.step"}
    b = B.Conv2dReparameterization(3, 3, 1, bias=True)
    nb = torch.nn.Sequential(b)
    nb.load_state_dict(torch.nn.Sequential(a).state_dict(), strict=True)      # the checkpoint proper: mu / rho only
    assert not torch.equal(b.prior_mu_weight, a.prior_mu_weight) and b.step == 0
    assert B.load_prior_state_dict(nb, side) == ["0"]
    assert torch.equal(b.prior_mu_weight, a.prior_mu_weight) and torch.equal(b.prior_rho_bias, a.prior_rho_bias) and b.step == 41
    assert torch.allclose(b.prior_sigma_bias, torch.log1p(torch.exp(a.prior_rho_bias)))


def test_integration_stub_declares_the_whole_scan_struct(bem):
    """the ctypes stub of INTEGRATION.md must end where include/bem_b200.h's struct ends (a short struct makes the
    library read dt_rank / dt_weight from whatever follows it)"""
    import ctypes
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class BemScanFwdParams\(ctypes\.Structure\):.*?\n(?=assert lib|lib\.)", text, flags=re.S)
    assert m, "INTEGRATION.md no longer contains the BemScanFwdParams stub"
    ns = {"ctypes": ctypes}
    exec(m.group(0), ns)
    doc = ns["BemScanFwdParams"]
    ours = bem._lib.BemScanFwdParams
    assert [f[0] for f in doc._fields_] == [f[0] for f in ours._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(ours)
    assert f"bem_abi_version() == {bem._lib.ABI_VERSION}" in text
