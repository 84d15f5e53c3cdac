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
    assert bem._lib.lib.bem_abi_version() == bem._lib.ABI_VERSION == 11
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
    res = bem_b200.mc.mc_infer(StubSampler(), x, 7, score_fn=score, monte_carlo_mean=True)
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
