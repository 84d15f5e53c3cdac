"""Monte-Carlo multi-sample inference + best-sample selection — the sample loop and selection of
Enhancement/eval.py:199-222, 268-297, re-organised for one-process-per-GPU execution.

Reference behaviour: `num_samples` sequential batch-1 forwards of the stage-1 network, every Bayesian layer drawing a
fresh eps (eval.py:199-211); every prediction is scored, and the first index attaining the best score wins
(`lst.index(max(lst))`, eval.py:270-274); optional Monte-Carlo mean (eval.py:224-225).

Here: samples are independent given the input, so sample i is owned by rank i % world_size. With the
counter-based eps source ("philox") sample i gets the same weights whatever the world size or batching, so the selected
image does not depend on how many GPUs took part. The only exchange is an all_gather of one score per sample plus a
broadcast of the winner (NCCL on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, bayesian
from ._lib import lib


def select_best(scores: torch.Tensor, take_min: bool = False):
    """scores: (n,) float32 CUDA tensor -> (index, value) 0-d CUDA tensors; first extremum, Python `list.index` semantics
    including NaN handling (Enhancement/eval.py:270-274). Runs bem_select_best; no host sync."""
    _lib.require_cuda(scores)
    s = scores.reshape(-1).to(torch.float32).contiguous()
    idx = torch.empty((), dtype=torch.int32, device=s.device)
    val = torch.empty((), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(lib.bem_select_best(_lib.ptr(s), s.numel(), int(bool(take_min)), _lib.ptr(idx), _lib.ptr(val),
                                       _lib.stream_ptr(s.device)), "select_best")
    _lib.profile.launches += 1
    return idx, val


def shard_samples(num_samples: int, rank: int, world_size: int) -> List[int]:
    """global sample indices owned by `rank`: i with i % world_size == rank (100 samples on 8 ranks -> 13/12 split)."""
    return list(range(rank, num_samples, world_size))


def owner_of(sample: int, world_size: int) -> int:
    return sample % world_size


def default_score(pred: torch.Tensor) -> torch.Tensor:
    """A device-resident no-reference score per image, (S, 3, H, W) -> (S,): luminance contrast (std) of the clamped
    prediction penalised by clipping. Stand-in for the CLIP-IQA / NIQE scorers of eval.py:229-260, which are separate
    pretrained models outside the hot path; any callable with this shape contract can be passed instead."""
    p = pred.clamp(0, 1)
    lum = 0.299 * p[:, 0] + 0.587 * p[:, 1] + 0.114 * p[:, 2]
    clipped = ((lum <= 0.0) | (lum >= 1.0)).float().mean(dim=(1, 2))
    return lum.std(dim=(1, 2)) - 0.5 * clipped


class MCArena:
    """All sampled tensors of one Monte-Carlo draw in one buffer, filled by ONE kernel launch (bem_bayes_sample_batched)
    instead of one launch per tensor per layer (conv.py:105-111 runs once per layer per forward in the reference).

    Entry order = module order x (weight, bias); stream ids are the layers' own (2 * layer_id [+ 1]), so a draw gives
    every layer exactly the numbers its per-layer Philox path would. The sample index can be read from a device word,
    which is what lets a captured CUDA graph of the forward be replayed for any sample."""

    def __init__(self, net, seed: int, n_sets: int = 1):
        from .bayesian.base_layer import BaseLayer_
        bayesian.set_mc_config(net)                       # assigns layer ids (= Philox stream ids) in module order
        self.layers = [m for m in net.modules() if isinstance(m, BaseLayer_)]
        if not self.layers:
            raise RuntimeError("MCArena: the network has no bem_b200.bayesian layers")
        self.seed = int(seed)
        self.S = S = max(1, int(n_sets))                  # weight sets per draw: the S Monte-Carlo samples of one batched forward
        self.device = self.layers[0].mu_weight.device
        _lib.require_cuda(self.layers[0].mu_weight)
        tensors = []   # (layer, which, mu, rho, stream_id, offset)
        total = 0
        for L in self.layers:
            for which in ("weight", "bias"):
                if which == "bias" and not L.bias:
                    continue
                mu, rho = getattr(L, "mu_" + which), getattr(L, "rho_" + which)
                if mu.dtype != torch.float32 or not mu.is_contiguous() or not rho.is_contiguous():
                    raise RuntimeError("MCArena: parameters must be contiguous float32")
                tensors.append((L, which, mu, rho, 2 * int(L.layer_id) + (1 if which == "bias" else 0), total))
                total += ((mu.numel() + 3) // 4 * 4) * S    # the S sets of a tensor lie back to back: the view (S, *shape) is contiguous
        self.buffer = torch.empty(total, dtype=torch.float32, device=self.device)
        base = self.buffer.data_ptr()
        rows, blocks = [[] for _ in range(S)], []
        self._ptrs = []
        self.views = {}    # id(layer) -> {"weight": view, "bias": view} into self.buffer (several arenas may serve one network)
        for e, (L, which, mu, rho, sid, off) in enumerate(tensors):
            n = mu.numel()
            pad = (n + 3) // 4 * 4
            for sset in range(S):
                rows[sset].append([mu.data_ptr(), rho.data_ptr(), base + 4 * (off + sset * pad), n, sid])
            self._ptrs.append((mu, rho, mu.data_ptr(), rho.data_ptr()))
            for b0 in range(0, (n + 3) // 4, 256):
                blocks.append([e, b0])
            # true view of the S sets of this tensor, (S, *shape); contiguous whenever numel is a multiple of 4 (every BEM tensor)
            view = torch.as_strided(self.buffer, (S,) + tuple(mu.shape), (pad,) + tuple(mu.stride()), off)
            self.views.setdefault(id(L), {})[which] = view
            L.__dict__["_arena_views"] = self.views[id(L)]      # the most recent arena's views (introspection, tests)
        self.entries = [torch.tensor(r, dtype=torch.int64).to(self.device) for r in rows]   # one table per weight set
        self.blocks = torch.tensor(blocks, dtype=torch.int32).to(self.device)
        self.n_blocks = len(blocks)
        self.sample0 = torch.zeros(S, dtype=torch.int64, device=self.device)   # device-side sample index of every set (graph replay)
        self._net = net
        self._stamp = self._net_stamp()
        self.plans = {}   # input shape -> functional.PackPlan: the pack steps of all Bayesian 1x1 layers in one launch per draw

    def _net_stamp(self):
        """(address, version) of every parameter and buffer of the network + the package's cache generation: what a captured
        graph of the forward depends on besides its input (packed constant weights, -exp(A_logs), split fusion weights are
        baked in by address and were derived from these tensors)"""
        ts = list(self._net.parameters()) + list(self._net.buffers())
        return (_lib.cache_generation(), tuple((t.data_ptr(), t._version) for t in ts))

    def valid(self) -> bool:
        """False once any weight of the network was replaced or modified (optimizer step, load_state_dict, .to(), or a `.data`
        write followed by bem_b200.invalidate_caches()): the arena's tables, pack plans and graphs are then rebuilt"""
        return (all(mu.data_ptr() == pm and rho.data_ptr() == pr for mu, rho, pm, pr in self._ptrs)
                and self._stamp == self._net_stamp())

    def attach(self, on: bool = True):
        for L in self.layers:
            L._arena = self.views[id(L)] if on else None

    def set_samples(self, sample_ids):
        """global sample indices of the S sets of the next device-indexed draw (a short list is padded with its last id);
        stream-ordered fills, no host synchronisation"""
        ids = [int(i) for i in sample_ids]
        ids = ids + [ids[-1]] * (self.S - len(ids))
        if self.S == 1:
            self.sample0.fill_(ids[0])
        else:
            for j in range(self.S):
                self.sample0[j].fill_(ids[j])

    def draw(self, sample_id=None, plan=None):
        """fill the arena: set s gets global sample `sample_id + s` (an int) / `sample_id[s]` (a sequence); None = read the indices
        from the device words `self.sample0` (what a captured graph does). One launch per weight set.
        `plan`: a built PackPlan whose layers are packed from the fresh draw (one more launch)."""
        from .bayesian import functional as BF
        for sset in range(self.S):
            if sample_id is None:
                BF.sample_batched(self.entries[sset], self.blocks, self.n_blocks, self.seed, 0, self.sample0[sset:])
            else:
                sid = sample_id[min(sset, len(sample_id) - 1)] if isinstance(sample_id, (list, tuple)) else int(sample_id) + sset
                BF.sample_batched(self.entries[sset], self.blocks, self.n_blocks, self.seed, int(sid), None)
        if plan is not None:
            plan.run()

    def forward_planned(self, key, sample_id, fn):
        """draw + fn() with the pack steps of the Bayesian 1x1 layers batched: the first call for `key` (the input's shape)
        runs fn() as it is and records the plan, later calls pack everything right after the draw."""
        from .bayesian import functional as BF
        plan = self.plans.get(key)
        if plan is None:
            plan = BF.PackPlan((self.buffer.data_ptr(), self.buffer.data_ptr() + 4 * self.buffer.numel()))
            self.draw(sample_id)
            with plan.recording():
                y = fn()
            plan.build(self.device)
            self.plans[key] = plan
            return y
        self.draw(sample_id, plan)
        with plan.playing():
            return fn()


class _McConfigScope:
    """The Monte-Carlo controls live on the (shared) layers; a sampler sets them for its forwards and puts back what it found,
    so a direct `net(x)` afterwards — e.g. the training step that follows an in-loop validation — draws fresh eps again exactly
    as before (the reference always does `eps.normal_()`, conv.py:107)."""
    KEYS = ("eps_source", "mc_seed", "mc_sample0", "mc_samples")

    def __init__(self, net):
        from .bayesian.base_layer import BaseLayer_
        self.layers = [m for m in net.modules() if isinstance(m, BaseLayer_)]

    def __enter__(self):
        self.saved = [{k: m.__dict__[k] for k in self.KEYS if k in m.__dict__} for m in self.layers]
        return self

    def __exit__(self, *exc):
        for m, sv in zip(self.layers, self.saved):
            for k in self.KEYS:
                if k in sv:
                    setattr(m, k, sv[k])
                else:
                    m.__dict__.pop(k, None)      # back to the class default
        return False


class _Lane:
    """One in-flight sample of an MCSampler: its own weight arena, captured graphs and streams. The library's scratch buffers
    are per stream, so the graphs of two lanes replay concurrently without sharing anything but the (read-only) parameters."""

    def __init__(self):
        self.arena = None
        self.graphs = {}
        self.capture_stream = None
        self.stream = None


class MCSampler:
    """Draws Monte-Carlo predictions of a Bayesian network for one input.

    net        : network whose Bayesian layers come from bem_b200.bayesian (e.g. network.build_bayesian_model())
    seed       : Philox seed shared by all ranks
    batch      : Monte-Carlo samples evaluated per forward: every launch carries `batch` images and `batch` weight sets
                 (S-batched kernels), which amortises the ramp of the ~150 kernels of a forward; the arena then draws `batch`
                 sets per forward. Sample i's prediction does not depend on the batch it was computed in (bit for bit).
                 1 reproduces the reference loop
    out_index  : which element of the network's output list is the prediction (eval.py:200 uses [-1])
    arena      : draw all layers' weights with one launch per sample set (philox source); same numbers as without
    graph      : capture the forward of one sample as a CUDA graph per input shape and replay it (needs `arena`); the
                 reference's ~700 launches per sample are otherwise bound by the host
    lanes      : samples in flight at once (needs `graph`): each lane replays its own graph on its own stream, so the ramp
                 and tail of one sample's kernels (148 per sample, 5-100 us each) fill with the other's. Same numbers.
    """

    def __init__(self, net, seed: int = 287128, batch: int = 1, eps_source: str = "philox", out_index: int = -1,
                 post: Optional[Callable] = None, arena: bool = True, graph: bool = False, lanes: int = 1):
        self.net = net
        self.seed = seed
        self.batch = max(1, int(batch))
        self.eps_source = eps_source
        self.out_index = out_index
        self.post = post or (lambda y: torch.clamp(y, 0, 1))   # eval.py:201
        self.use_arena = bool(arena) and eps_source == "philox"
        self.use_graph = bool(graph) and self.use_arena
        self._lanes = [_Lane() for _ in range(max(1, int(lanes)) if self.use_graph else 1)]
        bayesian.set_prediction_type(net, deterministic=False)

    @property
    def _arena(self):
        return self._lanes[0].arena

    @property
    def _graphs(self):
        return self._lanes[0].graphs

    # ------------------------------------------------------------------------------------------------
    def _get_arena(self, lane=None):
        lane = lane or self._lanes[0]
        if lane.arena is None or not lane.arena.valid():
            lane.arena = MCArena(self.net, self.seed, self.batch)
            lane.graphs = {}
        return lane.arena

    def _forward_one(self, x):
        """one forward of `batch` Monte-Carlo samples of the image x (1, C, H, W) -> (batch, C_out, H, W): the image is repeated,
        image s meets weight set s in every Bayesian layer (mc_samples = batch)"""
        if self.batch > 1:
            x = x.expand(self.batch, -1, -1, -1).contiguous()
        y = self.net(x)
        y = y[self.out_index] if isinstance(y, (list, tuple)) else y
        return self.post(y)

    def _groups(self, ids):
        """the ids of a request in runs of `batch` (one run per forward)"""
        return [ids[i:i + self.batch] for i in range(0, len(ids), self.batch)]

    def _tail_sampler(self, n):
        """a sampler of the same network / seed whose forwards carry exactly n samples: the ragged last group of a request
        (13 samples per rank on 8 GPUs = 3 x 4 + 1) runs at its own size instead of padding a full batch"""
        tails = self.__dict__.setdefault("_tails", {})
        t = tails.get(n)
        if t is None:
            t = MCSampler(self.net, self.seed, batch=n, eps_source=self.eps_source, out_index=self.out_index, post=self.post,
                          arena=self.use_arena, graph=self.use_graph, lanes=1)
            tails[n] = t
        return t

    def _graph_for(self, x, lane=None):
        """(graph, static input, static output) of `lane` for inputs like x; captured with the lane's arena attached"""
        lane = lane or self._lanes[0]
        key = (tuple(x.shape), x.dtype, x.device)
        rec = lane.graphs.get(key)
        if rec is None:
            arena = self._get_arena(lane)
            static_x = x.clone()
            if lane.capture_stream is None:
                lane.capture_stream = torch.cuda.Stream(device=x.device)
            side = lane.capture_stream           # warm up on the stream that captures: the library's scratch buffers are per
            side.wait_stream(torch.cuda.current_stream(x.device))   # stream, so they are created (and zero-filled) here, not as graph nodes
            with torch.cuda.stream(side):
                for _ in range(2):               # the first of them records the pack plan of this shape
                    arena.forward_planned(key, None, lambda: self._forward_one(static_x))
            torch.cuda.current_stream(x.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=lane.capture_stream):
                static_y = arena.forward_planned(key, None, lambda: self._forward_one(static_x))
            rec = (g, static_x, static_y)
            lane.graphs[key] = rec
        return rec

    def _lane_graphs(self, x, n):
        """the first min(lanes, n) lanes with their graphs for inputs like x (captured on first use)"""
        lanes = self._lanes[: max(1, min(len(self._lanes), (n + self.batch - 1) // self.batch))]
        bayesian.set_mc_config(self.net, mc_samples=self.batch, eps_source="philox", seed=self.seed, sample0=0)
        recs = []
        for lane in lanes:
            arena = self._get_arena(lane)
            key = (tuple(x.shape), x.dtype, x.device)
            if key not in lane.graphs:
                arena.attach(True)
                try:
                    self._graph_for(x, lane)
                finally:
                    arena.attach(False)
            recs.append(lane.graphs[key])
            if lane.stream is None:
                lane.stream = torch.cuda.Stream(device=x.device)
        return lanes, recs

    def samples_to_host(self, x_host, out_host, sample_ids):
        with _McConfigScope(self.net):
            return self._samples_to_host(x_host, out_host, sample_ids)

    def sample_to_host(self, x_host, out_host, sample_id):
        with _McConfigScope(self.net):
            return self._sample_to_host(x_host, out_host, sample_id)

    def sample(self, x, sample_ids):
        """x: (1, C, H, W) -> (len(sample_ids), C_out, H, W), one prediction per global sample index. The layers' Monte-Carlo
        controls are put back afterwards (see _McConfigScope)."""
        with _McConfigScope(self.net):
            return self._sample(x, sample_ids)

    @torch.no_grad()
    def _samples_to_host(self, x_host: torch.Tensor, out_host: torch.Tensor, sample_ids: Sequence[int]):
        """len(sample_ids) predictions of one (pinned) host image into the rows of a (pinned) host buffer, the lanes working
        side by side: per sample the image goes H2D into the lane's graph input and the prediction D2H from its output.
        Returns after everything has landed."""
        ids = list(sample_ids)
        if self.batch > 1 and len(ids) % self.batch and self.use_arena:     # ragged last group: its own batch size
            cut = len(ids) - len(ids) % self.batch
            if cut:
                self._samples_to_host(x_host, out_host[:cut], ids[:cut])
            self._tail_sampler(len(ids) - cut)._samples_to_host(x_host, out_host[cut:], ids[cut:])
            return out_host
        groups = self._groups(ids)
        if not (self.use_graph and x_host.shape[0] == 1 and (len(self._lanes) > 1 or self.batch > 1) and len(ids) > 1):
            if self.batch > 1:      # one full group without the graph path
                dev = next(self.net.parameters()).device
                out_host.copy_(self._sample(x_host.to(dev, non_blocking=True), ids), non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                return out_host
            for i, sid in enumerate(ids):
                self._sample_to_host(x_host, out_host[i:i + 1], sid)
            return out_host
        dev = next(self.net.parameters()).device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            lanes, recs = self._lane_graphs(torch.empty(x_host.shape, dtype=x_host.dtype, device=dev), len(ids))
            for lane in lanes:
                lane.stream.wait_stream(cur)
            row = 0
            for gi, grp in enumerate(groups):
                lane, (g, static_x, static_y) = lanes[gi % len(lanes)], recs[gi % len(lanes)]
                with torch.cuda.stream(lane.stream):
                    static_x.copy_(x_host, non_blocking=True)
                    lane.arena.set_samples(grp)
                    g.replay()
                    out_host[row:row + len(grp)].copy_(static_y[:len(grp)], non_blocking=True)
                row += len(grp)
            for lane in lanes:
                lane.stream.synchronize()
        return out_host

    @torch.no_grad()
    def _sample_to_host(self, x_host: torch.Tensor, out_host: torch.Tensor, sample_id: int):
        """One prediction from a (pinned) host image into a (pinned) host buffer, on the current stream and without
        intermediate device copies when the graph path is active: H2D straight into the graph's input, replay, D2H straight
        from the graph's output. Returns after the result has landed in `out_host`."""
        if self.use_graph and x_host.shape[0] == 1:
            arena = self._get_arena()
            dev = arena.device
            key = (tuple(x_host.shape), x_host.dtype, dev)
            rec = self._graphs.get(key)
            if rec is None:
                bayesian.set_mc_config(self.net, mc_samples=self.batch, eps_source="philox", seed=self.seed, sample0=0)
                arena.attach(True)
                try:
                    rec = self._graph_for(x_host.to(dev))
                finally:
                    arena.attach(False)
            g, static_x, static_y = rec
            static_x.copy_(x_host, non_blocking=True)
            arena.set_samples([int(sample_id)])
            g.replay()
            static_y = static_y[:1]
            out_host.copy_(static_y[0] if out_host.dim() == static_y.dim() - 1 else static_y, non_blocking=True)
        else:
            dev = next(self.net.parameters()).device
            y = self._sample(x_host.to(dev, non_blocking=True), [sample_id])
            out_host.copy_(y[0] if out_host.dim() == y.dim() - 1 else y, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    @torch.no_grad()
    def _sample(self, x: torch.Tensor, sample_ids: Sequence[int]) -> torch.Tensor:
        outs = []
        ids = list(sample_ids)
        if self.use_arena and ids and x.is_cuda and x.shape[0] == 1:
            if self.batch > 1 and len(ids) % self.batch:                    # ragged last group: its own batch size
                cut = len(ids) - len(ids) % self.batch
                tail = self._tail_sampler(len(ids) - cut)._sample(x, ids[cut:])
                return torch.cat([self._sample(x, ids[:cut]), tail], dim=0) if cut else tail
            groups = self._groups(ids)
            if self.use_graph and len(self._lanes) > 1 and len(groups) > 1:
                cur = torch.cuda.current_stream(x.device)
                lanes, recs = self._lane_graphs(x, len(ids))
                for lane, (g, static_x, static_y) in zip(lanes, recs):
                    lane.stream.wait_stream(cur)
                    with torch.cuda.stream(lane.stream):
                        static_x.copy_(x)
                outs = [None] * len(groups)
                for gi, grp in enumerate(groups):
                    lane, (g, static_x, static_y) = lanes[gi % len(lanes)], recs[gi % len(lanes)]
                    with torch.cuda.stream(lane.stream):
                        lane.arena.set_samples(grp)
                        g.replay()
                        outs[gi] = static_y[:len(grp)].clone()
                for lane in lanes:
                    cur.wait_stream(lane.stream)
                for o in outs:
                    o.record_stream(cur)
                return torch.cat(outs, dim=0)
            arena = self._get_arena()
            bayesian.set_mc_config(self.net, mc_samples=self.batch, eps_source="philox", seed=self.seed, sample0=0)
            arena.attach(True)
            try:
                if self.use_graph:
                    g, static_x, static_y = self._graph_for(x)
                    static_x.copy_(x)
                    for grp in groups:
                        arena.set_samples(grp)
                        g.replay()
                        outs.append(static_y[:len(grp)].clone())
                else:
                    key = (tuple(x.shape), x.dtype, x.device)
                    for grp in groups:
                        full = [int(i) for i in grp] + [int(grp[-1])] * (self.batch - len(grp))
                        outs.append(arena.forward_planned(key, full, lambda: self._forward_one(x))[:len(grp)])
            finally:
                arena.attach(False)
            return torch.cat(outs, dim=0)
        i = 0
        while i < len(ids):
            # batch runs of consecutive-by-stride ids cannot share a Philox `sample0` unless contiguous; batch only
            # contiguous runs, otherwise go one at a time
            j = i + 1
            while j < len(ids) and j - i < self.batch and ids[j] == ids[j - 1] + 1:
                j += 1
            S = j - i
            bayesian.set_mc_config(self.net, mc_samples=S, eps_source=self.eps_source, seed=self.seed, sample0=ids[i])
            xin = x.expand(S, *x.shape[1:]).contiguous() if S > 1 else x
            y = self.net(xin)
            y = y[self.out_index] if isinstance(y, (list, tuple)) else y
            outs.append(self.post(y))
            i = j
        bayesian.set_mc_config(self.net, mc_samples=1, sample0=0)
        return torch.cat(outs, dim=0) if outs else x.new_empty((0,) + tuple(x.shape[1:]))


@torch.no_grad()
def mc_infer(sampler: MCSampler, x: torch.Tensor, num_samples: int, score_fn: Callable = default_score,
             take_min: bool = False, monte_carlo_mean: bool = False, group=None, select_fn: Callable = None):
    """Distributed MC inference for one image. Every rank calls this with the same x / num_samples.

    Returns dict(best=(C,H,W) tensor on every rank, index=int, scores=(num_samples,) tensor[, mean=(C,H,W)]).
    Single process (no initialised process group) = all samples local.
    select_fn(scores, take_min) -> (index, value): the selection operator, bem_select_best by default (CUDA only — the
    host-logic tests inject a stand-in; there is no CPU selection in this package).
    """
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    mine = shard_samples(num_samples, rank, world)
    preds = sampler.sample(x, mine)                                 # (len(mine), C, H, W)
    local_scores = score_fn(preds).to(torch.float32) if len(mine) else preds.new_empty((0,), dtype=torch.float32)

    # scores -> every rank, in GLOBAL sample order (ragged shards: pad to the largest shard)
    per = (num_samples + world - 1) // world
    padded = torch.full((per,), float("nan"), dtype=torch.float32, device=x.device)
    padded[: len(mine)] = local_scores
    if distributed:
        gathered = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(gathered, padded, group=group)
        table = torch.stack(gathered, dim=1).reshape(-1)[:num_samples]   # [j, r] -> sample j * world + r
    else:
        table = padded[:num_samples]
    idx_t, _ = (select_fn or select_best)(table, take_min)
    index = int(idx_t)
    owner = owner_of(index, world)
    if owner == rank:
        best = preds[mine.index(index)].clone()
    else:   # receive buffer: every rank knows the prediction shape from its own shard (or from x when it has none)
        shape = preds.shape[1:] if preds.numel() else x.shape[1:]
        best = torch.empty(shape, dtype=preds.dtype, device=x.device)
    if distributed:
        dist.broadcast(best, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    out = dict(best=best, index=index, scores=table)
    if monte_carlo_mean:
        acc = preds.sum(dim=0) if len(mine) else torch.zeros(x.shape[1:], dtype=x.dtype, device=x.device)
        if distributed:
            dist.all_reduce(acc, group=group)
        out["mean"] = torch.clamp(acc / num_samples, 0, 1)           # eval.py:224-225
    return out
