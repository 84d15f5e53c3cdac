"""Experiment: two MCSamplers (own arena, graph, scratch) replaying on two streams vs one lane. `python tools/two_lanes.py [lanes]`"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bem_b200 import mc, network

torch.manual_seed(0)
dev = torch.device("cuda")
net = network.build_bayesian_model().to(dev).eval()
x = torch.rand(1, 3, 400, 600, device=dev)
N = 60
for lanes in (1, 2, 3):
    samplers = [mc.MCSampler(net, seed=1, arena=True, graph=True) for _ in range(lanes)]
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    recs = []
    with torch.no_grad():
        for s in samplers:
            y = s.sample(x, [0])          # builds arena, plan, graph
            recs.append(s._graphs[(tuple(x.shape), x.dtype, x.device)])
    ref = samplers[0].sample(x, [5])
    torch.cuda.synchronize()
    outs = [None] * lanes
    for rep in range(2):
        t0 = time.perf_counter()
        for i in range(N):
            l = i % lanes
            with torch.cuda.stream(streams[l]):
                samplers[l]._arena.sample0.fill_(5)
                recs[l][0].replay()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    ok = all(torch.equal(recs[l][2], ref) for l in range(lanes))
    print(f"lanes {lanes}: {N / dt:.1f} samples/s   outputs equal to single-lane result: {ok}")
