"""Probe: does keeping TWO forwards in flight (two workspaces, two streams, each with its own side stream) raise the throughput
over back-to-back forwards on one stream? The decoder third of a forward is latency-bound; the other forward's lifter could fill
it. Two model instances (same weights, separate engines / workspaces), one CUDA graph each; (a) graphs replayed alternately on ONE
stream, (b) graph A on stream 1 and graph B on stream 2, K pairs each; interleaved rounds, CUDA events.
Usage: two_in_flight.py [B] [N]   (N = engines / forwards in flight, default 2)"""
import json
import os
import statistics
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from pmce_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.B_PER_GPU
dev = torch.device("cuda")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
models = [bench.build_model(dev)[0] for _ in range(N)]
inputs = [[t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=3 + i)] for i in range(N)]
streams = [torch.cuda.Stream() for _ in range(N)]
graphs, outs = [], []
for m, (p2d, feat), s in zip(models, inputs, streams):
    m.engine().use_graph = False
    for _ in range(3):
        m(p2d, feat)
    torch.cuda.synchronize()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        o = m(p2d, feat)
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    graphs.append(g)
    outs.append([t.clone() for t in o])

main = torch.cuda.current_stream()
K = 10


def serial():
    for _ in range(K):
        for g in graphs:
            g.replay()


def overlapped():
    for s in streams:
        s.wait_stream(main)
    for _ in range(K):
        for g, s in zip(graphs, streams):
            with torch.cuda.stream(s):
                g.replay()
    for s in streams:
        main.wait_stream(s)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (N * K) * 1e3


for fn in (serial, overlapped):
    fn()
torch.cuda.synchronize()
t = {"serial": [], "overlapped": []}
for _ in range(20):
    t["serial"].append(timed(serial))
    t["overlapped"].append(timed(overlapped))
# the overlapped replays must leave the same results behind
ok = True
for m, (p2d, feat), ref in zip(models, inputs, outs):
    ok = ok and all(torch.equal(a, b) for a, b in zip(m(p2d, feat), ref))
print(json.dumps({"B": B, "in_flight": N, "us_per_forward_median": {k: round(statistics.median(v), 1) for k, v in t.items()},
                  "us_per_forward_min": {k: round(min(v), 1) for k, v in t.items()}, "results_unchanged": ok}))
