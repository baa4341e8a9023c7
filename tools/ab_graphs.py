"""Same-process A/B of live environment knobs (PMCE_PDL, PMCE_PDL_WPRE, PMCE_ATTN_FEWQ): one CUDA graph of the whole forward is
captured per configuration, then the graphs are replayed INTERLEAVED (A B C A B C ...) and timed with CUDA events, so every
configuration sees the same clocks / thermal state. Usage: ab_graphs.py [B] CONF [CONF ...]   with CONF = NAME=V[,NAME=V]"""
import json
import os
import statistics
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from pmce_b200 import synth  # noqa: E402

args = sys.argv[1:]
B = int(args.pop(0)) if args and args[0].isdigit() else bench.B_PER_GPU
confs = args or ["PMCE_PDL=0", "PMCE_PDL=5"]
labels = [f"{i}:{c}" for i, c in enumerate(confs)]          # the same configuration may be listed twice (graph-placement noise)
dev = torch.device("cuda")
model, sd = bench.build_model(dev)
model.engine().use_graph = False
p2d, feat = [t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=3)]
graphs = []
for lab, c in zip(labels, confs):
    kv = dict(x.split("=") for x in c.split(","))
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv)
    for _ in range(3):
        model(p2d, feat)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        out = model(p2d, feat)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    graphs.append((lab, g, [o.clone() for o in out]))
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
ref = graphs[0][2]
same = {c: all(torch.equal(a, b) for a, b in zip(o, ref)) for c, _, o in graphs}
times = {c: [] for c in labels}
n = 20
for r in range(30):
    for c, g, _ in graphs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        times[c].append(e0.elapsed_time(e1) / n * 1e3)
print(json.dumps({"B": B, "forward_us_median": {c: round(statistics.median(t), 1) for c, t in times.items()},
                  "forward_us_min": {c: round(min(t), 1) for c, t in times.items()}, "bit_identical_to_first": same}))
