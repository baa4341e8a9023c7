"""Same-process interleaved A/B of live knobs on the PIPELINED device loop (models.PMCE.forward_iter, two forwards in flight): one
model instance (own engine, graphs captured under the configuration's environment) per configuration, the loops timed in
interleaved rounds. Usage: ab_iter.py [B] CONF [CONF ...]   with CONF = NAME=V[,NAME=V]"""
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
labels = [f"{i}:{c}" for i, c in enumerate(confs)]
dev = torch.device("cuda")
sets = [tuple(t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=100 + i)) for i in range(4)]
K = 20


def run(model):
    for _ in model.forward_iter(sets[i % 4] for i in range(K)):
        pass


models = []
for c in confs:
    kv = dict(x.split("=") for x in c.split(","))
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv)
    m = bench.build_model(dev)[0]
    run(m)                      # captures the two slot graphs under this environment
    torch.cuda.synchronize()
    models.append(m)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
times = {lab: [] for lab in labels}
for r in range(15):
    for lab, m in zip(labels, models):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(m)
        e1.record()
        torch.cuda.synchronize()
        times[lab].append(e0.elapsed_time(e1) / K * 1e3)
print(json.dumps({"B": B, "us_per_step_median": {c: round(statistics.median(t), 1) for c, t in times.items()},
                  "us_per_step_min": {c: round(min(t), 1) for c, t in times.items()}}))
