"""Whole-forward time at the bench shape (CUDA-graph replay, CUDA events): one number per process, for A/B runs of an environment
knob on the same box (tools/ab_pdl.sh)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from pmce_b200 import synth  # noqa: E402
from tools.stage_times import timed  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.B_PER_GPU
dev = torch.device("cuda")
model, sd = bench.build_model(dev)
model.engine().use_graph = False
p2d, feat = [t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=3)]
t = [timed(lambda: model(p2d, feat), n=100) for _ in range(3)]
print(json.dumps({"B": B, "forward_us": [round(x, 1) for x in t], "knobs": {k: v for k, v in os.environ.items() if k.startswith("PMCE_")}}))
