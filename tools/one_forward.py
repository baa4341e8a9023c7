"""One eager PMCE.forward between cudaProfilerStart/Stop (for `ncu --profile-from-start off`): every kernel of ONE forward, in
launch order, at the headline workload (B=64 T=16 J=17 C=512) or `one_forward.py B`."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from pmce_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.B_PER_GPU
dev = torch.device("cuda")
model, sd = bench.build_model(dev)
model.engine().use_graph = False
p2d, feat = [t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=100)]
for _ in range(3):
    model(p2d, feat)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model(p2d, feat)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
