"""One SPIN feature extraction (16 frames = one clip) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pmce_b200 import synth  # noqa: E402
from pmce_b200 import build as _b  # noqa: E402
_b.build()
from pmce_b200.spin import HMR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = HMR()
m.load_state_dict(synth.make_spin_state_dict(17))
m = m.cuda()
x = synth.make_frames(B, 3).cuda()
for _ in range(3):
    m.feature_extractor(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.feature_extractor(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
