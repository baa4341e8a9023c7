"""Launch the fused vertex cross-attention kernel alone a few times (for `ncu -k regex:ca_vertex_fused`), and time it with
CUDA events over rotating buffer sets larger than L2. Usage: ca_one.py [B] [J]"""
import ctypes as C
import os
import sys
import tempfile

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda")
model, sd = bench.build_model(dev)
eng = model.engine()
from pmce_b200 import _lib  # noqa: E402
lib = _lib.load()
peaks = bench.load_peaks()
r = bench.cross_attn_roofline(lib, eng, dev, peaks, B, nsets=max(4, min(24, 16 * 64 // B)))
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
