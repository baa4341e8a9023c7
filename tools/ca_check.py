"""Quick GPU check of the fused cross-attention kernel alone (B200): parity of pmce_cross_attn_block (vertex stream) against the
unfused launch sequence is covered by pytest; this prints timing at several batch sizes. Usage: ca_check.py [B ...]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402

dev = torch.device("cuda")
model, sd = bench.build_model(dev)
eng = model.engine()
from pmce_b200 import _lib  # noqa: E402
lib = _lib.load()
peaks = bench.load_peaks()
for B in [int(v) for v in sys.argv[1:]] or [64, 256, 1024]:
    r = bench.cross_attn_roofline(lib, eng, dev, peaks, B, nsets=max(4, min(24, 16 * 64 // B)))
    print(B, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("us_per_launch", "frac", "frac_kernel_io", "achieved")},
          "embed mode:", {k: round(v, 3) for k, v in r["embed_mode"].items() if k in ("us_per_launch", "frac")}, flush=True)
