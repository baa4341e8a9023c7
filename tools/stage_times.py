"""Stage times of the forward at the bench shape (CUDA events around CUDA-graph replays of each C-ABI sub-path alone):
lifter (pose stream), GRU -> y[T//2] (image-feature stream), decoder (GRU + AdaLN gamma/beta + 3 blocks + mesh), whole forward.
Shows how much of the image-feature stream hides under the lifter."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from pmce_b200 import synth  # noqa: E402


def timed(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.B_PER_GPU
    dev = torch.device("cuda")
    model, sd = bench.build_model(dev)
    eng = model.engine()
    eng.use_graph = False
    p2d, feat = [t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=3)]
    joints = torch.randn(B, bench.J, 3, device=dev) * 0.3
    g = eng.gru_mid(feat)
    out = {"B": B,
           "lifter_us": timed(lambda: eng.lifter(p2d, feat)),
           "gru_mid_us": timed(lambda: eng.gru_mid(feat)),
           "adaln_gammabeta_us": timed(lambda: eng.adaln_gammabeta(g)),
           "decoder_us": timed(lambda: eng.decoder(joints, feat)),
           "forward_us": timed(lambda: model(p2d, feat))}
    out["decoder_minus_gru_us"] = out["decoder_us"] - out["gru_mid_us"]
    out["serial_sum_us"] = out["lifter_us"] + out["decoder_us"]
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
