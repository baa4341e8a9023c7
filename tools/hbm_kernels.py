"""The HBM-bound satellites of the path, each timed alone (CUDA events around CUDA-graph replays over rotating buffer sets that
exceed the 126 MB L2) and reported as achieved GB/s of ALGORITHMIC bytes against the measured copy peak:
  jregress (sparse J-regressor, core/base.py:225), eval_errors ((f)1 evaluation epilogue), SMPL LBS (smpl_layer.py:65-158).
Run under `ncu -k regex:<kernel>` for the DRAM-throughput counters (profiles/)."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import bench  # noqa: E402
from conftest import dense_regressor  # noqa: E402
from pmce_b200 import synth  # noqa: E402
from pmce_b200.engine import JRegressor  # noqa: E402
from pmce_b200.smpl_layer import SMPL_Layer  # noqa: E402


def timed_graph(fns, rounds=5, graph=True):
    """fns: list of zero-arg callables (one per buffer set). Returns seconds per call."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    if not graph:        # the call syncs with the host (SMPL_Layer.forward's `bool(norm == 0)` branches): plain event timing
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rounds):
            for f in fns:
                f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / (rounds * len(fns))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (rounds * len(fns))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    peaks = bench.load_peaks()
    dev = torch.device("cuda")
    V = 6890
    nsets = max(2, int(200e6 // (B * V * 3 * 4 * 2)) + 1)
    gen = torch.Generator(device=dev).manual_seed(1)
    mesh = [torch.randn(B, V, 3, device=dev, generator=gen) * 0.3 for _ in range(nsets)]
    gt = [m + 0.05 * torch.randn(B, V, 3, device=dev, generator=gen) for m in mesh]
    gtp = torch.randn(B, 17, 3, device=dev, generator=gen) * 300
    jr = JRegressor(dense_regressor("h36m"), "cuda")
    out = []

    def emit(o):
        o["frac_of_hbm_peak"] = o["gbs"] / peaks["hbm_gbs"]
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in o.items()}), flush=True)

    sec = timed_graph([lambda m=m: jr(m, scale=1000.0) for m in mesh])
    nnz = int(jr.vals.numel())
    byt = B * (nnz * 3 * 4 + 17 * 3 * 4)            # gathered vertex rows + output (the regressor itself is 1.3 KB, L2 resident)
    emit(dict(kernel="jregress_kernel", B=B, us=sec * 1e6, bytes=byt, gbs=byt / sec / 1e9,
                    note="latency-bound: %d gathered rows of 12 B per clip" % nnz))

    sec = timed_graph([lambda m=m, t=t: jr.eval_errors(m, t, gtp) for m, t in zip(mesh, gt)])
    byt = B * (2 * V * 3 * 4)
    emit(dict(kernel="jregress + eval_err_kernel + eval_mean_kernel (pmce_eval_errors)", B=B, us=sec * 1e6, bytes=byt, gbs=byt / sec / 1e9))

    layer = SMPL_Layer.from_buffers(synth.make_smpl_buffers(11)).cuda()
    pose, betas, trans = [t.cuda() for t in synth.make_smpl_inputs(B, seed=13)]
    sec = timed_graph([lambda: layer(pose, betas, trans) for _ in range(4)], graph=False)
    byt = B * (340 + 82968)                         # 340 B in + 82,968 B out per sample
    flops = B * 15.3e6
    emit(dict(kernel="smpl_pose_kernel + gemm_tn (blend shapes) + smpl_skin_kernel (smpl_lbs_forward)", B=B, us=sec * 1e6, bytes=byt,
                    gbs=byt / sec / 1e9, gflops=flops / sec / 1e9,
                    note="the v_posed intermediate (82,680 B/sample written + read) is extra traffic: 3x the algorithmic bytes"))


if __name__ == "__main__":
    main()
