"""(f)3: one track of N frames, all stride-1 windows (nwin = N - T + 1) through PMCE.forward on the materialised windows vs
PMCE.forward_sliding (per-frame work shared between windows). CUDA events around CUDA-graph replays. Usage: sliding_bench.py [nwin]"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def timed(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        out = fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, out


def main():
    nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T, J = bench.T, bench.J
    N = nwin + T - 1
    dev = torch.device("cuda")
    model, _ = bench.build_model(dev)
    model.engine().use_graph = False
    gen = torch.Generator(device=dev).manual_seed(4)
    pose_seq = torch.randn(N, J, 2, device=dev, generator=gen)
    feat_seq = torch.randn(N, 2048, device=dev, generator=gen)
    idx = torch.arange(nwin, device=dev)[:, None] + torch.arange(T, device=dev)[None, :]
    p_w, f_w = pose_seq[idx].contiguous(), feat_seq[idx].contiguous()
    us_mat, ref = timed(lambda: model(p_w, f_w))
    us_sl, out = timed(lambda: model.forward_sliding(pose_seq, feat_seq, 1))
    err = max(float((o - r).abs().max()) for o, r in zip(out, ref))
    print(json.dumps({"windows": nwin, "frames": N, "C": bench.C, "materialised_us": round(us_mat, 1), "sliding_us": round(us_sl, 1),
                      "speedup": round(us_mat / us_sl, 3), "max_abs_diff": err,
                      "windows_per_s_materialised": round(nwin / us_mat * 1e6, 1), "windows_per_s_sliding": round(nwin / us_sl * 1e6, 1)}))


if __name__ == "__main__":
    main()
