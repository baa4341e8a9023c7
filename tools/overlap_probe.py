"""Does the image-feature stream really run BESIDE the lifter's GEMMs? Stream A: n x the lifter fc1 GEMM (M=17408, N=1024, K=512,
GELU + split output); stream B: pmce_gru_mid (PMCE_GRU_FEW_STEPS / PMCE_GRU_FEW choose few-CTA or 128-CTA steps). Prints A alone,
B alone, A || B (CUDA events around replays of one CUDA graph with both branches)."""
import ctypes as C
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from pmce_b200 import _lib, synth  # noqa: E402


def main():
    nA = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dev = torch.device("cuda")
    lib = _lib.load()
    model, _ = bench.build_model(dev)
    eng = model.engine()
    B = 64
    _, feat = [t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=3)]
    M, N, K = 17408, 1024, 512
    P = lambda t: C.c_void_p(t.data_ptr())
    Z = C.c_void_p(0)
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    xs = [torch.empty(M, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    ws = [torch.empty(N, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    oh = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    st0 = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.pmce_split_bf16(P(x), M, K, P(xs[0]), P(xs[1]), st0) == 0
    assert lib.pmce_split_bf16(P(w), N, K, P(ws[0]), P(ws[1]), st0) == 0
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)

    def run_a():
        with torch.cuda.stream(sa):
            st = C.c_void_p(sa.cuda_stream)
            for _ in range(nA):
                assert lib.pmce_linear_tc_presplit(P(xs[0]), P(xs[1]), P(ws[0]), P(ws[1]), P(b), M, N, K, 1, Z, P(oh[0]), P(oh[1]), Z, st) == 0

    def run_b():
        with torch.cuda.stream(sb):
            eng.gru_mid(feat)

    def timed(fns, reps=20):
        # one CUDA graph with the two branches forked off the capture stream: replay has no CPU launch cost in the way
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            sa.wait_stream(cap); sb.wait_stream(cap)
            for f in fns:
                f()
            cap.wait_stream(sa); cap.wait_stream(sb)
        torch.cuda.current_stream().wait_stream(cap)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) * 1e3 / reps, 1)

    run_a(); run_b(); torch.cuda.synchronize()
    out = {"few_steps": os.environ.get("PMCE_GRU_FEW_STEPS", "-"), "few_ctas": os.environ.get("PMCE_GRU_FEW", "-"),
           "A_gemms_us": timed([run_a]), "B_gru_us": timed([run_b]), "A||B_us": timed([run_b, run_a]), "A||B_us (A first)": timed([run_a, run_b])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
