"""Time the tcgen05 split-bf16 GEMM alone (CUDA events) over the forward's shapes; PMCE_TC_BN forces the tile width.
Usage: [PMCE_TC_BN=128] python tools/gemm_sweep.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmce_b200 import _lib  # noqa: E402

SHAPES = [(17408, 1024, 512, 1), (17408, 512, 1024, 2), (17408, 1536, 512, 0), (17408, 512, 512, 2),
          (17408, 512, 256, 1), (17408, 256, 512, 2), (17408, 768, 256, 0),
          (1024, 6144, 2048, 0), (27584, 256, 64, 1), (27584, 64, 256, 0), (27584, 64, 64, 0), (27584, 192, 64, 0),
          (64, 3072, 2048, 0), (64, 20670, 2048, 0), (192, 6890, 1296, 0)]


def main():
    if os.environ.get("PMCE_SWEEP_SHAPES"):      # "M,N,K,act;M,N,K,act;..."
        global SHAPES
        SHAPES = [tuple(int(v) for v in t.split(",")) for t in os.environ["PMCE_SWEEP_SHAPES"].split(";")]
    lib = _lib.load()
    dev = torch.device("cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())
    print("PMCE_TC_BN =", os.environ.get("PMCE_TC_BN", "auto"))
    for M, N, K, act in SHAPES:
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * 0.05
        b = torch.randn(N, device=dev)
        o = torch.empty(M, N, device=dev)
        xs = [torch.empty(M, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        ws = [torch.empty(N, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        assert lib.pmce_split_bf16(P(x), M, K, P(xs[0]), P(xs[1]), st) == 0
        assert lib.pmce_split_bf16(P(w), N, K, P(ws[0]), P(ws[1]), st) == 0
        oh = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        Z = C.c_void_p(0)
        if act == 1:      # fc1-style: GELU + split-bf16 output
            call = lambda: lib.pmce_linear_tc_presplit(P(xs[0]), P(xs[1]), P(ws[0]), P(ws[1]), P(b), M, N, K, 1, Z, P(oh[0]), P(oh[1]), Z, st)
        elif act == 2:    # proj/fc2-style: fp32 out with in-place residual
            call = lambda: lib.pmce_linear_tc_presplit(P(xs[0]), P(xs[1]), P(ws[0]), P(ws[1]), P(b), M, N, K, 0, P(o), Z, Z, P(o), st)
        else:
            call = lambda: lib.pmce_linear_tc_presplit(P(xs[0]), P(xs[1]), P(ws[0]), P(ws[1]), P(b), M, N, K, 0, P(o), Z, Z, Z, st)
        for _ in range(3):
            assert call() == 0, lib.pmce_last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        tf = 2.0 * M * N * K / us / 1e6
        print(f"M={M:6d} N={N:6d} K={K:5d} act={act}: {us:8.1f} us  {tf:7.1f} TFLOP/s algorithmic ({3 * tf:7.1f} MMA)")


if __name__ == "__main__":
    main()
