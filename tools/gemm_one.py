"""Run the tcgen05 GEMM on one shape a few times (for `ncu -k regex:linear_tc_kernel`). Usage: gemm_one.py M N K act"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmce_b200 import _lib  # noqa: E402

M, N, K, act = [int(v) for v in sys.argv[1:5]]
lib = _lib.load()
dev = torch.device("cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
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
for _ in range(4):
    if act == 1:
        assert lib.pmce_linear_tc_presplit(P(xs[0]), P(xs[1]), P(ws[0]), P(ws[1]), P(b), M, N, K, 1, Z, P(oh[0]), P(oh[1]), Z, st) == 0
    else:
        assert lib.pmce_linear_tc_presplit(P(xs[0]), P(xs[1]), P(ws[0]), P(ws[1]), P(b), M, N, K, 0, P(o), Z, Z, P(o) if act == 2 else Z, st) == 0
torch.cuda.synchronize()
print("ok")
