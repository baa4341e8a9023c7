"""GPU (-m gpu): the tcgen05/TMA/TMEM GEMM (split-bf16 "bf16x3" operands) through its C-ABI entry point against a float64
matmul and against the fp32 CUDA-core GEMM, over the shapes the forward uses (incl. ragged M/N and K tails)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    (128, 128, 64), (128, 32, 64), (1, 64, 64), (300, 192, 256), (129, 200, 72),
    (17 * 16 * 4, 768, 256),      # lifter qkv, B=4
    (17 * 16 * 4, 512, 1024),     # lifter fc2 at C=512
    (1024, 6144, 2048),           # GRU layer-0 input projection, B=64
    (64, 3072, 2048),             # AdaLN gamma/beta, B=64
    (431 * 8, 64, 64), (431 * 8, 256, 64), (431 * 8, 64, 256), (431 * 8, 192, 64),   # co-evolution token GEMMs
    (192, 6890, 1296),            # upsample_conv as GEMM (K tail: 1296 = 20*64 + 16)
]


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 1])
def test_linear_tc_matches_fp64(lib, M, N, K, act):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    out_tc = torch.full((M, N), float("nan"), device="cuda")
    out_f32 = torch.empty(M, N, device="cuda")
    nbytes = lib.pmce_linear_tc_scratch_bytes(M, N, K)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.pmce_linear_tc(_p(x), _p(w), _p(b), M, N, K, act, _p(out_tc), _p(scratch), nbytes, st)
    assert rc == 0, lib.pmce_last_error()
    rc = lib.pmce_linear(_p(x), _p(w), _p(b), M, N, K, act, _p(out_f32), st)
    assert rc == 0, lib.pmce_last_error()
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t() + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    scale = float(ref.abs().max())
    e_tc = float((out_tc.double() - ref).abs().max()) / scale
    e_f32 = float((out_f32.double() - ref).abs().max()) / scale
    print(f"M={M} N={N} K={K} act={act}: rel err tc={e_tc:.2e} fp32-simt={e_f32:.2e}")
    assert torch.isfinite(out_tc).all()
    assert e_f32 < 2e-6
    assert e_tc < 3e-5          # bf16x3: ~2^-16 per product, far below single-pass TF32 (~5e-4)


def test_cta_pair_gemm_in_subprocess():
    """The cta_group::2 (CTA pair) variant of the GEMM on every shape above, forced on with 256-wide tiles (the library reads
    its knobs once per process, hence the subprocess)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PMCE_TC_PAIR="2", PMCE_TC_BN="256")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", "test_linear_tc_matches_fp64"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
