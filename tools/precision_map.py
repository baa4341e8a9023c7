"""Which of the lifter's projections need all three products of the split-bf16 scheme? (test-side tool, CPU, oracle emulation)
The forward runs at the box's POWER cap, so every MMA that can be dropped is clock headroom; a product A.W is computed as
A_lo.W_hi + A_hi.W_lo + A_hi.W_hi. For one class of lifter projection at a time (qkv / proj / fc1 / fc2, all six blocks) one
cross term is dropped - `w_hi`: the weights are bf16 (2 MMAs, half the weight bytes), `a_hi`: the activations are bf16 (2 MMAs,
half the activation bytes the producer writes and the GEMM reads), `both`: plain bf16 (1 MMA) - everything else stays at the
product path's precision, and the result is compared with the reference's golden outputs. Adoption bar as for TF32:
max|d mesh| and max|d pose3d|/1000 <= 2e-4 m on every fixture.   usage: python tools/precision_map.py"""
import glob
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import pmce_oracle as po      # noqa: E402
from pmce_b200 import synth               # noqa: E402


def split(x):
    h = x.bfloat16().float()
    return h, (x - h).bfloat16().float()


def product(x, w, mode):
    xh, xl = split(x)
    wh, wl = split(w)
    acc = xh.double() @ wh.double().t()
    if mode in ("full", "w_hi"):
        acc = acc + xl.double() @ wh.double().t()
    if mode in ("full", "a_hi"):
        acc = acc + xh.double() @ wl.double().t()
    return acc.float()


def run(path, cls, mode):
    g = np.load(path)
    J, C, depth, T, B = [int(v) for v in g["config"]]
    sd = synth.make_state_dict(int(g["weight_seed"]), init_vertices=g["init_vertices"], lifter_out_scale=float(g["lifter_out_scale"]),
                               num_joint=J, embed_dim=C, depth=depth, seqlen=T)
    p2d, feat = synth.make_inputs(B, T, J, seed=int(g["input_seed"]))
    orig = po._lin

    def lin(sd_, prefix, x):
        if not prefix.startswith("pose_lifter."):
            return orig(sd_, prefix, x)
        w, b = sd_[prefix + ".weight"], sd_[prefix + ".bias"]
        m = mode if prefix.endswith(cls) else "full"
        return product(x.reshape(-1, x.shape[-1]), w, m).reshape(*x.shape[:-1], -1) + b
    po._lin = lin
    try:
        with torch.no_grad():
            mesh, pose, p3 = po.pmce_forward(sd, p2d, feat, g["vj_relation"])
    finally:
        po._lin = orig
    return (float((mesh - torch.as_tensor(g["cam_mesh"])).abs().max()), float((p3 - torch.as_tensor(g["pose3d"])).abs().max()) / 1000.0)


if __name__ == "__main__":
    paths = [p for p in sorted(glob.glob(os.path.join(REPO, "tests", "golden", "pmce_*.npz"))) if "unitscale" not in p]
    print(f"{'fixture':28s} {'class':10s} {'mode':6s} {'max|d mesh| m':>14s} {'max|d pose3d|/1000 m':>21s}")
    for path in paths:
        e = run(path, ".none", "full")
        print(f"{os.path.basename(path)[:28]:28s} {'(all)':10s} {'full':6s} {e[0]:14.2e} {e[1]:21.2e}")
        for cls in (".attn.qkv", ".attn.proj", ".mlp.fc1", ".mlp.fc2"):
            for mode in ("w_hi", "a_hi", "both"):
                e = run(path, cls, mode)
                print(f"{os.path.basename(path)[:28]:28s} {cls:10s} {mode:6s} {e[0]:14.2e} {e[1]:21.2e}", flush=True)
