"""Which weight groups of the image-feature stream / decoder could be stored as bf16 only (one `hi` copy: half the bytes the forward
streams from HBM / L2 per batch, 2 MMAs instead of 3)? (test-side tool, CPU: the oracle run with one group of weights rounded to
bf16, everything else fp32, against the reference's golden outputs.) The groups are the big once-per-batch weight streams: the GRU
(`weight_hh` 25 MB as split-bf16 - with `hi` only a CTA's slice would fit in shared memory for the whole layer -, `weight_ih`
100 MB), the three `linear_cur` (169 MB), the 24 AdaLN gamma/beta projections (25 MB), `upsample_conv` (36 MB), and the lifter's
`imgfeat_embed`. Adoption bar: max|d mesh| <= 2e-4 m on every fixture.   usage: python tools/weight_precision_map.py"""
import glob
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import pmce_oracle as po      # noqa: E402
from pmce_b200 import synth               # noqa: E402

GROUPS = {
    "gru weight_hh (l0+l1)": lambda k: "gru_cur.weight_hh" in k,
    "gru weight_hh_l0": lambda k: "gru_cur.weight_hh_l0" in k,
    "gru weight_hh_l1": lambda k: "gru_cur.weight_hh_l1" in k,
    "gru weight_ih (l0+l1)": lambda k: "gru_cur.weight_ih" in k,
    "linear_cur1-3": lambda k: "linear_cur" in k and k.endswith("weight"),
    "AdaLN mlp_gamma/beta": lambda k: (".mlp_gamma." in k or ".mlp_beta." in k) and k.endswith("weight"),
    "upsample_conv": lambda k: "upsample_conv.weight" in k,
    "imgfeat_embed": lambda k: "imgfeat_embed.weight" in k,
}


def run(path, pick):
    g = np.load(path)
    J, C, depth, T, B = [int(v) for v in g["config"]]
    sd = synth.make_state_dict(int(g["weight_seed"]), init_vertices=g["init_vertices"], lifter_out_scale=float(g["lifter_out_scale"]),
                               num_joint=J, embed_dim=C, depth=depth, seqlen=T)
    n = 0
    if pick is not None:
        for k in list(sd.keys()):
            if pick(k):
                sd[k] = sd[k].bfloat16().float()
                n += 1
    p2d, feat = synth.make_inputs(B, T, J, seed=int(g["input_seed"]))
    with torch.no_grad():
        mesh, pose, p3 = po.pmce_forward(sd, p2d, feat, g["vj_relation"])
    return (n, float((mesh - torch.as_tensor(g["cam_mesh"])).abs().max()), float((pose - torch.as_tensor(g["cam_pose"])).abs().max()))


if __name__ == "__main__":
    paths = [p for p in sorted(glob.glob(os.path.join(REPO, "tests", "golden", "pmce_*.npz"))) if "unitscale" not in p]
    print(f"{'fixture':28s} {'bf16-only weights':24s} {'tensors':>7s} {'max|d mesh| m':>14s} {'max|d cam_pose| m':>18s}")
    for path in paths:
        for name, pick in [("(none)", None)] + list(GROUPS.items()):
            n, e0, e1 = run(path, pick)
            print(f"{os.path.basename(path)[:28]:28s} {name:24s} {n:7d} {e0:14.2e} {e1:18.2e}", flush=True)
