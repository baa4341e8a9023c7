"""Throughput of the other BASELINE.json configurations on one GPU (the headline config is bench.py's): full PMCE.forward at
B=256 (configs[2]), B=128 (the per-GPU share of configs[3]'s B=1024 over 8 GPUs) and the T=64 long-clip case at B=32 (configs[4]),
for C=256 and C=512. CUDA events around graph replays, inputs rotating over 4 resident sets. One JSON line per config."""
import json
import os
import sys
import tempfile

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pmce_b200 import synth  # noqa: E402

CONFIGS = [("configs[1] decoder+lifter B=64", 64, 16, 512), ("configs[2] B=256", 256, 16, 512), ("configs[2] B=256 C=256", 256, 16, 256),
           ("configs[3] per-GPU share B=128", 128, 16, 512), ("configs[4] T=64 B=32", 32, 64, 512), ("B=1 latency", 1, 16, 512)]


def main():
    root = tempfile.mkdtemp(prefix="pmce_cfg_")
    synth.prepare_data_root(root, os.path.join(REPO, "tests", "golden", "J_regressors_sparse.npz"))
    os.environ["PMCE_DATA_ROOT"] = root
    os.environ["PMCE_B200_STANDALONE_CFG"] = "1"
    from pmce_b200 import models
    from pmce_b200.config import cfg
    dev = torch.device("cuda")
    J = 17
    for name, B, T, C in CONFIGS:
        cfg.DATASET.seqlen = T
        m = models.PMCE.get_model(J, C, 3)
        sd = synth.make_state_dict(0, init_vertices=m.state_dict()["pose_mesh_coevo.init_vertices"].numpy(), lifter_out_scale=300.0,
                                   num_joint=J, embed_dim=C, depth=3, seqlen=T)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval()
        sets = [tuple(t.to(dev) for t in synth.make_inputs(B, T, J, seed=50 + i)) for i in range(4)]
        for i in range(4):
            out = m(*sets[i])
        assert all(torch.isfinite(o).all() for o in out)
        torch.cuda.synchronize()
        steps = 100 if B <= 64 else 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            m(*sets[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"config": name, "B": B, "T": T, "J": J, "C": C, "ms_per_step": round(ms, 4), "clips_per_s": round(B / ms * 1e3, 1),
                          "frames_per_s": round(B * T / ms * 1e3, 1)}))
        del m, sets, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
