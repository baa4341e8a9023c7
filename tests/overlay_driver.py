"""Child process of tests/test_overlay.py: executes INTEGRATION.md's overlay recipe for real.

A scratch copy of the reference's `lib/` is made (from /root/reference, or from oracle/_ref on the GPU box), the four model
files are overwritten exactly as INTEGRATION.md's shell loop does, `lib/` goes on sys.path the way the reference's
main/__init_path.py:14-26 does, and then the CALLER's code runs: `import models`, `models.PMCE.get_model(...)`
(lib/core/base.py:54), `load_state_dict` (:67), `.cuda()`, and the test-loop body (:218-227).
usage: overlay_driver.py <scratch_dir> <cpu|gpu>
"""
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scratch, mode = sys.argv[1], sys.argv[2]
src = os.environ.get("PMCE_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(src, "lib", "models")):
    src = os.path.join(REPO, "oracle", "_ref")
assert os.path.isdir(os.path.join(src, "lib", "models")), "no reference lib/ to overlay (run oracle/build_ref.py in the build container)"

shutil.copytree(os.path.join(src, "lib"), os.path.join(scratch, "lib"), ignore=shutil.ignore_patterns("__pycache__"))
os.makedirs(os.path.join(scratch, "experiment"))
os.makedirs(os.path.join(scratch, "data", "Human36M"))
shutil.copy(os.path.join(src, "data", "Human36M", "J_regressor_h36m_correct.npy"), os.path.join(scratch, "data", "Human36M"))
# "overlay the four model files (the only files that change)"
for f in ("PMCE", "PoseEstimation", "CoevoDecoder", "project_net"):
    with open(os.path.join(scratch, "lib", "models", f + ".py"), "w") as fh:
        fh.write("from pmce_b200.models.%s import *  # noqa\nfrom pmce_b200.models.%s import get_model  # noqa\n" % (f, f))

sys.path.insert(0, REPO)                                   # PYTHONPATH=/path/to/pmce-b200
sys.path.insert(0, os.path.join(REPO, "oracle", "shims"))  # easydict / matplotlib, absent from this image (the reference's own deps)
from pmce_b200 import synth                                # noqa: E402
synth.write_mesh_assets(scratch, seed=7)                   # data/base_data/* (absent from the reference checkout as well)
os.chdir(scratch)                                          # the reference resolves assets relative to the cwd
sys.path.insert(0, os.path.join(scratch, "lib"))           # main/__init_path.py:16-17

import numpy as np                                         # noqa: E402
import torch                                               # noqa: E402
import io, contextlib                                      # noqa: E402,E401
with contextlib.redirect_stdout(io.StringIO()):
    from core.config import cfg                            # the REFERENCE's config module (lib/core/config.py), unmodified
import models                                              # noqa: E402  the reference's lib/models/__init__.py:1-4, unmodified

import pmce_b200.config as pcfg                            # noqa: E402
assert pcfg.cfg is cfg, "the overlay must pick up the reference's global cfg"
assert models.PMCE.__file__.startswith(scratch) and models.PMCE.PMCE.__module__ == "pmce_b200.models.PMCE"

g = np.load(os.path.join(REPO, "tests", "golden", "pmce_J17_C256_T16_B2.npz"))
J, C, depth, T, B = [int(v) for v in g["config"]]
cfg.DATASET.seqlen = T                                     # what a YAML override does (core/config.py:107-121)
model = models.PMCE.get_model(J, cfg.MODEL.hpe_dim, cfg.MODEL.hpe_dep)        # core/base.py:54
sd = synth.make_state_dict(int(g["weight_seed"]), init_vertices=g["init_vertices"], lifter_out_scale=float(g["lifter_out_scale"]),
                           num_joint=J, embed_dim=C, depth=depth, seqlen=T)
assert sorted(model.state_dict().keys()) == sorted(sd.keys())
model.load_state_dict(sd)                                  # core/base.py:67
out = {"keys": len(sd), "vj_equal": bool(np.array_equal(np.asarray(model.pose_mesh_coevo.vj_relation).astype(np.int64), g["vj_relation"]))}
if mode == "gpu":
    J_regressor = torch.Tensor(np.load("data/Human36M/J_regressor_h36m_correct.npy")).cuda()     # core/base.py:196
    model = model.cuda()                                   # core/base.py:199
    model.eval()                                           # :208
    p2d, feat = synth.make_inputs(B, T, J, seed=int(g["input_seed"]))
    with torch.no_grad():                                  # :215-227
        input_pose, input_feat = p2d.cuda(), feat.cuda()
        pred_mesh, evo_pose, lift_pose3d = model(input_pose, input_feat)
        pred_mesh = pred_mesh * 1000
        pred_pose = torch.matmul(J_regressor[None, :, :], pred_mesh)
    out["mesh_err"] = float((pred_mesh.cpu() / 1000 - torch.as_tensor(g["cam_mesh"])).abs().max())
    out["pose_err_mm"] = float((pred_pose.cpu() - torch.as_tensor(g["pred_pose_h36m"])).abs().max())
    out["evo_err"] = float((evo_pose.cpu() - torch.as_tensor(g["cam_pose"])).abs().max())
print("OVERLAY_RESULT " + json.dumps(out))
