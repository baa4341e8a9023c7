"""Materialise `oracle/_ref/`: a runnable copy of the reference's CPU implementation of the hot path.  TEST INFRASTRUCTURE.

    python oracle/build_ref.py            (in the build container; needs /root/reference)

`/root/reference` does not exist on the GPU box, so the reference arm of `bench.py` (`--impl reference`) could only time the
oracle port there. This recipe copies the reference's own `lib/` package (the files `PMCE.forward` imports:
lib/models/{PMCE,PoseEstimation,CoevoDecoder,project_net}.py, lib/models/backbones/*, lib/core/config.py, lib/graph_utils.py,
lib/funcs_utils.py) and `smplpytorch/` UNMODIFIED into `oracle/_ref/`, adds the two shipped joint regressors,
seeded synthetic `data/base_data` assets and an `experiment/` directory (lib/core/config.py:20-38 creates its run
directories at import). `oracle/_ref/` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it
travels to the GPU box next to the built `.so`. `oracle/ref_harness.py` imports from it when `/root/reference` is absent.
Nothing in `pmce_b200/` reads it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
DEST = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("PMCE_REFERENCE_ROOT", "/root/reference")
COPY_TREES = ["lib", "smplpytorch"]
COPY_FILES = ["data/Human36M/J_regressor_h36m_correct.npy", "data/COCO/J_regressor_coco.npy"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()[:16]


def build(force=False, asset_seed=7):
    """-> DEST, or None when the reference tree is not present (the GPU box: use what travelled)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "models")):
        return DEST if os.path.isfile(os.path.join(DEST, "MANIFEST.json")) else None
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if os.path.isfile(manifest_path) and not force:
        return DEST
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from pmce_b200 import synth
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "experiment")
    for t in COPY_TREES:
        shutil.copytree(os.path.join(REFERENCE_ROOT, t), os.path.join(DEST, t), symlinks=False, ignore=ignore)
    for f in COPY_FILES:
        os.makedirs(os.path.dirname(os.path.join(DEST, f)), exist_ok=True)
        shutil.copy2(os.path.join(REFERENCE_ROOT, f), os.path.join(DEST, f))
    os.makedirs(os.path.join(DEST, "experiment"), exist_ok=True)
    synth.write_mesh_assets(DEST, seed=asset_seed)            # data/base_data/{mesh_downsampling.npz, smpl_mean_vertices.npy}
    files = {}
    for t in COPY_TREES:
        for root, _, names in os.walk(os.path.join(DEST, t)):
            for n in sorted(names):
                p = os.path.join(root, n)
                files[os.path.relpath(p, DEST)] = _sha(p)
    for f in COPY_FILES:
        files[f] = _sha(os.path.join(DEST, f))
    # every copied file is byte-identical to the reference (checked against the source tree, recorded for the GPU box)
    for rel, h in files.items():
        assert _sha(os.path.join(REFERENCE_ROOT, rel)) == h, rel
    with open(manifest_path, "w") as f:
        json.dump({"source": REFERENCE_ROOT, "asset_seed": asset_seed, "unmodified_files": files}, f, indent=1, sort_keys=True)
    for root, dirs, names in os.walk(DEST):
        for d in dirs:
            os.chmod(os.path.join(root, d), 0o755)
        for n in names:
            os.chmod(os.path.join(root, n), 0o644)
    return DEST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
