"""Import the UNMODIFIED reference (`/root/reference`) on a CPU-only host.  TEST INFRASTRUCTURE.

In the build container it imports from `/root/reference`: this is what `oracle/gen_golden.py` uses to produce
`tests/golden/*.npz`, and what `tests/test_oracle_vs_reference.py` re-runs (skipped when the reference is absent) to pin
the CPU restatement in `oracle/pmce_oracle.py`. On the GPU box, where `/root/reference` does not exist, it imports from
`oracle/_ref/` - the byte-identical copy `oracle/build_ref.py` materialises (git-ignored, travels with the snapshot) - so
`bench.py --impl reference` times the reference module itself.

No reference source enters the repository. A scratch "view" directory of *symlinks* to the read-only tree is
built so that the reference's import-time `mkdir experiment/...` (lib/core/config.py:20-38, paths
derived from `os.path.abspath(__file__)`, which does not resolve symlinks) lands in the scratch
dir, and `data/base_data` (a dangling symlink in the reference) is replaced by seeded synthetic
assets from `pmce_b200.synth`. Missing third-party modules (`timm`, `easydict`, `matplotlib`) come
from `oracle/shims/`; hard-coded `.cuda()` calls (CoevoDecoder.py:200,207; backbones/mesh.py:62)
are neutralised on CPU-only hosts.
"""
import os
import sys
import tempfile

import torch

REFERENCE_ROOT = os.environ.get("PMCE_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
REF_COPY = os.path.join(_HERE, "_ref")          # materialised by oracle/build_ref.py; travels to the GPU box (git-ignored)
_state = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "models"))


def copy_available():
    return os.path.isfile(os.path.join(REF_COPY, "MANIFEST.json"))


def which():
    """'tree' (the read-only reference tree, through a symlink view), 'copy' (oracle/_ref) or None."""
    return "tree" if available() else ("copy" if copy_available() else None)


def _import_reference(view):
    """chdir into `view`, neutralise `.cuda()` on CPU-only hosts, put the reference on sys.path and import `models`."""
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    sys.path.insert(0, os.path.join(_HERE, "shims"))
    sys.path.insert(0, os.path.join(view, "smplpytorch"))
    sys.path.insert(0, os.path.join(view, "data"))
    sys.path.insert(0, os.path.join(view, "lib"))
    os.chdir(view)

    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        from core.config import cfg  # noqa: creates experiment dirs inside the view
    from models.backbones import mesh as ref_mesh
    if not torch.cuda.is_available():
        defaults = list(ref_mesh.Mesh.__init__.__defaults__)
        defaults[-1] = torch.device("cpu")
        ref_mesh.Mesh.__init__.__defaults__ = tuple(defaults)
    import models  # noqa: reference lib/models/__init__.py
    return cfg, models


def setup(asset_seed=7):
    """Build the view dir (or use oracle/_ref), chdir into it, patch `.cuda()` and put the reference on sys.path."""
    if _state:
        return _state
    if not available():
        if not copy_available():
            raise RuntimeError("reference tree not found at %s and no oracle/_ref copy (python oracle/build_ref.py)" % REFERENCE_ROOT)
        # the materialised copy: real, writable files (the reference's import-time mkdirs land inside it)
        cfg, models = _import_reference(REF_COPY)
        _state.update(view=REF_COPY, cfg=cfg, models=models, assets=None, kind="copy")
        return _state
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    from pmce_b200 import synth

    view = tempfile.mkdtemp(prefix="pmce_ref_view_")
    # lib/ and lib/core/ must be real directories: config.py walks `<its dir>/../../` physically.
    os.makedirs(os.path.join(view, "lib", "core"))
    for name in os.listdir(os.path.join(REFERENCE_ROOT, "lib")):
        if name != "core":
            os.symlink(os.path.join(REFERENCE_ROOT, "lib", name), os.path.join(view, "lib", name))
    for name in os.listdir(os.path.join(REFERENCE_ROOT, "lib", "core")):
        os.symlink(os.path.join(REFERENCE_ROOT, "lib", "core", name), os.path.join(view, "lib", "core", name))
    os.symlink(os.path.join(REFERENCE_ROOT, "smplpytorch"), os.path.join(view, "smplpytorch"))
    os.makedirs(os.path.join(view, "data"))
    os.makedirs(os.path.join(view, "experiment"))
    for name in os.listdir(os.path.join(REFERENCE_ROOT, "data")):
        if name == "base_data":
            continue
        os.symlink(os.path.join(REFERENCE_ROOT, "data", name), os.path.join(view, "data", name))
    assets = synth.write_mesh_assets(view, seed=asset_seed)
    cfg, models = _import_reference(view)
    _state.update(view=view, cfg=cfg, models=models, assets=assets, kind="tree")
    return _state


def build_pmce(num_joint=17, embed_dim=256, depth=3, seqlen=16):
    """`models.PMCE.get_model` of the reference (lib/models/PMCE.py:23-26), eval mode."""
    st = setup()
    st["cfg"].DATASET.seqlen = seqlen
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = st["models"].PMCE.get_model(num_joint, embed_dim, depth)
    return model.eval()


def build_smpl_layer(buffers):
    """Reference `SMPL_Layer` with `__init__` bypassed (it needs licensed pkl files + chumpy,
    smpl_layer.py:30-37); `forward` (smpl_layer.py:65-158) then runs verbatim."""
    setup()
    from smplpytorch.pytorch.smpl_layer import SMPL_Layer
    layer = SMPL_Layer.__new__(SMPL_Layer)
    torch.nn.Module.__init__(layer)
    layer.center_idx = None
    layer.gender = "neutral"
    for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights"):
        layer.register_buffer(k, buffers[k].clone())
    layer.kintree_parents = list(buffers["kintree_parents"])
    layer.num_joints = len(layer.kintree_parents)
    return layer.eval()


def build_spin_trunk():
    """The reference's SPIN/HMR ResNet-50 TRUNK (lib/models/spin.py): `HMR.__init__` needs the licensed SMPL body model
    (spin.py:90-94) and `smpl_mean_params.npz` (:103), so the constructor is bypassed and only its trunk lines (:66-77) are
    replayed through the reference's own `_make_layer` / `Bottleneck`; `feature_extractor` (:129-143) then runs verbatim."""
    setup()
    import torch.nn as nn
    from models import spin as sp            # the reference module (smplx comes from oracle/shims)
    m = sp.HMR.__new__(sp.HMR)
    nn.Module.__init__(m)
    m.inplanes = 64
    m.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
    m.bn1 = nn.BatchNorm2d(64)
    m.relu = nn.ReLU(inplace=True)
    m.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
    m.layer1 = m._make_layer(sp.Bottleneck, 64, 3)
    m.layer2 = m._make_layer(sp.Bottleneck, 128, 4, stride=2)
    m.layer3 = m._make_layer(sp.Bottleneck, 256, 6, stride=2)
    m.layer4 = m._make_layer(sp.Bottleneck, 512, 3, stride=2)
    m.avgpool = nn.AvgPool2d(7, stride=1)
    return m.eval()
