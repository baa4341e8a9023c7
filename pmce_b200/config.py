"""`cfg` keys the hot-path modules read (reference lib/core/config.py:48,52,59-66).

When the package is dropped into the reference tree (`lib/` on sys.path) the reference's own global
`core.config.cfg` is used, so YAML overrides keep working; stand-alone, a local default object with the
same keys/values is used instead.
"""
import os


class _NS(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _default_cfg():
    c = _NS()
    c.DATASET = _NS(seqlen=16, BASE_DATA_DIR="data/base_data")
    c.MODEL = _NS(hpe_dim=256, hpe_dep=3, joint_dim=64, vertx_dim=64, posenet_pretrained=False,
                  posenet_path="./experiment/pretrained/pose_3dpw.pth.tar")
    return c


def _resolve():
    if os.environ.get("PMCE_B200_STANDALONE_CFG", "0") != "1":
        try:
            from core.config import cfg as ref_cfg  # reference tree present on sys.path
            return ref_cfg
        except Exception:
            pass
    return _default_cfg()


cfg = _resolve()


def data_root():
    """Directory relative to which `data/base_data/...` and `data/Human36M/...` are resolved
    (the reference uses the process cwd, lib/models/CoevoDecoder.py:194,207; backbones/mesh.py:61)."""
    return os.environ.get("PMCE_DATA_ROOT", ".")
