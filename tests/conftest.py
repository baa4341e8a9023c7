import os
import sys
import tempfile

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def dense_regressor(key):
    g = np.load(os.path.join(GOLDEN, "J_regressors_sparse.npz"))
    J = np.zeros(tuple(g[key + "_shape"]), dtype=np.float64)
    J[g[key + "_rows"], g[key + "_cols"]] = g[key + "_vals"]
    return J


@pytest.fixture(scope="session")
def assets_root():
    """Scratch data root with the seeded synthetic mesh assets + the (real, sparse-stored) H36M regressor;
    exported as PMCE_DATA_ROOT so `models.CoevoDecoder.get_model` finds `data/base_data/*` like the reference does
    relative to its cwd."""
    from pmce_b200 import synth
    root = tempfile.mkdtemp(prefix="pmce_assets_")
    synth.write_mesh_assets(root, seed=7)
    os.makedirs(os.path.join(root, "data", "Human36M"))
    np.save(os.path.join(root, "data", "Human36M", "J_regressor_h36m_correct.npy"), dense_regressor("h36m"))
    os.environ["PMCE_DATA_ROOT"] = root
    os.environ["PMCE_B200_STANDALONE_CFG"] = "1"
    return root


@pytest.fixture(scope="session")
def lib():
    """libpmce_b200.so, built on demand (nvcc cross-compiles for sm_100a without a GPU)."""
    from pmce_b200 import build, _lib
    build.build()
    return _lib.load()
