"""The drop-in claim, executed: INTEGRATION.md's overlay of the reference's four `lib/models/*.py` files, driven through the
reference's own `import models` / `get_model` / `load_state_dict` / test-loop lines (lib/core/base.py:54,67,218-227) in a child
process (tests/overlay_driver.py). CPU part: import + construction + schema; GPU part: the loop body against the golden."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO

HAVE_REF = os.path.isdir("/root/reference/lib/models") or os.path.isfile(os.path.join(REPO, "oracle", "_ref", "MANIFEST.json"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="needs the reference lib/ (the tree in the build container, oracle/_ref elsewhere)")


def _run(tmp_path, mode):
    env = {k: v for k, v in os.environ.items() if k not in ("PMCE_DATA_ROOT", "PMCE_B200_STANDALONE_CFG")}
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "overlay_driver.py"), str(tmp_path / "ref"), mode], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("OVERLAY_RESULT ")][-1]
    return json.loads(line[len("OVERLAY_RESULT "):])


@needs_ref
def test_overlay_imports_and_constructs_like_the_reference(lib, tmp_path):
    out = _run(tmp_path, "cpu")
    assert out["keys"] == 431 and out["vj_equal"]


@needs_ref
@pytest.mark.gpu
def test_overlay_runs_the_reference_test_loop(lib, tmp_path):
    out = _run(tmp_path, "gpu")
    assert out["vj_equal"]
    assert out["mesh_err"] < 1e-4 and out["evo_err"] < 1e-4 and out["pose_err_mm"] < 1e-1      # metres, metres, millimetres
