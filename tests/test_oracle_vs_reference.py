"""CPU (not gpu): the oracle <-> reference pin, re-verified on every run where the reference is present.

`oracle/gen_golden.py` is re-run into a scratch directory (a child process: the harness chdirs and patches `.cuda()`); it
loads the seeded weights into the UNMODIFIED reference, asserts inside that `oracle/pmce_oracle.py` reproduces the reference
(<= 2e-5) and that the restated state_dict schema / init geometry equal the reference's, and writes the fixtures. They must
equal the committed `tests/golden/*.npz` exactly: the committed vectors ARE what the reference produces, not a stale copy.
Also checks that `oracle/_ref` (the copy `bench.py --impl reference` runs on the GPU box) is byte-identical to the reference."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO

REFERENCE_ROOT = os.environ.get("PMCE_REFERENCE_ROOT", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "models")),
                               reason="the reference tree only exists in the build container")


@needs_ref
def test_goldens_regenerate_bit_identically(tmp_path):
    env = dict(os.environ, PMCE_GOLDEN_OUT=str(tmp_path), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(REPO, "oracle", "gen_golden.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    committed = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    fresh = sorted(os.path.basename(p) for p in glob.glob(os.path.join(str(tmp_path), "*.npz")))
    assert fresh == committed
    for name in committed:
        a, b = np.load(os.path.join(GOLDEN, name)), np.load(os.path.join(str(tmp_path), name))
        assert sorted(a.files) == sorted(b.files), name
        for k in a.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, (name, k)
            assert np.array_equal(a[k], b[k]), (name, k, float(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max()))
    assert "oracle-vs-reference" in r.stdout


@needs_ref
def test_ref_copy_is_byte_identical():
    import hashlib
    sys.path.insert(0, REPO)
    from oracle import build_ref
    dest = build_ref.build()
    man = json.load(open(os.path.join(dest, "MANIFEST.json")))
    assert len(man["unmodified_files"]) > 40
    for rel, h in man["unmodified_files"].items():
        for root in (dest, REFERENCE_ROOT):
            with open(os.path.join(root, rel), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest()[:16] == h, (root, rel)
    # git-ignored (no reference source in the history) but not gpurun-ignored (it travels to the GPU box)
    assert "oracle/_ref/" in open(os.path.join(REPO, ".gitignore")).read()
    gi = os.path.join(REPO, ".gpurunignore")
    assert not os.path.exists(gi) or "oracle/_ref" not in open(gi).read()
