"""(f)2: the SPIN / HMR ResNet-50 feature extractor (reference lib/models/spin.py:129-143).
CPU: the oracle restatement reproduces the fixture the unmodified reference produced (oracle/gen_golden.py::gen_spin).
GPU: `pmce_b200.spin.HMR.feature_extractor` (pmce_spin_features through the C ABI) against fixture and oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from pmce_b200 import synth


def _fixture():
    g = np.load(os.path.join(GOLDEN, "spin_B2.npz"))
    return g, synth.make_spin_state_dict(int(g["weight_seed"])), synth.make_frames(2, int(g["input_seed"]))


def test_spin_oracle_reproduces_reference_fixture():
    from oracle import spin_oracle as so
    g, sd, x = _fixture()
    with torch.no_grad():
        xf, inter = so.feature_extractor(sd, x, return_intermediates=True)
    assert torch.equal(xf, torch.as_tensor(g["xf"])) or float((xf - torch.as_tensor(g["xf"])).abs().max()) < 1e-4
    for k in ("layer1", "layer2", "layer3"):
        assert float((inter[k][:, :, ::7, ::7] - torch.as_tensor(g[k])).abs().max()) < 1e-4, k
    assert float((inter["layer4"][:, ::8] - torch.as_tensor(g["layer4"])).abs().max()) < 1e-4


def test_spin_schema_matches_mirror(lib):
    from pmce_b200.spin import HMR
    sch = synth.spin_state_dict_schema()
    sd = HMR().state_dict()
    assert set(sd) == set(sch) and all(tuple(sd[k].shape) == tuple(v) for k, v in sch.items())
    assert lib.pmce_spin_num_convs() == 53 and lib.pmce_spin_workspace_bytes(2) > 0 and lib.pmce_spin_workspace_bytes(0) == 0


@pytest.mark.gpu
def test_spin_features_vs_reference_golden(lib):
    from oracle import spin_oracle as so
    from pmce_b200.spin import HMR
    g, sd, x = _fixture()
    m = HMR()
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    xf = m.feature_extractor(x.cuda()).cpu()
    ref = torch.as_tensor(g["xf"])
    err = float((xf - ref).abs().max())
    print(f"spin: max|d xf| = {err:.3e} on max|xf| = {float(ref.abs().max()):.2f}")
    assert err < 1e-4 * float(ref.abs().max())           # 53 chained bf16x3 convolutions, relative to the feature scale
    # batch independence + a ragged batch (B=5 leaves partial 128-row GEMM tiles in layer4: 5 * 49 = 245 rows)
    idx = torch.tensor([1, 0, 1, 1, 0])
    xb = m.feature_extractor(x[idx].contiguous().cuda()).cpu()
    assert float((xb - xf[idx]).abs().max()) < 1e-5 * float(ref.abs().max())
    # a second input against the oracle directly
    x2 = synth.make_frames(1, 23)
    with torch.no_grad():
        r2 = so.feature_extractor(sd, x2)
    assert float((m.feature_extractor(x2.cuda()).cpu() - r2).abs().max()) < 1e-4 * float(r2.abs().max())
    from pmce_b200._lib import PmceError
    with pytest.raises(PmceError, match="CUDA"):
        m.feature_extractor(x)
