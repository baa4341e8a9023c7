"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/pmce_b200.h declares;
host-only entry points (layout, sizes, error paths) behave. No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from conftest import REPO
from pmce_b200 import synth


def header_functions():
    src = open(os.path.join(REPO, "include", "pmce_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\(", src)
    return sorted({n for n in names if n.startswith(("pmce_", "smpl_"))})


def test_exports_match_header(lib):
    from pmce_b200 import _lib
    names = header_functions()
    assert len(names) >= 19
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    for n in names:
        assert hasattr(lib, n)
    assert lib.pmce_abi_version() == 1


def test_sass_is_sm100a():
    import subprocess
    from pmce_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


@pytest.mark.parametrize("J,Cw,T", [(17, 256, 16), (19, 256, 16), (17, 512, 16), (17, 256, 64)])
def test_every_schema_tensor_has_a_slot(lib, J, Cw, T):
    from pmce_b200 import engine, _lib
    d = engine.make_dims(J, Cw, 3, T)
    total = lib.pmce_weights_bytes(C.byref(d)) // 4
    assert total > 0
    schema = synth.state_dict_schema(J, Cw, 3, T)
    slot = _lib.PmceSlot()
    live, dead, spans = 0, 0, []
    for name, shape in schema.items():
        rc = lib.pmce_weight_slot(C.byref(d), name.encode(), C.byref(slot))
        assert rc in (0, 1), (name, lib.pmce_last_error())
        if rc == 1:
            dead += 1
            assert "coevoblock1" in name or "coevoblock2" in name
            continue
        live += 1
        n = 1
        for s in shape:
            n *= s
        assert slot.rows * slot.cols == n, name
        assert slot.ld >= slot.cols and slot.offset % 4 == 0 and slot.ld % 4 == 0 or slot.rows == 1 or slot.cols < 4, name
        assert slot.offset + slot.rows * slot.ld <= total
        spans.append((slot.offset, slot.offset + (slot.rows - 1) * slot.ld + slot.cols))
    assert live + dead == 431 and dead == 100
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0, "weight slots overlap"
    assert lib.pmce_weight_slot(C.byref(d), b"not.a.tensor", C.byref(slot)) < 0
    assert b"not.a.tensor" in lib.pmce_last_error()


def test_sizes_and_error_paths(lib):
    from pmce_b200 import engine
    d = engine.make_dims()
    w1, w2 = lib.pmce_workspace_bytes(C.byref(d), 1), lib.pmce_workspace_bytes(C.byref(d), 64)
    assert 0 < w1 < w2
    assert lib.pmce_workspace_bytes(C.byref(d), 0) == 0
    assert lib.pmce_io_bytes(C.byref(d), 2) >= 2 * (16 * 17 * 2 + 16 * 2048 + 6890 * 3 + 2 * 17 * 3) * 4
    bad = engine.make_dims(embed_dim=200)
    assert lib.pmce_weights_bytes(C.byref(bad)) == 0
    assert b"embed_dim" in lib.pmce_last_error()
    # NULL weights / workspace are reported, not dereferenced
    rc = lib.pmce_forward(C.byref(d), None, None, None, None, 1, None, None, None, None, 0, None)
    assert rc != 0 and b"NULL" in lib.pmce_last_error()
    assert lib.pmce_adaln_slots() == 24 and lib.smpl_blend_ld() >= 217 and lib.smpl_workspace_bytes(2) > 0


def test_new_entry_points_validate_before_touching_the_gpu(lib):
    """Host-side argument checks of the round-1 additions: window arithmetic of pmce_forward_sliding ((f)3), the fold workspace
    of the fused cross-attention, and NULL / range errors of the attention-block, evaluation and SMPL entry points."""
    from pmce_b200 import engine
    d = engine.make_dims()                               # T = 16
    dp = C.byref(d)
    assert lib.pmce_sliding_windows(dp, 16, 1) == 1 and lib.pmce_sliding_windows(dp, 79, 1) == 64
    assert lib.pmce_sliding_windows(dp, 40, 3) == 9 and lib.pmce_sliding_windows(dp, 64, 16) == 4
    assert lib.pmce_sliding_windows(dp, 15, 1) == 0 and lib.pmce_sliding_windows(dp, 40, 0) == 0
    rc = lib.pmce_forward_sliding(dp, None, None, None, None, 8, 1, None, None, None, None, 0, None)
    assert rc != 0 and b"num_frames" in lib.pmce_last_error()
    rc = lib.pmce_forward_sliding(dp, None, None, None, None, 64, 17, None, None, None, None, 0, None)
    assert rc != 0 and b"stride" in lib.pmce_last_error()
    assert lib.pmce_ca_fold_bytes(0) == 0
    assert lib.pmce_ca_fold_bytes(2) == 2 * (4 * 48 * 64 * 2 + 48 * 4)      # KQ' | VP' (48 key slots x 64, bf16 hi + lo each) + sb' fp32
    blob = C.c_void_p(256)                               # non-NULL, never dereferenced: every call below fails in validation
    rc = lib.pmce_cross_attn_block(dp, blob, 4, 1, None, None, None, None, 1, None, None, 0, None)
    assert rc != 0 and b"block must be 1..3" in lib.pmce_last_error()
    rc = lib.pmce_cross_attn_block(dp, blob, 1, 0, None, None, None, None, 1, None, None, 0, None)
    assert rc != 0 and b"joint-branch" in lib.pmce_last_error()          # coevoblock1's joint branch is not stored
    rc = lib.pmce_self_attn_block(dp, blob, 3, 2, None, None, 1, None, None, 0, None)
    assert rc != 0 and b"which" in lib.pmce_last_error()
    rc = lib.pmce_self_attn_block(dp, blob, 3, 1, None, None, 1, None, None, 0, None)
    assert rc != 0 and b"NULL" in lib.pmce_last_error()
    rc = lib.pmce_eval_errors(None, None, None, 17, None, None, None, None, 14, 6890, 1, 1000.0, None, None, None, None)
    assert rc != 0 and b"NULL" in lib.pmce_last_error()
    rc = lib.smpl_lbs_forward_scaled(None, None, None, None, None, None, None, None, None, None, None, 1, 1.0, None, None, None, 0, None)
    assert rc != 0 and b"NULL" in lib.pmce_last_error()
