"""CPU: host-side mirror of the reference interface (module tree / state_dict schema / error behaviour) and the
batch-sharding + all-gather logic on gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, REPO
from pmce_b200 import synth


def test_module_schema_matches_reference(assets_root, lib):
    from pmce_b200 import models
    from pmce_b200.config import cfg
    cfg.DATASET.seqlen = 16
    m = models.PMCE.get_model(17, 256, 3)
    sd = m.state_dict()
    schema = synth.state_dict_schema(17, 256, 3, 16)
    assert set(sd) == set(schema)
    for k, shp in schema.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    assert sum(p.numel() for p in m.parameters()) == 103064752
    g = np.load(os.path.join(GOLDEN, "pmce_J17_C256_T16_B2.npz"))
    assert np.array_equal(m.pose_mesh_coevo.vj_relation, g["vj_relation"])
    assert np.abs(sd["pose_mesh_coevo.init_vertices"].numpy() - g["init_vertices"]).max() < 1e-6
    # strict load of a reference-schema checkpoint works
    w = synth.make_state_dict(0, init_vertices=g["init_vertices"], num_joint=17, embed_dim=256, depth=3, seqlen=16)
    m.load_state_dict(w, strict=True)
    assert hasattr(models, "project_net") and hasattr(models.PoseEstimation, "get_model") and hasattr(models.CoevoDecoder, "get_model")


def test_no_cpu_fallback(assets_root, lib):
    from pmce_b200 import models
    from pmce_b200._lib import PmceError
    from pmce_b200.config import cfg
    cfg.DATASET.seqlen = 16
    m = models.PoseEstimation.get_model(17, 256, 3).eval()
    with pytest.raises(PmceError, match="CUDA only"):
        m(torch.zeros(1, 16, 17, 2), torch.zeros(1, 16, 2048))


def test_missing_library_fails_loudly(monkeypatch):
    from pmce_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpmce_b200.so")
    with pytest.raises(_lib.PmceError, match="no CPU/PyTorch fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pmce_b200/ may reference it."""
    for root, _, files in os.walk(os.path.join(REPO, "pmce_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/gen_golden.py", ""), f


def test_shard_bounds_cover_batch():
    from pmce_b200.dist import shard_bounds
    for total in (1, 2, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_bounds(total, r, world)
                assert 0 <= lo <= hi <= total and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(total))


def _fake_forward(p2d, feat):
    """Deterministic per-clip stand-in for the GPU forward (depends only on that clip's inputs)."""
    b, j = p2d.shape[0], p2d.shape[2]
    s = p2d.sum(dim=(1, 2, 3)) + feat.sum(dim=(1, 2))
    mesh = s[:, None, None] + torch.arange(6890 * 3, dtype=torch.float32).reshape(1, 6890, 3)
    return mesh, s[:, None, None].expand(b, j, 3).clone(), (2 * s)[:, None, None].expand(b, j, 3).clone()


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pmce_b200.dist import sharded_forward
        p2d, feat = synth.make_inputs(total, 16, 17, seed=5)
        out = sharded_forward(_fake_forward, p2d, feat)
        ref = _fake_forward(p2d, feat)
        ok = all(torch.equal(a, b) for a, b in zip(out, ref))
        q.put((rank, ok, tuple(out[0].shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8, 1])
def test_sharded_forward_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + total) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (total, 6890, 3)


def _fake_forward_out(p2d, feat, out=None):
    m, a, b = _fake_forward(p2d, feat)
    out[0].copy_(m); out[1].copy_(a); out[2].copy_(b)
    return out


def _worker_direct(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pmce_b200.dist import ShardedForward
        per = 3
        sf = ShardedForward(_fake_forward_out, per, 17, "cpu")
        ok = True
        for step in range(3):                       # three steps over two slots: slot reuse
            p2d, feat = synth.make_inputs(world * per, 16, 17, seed=40 + step)
            k = sf.step(p2d[rank * per:(rank + 1) * per].contiguous(), feat[rank * per:(rank + 1) * per].contiguous())
            out = sf.unpack(k)
            ref = _fake_forward(p2d, feat)
            ok = ok and all(torch.equal(a, b) for a, b in zip(out, ref))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_forward_direct_outputs_gloo_world2():
    """ShardedForward: outputs written straight into the rank's row of the gather buffer, in-place all-gather, slot rotation."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_direct, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
