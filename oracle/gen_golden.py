"""Generate `tests/golden/*.npz` from the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Run in the build container only (`python oracle/gen_golden.py`); needs `/root/reference`.
For every config: seeded synthetic weights (`pmce_b200.synth.make_state_dict`) are loaded with
`load_state_dict(strict=True)` into the reference `models.PMCE.get_model(...)`
(reference lib/models/PMCE.py:23-26), seeded inputs are pushed through `PMCE.forward`
(PMCE.py:15-20) on CPU fp32, and outputs + a few hooked intermediates are stored. The same script
asserts that the schema restated in `pmce_b200.synth.state_dict_schema` equals the reference
`state_dict()` and that the CPU restatement `oracle/pmce_oracle.py` reproduces the reference.
It also stores the shipped H36M / COCO joint regressors in sparse (exact) form, and a golden for
`SMPL_Layer.forward` (smpl_layer.py:65-158) on synthetic model buffers.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import ref_harness as rh          # noqa: E402
from oracle import pmce_oracle as po          # noqa: E402
from pmce_b200 import synth                   # noqa: E402

# PMCE_GOLDEN_OUT redirects the output (tests/test_oracle_vs_reference.py regenerates into a scratch dir and diffs)
GOLDEN = os.environ.get("PMCE_GOLDEN_OUT") or os.path.join(REPO, "tests", "golden")

CONFIGS = [
    # name, J, C, T, B, lifter_out_scale
    ("pmce_J17_C256_T16_B2", 17, 256, 16, 2, 300.0),
    ("pmce_J19_C256_T16_B2", 19, 256, 16, 2, 300.0),
    ("pmce_J17_C512_T16_B2", 17, 512, 16, 2, 300.0),
    ("pmce_J17_C256_T64_B1", 17, 256, 64, 1, 300.0),
    ("pmce_J17_C256_T16_B3_unitscale", 17, 256, 16, 3, 1.0),
]


def weight_checksum(sd):
    """Order-independent fingerprint of the synthetic weights (float64 sums of |w| per tensor)."""
    return np.array([float(v.double().abs().sum()) for v in sd.values()], dtype=np.float64)


def gen_pmce(name, J, C, T, B, scale):
    model = rh.build_pmce(J, C, 3, T)
    ref_sd = model.state_dict()
    schema = synth.state_dict_schema(J, C, 3, T)
    assert set(schema.keys()) == set(ref_sd.keys()) and len(schema) == len(ref_sd), "schema name mismatch"
    for k, shp in schema.items():
        assert tuple(ref_sd[k].shape) == tuple(shp), (k, ref_sd[k].shape, shp)

    init_vertices = ref_sd["pose_mesh_coevo.init_vertices"].numpy().copy()
    vj = np.asarray(model.pose_mesh_coevo.vj_relation)
    assert vj.dtype == np.float64                       # the reference indexes with a float64 ndarray
    vj_int = vj.astype(np.int64)
    assert np.array_equal(vj_int.astype(np.float64), vj)

    sd = synth.make_state_dict(0, init_vertices=init_vertices, lifter_out_scale=scale,
                               num_joint=J, embed_dim=C, depth=3, seqlen=T)
    model.load_state_dict(sd, strict=True)
    pose2d, img_feat = synth.make_inputs(B, T, J, seed=1)

    cap = {}
    coevo = model.pose_mesh_coevo
    hooks = [coevo.gru_cur.register_forward_hook(lambda m, i, o: cap.__setitem__("y", o[0].detach().clone()))]
    for k in (1, 2, 3):
        hooks.append(getattr(coevo, f"coevoblock{k}").register_forward_hook(
            lambda m, i, o, k=k: cap.__setitem__(f"blk{k}", (o[0].detach().clone(), o[1].detach().clone()))))
    with torch.no_grad():
        mesh, cam_pose, pose3d = model(pose2d, img_feat)
    for h in hooks:
        h.remove()
    g = cap["y"][T // 2]

    # the CPU restatement must reproduce the reference (same torch CPU kernels -> tiny differences only)
    with torch.no_grad():
        o_mesh, o_pose, o_p3, inter = po.pmce_forward(sd, pose2d, img_feat, vj_int, return_intermediates=True)
    errs = dict(mesh=float((o_mesh - mesh).abs().max()), cam_pose=float((o_pose - cam_pose).abs().max()),
                pose3d_rel=float(((o_p3 - pose3d).abs().max() / pose3d.abs().max())),
                g=float((inter["g"] - g).abs().max()))
    print(name, "oracle-vs-reference max abs:", errs)
    assert errs["mesh"] < 2e-5 and errs["cam_pose"] < 2e-5 and errs["pose3d_rel"] < 1e-5 and errs["g"] < 1e-5

    # init-time geometry restatement
    assets = rh.setup()["assets"]
    jreg = np.load(os.path.join(rh.REFERENCE_ROOT, "data", "Human36M", "J_regressor_h36m_correct.npy")).astype(np.float32)
    iv, vj2 = po.init_geometry(assets["verts"], assets["D"], jreg)
    assert np.array_equal(vj2, vj_int), "vj_relation restatement differs"
    assert float((iv - torch.from_numpy(init_vertices)).abs().max()) < 1e-6

    # J-regressor post-step (core/base.py:223-225)
    pred_pose = torch.matmul(torch.from_numpy(jreg)[None], mesh * 1000)

    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        config=np.array([J, C, 3, T, B], dtype=np.int64), lifter_out_scale=np.float64(scale),
        weight_seed=np.int64(0), input_seed=np.int64(1), asset_seed=np.int64(7),
        weight_checksum=weight_checksum(sd),
        input_checksum=np.array([float(pose2d.double().sum()), float(img_feat.double().sum())]),
        init_vertices=init_vertices, vj_relation=vj_int,
        cam_mesh=mesh.numpy(), cam_pose=cam_pose.numpy(), pose3d=pose3d.numpy(), gru_mid=g.numpy(),
        verts1=cap["blk1"][1].numpy(), verts2=cap["blk2"][1].numpy(), verts3=cap["blk3"][1].numpy(),
        joints3=cap["blk3"][0].numpy(), pred_pose_h36m=pred_pose.numpy())


def gen_jregressors():
    out = {}
    for key, rel in (("h36m", "Human36M/J_regressor_h36m_correct.npy"), ("coco", "COCO/J_regressor_coco.npy")):
        J = np.load(os.path.join(rh.REFERENCE_ROOT, "data", rel))
        r, c = np.nonzero(J)
        out[key + "_rows"], out[key + "_cols"], out[key + "_vals"] = r.astype(np.int32), c.astype(np.int32), J[r, c]
        out[key + "_shape"] = np.array(J.shape, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "J_regressors_sparse.npz"), **out)


def gen_smpl():
    buf = synth.make_smpl_buffers(11)
    layer = rh.build_smpl_layer(buf)
    pose, betas, trans = synth.make_smpl_inputs(4, seed=13)
    with torch.no_grad():
        v, j = layer(pose, betas, trans)
        v0, j0 = layer(pose)                      # default zero betas / zero trans branch (:87-91,:148)
        ov, oj = po.smpl_lbs(buf, pose, betas, trans)
        ov0, oj0 = po.smpl_lbs(buf, pose)
    e = [float((ov - v).abs().max()), float((oj - j).abs().max()), float((ov0 - v0).abs().max()), float((oj0 - j0).abs().max())]
    print("smpl oracle-vs-reference max abs:", e)
    assert max(e) < 2e-5
    np.savez_compressed(os.path.join(GOLDEN, "smpl_lbs_B4.npz"), buffer_seed=np.int64(11), input_seed=np.int64(13),
                        verts=v.numpy(), joints=j.numpy(), verts_default=v0.numpy(), joints_default=j0.numpy())


def reference_method(rel_path, cls, name):
    """Compile ONE method of a reference class from its source file where it lies (nothing is copied into the repo): the
    dataset modules cannot be imported here (pycocotools, licensed SMPL files), but `compute_both_err` only needs numpy."""
    import ast
    src = open(os.path.join(rh.REFERENCE_ROOT, rel_path)).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == name:
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ns = {"np": np, "torch": torch}
                    exec(compile(mod, os.path.join(rh.REFERENCE_ROOT, rel_path), "exec"), ns)
                    return ns[name]
    raise KeyError((rel_path, cls, name))


def gen_eval():
    """Golden vector of the evaluation epilogue (f)1: the reference's own PW3D.compute_both_err on seeded meshes."""
    fn = reference_method("data/PW3D/dataset.py", "PW3D", "compute_both_err")

    class Stub:
        human36_eval_joint = (1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16)   # data/PW3D/dataset.py:35
    B = 6
    gen = torch.Generator().manual_seed(21)
    cam_mesh = torch.randn(B, 6890, 3, generator=gen) * 0.3
    gt_mesh = cam_mesh + torch.randn(B, 6890, 3, generator=gen) * 0.05
    gt_pose = torch.randn(B, 17, 3, generator=gen) * 300
    jreg = torch.from_numpy(np.load(os.path.join(rh.REFERENCE_ROOT, "data", "Human36M", "J_regressor_h36m_correct.npy")).astype(np.float32))
    pred_mesh, gt_mm = cam_mesh * 1000, gt_mesh * 1000                    # lib/core/base.py:223
    pred_pose = torch.matmul(jreg[None, :, :], pred_mesh)                 # :225
    j_err, s_err = fn(Stub(), pred_mesh, gt_mm, pred_pose, gt_pose)       # :227
    o_pose, oj, os_ = po.eval_step(jreg, cam_mesh, gt_mesh, gt_pose)
    print("eval oracle-vs-reference:", float(j_err), float(oj), float(s_err), float(os_))
    assert abs(float(oj) - float(j_err)) < 1e-4 and abs(float(os_) - float(s_err)) < 1e-4 and torch.equal(o_pose, pred_pose)
    np.savez_compressed(os.path.join(GOLDEN, "eval_err_B6.npz"), seed=np.int64(21), B=np.int64(B), joint_mean_error=np.float64(j_err),
                        mesh_mean_error=np.float64(s_err), pred_pose=pred_pose.numpy())


def gen_spin():
    """(f)2 golden: the reference's own `HMR.feature_extractor` (lib/models/spin.py:129-143) on seeded trunk weights and frames."""
    from oracle import spin_oracle as so
    model = rh.build_spin_trunk()
    ref_sd = model.state_dict()
    schema = synth.spin_state_dict_schema()
    assert set(schema) == set(ref_sd) and all(tuple(ref_sd[k].shape) == tuple(v) for k, v in schema.items()), "spin schema mismatch"
    sd = synth.make_spin_state_dict(17)
    model.load_state_dict(sd, strict=True)
    x = synth.make_frames(2, 19)
    cap = {}
    hooks = [getattr(model, f"layer{i}").register_forward_hook(lambda m, i_, o, i=i: cap.__setitem__(f"layer{i}", o.detach().clone())) for i in (1, 2, 3, 4)]
    with torch.no_grad():
        xf = model.feature_extractor(x)
        oxf, inter = so.feature_extractor(sd, x, return_intermediates=True)
    for h in hooks:
        h.remove()
    errs = {k: float((inter[k] - cap[k]).abs().max()) for k in cap}
    errs["xf"] = float((oxf - xf).abs().max())
    print("spin oracle-vs-reference max abs:", errs, "| max|xf| = %.3f" % float(xf.abs().max()))
    assert max(errs.values()) < 1e-4
    # layer outputs are large: keep a strided sub-sample (every 7th position of both spatial axes) next to the full feature vector
    np.savez_compressed(os.path.join(GOLDEN, "spin_B2.npz"), weight_seed=np.int64(17), input_seed=np.int64(19), xf=xf.numpy(),
                        layer1=cap["layer1"][:, :, ::7, ::7].numpy(), layer2=cap["layer2"][:, :, ::7, ::7].numpy(),
                        layer3=cap["layer3"][:, :, ::7, ::7].numpy(), layer4=cap["layer4"][:, ::8].numpy())


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    gen_spin()
    gen_jregressors()
    gen_smpl()
    gen_eval()
    for cfg_ in CONFIGS:
        gen_pmce(*cfg_)
    print("golden fixtures written to", GOLDEN)
