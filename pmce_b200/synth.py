"""Seeded synthetic weights, inputs and mesh assets for the PMCE hot path.

The reference ships no checkpoints, no `data/base_data` assets and no SMPL model files
(SURVEY.md §8c), so every parity / benchmark run uses the generators below. They are numpy
`default_rng` (PCG64) based, hence bit-reproducible on any host: the golden fixtures under
`tests/golden/` were produced by loading exactly these tensors into the *reference* modules
(see `oracle/gen_golden.py`).

`state_dict_schema` restates the reference checkpoint schema (names and shapes of the 431
tensors; reference lib/models/PoseEstimation.py:31-66, lib/models/CoevoDecoder.py:16-224) and
is asserted equal to the real reference `state_dict()` in `oracle/gen_golden.py`.
"""
from collections import OrderedDict
import math
import os

import numpy as np
import torch

N_VERT = 6890      # SMPL vertices
N_VERT_DS = 431    # after two mesh down-samplings (6890 -> 1723 -> 431)
N_VERT_MID = 1723
F_IMG = 2048       # ResNet-50 feature width
H_GRU = 1024
D_COEVO = 64


def state_dict_schema(num_joint=17, embed_dim=256, depth=3, seqlen=16, num_vert_ds=N_VERT_DS,
                      coevo_dim=D_COEVO):
    """name -> shape for `PMCE.state_dict()` (reference lib/models/PMCE.py:8-13)."""
    J, C, T, Vd, D = num_joint, embed_dim, seqlen, num_vert_ds, coevo_dim
    s = OrderedDict()

    def lin(prefix, n_out, n_in):
        s[prefix + ".weight"] = (n_out, n_in)
        s[prefix + ".bias"] = (n_out,)

    def ln(prefix, n):
        s[prefix + ".weight"] = (n,)
        s[prefix + ".bias"] = (n,)

    p = "pose_lifter."
    lin(p + "joint_embed", C, 2)
    lin(p + "imgfeat_embed", C, F_IMG)
    s[p + "spatial_pos_embed"] = (1, J, C)
    s[p + "temporal_pos_embed"] = (1, T, C)
    for kind in ("SpatialBlocks", "TemporalBlocks"):
        for i in range(depth):
            b = f"{p}{kind}.{i}."
            ln(b + "norm1", C)
            lin(b + "attn.qkv", 3 * C, C)
            lin(b + "attn.proj", C, C)
            ln(b + "norm2", C)
            lin(b + "mlp.fc1", 2 * C, C)
            lin(b + "mlp.fc2", C, 2 * C)
    ln(p + "norm_s", C)
    ln(p + "norm_t", C)
    ln(p + "regression.0", C)
    lin(p + "regression.1", 3, C)
    s[p + "fusion.weight"] = (1, T, 1, 1)
    s[p + "fusion.bias"] = (1,)

    p = "pose_mesh_coevo."
    s[p + "init_vertices"] = (Vd, 3)

    def adaln(prefix):
        lin(prefix + ".mlp_gamma", D, F_IMG)
        lin(prefix + ".mlp_beta", D, F_IMG)

    for k in (1, 2, 3):
        b = f"{p}coevoblock{k}."
        for nm in ("joint_pos_embed", "j_Q_embed", "j2v_K_embed"):
            s[b + nm] = (1, J, D)
        for nm in ("vertx_pos_embed", "v_Q_embed", "v2j_K_embed"):
            s[b + nm] = (1, Vd, D)
        lin(b + "joint_proj", D, 3)
        lin(b + "vertx_proj", D, 3)
        lin(b + "proj_v2j_dim", D, D)
        lin(b + "proj_j2v_dim", D, D)
        for st in ("joint", "vertx"):
            sa = f"{b}{st}_SA_FFN."
            adaln(sa + "norm1")
            lin(sa + "attn.qkv", 3 * D, D)
            lin(sa + "attn.proj", D, D)
            adaln(sa + "norm2")
            lin(sa + "mlp.fc1", 4 * D, D)
            lin(sa + "mlp.fc2", D, 4 * D)
            ca = f"{b}{st}_CA_FFN."
            for nm in ("normq", "normk", "normv"):
                adaln(ca + nm)
            for nm in ("wq", "wk", "wv", "proj"):
                lin(ca + "attn." + nm, D, D)
            adaln(ca + "norm2")
            lin(ca + "mlp.fc1", 4 * D, D)
            lin(ca + "mlp.fc2", D, 4 * D)
        lin(b + "proj_joint_feat2coor", 3, D)
        lin(b + "proj_vertx_feat2coor", 3, D)
    s[p + "upsample_conv.weight"] = (N_VERT, Vd, 3)
    s[p + "upsample_conv.bias"] = (N_VERT,)
    for layer in (0, 1):
        for sfx in ("", "_reverse"):
            s[f"{p}gru_cur.weight_ih_l{layer}{sfx}"] = (3 * H_GRU, F_IMG if layer == 0 else 2 * H_GRU)
            s[f"{p}gru_cur.weight_hh_l{layer}{sfx}"] = (3 * H_GRU, H_GRU)
            s[f"{p}gru_cur.bias_ih_l{layer}{sfx}"] = (3 * H_GRU,)
            s[f"{p}gru_cur.bias_hh_l{layer}{sfx}"] = (3 * H_GRU,)
    for i in (1, 2, 3):
        lin(f"{p}linear_cur{i}", N_VERT, 2 * H_GRU)
    return s


def _fan_in(name, shape):
    if name.endswith("upsample_conv.weight"):
        return shape[1] * shape[2]
    if name.endswith("fusion.weight"):
        return shape[1]
    if "gru_cur" in name:
        return H_GRU
    return shape[-1]


def make_state_dict(seed=0, init_vertices=None, lifter_out_scale=1.0, **dims):
    """Seeded weights with PyTorch-default-like magnitudes but *no* trivial tensors.

    LayerNorm affine = 1 + 0.1 n / 0.1 n, lifter pos-embeds 0.02 n (reference initialises them to
    zeros, PoseEstimation.py:42-43, which would hide indexing bugs), CoevoBlock embeds ~ N(0,1)
    (CoevoDecoder.py:151-160). `lifter_out_scale` multiplies the regression head so `pose3d`
    can be put on the trained (millimetre) scale that `PMCE.forward` divides by 1000.
    """
    schema = state_dict_schema(**dims)
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for name, shape in schema.items():
        if name.endswith("init_vertices"):
            if init_vertices is None:
                arr = 0.3 * rng.standard_normal(shape)
            else:
                rng.standard_normal(shape)  # keep the stream position independent of the argument
                arr = np.asarray(init_vertices, dtype=np.float32).reshape(shape)
        elif name.endswith("_embed") and "coevoblock" in name:
            arr = rng.standard_normal(shape)
        elif name.endswith("pos_embed"):
            arr = 0.02 * rng.standard_normal(shape)
        elif (".norm" in name or "regression.0" in name) and "mlp_" not in name:
            n = rng.standard_normal(shape)
            arr = 1.0 + 0.1 * n if name.endswith("weight") else 0.1 * n
        else:
            if name.endswith(".bias") or "bias_" in name:
                wname = name.replace(".bias", ".weight")
                if "gru_cur" in name:
                    fan = H_GRU
                else:
                    fan = _fan_in(wname, schema[wname])
            else:
                fan = _fan_in(name, shape)
            k = 1.0 / math.sqrt(fan)
            arr = rng.uniform(-k, k, size=shape)
            if "regression.1" in name:
                arr = arr * lifter_out_scale
        sd[name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return sd


def make_inputs(batch, seqlen=16, num_joint=17, seed=1):
    """`pose2d ~ N(0,1) [B,T,J,2]`, `img_feat ~ N(0,1) [B,T,2048]`."""
    rng = np.random.default_rng(seed)
    pose2d = rng.standard_normal((batch, seqlen, num_joint, 2)).astype(np.float32)
    img_feat = rng.standard_normal((batch, seqlen, F_IMG)).astype(np.float32)
    return torch.from_numpy(pose2d), torch.from_numpy(img_feat)


# ----------------------------------------------------------------------------------------------
# Mesh assets (`data/base_data/*`), synthetic because the real ones are not shipped
# ----------------------------------------------------------------------------------------------

def make_mesh_assets(seed=7):
    """Synthetic `mesh_downsampling.npz` members + `smpl_mean_vertices` (fp32 [6890,3]).

    Layout follows what the reference loader expects (lib/models/backbones/mesh.py:49-57): object
    arrays `A` (3 adjacency matrices), `U` (2 up-sampling), `D` (2 down-sampling: 1723x6890,
    431x1723), all scipy sparse. D rows are convex combinations of 1-3 source vertices.
    """
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    sizes = [N_VERT, N_VERT_MID, N_VERT_DS]
    verts = (rng.standard_normal((N_VERT, 3)) * np.array([0.25, 0.45, 0.12])).astype(np.float32)

    def adjacency(n):
        r = rng.integers(0, n, size=3 * n)
        c = rng.integers(0, n, size=3 * n)
        a = sp.coo_matrix((np.ones(3 * n), (r, c)), shape=(n, n))
        a = ((a + a.T) > 0).astype(np.float64)
        return sp.coo_matrix(a)

    def down(n_out, n_in):
        rows, cols, vals = [], [], []
        for r in range(n_out):
            k = int(rng.integers(1, 4))
            cc = rng.choice(n_in, size=k, replace=False)
            w = rng.random(k) + 0.1
            w = w / w.sum()
            rows += [r] * k
            cols += list(cc)
            vals += list(w)
        return sp.coo_matrix((np.array(vals), (np.array(rows), np.array(cols))), shape=(n_out, n_in))

    A = [adjacency(n) for n in sizes]
    D = [down(sizes[1], sizes[0]), down(sizes[2], sizes[1])]
    U = [sp.coo_matrix(D[0].T), sp.coo_matrix(D[1].T)]
    return dict(A=A, U=U, D=D, verts=verts)


def write_mesh_assets(root, seed=7):
    """Write `<root>/data/base_data/{mesh_downsampling.npz,smpl_mean_vertices.npy}`."""
    assets = make_mesh_assets(seed)
    base = os.path.join(root, "data", "base_data")
    os.makedirs(base, exist_ok=True)

    def obj(lst):
        a = np.empty(len(lst), dtype=object)
        for i, m in enumerate(lst):
            a[i] = m
        return a

    np.savez(os.path.join(base, "mesh_downsampling.npz"), A=obj(assets["A"]), U=obj(assets["U"]),
             D=obj(assets["D"]))
    np.save(os.path.join(base, "smpl_mean_vertices.npy"), assets["verts"])
    return assets


# ----------------------------------------------------------------------------------------------
# SMPL model buffers (licensed pkl files are absent) and LBS inputs
# ----------------------------------------------------------------------------------------------

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def make_smpl_buffers(seed=11):
    """Synthetic SMPL buffers with the shapes `SMPL_Layer.__init__` registers
    (reference smplpytorch/smplpytorch/pytorch/smpl_layer.py:40-63)."""
    rng = np.random.default_rng(seed)
    v_template = (rng.standard_normal((1, N_VERT, 3)) * np.array([0.25, 0.45, 0.12])).astype(np.float32)
    shapedirs = (0.01 * rng.standard_normal((N_VERT, 3, 10))).astype(np.float32)
    posedirs = (0.002 * rng.standard_normal((N_VERT, 3, 207))).astype(np.float32)
    jreg = np.zeros((24, N_VERT), dtype=np.float32)
    for j in range(24):
        idx = rng.choice(N_VERT, size=12, replace=False)
        w = rng.random(12).astype(np.float32)
        jreg[j, idx] = w / w.sum()
    weights = np.zeros((N_VERT, 24), dtype=np.float32)
    for v in range(N_VERT):
        idx = rng.choice(24, size=4, replace=False)
        w = rng.random(4).astype(np.float32)
        weights[v, idx] = w / w.sum()
    betas = np.zeros((1, 10), dtype=np.float32)
    return dict(th_betas=torch.from_numpy(betas), th_shapedirs=torch.from_numpy(shapedirs),
                th_posedirs=torch.from_numpy(posedirs), th_v_template=torch.from_numpy(v_template),
                th_J_regressor=torch.from_numpy(jreg), th_weights=torch.from_numpy(weights),
                kintree_parents=list(SMPL_PARENTS))


def make_smpl_inputs(batch, seed=13):
    rng = np.random.default_rng(seed)
    pose = (0.2 * rng.standard_normal((batch, 72))).astype(np.float32)
    betas = (0.5 * rng.standard_normal((batch, 10))).astype(np.float32)
    trans = rng.standard_normal((batch, 3)).astype(np.float32)
    return torch.from_numpy(pose), torch.from_numpy(betas), torch.from_numpy(trans)


def prepare_data_root(root, sparse_regressors_npz, asset_seed=7):
    """Populate `<root>/data/{base_data,Human36M}` with the synthetic mesh assets and the H36M joint regressor
    rebuilt (exactly) from its sparse fixture; returns `root`. Point PMCE_DATA_ROOT at it."""
    write_mesh_assets(root, seed=asset_seed)
    g = np.load(sparse_regressors_npz)
    J = np.zeros(tuple(g["h36m_shape"]), dtype=np.float64)
    J[g["h36m_rows"], g["h36m_cols"]] = g["h36m_vals"]
    os.makedirs(os.path.join(root, "data", "Human36M"), exist_ok=True)
    np.save(os.path.join(root, "data", "Human36M", "J_regressor_h36m_correct.npy"), J)
    return root


# ---- (f)2: SPIN / HMR ResNet-50 trunk (reference lib/models/spin.py:66-77) --------------------------------------------------
SPIN_LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))      # (planes, blocks, stride) of layer1..4


def spin_state_dict_schema():
    """name -> shape of the trunk's tensors (the keys `HMR.state_dict()` has for conv1/bn1/layer1-4)."""
    sch = {"conv1.weight": (64, 3, 7, 7)}

    def bn(prefix, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            sch[f"{prefix}.{k}"] = (c,)
        sch[f"{prefix}.num_batches_tracked"] = ()
    bn("bn1", 64)
    inpl = 64
    for li, (planes, blocks, stride) in enumerate(SPIN_LAYERS, 1):
        for bi in range(blocks):
            p = f"layer{li}.{bi}"
            sch[p + ".conv1.weight"] = (planes, inpl, 1, 1); bn(p + ".bn1", planes)
            sch[p + ".conv2.weight"] = (planes, planes, 3, 3); bn(p + ".bn2", planes)
            sch[p + ".conv3.weight"] = (planes * 4, planes, 1, 1); bn(p + ".bn3", planes * 4)
            if bi == 0:
                sch[p + ".downsample.0.weight"] = (planes * 4, inpl, 1, 1); bn(p + ".downsample.1", planes * 4)
            inpl = planes * 4
    return sch


def make_spin_state_dict(seed):
    """Seeded trunk weights at a trained-like scale: He-normal convolutions, BatchNorm running statistics / affine parameters
    away from the identity, so BN folding, ReLU sparsity and the residual adds are all exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in spin_state_dict_schema().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100, dtype=torch.int64)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shp, generator=g) * 1.0 + 0.5
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif len(shp) == 1:                                  # BN weight: the last BN of a bottleneck is damped like a trained net's
            sd[k] = (torch.rand(shp, generator=g) * 0.5 + 0.75) * (0.2 if ".bn3." in k else 1.0)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def make_frames(B, seed):
    """[B,3,224,224] crops, normalised-image-like values."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, 224, 224, generator=g)
