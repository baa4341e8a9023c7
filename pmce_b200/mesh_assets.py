"""Init-time template geometry for the decoder (host side, runs once per model construction).

Restates what `Pose2Mesh.__init__` does with the `Mesh` helper in the reference
(lib/models/CoevoDecoder.py:197-208, lib/models/backbones/mesh.py:49-96, lib/graph_utils.py:27-46):
down-sample the SMPL mean template 6890 -> 1723 -> 431 with the sparse `D` matrices of
`mesh_downsampling.npz`, regress the template joints with the H36M regressor and assign every
down-sampled vertex to its nearest joint. This is not worth a kernel (last row).
"""
import os

import numpy as np
import torch

from .config import cfg, data_root


def load_downsampling(path):
    """`D` matrices (scipy sparse, object array) of mesh_downsampling.npz (mesh.py:49-57)."""
    data = np.load(path, encoding="latin1", allow_pickle=True)
    return [d for d in data["D"]]


def _spmm_f32(d, x):
    d = d.tocoo()
    idx = torch.from_numpy(np.stack([d.row, d.col]).astype(np.int64))
    m = torch.sparse_coo_tensor(idx, torch.from_numpy(d.data.astype(np.float32)), d.shape, check_invariants=True)
    return torch.matmul(m, x)


def template_geometry(mean_vertices, D_list, J_regressor):
    """-> (init_vertices fp32 [431,3], vj_relation int64 [431])."""
    v = torch.as_tensor(np.asarray(mean_vertices), dtype=torch.float32)
    ds = v
    # the reference applies exactly two levels, n1=0 -> n2=2 (CoevoDecoder.py:200-202: 6890 -> 1723 -> 431), whatever the npz holds
    if len(D_list) < 2:
        raise ValueError(f"mesh_downsampling.npz holds {len(D_list)} down-sampling levels; the decoder needs 2 (6890 -> 1723 -> 431)")
    for d in D_list[:2]:
        ds = _spmm_f32(d, ds)
    joints_t = torch.matmul(torch.as_tensor(np.asarray(J_regressor), dtype=torch.float32), v).numpy()
    dsn = ds.numpy()
    d2 = ((dsn[:, None, :] - joints_t[None, :, :]) ** 2).sum(-1)
    return ds, np.argmin(d2, axis=1).astype(np.int64)


def default_paths():
    root = data_root()
    base = os.path.join(root, cfg.DATASET.BASE_DATA_DIR)
    return dict(mean_vertices=os.path.join(base, "smpl_mean_vertices.npy"),
                downsampling=os.path.join(root, "data", "base_data", "mesh_downsampling.npz"),
                j_regressor=os.path.join(root, "data", "Human36M", "J_regressor_h36m_correct.npy"))
