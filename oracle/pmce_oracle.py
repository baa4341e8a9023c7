"""CPU restatement of the PMCE per-clip forward hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this file; the product (`pmce_b200/`) never does and fails loudly without its CUDA
library.

It is a *functional* restatement (plain tensor math over a `state_dict`, no nn.Module tree) of:
  * `PMCE.forward`            reference lib/models/PMCE.py:15-20
  * `GraphormerNet.forward`   reference lib/models/PoseEstimation.py:76-115 (+ `Block` :13-29)
  * `Pose2Mesh.forward`       reference lib/models/CoevoDecoder.py:226-246
  * `CoevoBlock.forward`      reference lib/models/CoevoDecoder.py:175-191
  * `AdaLayerNorm.forward`    reference lib/models/CoevoDecoder.py:23-29
  * `CrossAttention.forward`  reference lib/models/CoevoDecoder.py:47-62
  * `Attention.forward`       reference lib/models/CoevoDecoder.py:119-131 (same math as timm's)
  * init-time template down-sampling + nearest-joint relation
                              reference lib/models/CoevoDecoder.py:199-208, lib/graph_utils.py:27-46,
                              lib/models/backbones/mesh.py:81-96
  * J-regressor matvec        reference lib/core/base.py:225
  * evaluation epilogue       reference lib/core/base.py:223-227 + data/PW3D/dataset.py:269-282 (`compute_both_err`)
  * `SMPL_Layer.forward`      reference smplpytorch/smplpytorch/pytorch/smpl_layer.py:65-158
The GRU follows PyTorch's documented `nn.GRU` gate equations (rows of weight_ih/hh ordered r,z,n).

Pinning: the reference ships NO golden vectors or tests. This oracle is pinned
against outputs of the reference itself, run in the build container through
`oracle/ref_harness.py`; the resulting fixtures are `tests/golden/*.npz` (generator:
`oracle/gen_golden.py`) and `tests/test_oracle_golden.py` re-checks the oracle against them on any
host. The `timm` Mlp/Attention arithmetic is third-party, unpinned by the reference and restated
from timm's published layers: parity is "unpinned" at that one boundary (see shims/timm).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------

def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(sd, prefix, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _mhsa(q, k, v, heads):
    """softmax(q k^T d^-1/2) v with heads split h-major over the feature axis. q:[B,N1,C] k,v:[B,N2,C]."""
    B, N1, C = q.shape
    N2 = k.shape[1]
    d = C // heads
    qh = q.reshape(B, N1, heads, d).permute(0, 2, 1, 3)
    kh = k.reshape(B, N2, heads, d).permute(0, 2, 1, 3)
    vh = v.reshape(B, N2, heads, v.shape[-1] // heads).permute(0, 2, 1, 3)
    attn = (qh @ kh.transpose(-2, -1)) * (d ** -0.5)
    attn = attn.softmax(dim=-1)
    return (attn @ vh).transpose(1, 2).reshape(B, N1, -1)


def _self_attention(sd, prefix, x, heads):
    """Fused-qkv attention (CoevoDecoder.py:119-131 / timm Attention)."""
    B, N, C = x.shape
    qkv = _lin(sd, prefix + ".qkv", x)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    return _lin(sd, prefix + ".proj", _mhsa(q, k, v, heads))


def _mlp(sd, prefix, x):
    return _lin(sd, prefix + ".fc2", F.gelu(_lin(sd, prefix + ".fc1", x)))


def _vit_block(sd, prefix, x, heads, eps=1e-6):
    """Pre-LN ViT block (PoseEstimation.py:26-29)."""
    x = x + _self_attention(sd, prefix + ".attn", _ln(sd, prefix + ".norm1", x, eps), heads)
    x = x + _mlp(sd, prefix + ".mlp", _ln(sd, prefix + ".norm2", x, eps))
    return x


# ----------------------------------------------------------------------------------------------
# pose lifter (PoseEstimation.py:76-115)
# ----------------------------------------------------------------------------------------------

def lifter_forward(sd, pose2d, img_feat, depth=3, heads=8, prefix="pose_lifter."):
    B, T, J, _ = pose2d.shape
    p = prefix
    C = sd[p + "joint_embed.weight"].shape[0]
    x = _lin(sd, p + "joint_embed", pose2d)                                   # [B,T,J,C]
    x = x + _lin(sd, p + "imgfeat_embed", img_feat)[:, :, None, :]
    x = x + sd[p + "spatial_pos_embed"].reshape(1, 1, J, C)
    for i in range(depth):
        xs = x.reshape(B * T, J, C)
        xs = _vit_block(sd, f"{p}SpatialBlocks.{i}", xs, heads)
        xs = _ln(sd, p + "norm_s", xs, 1e-6)
        x = xs.reshape(B, T, J, C)
        xt = x.permute(0, 2, 1, 3).reshape(B * J, T, C)
        if i == 0:
            xt = xt + sd[p + "temporal_pos_embed"].reshape(1, T, C)
        xt = _vit_block(sd, f"{p}TemporalBlocks.{i}", xt, heads)
        xt = _ln(sd, p + "norm_t", xt, 1e-6)
        x = xt.reshape(B, J, T, C).permute(0, 2, 1, 3)
    r = _lin(sd, p + "regression.1", _ln(sd, p + "regression.0", x, 1e-5))    # [B,T,J,3]
    w = sd[p + "fusion.weight"].reshape(T)
    return torch.einsum("t,btjc->bjc", w, r) + sd[p + "fusion.bias"]


# ----------------------------------------------------------------------------------------------
# GRU image-feature aggregation (CoevoDecoder.py:216-221,228-229)
# ----------------------------------------------------------------------------------------------

def _gru_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one GRU layer. x: [T,B,I] -> [T,B,H]; h0 = 0."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    out = [None] * T
    gi_all = F.linear(x, w_ih, b_ih)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gi = gi_all[t]
        gh = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        out[t] = h
    return torch.stack(out, 0)


def gru_mid(sd, img_feat, prefix="pose_mesh_coevo.gru_cur."):
    """y[T//2] of the 2-layer bidirectional GRU; img_feat [B,T,2048] -> [B,2048]."""
    x = img_feat.permute(1, 0, 2)
    T = x.shape[0]
    for layer in (0, 1):
        outs = []
        for sfx, rev in (("", False), ("_reverse", True)):
            outs.append(_gru_dir(x, sd[f"{prefix}weight_ih_l{layer}{sfx}"], sd[f"{prefix}weight_hh_l{layer}{sfx}"],
                                 sd[f"{prefix}bias_ih_l{layer}{sfx}"], sd[f"{prefix}bias_hh_l{layer}{sfx}"], rev))
        x = torch.cat(outs, dim=-1)
    return x[T // 2]


# ----------------------------------------------------------------------------------------------
# co-evolution decoder (CoevoDecoder.py:16-246)
# ----------------------------------------------------------------------------------------------

def adaln(sd, prefix, x, g, eps=1e-6):
    """gamma(g) * (x - mean) / (unbiased_std + eps) + beta(g)  (CoevoDecoder.py:23-29)."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)  # unbiased (n-1)
    gamma = _lin(sd, prefix + ".mlp_gamma", g)[:, None, :]
    beta = _lin(sd, prefix + ".mlp_beta", g)[:, None, :]
    return gamma * (x - mean) / (std + eps) + beta


def cross_attention_block(sd, prefix, xq, xk, xv, g, heads):
    """CrossAttentionBlock.forward (CoevoDecoder.py:82-87) with CrossAttention (:47-62)."""
    q = _lin(sd, prefix + ".attn.wq", adaln(sd, prefix + ".normq", xq, g))
    k = _lin(sd, prefix + ".attn.wk", adaln(sd, prefix + ".normk", xk, g))
    v = _lin(sd, prefix + ".attn.wv", adaln(sd, prefix + ".normv", xv, g))
    xq = xq + _lin(sd, prefix + ".attn.proj", _mhsa(q, k, v, heads))
    xq = xq + _mlp(sd, prefix + ".mlp", adaln(sd, prefix + ".norm2", xq, g))
    return xq


def self_attention_block(sd, prefix, x, g, heads):
    """Block.forward (CoevoDecoder.py:102-105)."""
    x = x + _self_attention(sd, prefix + ".attn", adaln(sd, prefix + ".norm1", x, g), heads)
    x = x + _mlp(sd, prefix + ".mlp", adaln(sd, prefix + ".norm2", x, g))
    return x


def coevo_block(sd, prefix, joints, verts, g, joint_heads=8, vertx_heads=2):
    """CoevoBlock.forward (CoevoDecoder.py:175-191): both cross-attentions read the PRE-update features."""
    jf = _lin(sd, prefix + "joint_proj", joints) + sd[prefix + "joint_pos_embed"]
    vf = _lin(sd, prefix + "vertx_proj", verts) + sd[prefix + "vertx_pos_embed"]
    jf1 = cross_attention_block(sd, prefix + "joint_CA_FFN", jf + sd[prefix + "j_Q_embed"],
                                _lin(sd, prefix + "proj_v2j_dim", vf) + sd[prefix + "v2j_K_embed"], vf, g, joint_heads)
    vf1 = cross_attention_block(sd, prefix + "vertx_CA_FFN", vf + sd[prefix + "v_Q_embed"],
                                _lin(sd, prefix + "proj_j2v_dim", jf) + sd[prefix + "j2v_K_embed"], jf, g, vertx_heads)
    jf2 = self_attention_block(sd, prefix + "joint_SA_FFN", jf1, g, joint_heads)
    vf2 = self_attention_block(sd, prefix + "vertx_SA_FFN", vf1, g, vertx_heads)
    return (_lin(sd, prefix + "proj_joint_feat2coor", jf2) + joints,
            _lin(sd, prefix + "proj_vertx_feat2coor", vf2) + verts)


def decoder_forward(sd, joints, img_feat, vj_relation, prefix="pose_mesh_coevo.", return_intermediates=False):
    """Pose2Mesh.forward (CoevoDecoder.py:226-246). joints [B,J,3] (metres), img_feat [B,T,2048]."""
    p = prefix
    g = gru_mid(sd, img_feat, p + "gru_cur.")
    idx = torch.as_tensor(np.asarray(vj_relation), dtype=torch.long)
    verts = joints[:, idx, :3]                                   # pure copy -> bit exact
    inter = {"g": g, "verts0": verts}
    j_out = None
    for k in (1, 2, 3):
        j_out, verts = coevo_block(sd, f"{p}coevoblock{k}.", joints, verts, g)   # always the ORIGINAL joints
        inter[f"verts{k}"] = verts
        inter[f"joints{k}"] = j_out
    mesh = F.conv1d(verts, sd[p + "upsample_conv.weight"], sd[p + "upsample_conv.bias"], padding=1)
    rg = F.relu(g)
    res = torch.stack([_lin(sd, f"{p}linear_cur{i}", rg) for i in (1, 2, 3)], dim=-1)
    mesh = mesh + res
    if return_intermediates:
        return j_out, mesh, inter
    return j_out, mesh


def pmce_forward(sd, pose2d, img_feat, vj_relation, depth=3, return_intermediates=False):
    """PMCE.forward (PMCE.py:15-20) -> (cam_mesh [B,6890,3], cam_pose [B,J,3], pose3d [B,J,3])."""
    pose3d = lifter_forward(sd, pose2d, img_feat, depth=depth)
    out = decoder_forward(sd, pose3d / 1000, img_feat, vj_relation, return_intermediates=return_intermediates)
    if return_intermediates:
        return out[1], out[0], pose3d, out[2]
    return out[1], out[0], pose3d


def j_regress(J_regressor, mesh):
    """`torch.matmul(J_regressor[None], pred_mesh)` (lib/core/base.py:225), dense fp32."""
    return torch.matmul(J_regressor[None, :, :], mesh)


H36M_EVAL_JOINTS = (1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16)   # data/PW3D/dataset.py:35


def compute_both_err(pred_mesh, target_mesh, pred_joint, target_joint, eval_joints=H36M_EVAL_JOINTS):
    """The test loop's evaluation epilogue (data/PW3D/dataset.py:269-282, called from lib/core/base.py:227): root-align
    meshes and joints on joint 0, keep the 14 evaluation joints, mean per-joint / per-vertex L2 error (numpy fp32).
    Returns (joint_mean_error, mesh_mean_error)."""
    pred_mesh, target_mesh = pred_mesh - pred_joint[:, :1, :], target_mesh - target_joint[:, :1, :]
    pred_joint, target_joint = pred_joint - pred_joint[:, :1, :], target_joint - target_joint[:, :1, :]
    pm, tm = pred_mesh.detach().cpu().numpy(), target_mesh.detach().cpu().numpy()
    pj, tj = pred_joint.detach().cpu().numpy(), target_joint.detach().cpu().numpy()
    pj, tj = pj[:, eval_joints, :], tj[:, eval_joints, :]
    mesh_mean_error = np.power((np.power((pm - tm), 2)).sum(axis=2), 0.5).mean()
    joint_mean_error = np.power((np.power((pj - tj), 2)).sum(axis=2), 0.5).mean()
    return joint_mean_error, mesh_mean_error


def eval_step(J_regressor, cam_mesh, gt_mesh, gt_pose3d, eval_joints=H36M_EVAL_JOINTS):
    """lib/core/base.py:223-227: metres -> millimetres, J-regressor, compute_both_err. -> (pred_pose, j_error, s_error)."""
    pred_mesh, gt_mesh = cam_mesh * 1000, gt_mesh * 1000
    pred_pose = torch.matmul(J_regressor[None, :, :], pred_mesh)
    j_error, s_error = compute_both_err(pred_mesh, gt_mesh, pred_pose, gt_pose3d, eval_joints)
    return pred_pose, j_error, s_error


# ----------------------------------------------------------------------------------------------
# init-time geometry (host side in the reference too)
# ----------------------------------------------------------------------------------------------

def downsample_template(verts, D_list):
    """6890 -> 1723 -> 431 by sparse D matrices in fp32 (CoevoDecoder.py:200-202, mesh.py:81-96)."""
    x = torch.as_tensor(np.asarray(verts), dtype=torch.float32)
    for d in D_list:
        d = d.tocoo()
        m = torch.sparse_coo_tensor(np.array([d.row, d.col]), torch.as_tensor(d.data, dtype=torch.float32), d.shape)
        x = torch.matmul(m, x)
    return x


def nearest_joint_relation(joints_template, verts_ds):
    """argmin_j |v - joint_j|^2 per down-sampled vertex (graph_utils.py:27-46), int64 indices."""
    jt = np.asarray(joints_template)
    out = np.zeros(len(verts_ds), dtype=np.int64)
    for i, v in enumerate(np.asarray(verts_ds)):
        out[i] = int(np.argmin(((v - jt) ** 2).sum(1)))
    return out


def init_geometry(verts, D_list, J_regressor_h36m):
    """(init_vertices[431,3] fp32, vj_relation[431] int64) as `Pose2Mesh.__init__` builds them."""
    v = torch.as_tensor(np.asarray(verts), dtype=torch.float32)
    init_vertices = downsample_template(v, D_list)
    jt = torch.matmul(torch.as_tensor(np.asarray(J_regressor_h36m), dtype=torch.float32), v)
    return init_vertices, nearest_joint_relation(jt.numpy(), init_vertices.numpy())


# ----------------------------------------------------------------------------------------------
# SMPL linear blend skinning (smplpytorch/smplpytorch/pytorch/smpl_layer.py:65-158)
# ----------------------------------------------------------------------------------------------

def _rodrigues(aa):
    """axis-angle [N,3] -> rotmat [N,3,3] via quaternion (rodrigues_layer.py:13-52): theta = |a+1e-8|."""
    theta = torch.norm(aa + 1e-8, p=2, dim=1, keepdim=True)
    axis = aa / theta
    half = theta * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def smpl_lbs(buf, pose, betas=None, trans=None):
    """SMPL_Layer.forward -> (verts [B,6890,3], joints [B,24,3]); center_idx is None in PMCE's use."""
    B = pose.shape[0]
    parents = list(buf["kintree_parents"])
    nj = len(parents)
    R = _rodrigues(pose.reshape(B * nj, 3)).view(B, nj, 3, 3)
    pose_map = (R[:, 1:] - torch.eye(3)).reshape(B, (nj - 1) * 9)
    if betas is None or bool(torch.norm(betas) == 0):
        betas_used = buf["th_betas"].expand(B, -1)
    else:
        betas_used = betas
    v_shaped = buf["th_v_template"] + torch.einsum("vck,bk->bvc", buf["th_shapedirs"], betas_used)
    Jr = torch.einsum("jv,bvc->bjc", buf["th_J_regressor"], v_shaped)
    v_posed = v_shaped + torch.einsum("vck,bk->bvc", buf["th_posedirs"], pose_map)

    def with_zeros(Rm, t):
        top = torch.cat([Rm, t.reshape(B, 3, 1)], dim=2)
        bottom = torch.tensor([0.0, 0.0, 0.0, 1.0]).view(1, 1, 4).expand(B, 1, 4)
        return torch.cat([top, bottom], dim=1)

    G = [with_zeros(R[:, 0], Jr[:, 0])]
    for i in range(1, nj):
        G.append(torch.matmul(G[parents[i]], with_zeros(R[:, i], Jr[:, i] - Jr[:, parents[i]])))
    A = []
    for i in range(nj):
        jh = torch.cat([Jr[:, i], Jr.new_zeros(B, 1)], dim=1).unsqueeze(2)
        corr = torch.bmm(G[i], jh)                                            # [B,4,1]
        A.append(G[i] - torch.cat([Jr.new_zeros(B, 4, 3), corr], dim=2))
    A = torch.stack(A, dim=3)                                                 # [B,4,4,nj]
    Tm = torch.matmul(A, buf["th_weights"].t())                               # [B,4,4,V]
    vh = torch.cat([v_posed.transpose(2, 1), v_posed.new_ones(B, 1, v_posed.shape[1])], dim=1)
    verts = (Tm * vh.unsqueeze(1)).sum(2).transpose(2, 1)[:, :, :3]
    jtr = torch.stack(G, dim=1)[:, :, :3, 3]
    if not (trans is None or bool(torch.norm(trans) == 0)):
        jtr = jtr + trans.unsqueeze(1)
        verts = verts + trans.unsqueeze(1)
    return verts, jtr
