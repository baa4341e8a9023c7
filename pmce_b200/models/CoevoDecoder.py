"""Co-evolution decoder: GRU image-feature aggregation, 3 joint<->vertex co-evolution blocks, 431->6890
up-sampling with image-feature residual.

Drop-in for reference lib/models/CoevoDecoder.py: `Pose2Mesh` (:193-246), `get_model` (:249-252) and the
parameter schema of `AdaLayerNorm` (:16-29), `CrossAttention` (:31-62), `CrossAttentionBlock` (:64-87),
`Block` (:89-105), `Attention` (:107-131), `CoevoBlock` (:133-191). The classes below only hold parameters
under the reference's names; `Pose2Mesh.forward` runs `pmce_decoder_forward` (include/pmce_b200.h).
"""
import numpy as np
import torch
import torch.nn as nn

from ..config import cfg
from .. import mesh_assets
from ._base import EngineModule, make_dims


class AdaLayerNorm(nn.Module):
    def __init__(self, num_features, eps=1e-6):
        super().__init__()
        self.mlp_gamma = nn.Linear(2048, num_features)
        self.mlp_beta = nn.Linear(2048, num_features)
        self.eps = eps


class _MlpParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class CrossAttention(nn.Module):
    def __init__(self, dim, v_dim):
        super().__init__()
        self.wq = nn.Linear(dim, dim, bias=True)
        self.wk = nn.Linear(dim, dim, bias=True)
        self.wv = nn.Linear(v_dim, v_dim, bias=True)
        self.proj = nn.Linear(v_dim, dim)


class CrossAttentionBlock(nn.Module):
    def __init__(self, q_dim, k_dim, v_dim, mlp_ratio=4.0):
        super().__init__()
        self.normq = AdaLayerNorm(q_dim)
        self.normk = AdaLayerNorm(k_dim)
        self.normv = AdaLayerNorm(v_dim)
        self.attn = CrossAttention(q_dim, v_dim)
        self.norm2 = AdaLayerNorm(q_dim)
        self.mlp = _MlpParams(q_dim, int(q_dim * mlp_ratio))


class Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class Block(nn.Module):
    def __init__(self, dim, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = AdaLayerNorm(dim)
        self.attn = Attention(dim)
        self.norm2 = AdaLayerNorm(dim)
        self.mlp = _MlpParams(dim, int(dim * mlp_ratio))


class CoevoBlock(nn.Module):
    """Parameters of one co-evolution block; joint stream 8 heads x 8, vertex stream 2 heads x 32 (:139-140)."""

    def __init__(self, num_joint, num_vertx, joint_dim=64, vertx_dim=64):
        super().__init__()
        if joint_dim != 64 or vertx_dim != 64:
            raise ValueError("libpmce_b200 implements cfg.MODEL.joint_dim == vertx_dim == 64")
        self.num_joint, self.num_vertx = num_joint, num_vertx
        self.joint_proj = nn.Linear(3, joint_dim)
        self.vertx_proj = nn.Linear(3, vertx_dim)
        self.joint_pos_embed = nn.Parameter(torch.randn(1, num_joint, joint_dim))
        self.vertx_pos_embed = nn.Parameter(torch.randn(1, num_vertx, vertx_dim))
        self.j_Q_embed = nn.Parameter(torch.randn(1, num_joint, joint_dim))
        self.v_Q_embed = nn.Parameter(torch.randn(1, num_vertx, vertx_dim))
        self.proj_v2j_dim = nn.Linear(vertx_dim, joint_dim)
        self.proj_j2v_dim = nn.Linear(joint_dim, vertx_dim)
        self.v2j_K_embed = nn.Parameter(torch.randn(1, num_vertx, joint_dim))
        self.j2v_K_embed = nn.Parameter(torch.randn(1, num_joint, vertx_dim))
        self.joint_SA_FFN = Block(joint_dim)
        self.vertx_SA_FFN = Block(vertx_dim)
        self.joint_CA_FFN = CrossAttentionBlock(joint_dim, joint_dim, vertx_dim)
        self.vertx_CA_FFN = CrossAttentionBlock(vertx_dim, vertx_dim, joint_dim)
        self.proj_joint_feat2coor = nn.Linear(joint_dim, 3)
        self.proj_vertx_feat2coor = nn.Linear(vertx_dim, 3)


class Pose2Mesh(EngineModule):
    _engine_prefix = "pose_mesh_coevo."

    def __init__(self, num_joint, embed_dim=256, SMPL_MEAN_vertices=None, mesh_downsampling=None, J_regressor=None):
        super().__init__()
        paths = mesh_assets.default_paths()
        mean_v = np.load(SMPL_MEAN_vertices or paths["mean_vertices"])
        D_list = mesh_assets.load_downsampling(mesh_downsampling or paths["downsampling"])
        jreg = np.load(J_regressor or paths["j_regressor"]).astype(np.float32)
        init_vertices, vj = mesh_assets.template_geometry(mean_v, D_list, jreg)
        self.register_buffer("init_vertices", init_vertices)
        self.num_verts = init_vertices.shape[0]
        self.num_joint = num_joint
        self.seqlen = cfg.DATASET.seqlen
        # the reference keeps a float64 ndarray (graph_utils.py:33); same values, integer dtype here
        self.vj_relation = vj

        jd, vd = cfg.MODEL.joint_dim, cfg.MODEL.vertx_dim
        self.coevoblock1 = CoevoBlock(num_joint, self.num_verts, jd, vd)
        self.coevoblock2 = CoevoBlock(num_joint, self.num_verts, jd, vd)
        self.coevoblock3 = CoevoBlock(num_joint, self.num_verts, jd, vd)
        self.upsample_conv = nn.Conv1d(self.num_verts, 6890, kernel_size=3, padding=1)
        self.gru_cur = nn.GRU(input_size=2048, hidden_size=1024, bidirectional=True, num_layers=2)
        self.linear_cur1 = nn.Linear(2048, 6890)
        self.linear_cur2 = nn.Linear(2048, 6890)
        self.linear_cur3 = nn.Linear(2048, 6890)

    def _engine_dims(self):
        return make_dims(self.num_joint, 256, 3, self.seqlen)   # lifter dims unused by the decoder entry points

    def _engine_vj(self):
        return self.vj_relation

    @torch.no_grad()
    def forward(self, joints, img_feats):
        """joints [B,J,3] (metres), img_feats [B,T,2048] -> (joints3 [B,J,3], mesh [B,6890,3]) (:226-246)."""
        return self.engine().decoder(joints, img_feats)


def get_model(num_joint, embed_dim):
    return Pose2Mesh(num_joint, embed_dim)
