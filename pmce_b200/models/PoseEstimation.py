"""2D->3D pose lifter (spatial/temporal transformer over T-frame 2D pose + image-feature tokens).

Drop-in for reference lib/models/PoseEstimation.py: `GraphormerNet` (:31-115), `get_model` (:118-120).
The sub-module names below exist only to reproduce the checkpoint schema
(`SpatialBlocks.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}` ...); `forward` runs
`pmce_lifter_forward` (include/pmce_b200.h).
"""
import torch
import torch.nn as nn

from ..config import cfg
from ._base import EngineModule, make_dims


class _AttnParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class Block(nn.Module):
    """Parameter container of one pre-LN ViT block (reference PoseEstimation.py:13-29), mlp_ratio 2."""

    def __init__(self, dim, mlp_ratio=2.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _AttnParams(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _MlpParams(dim, int(dim * mlp_ratio))


class GraphormerNet(EngineModule):
    _engine_prefix = "pose_lifter."

    def __init__(self, num_frames=16, num_joints=17, embed_dim=256, depth=3, num_heads=8, mlp_ratio=2.0,
                 pretrained=False):
        super().__init__()
        if num_heads != 8 or mlp_ratio != 2.0:
            raise ValueError("libpmce_b200 implements the reference configuration: 8 heads, mlp_ratio 2")
        self.num_frames, self.num_joints, self.embed_dim, self.depth = num_frames, num_joints, embed_dim, depth
        self.joint_embed = nn.Linear(2, embed_dim)
        self.imgfeat_embed = nn.Linear(2048, embed_dim)
        self.spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, embed_dim))
        self.temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
        self.SpatialBlocks = nn.ModuleList([Block(embed_dim, mlp_ratio) for _ in range(depth)])
        self.TemporalBlocks = nn.ModuleList([Block(embed_dim, mlp_ratio) for _ in range(depth)])
        self.norm_s = nn.LayerNorm(embed_dim, eps=1e-6)
        self.norm_t = nn.LayerNorm(embed_dim, eps=1e-6)
        self.regression = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, 3))
        self.fusion = nn.Conv2d(in_channels=num_frames, out_channels=1, kernel_size=1)
        if pretrained:
            self._load_pretrained_model()

    def _load_pretrained_model(self):
        """reference PoseEstimation.py:71-74 (checkpoint dict with 'model_state_dict')."""
        ckpt = torch.load(cfg.MODEL.posenet_path, map_location="cpu")
        self.load_state_dict(ckpt["model_state_dict"])

    def _engine_dims(self):
        return make_dims(self.num_joints, self.embed_dim, self.depth, self.num_frames)

    @torch.no_grad()
    def forward(self, x, img_feat):
        """x [B,T,J,2], img_feat [B,T,2048] -> [B,J,3] (reference PoseEstimation.py:95-115)."""
        return self.engine().lifter(x, img_feat)


def get_model(num_joint=17, embed_dim=256, depth=3, pretrained=False):
    return GraphormerNet(num_frames=cfg.DATASET.seqlen, num_joints=num_joint, embed_dim=embed_dim, depth=depth,
                         pretrained=pretrained)
