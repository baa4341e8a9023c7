"""Top-level module: pose lifter -> /1000 -> co-evolution decoder.

Drop-in for reference lib/models/PMCE.py (`PMCE` :7-20, `get_model` :23-26). One `pmce_forward` call
(include/pmce_b200.h) covers the whole of `PMCE.forward`, replayed from a CUDA graph per batch size.
"""
import torch
import torch.nn as nn

from ..config import cfg
from . import CoevoDecoder, PoseEstimation
from ._base import EngineModule, make_dims


class PMCE(EngineModule):
    def __init__(self, num_joint, embed_dim, depth):
        super().__init__()
        self.num_joint = num_joint
        self.pose_lifter = PoseEstimation.get_model(num_joint, embed_dim, depth, pretrained=cfg.MODEL.posenet_pretrained)
        self.pose_mesh_coevo = CoevoDecoder.get_model(num_joint, embed_dim)

    def _engine_dims(self):
        pl = self.pose_lifter
        return make_dims(self.num_joint, pl.embed_dim, pl.depth, pl.num_frames)

    def _engine_vj(self):
        return self.pose_mesh_coevo.vj_relation

    @torch.no_grad()
    def forward(self, pose2d, img_feat, out=None):
        """pose2d [B,T,J,2], img_feat [B,T,2048] -> (cam_mesh [B,6890,3], cam_pose [B,J,3], pose3d [B,J,3]).
        `out` (extension, optional): caller-owned output tensors the kernels write into directly (Engine.forward)."""
        return self.engine().forward(pose2d, img_feat, out)


    @torch.no_grad()
    def forward_host(self, pose2d_cpu, img_feat_cpu, out=None):
        """Host-buffer variant of `forward` (pinned CPU tensors in, pinned CPU tensors out): what the reference's test
        loop does around `forward` with `.cuda()` / `.cpu()` (lib/core/base.py:218-238), as one call
        (`pmce_forward_host` / graph replay with direct host<->device copies)."""
        return self.engine().forward_host(pose2d_cpu, img_feat_cpu, out)


    @torch.no_grad()
    def forward_sliding(self, pose2d_seq, img_feat_seq, stride=1):
        """All stride-`stride` windows of one track in one call (what the reference's demo does window by window,
        main/run_demo.py:145 over lib/_img_utils.py:58-92 chunks): pose2d_seq [N,J,2], img_feat_seq [N,2048] ->
        (cam_mesh [nwin,6890,3], cam_pose [nwin,J,3], pose3d [nwin,J,3]); per-frame work is shared between windows."""
        return self.engine().forward_sliding(pose2d_seq, img_feat_seq, stride)

    @torch.no_grad()
    def forward_host_iter(self, batches, **hooks):
        """Pipelined `forward_host` over an iterable of (pose2d, img_feat) pinned CPU batches: copies of neighbouring batches
        overlap the forward (Engine.forward_host_iter). Yields (cam_mesh, cam_pose, pose3d) pinned CPU tensors in order."""
        return self.engine().forward_host_iter(batches, **hooks)

    @torch.no_grad()
    def forward_iter(self, batches, **hooks):
        """Device-resident pipelined loop: `for mesh, cam_pose, pose3d in model.forward_iter(cuda_batches)` - the forwards of two
        consecutive batches are in flight together (Engine.forward_iter); results are bit-identical to calling the module."""
        return self.engine().forward_iter(batches, **hooks)


def get_model(num_joint, embed_dim, depth):
    return PMCE(num_joint, embed_dim, depth)
