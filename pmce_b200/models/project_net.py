"""Weak-perspective camera layer used by the demo only (reference lib/models/project_net.py:6-22).
Kept so `import models` exposes the same four sub-modules; three parameters, plain torch, off the hot path."""
import torch
import torch.nn as nn


class OptimzeCamLayer(nn.Module):
    def __init__(self, crop_size):
        super().__init__()
        self.img_res = crop_size / 2
        self.cam_param = nn.Parameter(torch.rand((1, 3)))

    def forward(self, pose3d):
        xy = pose3d[:, :, :2] + self.cam_param[None, :, 1:]
        return xy * self.cam_param[None, :, :1] * self.img_res + self.img_res


def get_model(crop_size):
    return OptimzeCamLayer(crop_size)
