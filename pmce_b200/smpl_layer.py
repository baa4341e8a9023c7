"""SMPL linear-blend-skinning layer on libpmce_b200 (`smpl_lbs_forward`, include/pmce_b200.h).

Drop-in for reference smplpytorch/smplpytorch/pytorch/smpl_layer.py: `SMPL_Layer.__init__` (:14-63) registers
the same buffers (`th_betas`, `th_shapedirs`, `th_posedirs`, `th_v_template`, `th_J_regressor`, `th_weights`,
`th_faces`) and `forward(th_pose_axisang, th_betas, th_trans)` (:65-158) returns `(verts [B,6890,3],
joints [B,24,3])`. The licensed `basicModel_*_lbs_10_207_0_v1.0.0.pkl` files are not shipped with the
reference; `from_buffers` builds the layer from already-loaded arrays (used with synthetic buffers in tests).
"""
import ctypes as C
import os
import pickle

import numpy as np
import torch
from torch.nn import Module

from . import _lib
from ._lib import PmceError, check
from .engine import _ptr, _stream, _require_cuda_f32


class SMPL_Layer(Module):
    __constants__ = ["kintree_parents", "gender", "center_idx", "num_joints"]

    def __init__(self, center_idx=None, gender="neutral", model_root="smpl/native/models", _buffers=None):
        super().__init__()
        self.center_idx = center_idx
        self.gender = gender
        if _buffers is None:
            name = {"neutral": "basicModel_neutral_lbs_10_207_0_v1.0.0.pkl", "female": "basicModel_f_lbs_10_207_0_v1.0.0.pkl",
                    "male": "basicModel_m_lbs_10_207_0_v1.0.0.pkl"}[gender]
            self.model_path = os.path.join(model_root, name)
            _buffers = self._load_pkl(self.model_path)
        for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights"):
            self.register_buffer(k, torch.as_tensor(_buffers[k], dtype=torch.float32).clone())
        if "th_faces" in _buffers:
            self.register_buffer("th_faces", torch.as_tensor(_buffers["th_faces"]).long())
        self.kintree_parents = [int(p) for p in _buffers["kintree_parents"]]
        self.num_joints = len(self.kintree_parents)
        if self.num_joints != 24 or tuple(self.th_v_template.shape) != (1, 6890, 3):
            raise ValueError("libpmce_b200 implements the 24-joint / 6890-vertex SMPL body model")
        self.vertice_segmentation = torch.argmax(self.th_weights, dim=1)
        object.__setattr__(self, "_packed", None)

    @classmethod
    def from_buffers(cls, buffers, center_idx=None, gender="neutral"):
        return cls(center_idx=center_idx, gender=gender, _buffers=buffers)

    @staticmethod
    def _load_pkl(path):
        """Equivalent of `ready_arguments` (reference native/webuser/serialization.py:1-39) for the fields LBS needs."""
        if not os.path.exists(path):
            raise FileNotFoundError(f"SMPL model file {path} not found (licensed asset, not shipped)")
        with open(path, "rb") as f:
            dd = pickle.load(f, encoding="latin1")   # needs chumpy importable, as in the reference

        def arr(x):
            return np.array(getattr(x, "r", x))
        J = dd["J_regressor"]
        J = np.array(J.toarray()) if hasattr(J, "toarray") else np.array(J)
        return dict(th_betas=np.zeros((1, 10), np.float32) if "betas" not in dd else arr(dd["betas"]).reshape(1, -1)[:, :10],
                    th_shapedirs=arr(dd["shapedirs"])[:, :, :10], th_posedirs=arr(dd["posedirs"]),
                    th_v_template=arr(dd["v_template"])[None], th_J_regressor=J, th_weights=arr(dd["weights"]),
                    th_faces=np.array(dd["f"]).astype(np.int64), kintree_parents=list(np.array(dd["kintree_table"])[0].tolist()))

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        object.__setattr__(self, "_packed", None)
        return out

    def _pack(self):
        dev = self.th_v_template.device
        if dev.type != "cuda":
            raise PmceError("SMPL_Layer runs on CUDA only: call .cuda() first (there is no CPU path)")
        if self._packed is not None and self._packed["dev"] == dev:
            return self._packed
        lib = _lib.load()
        ld = lib.smpl_blend_ld()
        V = 6890
        blend = torch.zeros(V * 3, ld, device=dev)
        blend[:, :10] = self.th_shapedirs.reshape(V * 3, 10)
        blend[:, 10:217] = self.th_posedirs.reshape(V * 3, 207)
        vt = self.th_v_template.reshape(V * 3).contiguous()
        # tensor-core copies: blend matrix and template padded to 20672 rows (the GEMM epilogue's 16-column chunk), split-bf16
        npad = 20672
        blend_pad = torch.zeros(npad, ld, device=dev)
        blend_pad[:V * 3] = blend
        vt_pad = torch.zeros(npad, device=dev)
        vt_pad[:V * 3] = vt
        blend_hi = torch.empty(npad, ld, dtype=torch.bfloat16, device=dev)
        blend_lo = torch.empty(npad, ld, dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            check(lib.pmce_split_bf16(_ptr(blend_pad), npad, ld, _ptr(blend_hi), _ptr(blend_lo), _stream()), "pmce_split_bf16")
        Jreg = self.th_J_regressor.double()
        j_template = (Jreg @ self.th_v_template[0].double()).float().contiguous()                       # [24,3]
        j_shapedirs = torch.einsum("jv,vck->jck", Jreg, self.th_shapedirs.double()).float().contiguous()  # [24,3,10]
        parents = torch.as_tensor(np.array([max(p, 0) if i else 0 for i, p in enumerate(self.kintree_parents)],
                                           dtype=np.int32) % 24, device=dev)
        # sparse skinning table: the <= 4 (joint, weight) pairs of each vertex, ascending joint order (dense fallback otherwise)
        wts = self.th_weights
        idx4 = w4 = None
        if int((wts != 0).sum(dim=1).max()) <= 4:
            order = torch.argsort((wts == 0).to(torch.int8), dim=1, stable=True)[:, :4]      # non-zeros first, joint order kept
            order, _ = torch.sort(order, dim=1)
            w4 = torch.gather(wts, 1, order).contiguous()
            idx4 = order.to(torch.int32).contiguous()
        packed = dict(dev=dev, lib=lib, blend=blend, vt=vt, jt=j_template, js=j_shapedirs, parents=parents,
                      weights=self.th_weights.contiguous(), ws=None, blend_hi=blend_hi, blend_lo=blend_lo, vt_pad=vt_pad, idx4=idx4, w4=w4)
        object.__setattr__(self, "_packed", packed)
        return packed

    @torch.no_grad()
    def get_smpl_coord(self, pose_param, shape_param, trans_param):
        """Batched form of the datasets' `get_smpl_coord` (reference data/PW3D/dataset.py:70-88, data/Human36M/dataset.py,
        data/MPII3D/dataset.py ...): pose [B,72], shape [B,10], trans [B,3] CUDA tensors ->
        (smpl_mesh_coord [B,6890,3], smpl_joint_coord [B,24,3]) in MILLIMETRES, one launch sequence for the whole batch
        instead of one CPU SMPL_Layer call per sample inside the DataLoader workers."""
        return self.forward(pose_param, shape_param, trans_param, _out_scale=1000.0)

    @torch.no_grad()
    def forward(self, th_pose_axisang, th_betas=torch.zeros(1), th_trans=torch.zeros(1), _out_scale=1.0):
        p = self._pack()
        lib, dev = p["lib"], p["dev"]
        B = th_pose_axisang.shape[0]
        pose = _require_cuda_f32(th_pose_axisang, "th_pose_axisang", (B, 72))
        # same branch conditions as the reference (smpl_layer.py:87,148); the 1-element default never syncs
        if th_betas is None or th_betas.numel() == 1 or bool(torch.norm(th_betas) == 0):
            betas = self.th_betas.expand(B, 10).contiguous()
        else:
            betas = _require_cuda_f32(th_betas, "th_betas", (B, 10))
        if th_trans is None or th_trans.numel() == 1 or bool(torch.norm(th_trans) == 0):
            trans = None
        else:
            trans = _require_cuda_f32(th_trans, "th_trans", (B, 3))
        with torch.cuda.device(dev):
            need = lib.smpl_workspace_bytes(B)
            if p["ws"] is None or p["ws"].numel() < need:
                p["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
            verts = torch.empty(B, 6890, 3, device=dev)
            joints = torch.empty(B, 24, 3, device=dev)
            tcp = os.environ.get("PMCE_SMPL_FP32", "0") != "1"      # PMCE_SMPL_FP32=1: exact fp32 CUDA-core blend-shape GEMM
            dense = os.environ.get("PMCE_SMPL_DENSE_SKIN", "0") == "1"    # A/B knob: the dense 24-joint blend
            check(lib.smpl_lbs_forward_sparse(_ptr(p["blend"]), _ptr(p["blend_hi"] if tcp else None), _ptr(p["blend_lo"] if tcp else None),
                                              _ptr(p["vt_pad"] if tcp else p["vt"]), _ptr(p["jt"]), _ptr(p["js"]), _ptr(p["weights"]),
                                              _ptr(None if dense else p["idx4"]), _ptr(None if dense else p["w4"]),
                                              _ptr(p["parents"]), _ptr(pose), _ptr(betas), _ptr(trans), B, float(_out_scale),
                                              _ptr(verts), _ptr(joints), _ptr(p["ws"]), p["ws"].numel(), _stream()),
                  "smpl_lbs_forward")
        if trans is None and self.center_idx is not None:   # not used by PMCE (lib/smpl.py:50-51 passes none)
            center = joints[:, self.center_idx].unsqueeze(1).clone()
            joints -= center
            verts -= center
        return verts, joints
