"""(f)2: the SPIN / HMR ResNet-50 feature extractor on the device (reference lib/models/spin.py:60-143).

`HMR` here mirrors the part of the reference class the PMCE pipeline uses: the trunk's parameter schema (`conv1`, `bn1`,
`layer1..4` with `Bottleneck` blocks, so a SPIN checkpoint's trunk tensors load with `load_state_dict(strict=False)` exactly as
in the reference) and `feature_extractor(x)` (`main/run_demo.py:315`): frame crops `[N,3,224,224]` -> features `[N,2048]`.
The regression head (fc1/fc2/decpose/... + the SMPL layer, spin.py:78-94,145-202) is OUT OF SCOPE: the demo never calls it.

Pack time (host, torch): eval-mode BatchNorm is folded into each convolution (`w * gamma / sqrt(var + eps)`, bias `beta - mean *
gamma / sqrt(var + eps)`), weights are laid out `[cout, kh, kw, cin]` (the im2col order of the NHWC kernels) and split into
bf16 hi/lo. Run time: `pmce_spin_features` (csrc/spin.cuh + the tcgen05 GEMM). No CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import PmceError, PmceSpinConv, check
from .synth import SPIN_LAYERS

STEM_K = 152


class _BN(nn.Module):
    """Parameter container with nn.BatchNorm2d's state_dict keys."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Conv(nn.Module):
    def __init__(self, cout, cin, k):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, k, k))


class Bottleneck(nn.Module):
    """Parameter schema of reference spin.py:17-36 (the arithmetic runs in libpmce_b200)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=False):
        super().__init__()
        self.conv1, self.bn1 = _Conv(planes, inplanes, 1), _BN(planes)
        self.conv2, self.bn2 = _Conv(planes, planes, 3), _BN(planes)
        self.conv3, self.bn3 = _Conv(planes * 4, planes, 1), _BN(planes * 4)
        self.downsample = nn.Sequential(_Conv(planes * 4, inplanes, 1), _BN(planes * 4)) if downsample else None
        self.stride = stride


class HMR(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.bn1 = _Conv(64, 3, 7), _BN(64)
        inpl = 64
        for li, (planes, blocks, stride) in enumerate(SPIN_LAYERS, 1):
            layer = [Bottleneck(inpl, planes, stride, downsample=True)]
            inpl = planes * 4
            layer += [Bottleneck(inpl, planes) for _ in range(1, blocks)]
            setattr(self, f"layer{li}", nn.Sequential(*layer))
        self._packed = None
        self._ws = None

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._packed = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    # ---- pack: fold BN, reorder, split ------------------------------------------------------------------------------
    @staticmethod
    def _fold(conv, bn):
        w = conv.weight.detach().float()
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + 1e-5)
        bias = bn.bias.detach().float() - bn.running_mean.float() * scale
        w = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(w.shape[0], -1)        # [cout, kh*kw*cin]
        return w, bias

    def _pack(self, dev):
        lib = _lib.load()
        mats = [self._fold(self.conv1, self.bn1)]
        mats[0] = (torch.nn.functional.pad(mats[0][0], (0, STEM_K - 147)), mats[0][1])
        for li in range(1, 5):
            for blk in getattr(self, f"layer{li}"):
                mats += [self._fold(blk.conv1, blk.bn1), self._fold(blk.conv2, blk.bn2), self._fold(blk.conv3, blk.bn3)]
                if blk.downsample is not None:
                    mats.append(self._fold(blk.downsample[0], blk.downsample[1]))
        if len(mats) != lib.pmce_spin_num_convs():
            raise PmceError("spin: unexpected number of convolutions")
        table = (PmceSpinConv * len(mats))()
        off, parts = 0, []
        for i, (w, b) in enumerate(mats):          # weights first (64-element aligned blocks), biases after
            table[i].w_off, table[i].cout, table[i].k = off, w.shape[0], w.shape[1]
            n = (w.numel() + 63) // 64 * 64
            parts.append(torch.nn.functional.pad(w.reshape(-1), (0, n - w.numel())))
            off += n
        for i, (w, b) in enumerate(mats):
            table[i].b_off = off
            n = (b.numel() + 63) // 64 * 64
            parts.append(torch.nn.functional.pad(b, (0, n - b.numel())))
            off += n
        blob = torch.cat(parts).to(dev).contiguous()
        hi = torch.empty(off, dtype=torch.bfloat16, device=dev)
        lo = torch.empty(off, dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(lib.pmce_split_bf16(C.c_void_p(blob.data_ptr()), off // 64, 64, C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()), st), "pmce_split_bf16")
        self._packed = dict(lib=lib, blob=blob, hi=hi, lo=lo, table=table, n=len(mats), dev=dev)
        return self._packed

    # ---- the reference's entry point ------------------------------------------------------------------------------------
    @torch.no_grad()
    def feature_extractor(self, x):
        """x [N,3,224,224] fp32 CUDA (the crops of main/run_demo.py:315) -> xf [N,2048] (spin.py:129-143)."""
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise PmceError("spin.HMR.feature_extractor: needs a CUDA tensor (pmce_b200 has no CPU path)")
        if x.dtype != torch.float32 or x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise PmceError(f"spin.HMR.feature_extractor: expected float32 [N,3,224,224], got {x.dtype} {tuple(x.shape)}")
        x = x.contiguous()
        dev = x.device
        pk = self._packed if self._packed is not None and self._packed["dev"] == dev else self._pack(dev)
        lib, B = pk["lib"], x.shape[0]
        need = lib.pmce_spin_workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.empty(B, 2048, device=dev)
        P = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            check(lib.pmce_spin_features(P(pk["blob"]), P(pk["hi"]), P(pk["lo"]), C.cast(pk["table"], C.c_void_p), pk["n"], P(x), B, P(out),
                                         P(self._ws), self._ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pmce_spin_features")
        return out

    forward = feature_extractor


def hmr(pretrained=False, **kwargs):
    """reference `hmr(smpl_mean_params, pretrained=True)` (spin.py:296-304) builds the ResNet-50 HMR; here: the trunk only."""
    return HMR()
