"""Host-side driver of libpmce_b200: packed weights, workspace, launches, CUDA-graph replay.

PyTorch is used for device memory (caching allocator), streams and graphs only; all arithmetic on
the hot path happens inside the C-ABI library (`include/pmce_b200.h`). Nothing here falls back to
PyTorch ops or to the CPU: non-CUDA inputs raise.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import PmceDims, PmceSlot, PmceError, check

N_VERT, N_VERT_DS, F_IMG, H_GRU, D_COEVO = 6890, 431, 2048, 1024, 64


def make_dims(num_joint=17, embed_dim=256, depth=3, seqlen=16):
    return PmceDims(num_joint=num_joint, embed_dim=embed_dim, depth=depth, seqlen=seqlen, num_vert_ds=N_VERT_DS,
                    num_vert=N_VERT, feat_dim=F_IMG, gru_hidden=H_GRU, coevo_dim=D_COEVO, lifter_heads=8)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda_f32(t, name, shape=None):
    if not isinstance(t, torch.Tensor):
        raise PmceError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise PmceError(f"{name}: must be a CUDA tensor — pmce_b200 has no CPU path (got device {t.device})")
    if t.dtype != torch.float32:
        raise PmceError(f"{name}: must be float32 (got {t.dtype})")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise PmceError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.contiguous()


class Engine:
    """One model instance's packed weights + scratch on one device."""

    def __init__(self, dims, use_graph=None):
        self.lib = _lib.load()
        self.dims = dims
        self._dp = C.byref(self.dims)
        nbytes = self.lib.pmce_weights_bytes(self._dp)
        if nbytes == 0:
            check(1, "pmce_weights_bytes")
        self.weight_bytes = nbytes
        self.weights = None
        self.vj = None
        self._ws = None
        self._graphs = {}
        self._pipe = {}
        if use_graph is None:
            use_graph = os.environ.get("PMCE_B200_GRAPH", "1") != "0"
        self.use_graph = use_graph
        self.launch_count = 0

    # ---- weights -------------------------------------------------------------------------------------
    def pack(self, named_tensors, device, vj_relation=None):
        """Copy reference-schema tensors (`state_dict` names) into the packed device blob."""
        device = torch.device(device)
        if device.type != "cuda":
            raise PmceError("weights must be packed onto a CUDA device")
        blob = torch.zeros(self.weight_bytes // 4, dtype=torch.float32, device=device)
        slot = PmceSlot()
        for name, t in named_tensors:
            rc = self.lib.pmce_weight_slot(self._dp, name.encode(), C.byref(slot))
            if rc == 1:
                continue  # never reaches an output (joint branch of coevoblock1/2)
            if rc != 0:
                check(rc, f"pmce_weight_slot({name})")
            src = t.detach().to(device=device, dtype=torch.float32).reshape(slot.rows, slot.cols)
            dst = blob[slot.offset: slot.offset + slot.rows * slot.ld].view(slot.rows, slot.ld)[:, :slot.cols]
            dst.copy_(src)
        with torch.cuda.device(device):
            check(self.lib.pmce_pack_weights(self._dp, _ptr(blob), _stream()), "pmce_pack_weights")
        self.weights = blob
        if vj_relation is not None:
            self.vj = torch.as_tensor(np.asarray(vj_relation).astype(np.int32), device=device)
        # captured graphs (single-shot and pipelined) have the old blob / vj pointers baked in
        self._graphs.clear()
        self._pipe.clear()
        return self

    def _workspace(self, B, device):
        need = self.lib.pmce_workspace_bytes(self._dp, B)
        if need == 0:
            check(1, "pmce_workspace_bytes")
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._graphs.clear()
            self._pipe.clear()
        return self._ws

    def _workspace2(self, B, device):
        """The second workspace: the pipelined iterators keep TWO forwards in flight (one per slot, each on its own stream)."""
        need = self.lib.pmce_workspace_bytes(self._dp, B)
        if need == 0:
            check(1, "pmce_workspace_bytes")
        ws2 = getattr(self, "_ws2", None)
        if ws2 is None or ws2.numel() < need or ws2.device != device:
            self._ws2 = torch.empty(need, dtype=torch.uint8, device=device)
            self._pipe.clear()
        return self._ws2

    def _ready(self, need_vj=False):
        if self.weights is None:
            raise PmceError("weights have not been packed (call Engine.pack / load_state_dict first)")
        if need_vj and self.vj is None:
            raise PmceError("vj_relation has not been set")

    # ---- whole forward (a1) --------------------------------------------------------------------------
    def _forward_eager(self, pose2d, img_feat, mesh, cam_pose, pose3d, ws=None):
        B = pose2d.shape[0]
        if ws is None:
            ws = self._workspace(B, pose2d.device)
        check(self.lib.pmce_forward(self._dp, _ptr(self.weights), _ptr(pose2d), _ptr(img_feat), _ptr(self.vj), B,
                                    _ptr(mesh), _ptr(cam_pose), _ptr(pose3d), _ptr(ws), ws.numel(), _stream()),
              "pmce_forward")

    def forward(self, pose2d, img_feat, out=None):
        """PMCE.forward: ([B,T,J,2], [B,T,2048]) -> (cam_mesh [B,6890,3], cam_pose [B,J,3], pose3d [B,J,3]).

        `out` = (cam_mesh, cam_pose, pose3d) caller-owned contiguous CUDA tensors: the kernels write straight into them (in
        graph mode a graph is captured per (B, output buffers) and replayed without the final clones) - e.g. a rank's slot of an
        all-gather buffer (`pmce_b200.dist.ShardedForward`)."""
        self._ready(need_vj=True)
        d = self.dims
        B = pose2d.shape[0] if isinstance(pose2d, torch.Tensor) and pose2d.dim() == 4 else -1
        pose2d = _require_cuda_f32(pose2d, "pose2d", (B, d.seqlen, d.num_joint, 2))
        img_feat = _require_cuda_f32(img_feat, "img_feat", (B, d.seqlen, d.feat_dim))
        dev = pose2d.device
        if dev != self.weights.device:
            raise PmceError(f"inputs on {dev} but weights on {self.weights.device}")
        if out is not None:
            shapes = ((B, d.num_vert, 3), (B, d.num_joint, 3), (B, d.num_joint, 3))
            for t, n, sh in zip(out, ("cam_mesh", "cam_pose", "pose3d"), shapes):
                if _require_cuda_f32(t, "out." + n, sh) is not t or t.device != dev:
                    raise PmceError(f"out.{n}: must be a contiguous float32 tensor of shape {sh} on {dev}")
        with torch.cuda.device(dev):
            if not self.use_graph or torch.cuda.is_current_stream_capturing():
                mesh, cam_pose, pose3d = out if out is not None else (
                    torch.empty(B, d.num_vert, 3, device=dev), torch.empty(B, d.num_joint, 3, device=dev), torch.empty(B, d.num_joint, 3, device=dev))
                self._forward_eager(pose2d, img_feat, mesh, cam_pose, pose3d)
                return mesh, cam_pose, pose3d
            key = B if out is None else (B,) + tuple(t.data_ptr() for t in out)
            g = self._graphs.get(key)
            if g is None:
                g = self._capture(B, dev, key, out)
            g["p2d"].copy_(pose2d)
            g["feat"].copy_(img_feat)
            g["graph"].replay()
            if out is not None:
                return out
            return g["mesh"].clone(), g["cam_pose"].clone(), g["pose3d"].clone()

    def _capture(self, B, dev, key=None, out=None):
        d = self.dims
        key = B if key is None else key
        mesh, cam_pose, pose3d = out if out is not None else (
            torch.empty(B, d.num_vert, 3, device=dev), torch.empty(B, d.num_joint, 3, device=dev), torch.empty(B, d.num_joint, 3, device=dev))
        st = dict(p2d=torch.zeros(B, d.seqlen, d.num_joint, 2, device=dev),
                  feat=torch.zeros(B, d.seqlen, d.feat_dim, device=dev),
                  mesh=mesh, cam_pose=cam_pose, pose3d=pose3d)
        self._workspace(B, dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):      # warm-up outside capture (func attributes, lazy module load)
            self._forward_eager(st["p2d"], st["feat"], st["mesh"], st["cam_pose"], st["pose3d"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._forward_eager(st["p2d"], st["feat"], st["mesh"], st["cam_pose"], st["pose3d"])
        st["graph"] = graph
        self._graphs[key] = st
        return st

    def _require_host(self, pose2d_cpu, img_feat_cpu, what):
        """Host-buffer inputs: contiguous float32 CPU tensors of exactly (B,T,J,2) / (B,T,feat_dim) — `Tensor.copy_` into the
        static graph buffers would silently broadcast a wrong-but-broadcastable shape."""
        d = self.dims
        B = pose2d_cpu.shape[0] if isinstance(pose2d_cpu, torch.Tensor) and pose2d_cpu.dim() == 4 else -1
        for t, n, shape in ((pose2d_cpu, "pose2d", (B, d.seqlen, d.num_joint, 2)), (img_feat_cpu, "img_feat", (B, d.seqlen, d.feat_dim))):
            if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise PmceError(f"{n}: {what} expects contiguous float32 CPU tensors")
            if B < 1 or tuple(t.shape) != shape:
                raise PmceError(f"{n}: {what} expected shape {shape}, got {tuple(t.shape)}")

    def forward_host(self, pose2d_cpu, img_feat_cpu, out=None):
        """The C-ABI host-buffer call (`pmce_forward_host`): pinned/pageable CPU tensors in, CPU tensors out."""
        self._ready(need_vj=True)
        d = self.dims
        B = pose2d_cpu.shape[0]
        dev = self.weights.device
        self._require_host(pose2d_cpu, img_feat_cpu, "forward_host")
        if out is None:
            out = (torch.empty(B, d.num_vert, 3).pin_memory(), torch.empty(B, d.num_joint, 3).pin_memory(),
                   torch.empty(B, d.num_joint, 3).pin_memory())
        with torch.cuda.device(dev):
            if self.use_graph and not torch.cuda.is_current_stream_capturing():
                # host buffers straight into / out of the captured graph's static device buffers: one H2D per input,
                # one replay, one D2H per output, one synchronise
                g = self._graphs.get(B) or self._capture(B, dev)
                g["p2d"].copy_(pose2d_cpu, non_blocking=True)
                g["feat"].copy_(img_feat_cpu, non_blocking=True)
                g["graph"].replay()
                out[0].copy_(g["mesh"], non_blocking=True)
                out[1].copy_(g["cam_pose"], non_blocking=True)
                out[2].copy_(g["pose3d"], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return out
            ws = self._workspace(B, dev)
            io_bytes = self.lib.pmce_io_bytes(self._dp, B)
            if getattr(self, "_io", None) is None or self._io.numel() < io_bytes:
                self._io = torch.empty(io_bytes, dtype=torch.uint8, device=dev)
            check(self.lib.pmce_forward_host(self._dp, _ptr(self.weights), _ptr(pose2d_cpu), _ptr(img_feat_cpu),
                                             _ptr(self.vj), B, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(self._io),
                                             _ptr(ws), ws.numel(), _stream()), "pmce_forward_host")
        return out

    def forward_sliding(self, pose2d_seq, img_feat_seq, stride=1):
        """Overlapping windows of one track (reference lib/_img_utils.py:58-92; the demo's per-window loop, main/run_demo.py:145):
        pose2d_seq [N,J,2], img_feat_seq [N,2048] -> the forward's three outputs for the nwin = (N-T)//stride + 1 windows
        [w*stride, w*stride+T). Per-frame work (imgfeat_embed, token embedding, SpatialBlocks[0], GRU layer-0 input
        projection) runs once per frame instead of once per window (`pmce_forward_sliding`)."""
        self._ready(need_vj=True)
        d = self.dims
        N = pose2d_seq.shape[0] if isinstance(pose2d_seq, torch.Tensor) and pose2d_seq.dim() == 3 else -1
        pose2d_seq = _require_cuda_f32(pose2d_seq, "pose2d_seq", (N, d.num_joint, 2))
        img_feat_seq = _require_cuda_f32(img_feat_seq, "img_feat_seq", (N, d.feat_dim))
        nwin = self.lib.pmce_sliding_windows(self._dp, N, int(stride))
        if nwin < 1 or stride > d.seqlen:
            raise PmceError(f"forward_sliding: need at least seqlen={d.seqlen} frames and 1 <= stride <= seqlen (got {N} frames, stride {stride})")
        dev = pose2d_seq.device
        with torch.cuda.device(dev):
            ws = self._workspace(nwin, dev)
            mesh = torch.empty(nwin, d.num_vert, 3, device=dev)
            cam_pose = torch.empty(nwin, d.num_joint, 3, device=dev)
            pose3d = torch.empty(nwin, d.num_joint, 3, device=dev)
            check(self.lib.pmce_forward_sliding(self._dp, _ptr(self.weights), _ptr(pose2d_seq), _ptr(img_feat_seq), _ptr(self.vj), N,
                                                int(stride), _ptr(mesh), _ptr(cam_pose), _ptr(pose3d), _ptr(ws), ws.numel(), _stream()),
                  "pmce_forward_sliding")
        return mesh, cam_pose, pose3d

    # ---- pipelined host loop ---------------------------------------------------------------------------
    def _pipeline(self, B, dev, out_slots=None):
        """Two slots of static device buffers + captured graphs + pinned host outputs; streams: H2D, one forward stream PER SLOT,
        D2H. Each slot has its own workspace, so the forwards of two consecutive batches are IN FLIGHT TOGETHER: the decoder
        third of a forward is a chain of latency-bound kernels, and the other batch's lifter GEMMs fill what it leaves idle
        (B=64: 2,486 vs 2,748 us per forward, tools/two_in_flight.py). `out_slots`: two caller-owned (cam_mesh, cam_pose, pose3d)
        triples the two graphs write into (e.g. the rank's rows of two all-gather buffers)."""
        wss = (self._workspace(B, dev), self._workspace2(B, dev))
        ws = wss[0]
        baked = (ws.data_ptr(), wss[1].data_ptr(), self.weights.data_ptr(), self.vj.data_ptr()) + (        # pointers the captured graphs hold
            tuple(t.data_ptr() for sl in out_slots for t in sl) if out_slots is not None else ())
        pipe = self._pipe.get(B)
        if pipe is not None and pipe["baked"] == baked:
            return pipe
        d = self.dims
        slots = []
        for k in range(2):
            o = out_slots[k] if out_slots is not None else (torch.empty(B, d.num_vert, 3, device=dev), torch.empty(B, d.num_joint, 3, device=dev),
                                                            torch.empty(B, d.num_joint, 3, device=dev))
            sl = dict(p2d=torch.zeros(B, d.seqlen, d.num_joint, 2, device=dev), feat=torch.zeros(B, d.seqlen, d.feat_dim, device=dev),
                      mesh=o[0], cam_pose=o[1], pose3d=o[2],
                      # staging copies for the host loop: the H2D of the slot's NEXT batch and the D2H of its PREVIOUS result must
                      # not wait for / hold up the forward that owns the graph's static buffers (a 14 MB device copy per step buys
                      # ~170 us of H2D + ~100 us of D2H off the slot's critical path)
                      p2d_in=torch.zeros(B, d.seqlen, d.num_joint, 2, device=dev), feat_in=torch.zeros(B, d.seqlen, d.feat_dim, device=dev),
                      mesh_out=torch.empty(B, d.num_vert, 3, device=dev), cam_pose_out=torch.empty(B, d.num_joint, 3, device=dev),
                      pose3d_out=torch.empty(B, d.num_joint, 3, device=dev),
                      h2d=torch.cuda.Event(), in_free=torch.cuda.Event(), fwd=torch.cuda.Event(), d2h=torch.cuda.Event(), used=False)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up outside capture
                self._forward_eager(sl["p2d"], sl["feat"], sl["mesh"], sl["cam_pose"], sl["pose3d"], ws=wss[k])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._forward_eager(sl["p2d"], sl["feat"], sl["mesh"], sl["cam_pose"], sl["pose3d"], ws=wss[k])
            sl["graph"] = graph
            slots.append(sl)
        # THREE pinned host output sets for two device slots: result i (host set i % 3) is not written again before batch i+3
        # is submitted, i.e. before the consumer asks for result i+2
        hosts = [(torch.empty(B, d.num_vert, 3).pin_memory(), torch.empty(B, d.num_joint, 3).pin_memory(),
                  torch.empty(B, d.num_joint, 3).pin_memory()) for _ in range(3)]
        pipe = dict(ws=ws, baked=baked, slots=slots, hosts=hosts, s_in=torch.cuda.Stream(device=dev),
                    s_fwd=[torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)], s_out=torch.cuda.Stream(device=dev))
        self._pipe[B] = pipe
        return pipe

    def forward_host_iter(self, batches, out_slots=None, before_forward=None, after_forward=None):
        """Pipelined form of the reference's test loop (lib/core/base.py:218-238: `.cuda()` -> forward -> `.cpu()` per batch).

        `batches` yields (pose2d [B,T,J,2], img_feat [B,T,2048]) contiguous float32 CPU tensors (pinned for asynchronous
        copies) of one batch size; the generator yields (cam_mesh, cam_pose, pose3d) pinned CPU tensors in the same order.
        The H2D copy of batch i+1 and the D2H copy of batch i-1 run on the copy engines while batch i is in the forward, and the
        forwards of batches i and i+1 are in flight together on the two slots' own streams and workspaces (two device buffer
        slots with one captured graph each, three pinned host output sets). A yielded
        triple stays valid while the NEXT result is requested and consumed; it is overwritten once the result after next is
        requested, so consume (or copy) it before asking for the batch after next. `list(forward_host_iter(...))` therefore
        aliases buffers: copy each triple as it arrives.

        Multi-GPU hooks (pmce_b200.dist.ShardedForward): `out_slots` = the two device output triples the forwards write into;
        `before_forward(k)` / `after_forward(k)` run on slot k's forward stream around the replay of slot k (wait until the slot's
        previous all-gather has finished / issue this step's all-gather on the communication stream)."""
        self._ready(need_vj=True)
        dev = self.weights.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream()
            pipe, pending, i = None, [], 0
            try:
                for hp, hf in batches:
                    self._require_host(hp, hf, "forward_host_iter")
                    if pipe is None:
                        B = hp.shape[0]
                        pipe = self._pipeline(B, dev, out_slots)
                        for s in (pipe["s_in"], *pipe["s_fwd"], pipe["s_out"]):
                            s.wait_stream(cur)
                    elif hp.shape[0] != B:
                        raise PmceError("forward_host_iter: every batch must have the same size (pad or run the tail through forward_host)")
                    sl = pipe["slots"][i & 1]
                    host = pipe["hosts"][i % 3]
                    with torch.cuda.stream(pipe["s_in"]):
                        if sl["used"]:
                            pipe["s_in"].wait_event(sl["in_free"])   # the slot's previous batch has left the staging inputs
                        sl["p2d_in"].copy_(hp, non_blocking=True)
                        sl["feat_in"].copy_(hf, non_blocking=True)
                        sl["h2d"].record(pipe["s_in"])
                    s_fwd = pipe["s_fwd"][i & 1]
                    with torch.cuda.stream(s_fwd):
                        s_fwd.wait_event(sl["h2d"])
                        sl["p2d"].copy_(sl["p2d_in"], non_blocking=True)     # staging -> the graph's static inputs (device copy)
                        sl["feat"].copy_(sl["feat_in"], non_blocking=True)
                        sl["in_free"].record(s_fwd)
                        if before_forward is not None:
                            before_forward(i & 1)
                        sl["graph"].replay()
                        if after_forward is not None:
                            after_forward(i & 1)
                        if sl["used"]:
                            s_fwd.wait_event(sl["d2h"])              # the slot's previous result has left the staging outputs
                        sl["mesh_out"].copy_(sl["mesh"], non_blocking=True)
                        sl["cam_pose_out"].copy_(sl["cam_pose"], non_blocking=True)
                        sl["pose3d_out"].copy_(sl["pose3d"], non_blocking=True)
                        sl["fwd"].record(s_fwd)
                    with torch.cuda.stream(pipe["s_out"]):
                        pipe["s_out"].wait_event(sl["fwd"])
                        host[0].copy_(sl["mesh_out"], non_blocking=True)
                        host[1].copy_(sl["cam_pose_out"], non_blocking=True)
                        host[2].copy_(sl["pose3d_out"], non_blocking=True)
                        sl["d2h"].record(pipe["s_out"])
                    sl["used"] = True
                    pending.append((sl["d2h"], host))     # the slot's event is re-recorded only after this entry was popped
                    i += 1
                    if len(pending) == 2:
                        ev, res = pending.pop(0)
                        ev.synchronize()
                        yield res
                while pending:
                    ev, res = pending.pop(0)
                    ev.synchronize()
                    yield res
            finally:
                # also on early exit (the consumer stopped iterating, or a batch was rejected): drain and rejoin the caller's stream
                if pipe is not None:
                    for s in (pipe["s_in"], *pipe["s_fwd"], pipe["s_out"]):
                        cur.wait_stream(s)
                    for sl in pipe["slots"]:
                        sl["used"] = False

    def forward_iter(self, batches, out_slots=None, before_forward=None, after_forward=None):
        """Device-resident form of the pipelined loop: `batches` yields (pose2d [B,T,J,2], img_feat [B,T,2048]) CUDA tensors of one
        batch size, the generator yields (cam_mesh, cam_pose, pose3d) CUDA tensors in the same order, with TWO forwards in flight
        (the slots of `_pipeline`: own static buffers, workspace, captured graph and stream each). Results are bit-identical to
        `forward()`. A yielded triple is the slot's output buffers: the caller's current stream has been made to wait for them,
        and they stay valid until the NEXT result is requested - the batch submitted then reuses the slot, after everything the
        consumer queued on its stream (copy what must live longer). Inputs are read on the slot's stream after everything
        already queued on the caller's stream. `out_slots` / hooks: as `forward_host_iter`."""
        self._ready(need_vj=True)
        d = self.dims
        dev = self.weights.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream()
            pipe, pending, i = None, [], 0
            try:
                for p2d, feat in batches:
                    B = p2d.shape[0] if isinstance(p2d, torch.Tensor) and p2d.dim() == 4 else -1
                    p2d = _require_cuda_f32(p2d, "pose2d", (B, d.seqlen, d.num_joint, 2))
                    feat = _require_cuda_f32(feat, "img_feat", (B, d.seqlen, d.feat_dim))
                    if p2d.device != dev or feat.device != dev:
                        raise PmceError(f"forward_iter: inputs on {p2d.device} / {feat.device} but weights on {dev}")
                    if pipe is None:
                        B0 = B
                        pipe = self._pipeline(B, dev, out_slots)
                    elif B != B0:
                        raise PmceError("forward_iter: every batch must have the same size (run the tail through forward)")
                    sl = pipe["slots"][i & 1]
                    s_fwd = pipe["s_fwd"][i & 1]
                    # the inputs were produced, and the slot's previous outputs consumed, by work queued on the caller's stream
                    s_fwd.wait_stream(cur)
                    with torch.cuda.stream(s_fwd):
                        sl["p2d"].copy_(p2d, non_blocking=True)
                        sl["feat"].copy_(feat, non_blocking=True)
                        if before_forward is not None:
                            before_forward(i & 1)
                        sl["graph"].replay()
                        sl["fwd"].record(s_fwd)
                        if after_forward is not None:
                            after_forward(i & 1)
                    pending.append(sl)
                    i += 1
                    if len(pending) == 2:
                        done = pending.pop(0)
                        cur.wait_event(done["fwd"])
                        yield done["mesh"], done["cam_pose"], done["pose3d"]
                while pending:
                    done = pending.pop(0)
                    cur.wait_event(done["fwd"])
                    yield done["mesh"], done["cam_pose"], done["pose3d"]
            finally:
                if pipe is not None:        # also on early exit: nothing of the pipeline outlives the call un-joined
                    for s in pipe["s_fwd"]:
                        cur.wait_stream(s)

    # ---- sub-paths (each is a C-ABI entry point; used by the module API and by the parity tests) --------
    def lifter(self, pose2d, img_feat):
        self._ready()
        d = self.dims
        B = pose2d.shape[0]
        pose2d = _require_cuda_f32(pose2d, "pose2d", (B, d.seqlen, d.num_joint, 2))
        img_feat = _require_cuda_f32(img_feat, "img_feat", (B, d.seqlen, d.feat_dim))
        with torch.cuda.device(pose2d.device):
            ws = self._workspace(B, pose2d.device)
            out = torch.empty(B, d.num_joint, 3, device=pose2d.device)
            check(self.lib.pmce_lifter_forward(self._dp, _ptr(self.weights), _ptr(pose2d), _ptr(img_feat), B, _ptr(out),
                                               _ptr(ws), ws.numel(), _stream()), "pmce_lifter_forward")
        return out

    def gru_mid(self, img_feat):
        self._ready()
        d = self.dims
        B = img_feat.shape[0]
        img_feat = _require_cuda_f32(img_feat, "img_feat", (B, d.seqlen, d.feat_dim))
        with torch.cuda.device(img_feat.device):
            ws = self._workspace(B, img_feat.device)
            g = torch.empty(B, d.feat_dim, device=img_feat.device)
            check(self.lib.pmce_gru_mid(self._dp, _ptr(self.weights), _ptr(img_feat), B, _ptr(g), _ptr(ws), ws.numel(),
                                        _stream()), "pmce_gru_mid")
        return g

    def adaln_gammabeta(self, g):
        self._ready()
        B = g.shape[0]
        g = _require_cuda_f32(g, "g", (B, self.dims.feat_dim))
        with torch.cuda.device(g.device):
            ws = self._workspace(B, g.device)
            gb = torch.empty(B, self.lib.pmce_adaln_slots(), 2, self.dims.coevo_dim, device=g.device)
            check(self.lib.pmce_adaln_gammabeta(self._dp, _ptr(self.weights), _ptr(g), B, _ptr(gb), _ptr(ws), ws.numel(),
                                                _stream()), "pmce_adaln_gammabeta")
        return gb

    def coevo_block(self, block, joints, verts, gb, want_joints=False):
        self._ready()
        d = self.dims
        B = joints.shape[0]
        joints = _require_cuda_f32(joints, "joints", (B, d.num_joint, 3))
        verts = _require_cuda_f32(verts, "verts", (B, d.num_vert_ds, 3))
        gb = _require_cuda_f32(gb, "gb")
        with torch.cuda.device(joints.device):
            ws = self._workspace(B, joints.device)
            jout = torch.empty_like(joints) if want_joints else None
            vout = torch.empty_like(verts)
            check(self.lib.pmce_coevo_block(self._dp, _ptr(self.weights), block, _ptr(joints), _ptr(verts), _ptr(gb), B,
                                            _ptr(jout), _ptr(vout), _ptr(ws), ws.numel(), _stream()), "pmce_coevo_block")
        return jout, vout

    def cross_attn_block(self, block, which, xq, xk, xv, gb):
        """CrossAttentionBlock.forward of coevoblock<block> (which: 0 = joint_CA_FFN, 1 = vertx_CA_FFN) -> updated xq."""
        self._ready()
        d = self.dims
        B = xq.shape[0]
        n1, n2 = (d.num_vert_ds, d.num_joint) if which else (d.num_joint, d.num_vert_ds)
        xq = _require_cuda_f32(xq, "xq", (B, n1, d.coevo_dim))
        xk = _require_cuda_f32(xk, "xk", (B, n2, d.coevo_dim))
        xv = _require_cuda_f32(xv, "xv", (B, n2, d.coevo_dim))
        gb = _require_cuda_f32(gb, "gb")
        with torch.cuda.device(xq.device):
            ws = self._workspace(B, xq.device)
            out = torch.empty_like(xq)
            check(self.lib.pmce_cross_attn_block(self._dp, _ptr(self.weights), block, which, _ptr(xq), _ptr(xk), _ptr(xv), _ptr(gb), B,
                                                 _ptr(out), _ptr(ws), ws.numel(), _stream()), "pmce_cross_attn_block")
        return out

    def self_attn_block(self, block, which, x, gb):
        """Block.forward of coevoblock<block> (which: 0 = joint_SA_FFN, 1 = vertx_SA_FFN) -> updated x."""
        self._ready()
        d = self.dims
        B = x.shape[0]
        x = _require_cuda_f32(x, "x", (B, d.num_vert_ds if which else d.num_joint, d.coevo_dim))
        gb = _require_cuda_f32(gb, "gb")
        with torch.cuda.device(x.device):
            ws = self._workspace(B, x.device)
            out = torch.empty_like(x)
            check(self.lib.pmce_self_attn_block(self._dp, _ptr(self.weights), block, which, _ptr(x), _ptr(gb), B, _ptr(out),
                                                _ptr(ws), ws.numel(), _stream()), "pmce_self_attn_block")
        return out

    def mesh_epilogue(self, verts3, g):
        self._ready()
        d = self.dims
        B = verts3.shape[0]
        verts3 = _require_cuda_f32(verts3, "verts3", (B, d.num_vert_ds, 3))
        g = _require_cuda_f32(g, "g", (B, d.feat_dim))
        with torch.cuda.device(g.device):
            ws = self._workspace(B, g.device)
            mesh = torch.empty(B, d.num_vert, 3, device=g.device)
            check(self.lib.pmce_mesh_epilogue(self._dp, _ptr(self.weights), _ptr(verts3), _ptr(g), B, _ptr(mesh), _ptr(ws),
                                              ws.numel(), _stream()), "pmce_mesh_epilogue")
        return mesh

    def decoder(self, joints, img_feat, want_verts0=False):
        """Pose2Mesh.forward: (joints [B,J,3] metres, img_feat [B,T,2048]) -> (cam_pose, cam_mesh[, verts0])."""
        self._ready(need_vj=True)
        d = self.dims
        B = joints.shape[0]
        joints = _require_cuda_f32(joints, "joints", (B, d.num_joint, 3))
        img_feat = _require_cuda_f32(img_feat, "img_feat", (B, d.seqlen, d.feat_dim))
        with torch.cuda.device(joints.device):
            ws = self._workspace(B, joints.device)
            cam_pose = torch.empty(B, d.num_joint, 3, device=joints.device)
            mesh = torch.empty(B, d.num_vert, 3, device=joints.device)
            v0 = torch.empty(B, d.num_vert_ds, 3, device=joints.device) if want_verts0 else None
            check(self.lib.pmce_decoder_forward(self._dp, _ptr(self.weights), _ptr(joints), _ptr(img_feat), _ptr(self.vj), B,
                                                _ptr(cam_pose), _ptr(mesh), _ptr(v0), _ptr(ws), ws.numel(), _stream()),
                  "pmce_decoder_forward")
        return (cam_pose, mesh, v0) if want_verts0 else (cam_pose, mesh)


class JRegressor:
    """Sparse joint regressor (`torch.matmul(J_regressor[None], pred_mesh)`, reference lib/core/base.py:225).

    The shipped regressors have 105-107 non-zeros out of 17x6890; stored as CSR on the device.
    """

    def __init__(self, J_dense, device):
        self.lib = _lib.load()
        J = np.asarray(J_dense, dtype=np.float32)
        self.R, self.V = J.shape
        rows, cols = np.nonzero(J)
        order = np.lexsort((cols, rows))
        rows, cols = rows[order], cols[order]
        row_ptr = np.zeros(self.R + 1, dtype=np.int32)
        np.add.at(row_ptr, rows + 1, 1)
        self.row_ptr = torch.as_tensor(np.cumsum(row_ptr).astype(np.int32), device=device)
        self.cols = torch.as_tensor(cols.astype(np.int32), device=device)
        self.vals = torch.as_tensor(J[rows, cols], device=device)

    def __call__(self, mesh, scale=1.0):
        B = mesh.shape[0]
        mesh = _require_cuda_f32(mesh, "mesh", (B, self.V, 3))
        with torch.cuda.device(mesh.device):
            out = torch.empty(B, self.R, 3, device=mesh.device)
            check(self.lib.pmce_jregress(_ptr(self.row_ptr), _ptr(self.cols), _ptr(self.vals), self.R, _ptr(mesh), self.V, B,
                                         float(scale), _ptr(out), _stream()), "pmce_jregress")
        return out

    def eval_errors(self, cam_mesh, gt_mesh, gt_pose, eval_joints=(1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15, 16), scale=1000.0):
        """The test loop's evaluation epilogue (reference lib/core/base.py:223-227 + data/PW3D/dataset.py:269-282) on the
        device: cam_mesh / gt_mesh [B,V,3] in metres, gt_pose [B,R,3] in mm ->
        (pred_pose [B,R,3] mm, clip_err [B,2] = per-clip (joint, mesh) mean error, mean_err [2] = (j_error, s_error))."""
        B = cam_mesh.shape[0]
        cam_mesh = _require_cuda_f32(cam_mesh, "cam_mesh", (B, self.V, 3))
        gt_mesh = _require_cuda_f32(gt_mesh, "gt_mesh", (B, self.V, 3))
        gt_pose = _require_cuda_f32(gt_pose, "gt_pose", (B, self.R, 3))
        dev = cam_mesh.device
        key = tuple(int(j) for j in eval_joints)
        if getattr(self, "_eval_key", None) != (key, dev):
            if not key or min(key) < 0 or max(key) >= self.R:
                raise PmceError(f"eval_joints must index the {self.R} regressor rows")
            self._eval_idx = torch.tensor(key, dtype=torch.int32, device=dev)
            self._eval_key = (key, dev)
        with torch.cuda.device(dev):
            pred_pose = torch.empty(B, self.R, 3, device=dev)
            clip_err = torch.empty(B, 2, device=dev)
            mean_err = torch.empty(2, device=dev)
            check(self.lib.pmce_eval_errors(_ptr(self.row_ptr), _ptr(self.cols), _ptr(self.vals), self.R, _ptr(cam_mesh), _ptr(gt_mesh),
                                            _ptr(gt_pose), _ptr(self._eval_idx), len(key), self.V, B, float(scale), _ptr(pred_pose),
                                            _ptr(clip_err), _ptr(mean_err), _stream()), "pmce_eval_errors")
        return pred_pose, clip_err, mean_err
