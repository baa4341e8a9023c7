"""Batch-axis sharding of the per-clip forward across the GPUs of one box.

Clips are independent (no cross-clip op anywhere in PMCE.forward), so rank r owns clips
[r*ceil(B/G), (r+1)*ceil(B/G)) with replicated weights and there is exactly ONE collective: an
all-gather of the per-rank packed output block `cam_mesh | cam_pose | pose3d` (83,088 B per clip for J=17).
The reference has no multi-GPU support at all; this is the harness around the drop-in module.
One process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch; gloo for the CPU logic tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous shard [lo, hi) of `total` clips for `rank`; shards are ceil(total/world) long, the tail may be short/empty."""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    hi = min(lo + per, total)
    return lo, hi, per


def pack_outputs(mesh, cam_pose, pose3d, per):
    """-> [per, 6890*3 + 2*J*3] block, zero padded to `per` clips so every rank contributes the same size."""
    n = mesh.shape[0]
    width = mesh[0].numel() + cam_pose[0].numel() + pose3d[0].numel() if n else None
    if n == 0:
        raise ValueError("pack_outputs needs at least one clip (use empty_block for empty shards)")
    block = mesh.new_zeros(per, width)
    a, b = mesh[0].numel(), mesh[0].numel() + cam_pose[0].numel()
    block[:n, :a] = mesh.reshape(n, -1)
    block[:n, a:b] = cam_pose.reshape(n, -1)
    block[:n, b:] = pose3d.reshape(n, -1)
    return block


def unpack_outputs(gathered, total, num_vert, num_joint):
    """[world*per, width] -> (cam_mesh [B,6890,3], cam_pose [B,J,3], pose3d [B,J,3]) with padding dropped."""
    a, b = num_vert * 3, num_vert * 3 + num_joint * 3
    g = gathered[:total]
    return (g[:, :a].reshape(total, num_vert, 3), g[:, a:b].reshape(total, num_joint, 3),
            g[:, b:].reshape(total, num_joint, 3))


def all_gather_blocks(block, group=None):
    world = dist.get_world_size(group)
    out = block.new_empty(world * block.shape[0], block.shape[1])
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, block.contiguous(), group=group)
    else:
        chunks = list(out.chunk(world, dim=0))
        dist.all_gather(chunks, block.contiguous(), group=group)
    return out


def sharded_forward(forward_fn, pose2d, img_feat, num_vert=6890, group=None, gather=True):
    """Run `forward_fn` (e.g. a `models.PMCE.PMCE` instance) on this rank's clip shard and all-gather.

    `pose2d` / `img_feat` hold the GLOBAL batch on every rank (or at least this rank's rows); returns the global
    outputs on every rank when `gather`, else this rank's shard outputs.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, J = pose2d.shape[0], pose2d.shape[2]
    lo, hi, per = shard_bounds(B, rank, world)
    width = num_vert * 3 + 2 * J * 3
    if hi > lo:
        mesh, cam_pose, pose3d = forward_fn(pose2d[lo:hi].contiguous(), img_feat[lo:hi].contiguous())
        if not gather:
            return mesh, cam_pose, pose3d
        block = pack_outputs(mesh, cam_pose, pose3d, per)
    else:
        if not gather:
            e = pose2d.new_zeros(0)
            return e.reshape(0, num_vert, 3), e.reshape(0, J, 3), e.reshape(0, J, 3)
        block = pose2d.new_zeros(per, width)
    return unpack_outputs(all_gather_blocks(block, group), B, num_vert, J)


class ShardedForward:
    """The multi-GPU step of the path, without staging copies and with the collective off the critical path.

    Every rank owns `slots` gather buffers of shape [world, per * width]. Row `rank` of a buffer is three contiguous sections
    `cam_mesh [per,6890,3] | cam_pose [per,J,3] | pose3d [per,J,3]`, and the forward writes its outputs STRAIGHT into them
    (`forward_fn(pose2d, img_feat, out=...)`, no pack kernel). The single collective, an in-place all-gather of the rank rows,
    is issued on a communication stream: step i+1's forward (other slot) runs under step i's gather. `result(i)` waits for the
    gather of step i and returns per-rank views (rank-major; `unpack` concatenates them when one global tensor is wanted).
    """

    def __init__(self, forward_fn, per, num_joint, device, num_vert=6890, group=None, slots=2):
        self.fn, self.per, self.J, self.V, self.group = forward_fn, per, num_joint, num_vert, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.sizes = (per * num_vert * 3, per * num_joint * 3, per * num_joint * 3)
        self.width = sum(self.sizes)
        self.bufs = [torch.zeros(self.world, self.width, device=device) for _ in range(slots)]
        self.cuda = torch.device(device).type == "cuda"
        if self.cuda:
            self.comm = torch.cuda.Stream(device=device)
            self.ev_fwd = [torch.cuda.Event() for _ in range(slots)]
            self.ev_comm = [torch.cuda.Event() for _ in range(slots)]
        self.steps = 0

    def views(self, buf, r):
        a, b, _ = self.sizes
        row = buf[r]
        return (row[:a].view(self.per, self.V, 3), row[a:a + b].view(self.per, self.J, 3), row[a + b:].view(self.per, self.J, 3))

    def step(self, pose2d, img_feat):
        """One forward on this rank's `per` clips + the (asynchronous) all-gather of the step. Returns the slot index."""
        k = self.steps % len(self.bufs)
        self.steps += 1
        buf = self.bufs[k]
        out = self.views(buf, self.rank)
        if self.cuda:
            self.wait_slot_free(k)
            self.fn(pose2d, img_feat, out=out)
            self.gather_async(k)
        else:                                              # gloo / CPU: same data flow, synchronous
            self.fn(pose2d, img_feat, out=out)
            chunks = [torch.empty_like(buf[0]) for _ in range(self.world)]
            dist.all_gather(chunks, buf[self.rank].clone(), group=self.group)
            for r, c in enumerate(chunks):
                buf[r].copy_(c)
        return k

    def wait_slot_free(self, k):
        """Make the current stream wait until the previous gather out of slot k has finished (before writing the slot again)."""
        torch.cuda.current_stream().wait_event(self.ev_comm[k])

    def gather_async(self, k):
        """Issue slot k's in-place all-gather on the communication stream, after everything queued on the current stream."""
        buf = self.bufs[k]
        self.ev_fwd[k].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ev_fwd[k])
            dist.all_gather_into_tensor(buf.view(-1), buf[self.rank], group=self.group)     # in place: row `rank` is the send buffer
            self.ev_comm[k].record(self.comm)

    def result(self, k):
        """Per-rank (cam_mesh, cam_pose, pose3d) views of slot k, valid once its gather has completed (waited for here)."""
        if self.cuda:
            torch.cuda.current_stream().wait_event(self.ev_comm[k])
        return [self.views(self.bufs[k], r) for r in range(self.world)]

    def unpack(self, k, total=None):
        """Global (cam_mesh [B,6890,3], cam_pose [B,J,3], pose3d [B,J,3]) of slot k (a concatenating copy), padding dropped."""
        parts = self.result(k)
        total = self.world * self.per if total is None else total
        return tuple(torch.cat([p[i] for p in parts], dim=0)[:total] for i in range(3))
