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
