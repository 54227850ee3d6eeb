"""Frame sharding for multi-GPU runs.  Frames are independent (module/unet.py:314-345 has no cross-frame
state; eval-mode BatchNorm uses running statistics), so rank r of W takes a contiguous frame range and
the only collective is the ordered gather of output frames to rank 0."""
from __future__ import annotations

import torch


def shard_sizes(n_frames: int, world: int):
    """Contiguous partition: ceil(n/W) frames per rank, trailing ranks may get fewer (or zero)."""
    per = -(-n_frames // world)
    return [max(0, min(n_frames, (r + 1) * per) - min(n_frames, r * per)) for r in range(world)]


def frame_shard(n_frames: int, rank: int, world: int):
    """(start, stop) of rank's frames."""
    sizes = shard_sizes(n_frames, world)
    start = sum(sizes[:rank])
    return start, start + sizes[rank]


def gather_frames(local: torch.Tensor, n_frames: int, dst: int = 0, group=None):
    """Ordered gather of per-rank output frames [n_r, ...] to `dst` (concatenation in rank order ==
    frame order because shards are contiguous).  Works on NCCL (device tensors) and gloo (CPU tensors).
    Returns the [n_frames, ...] tensor on `dst`, None elsewhere."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_frames, world)
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    if world == 1:
        return local
    per = max(sizes)
    pad = local
    if local.shape[0] < per:  # equal-size buffers keep this a single collective
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
