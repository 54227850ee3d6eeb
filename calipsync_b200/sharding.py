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


def gather_chunked(produce, n_frames: int, batch: int, frame_shape, dtype, device, dst: int = 0, group=None, out=None):
    """Rank's shard of an n_frames clip, produced in batches of `batch` frames; every finished batch is gathered to
    `dst` while the next one is being computed ("chunked and issued on a side stream so it overlaps
    the next chunk's compute").  `produce(lo, hi, out)` fills out[: hi - lo] with frames [lo, hi) of the clip (global
    indices).  All ranks issue the same number of equal-size collectives (ragged chunks travel zero padded).
    `out` (multi-rank runs only): a caller-owned staging buffer [chunks, batch, *frame_shape] to reuse across calls --
    stable addresses keep the CUDA graphs of the producing forwards valid.
    Returns the ordered [n_frames, *frame_shape] tensor on `dst`, None elsewhere."""
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    sizes = shard_sizes(n_frames, world)
    lo = sum(sizes[:rank])
    n_local = sizes[rank]
    nchunks = -(-max(sizes) // batch) if max(sizes) else 0
    cuda = torch.device(device).type == "cuda"
    shape = (max(nchunks, 1), batch) + tuple(frame_shape)
    if out is None or not multi or tuple(out.shape) != shape or out.dtype != dtype:
        out = torch.zeros(shape, dtype=dtype, device=device)
    recv = ([torch.empty_like(out) for _ in range(world)] if (multi and rank == dst) else None)
    side = torch.cuda.Stream(device) if (cuda and multi) else None
    for c in range(nchunks):
        a, b = min(n_local, c * batch), min(n_local, (c + 1) * batch)
        if b > a:
            produce(lo + a, lo + b, out[c])
        if not multi:
            continue
        glist = [r[c] for r in recv] if rank == dst else None
        if side is None:
            dist.gather(out[c], glist, dst=dst, group=group)
            continue
        ev = torch.cuda.Event()
        ev.record()
        side.wait_event(ev)
        with torch.cuda.stream(side):                 # NCCL orders the collective behind this chunk's kernels only
            dist.gather(out[c], glist, dst=dst, group=group)
    if side is not None:
        torch.cuda.current_stream(device).wait_stream(side)
    flat = (nchunks * batch,) + tuple(frame_shape)
    if not multi:
        return out.reshape(flat)[:n_local]
    if rank != dst:
        return None
    return torch.cat([r.reshape(flat)[:n] for r, n in zip(recv, sizes)], dim=0)


def synthesize_clip(model, crops_u8, hubert_feats, n_frames: int, batch: int = 64, dst: int = 0, group=None):
    """BASELINE config 3: every rank holds the crops of ITS shard (uint8 [n_r,160,160,3], frame order) and the clip's
    HuBERT features (replicated, [n_frames,2,1024]); frames are synthesised in batches through Model.forward_frames and
    gathered in order to `dst` with the per-batch gather overlapping the next batch (gather_chunked)."""
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    lo0, hi0 = frame_shard(n_frames, rank, world)
    assert crops_u8.shape[0] == hi0 - lo0, (crops_u8.shape, lo0, hi0)
    dev = crops_u8.device

    def produce(lo, hi, out):
        idx = torch.arange(lo, hi, device=dev, dtype=torch.int32)
        model.forward_frames(crops_u8[lo - lo0: hi - lo0], hubert_feats, idx, out=out[: hi - lo])

    staging = None
    if multi:   # one staging buffer per clip geometry, kept on the model (see gather_chunked)
        cache = model.__dict__.setdefault("_clip_out", {})
        key = (n_frames, batch, world, str(dev))
        staging = cache.get(key)
        if staging is None:
            if len(cache) >= 2:
                cache.pop(next(iter(cache)))
            nchunks = max(1, -(-max(shard_sizes(n_frames, world)) // batch))
            staging = cache[key] = torch.zeros((nchunks, batch, 160, 160, 3), dtype=torch.uint8, device=dev)
    return gather_chunked(produce, n_frames, batch, (160, 160, 3), torch.uint8, dev, dst, group, out=staging)
