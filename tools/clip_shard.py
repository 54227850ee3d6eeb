"""BASELINE config 3: a synthetic 60 s clip (1500 frames @25 fps) frame-sharded over the ranks of one box, every batch
of output frames gathered IN ORDER to rank 0 over NCCL on a side stream (calipsync_b200.synthesize_clip), timed with the
gather inside the region, and checked bit-exact against rank 0 synthesising the whole clip alone.

    python tools/clip_shard.py [n_frames] [batch]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/clip_shard.py [n_frames] [batch]

Inputs (uint8 crops per frame index, HuBERT-like features) are generated and the communicator / CUDA graphs are warmed
up BEFORE the event pair; the timed region is synthesize_clip only (window gather + crop assembly + forward + uint8
epilogue + gather), max over ranks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from calipsync_b200 import frame_shard, synthesize_clip  # noqa: E402


def clip_crops(lo, hi, dev):
    """Frame i depends only on its index, so any rank can generate its shard (one generator per 64-frame block)."""
    out = torch.empty(hi - lo, 160, 160, 3, dtype=torch.uint8, device=dev)
    for blk in range(lo // 64, (hi + 63) // 64):
        g = torch.Generator(device=dev).manual_seed(10_000 + blk)
        full = torch.randint(0, 256, (64, 160, 160, 3), dtype=torch.uint8, device=dev, generator=g)
        a, b = max(lo, blk * 64), min(hi, blk * 64 + 64)
        out[a - lo: b - lo] = full[a - blk * 64: b - blk * 64]
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = bench.build_model(dev)
    feats = torch.randn(n, 2, 1024, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
    lo, hi = frame_shard(n, rank, world)
    crops = clip_crops(lo, hi, dev)
    for _ in range(2):                                          # warm-up: NCCL communicator, graph capture, allocator
        synthesize_clip(net, crops, feats, n, batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 3
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        full = synthesize_clip(net, crops, feats, n, batch)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref = net.forward_frames(clip_crops(0, n, dev), feats, torch.arange(n, device=dev, dtype=torch.int32)) if world > 1 \
            else None
        same = True if ref is None else bool(torch.equal(full, ref))
        print(json.dumps({"config": "clip of %d frames, batches of %d, frame-sharded dp%d, per-batch ordered gather of "
                                    "uint8 HWC frames to rank 0 (overlapped)" % (n, batch, world),
                          "n_gpus": world, "frames": n, "batch": batch, "ms_per_clip_incl_gather": float(ms),
                          "frames_per_s": n / (float(ms) / 1e3),
                          "gathered_equals_single_gpu_bit_exact": same}))
        assert same
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
