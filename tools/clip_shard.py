"""BASELINE config 3: a synthetic 60 s clip (1500 frames @25 fps) frame-sharded over the ranks of one box, outputs
gathered IN ORDER to rank 0 over NCCL, and checked bit-exact against rank 0 synthesising the whole clip alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/clip_shard.py [n_frames]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from calipsync_b200 import frame_shard  # noqa: E402
from calipsync_b200.sharding import gather_frames  # noqa: E402


def window(features, lo, hi):
    """HuBERT windowing of infer_api.py:99-145 on device: rows idx-8..idx+8 of [T,2,1024], zero-padded, -> [n,32,32,32]."""
    T = features.shape[0]
    idx = torch.arange(lo, hi, device=features.device)[:, None] + torch.arange(-8, 8, device=features.device)[None, :]
    ok = (idx >= 0) & (idx < T)
    w = features[idx.clamp(0, T - 1)] * ok[:, :, None, None]
    return w.reshape(hi - lo, 32, 32, 32).contiguous()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = bench.build_model(dev)
    g = torch.Generator(device=dev).manual_seed(1234)          # same clip on every rank (replicated features)
    feats = torch.randn(n, 2, 1024, device=dev, generator=g)
    lo, hi = frame_shard(n, rank, world)

    def frames(a, b):                                           # frame i depends only on its index
        xs = []
        for i in range(a, b):
            gi = torch.Generator(device=dev).manual_seed(10_000 + i)
            x = torch.rand(6, 160, 160, device=dev, generator=gi)
            x[3:6, 5:150, 5:155] = 0
            xs.append(x)
        return torch.stack(xs) if xs else torch.empty(0, 6, 160, 160, device=dev)

    def synth(a, b, bs=256):
        outs = [net.forward_uint8(frames(s, min(s + bs, b)), window(feats, s, min(s + bs, b))) for s in range(a, b, bs)]
        return torch.cat(outs) if outs else torch.empty(0, 160, 160, 3, dtype=torch.uint8, device=dev)

    synth(lo, min(hi, lo + 8))                                  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    local_out = synth(lo, hi)
    full = gather_frames(local_out, n, dst=0) if world > 1 else local_out
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref = synth(0, n)
        same = bool(torch.equal(full, ref))
        print(json.dumps({"config": "clip of %d frames, frame-sharded dp%d, ordered gather of uint8 HWC frames to rank 0" % (n, world),
                          "n_gpus": world, "frames": n, "ms_incl_input_synthesis_and_gather": float(ms), 
                          "gathered_equals_single_gpu_bit_exact": same}))
        assert same
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
