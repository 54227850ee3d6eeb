"""BASELINE config 3 on real GPUs: a clip frame-sharded over 2 ranks with the per-batch ordered NCCL gather must equal
the single-GPU synthesis bit for bit.  Needs two GPUs; skipped cleanly on a one-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _clip(n, dev):
    g = torch.Generator().manual_seed(77)
    crops = torch.randint(0, 256, (n, 160, 160, 3), dtype=torch.uint8, generator=g)
    feats = torch.randn(n, 2, 1024, generator=g)
    return crops.to(dev), feats.to(dev)


def _model(dev):
    from calipsync_b200 import Model
    from oracle import casync_oracle as O
    m = Model(6, "hubert")
    m.load_state_dict(O.make_state_dict(3, "R1"))
    return m.to(dev).eval()


def _worker(rank, world, port, n, batch, path):
    from calipsync_b200 import frame_shard, synthesize_clip
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        crops, feats = _clip(n, dev)
        lo, hi = frame_shard(n, rank, world)
        full = synthesize_clip(_model(dev), crops[lo:hi].contiguous(), feats, n, batch)
        torch.cuda.synchronize()
        if rank == 0:
            torch.save(full.cpu(), path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("n,batch", [(37, 8), (50, 32)])
def test_sharded_clip_with_chunked_nccl_gather_equals_single_gpu(n, batch, tmp_path):
    from calipsync_b200 import synthesize_clip
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    path = str(tmp_path / "full.pt")
    mp.spawn(_worker, args=(2, port, n, batch, path), nprocs=2, join=True)
    dev = torch.device("cuda", 0)
    crops, feats = _clip(n, dev)
    ref = synthesize_clip(_model(dev), crops, feats, n, batch)          # world size 1: no process group
    assert torch.equal(torch.load(path), ref.cpu())


def test_single_gpu_clip_equals_direct_forward_frames():
    from calipsync_b200 import synthesize_clip
    dev = torch.device("cuda", 0)
    crops, feats = _clip(21, dev)
    m = _model(dev)
    got = synthesize_clip(m, crops, feats, 21, 8)
    want = m.forward_frames(crops, feats, torch.arange(21, device=dev, dtype=torch.int32))
    assert torch.equal(got, want)
