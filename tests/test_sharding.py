"""Frame sharding + ordered gather (host logic), world_size 2 on gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from calipsync_b200.sharding import frame_shard, gather_chunked, gather_frames, shard_sizes


@pytest.mark.parametrize("n,w", [(1500, 1), (1500, 2), (1500, 4), (1500, 8), (5, 8), (0, 2), (7, 2)])
def test_shards_partition_frames_in_order(n, w):
    sizes = shard_sizes(n, w)
    assert sum(sizes) == n and len(sizes) == w
    spans = [frame_shard(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = frame_shard(n, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(hi - lo, 3, 4).contiguous()
        full = gather_frames(local, n, dst=0)
        if rank == 0:
            assert full.shape == (n, 3, 4)
            assert torch.equal(full[:, 0, 0], torch.arange(n, dtype=torch.float32))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 10])
def test_ordered_gather_world2_gloo(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def _worker_chunked(rank, world, port, n, batch):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def produce(lo, hi, out):                    # frame i of the clip = the constant i
            calls.append((lo, hi))
            out[: hi - lo] = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1)

        full = gather_chunked(produce, n, batch, (3, 4), torch.float32, "cpu", dst=0)
        lo, hi = frame_shard(n, rank, world)
        assert calls == [(a, min(a + batch, hi)) for a in range(lo, hi, batch)]     # own shard only, in order
        if rank == 0:
            assert full.shape == (n, 3, 4)
            assert torch.equal(full[:, 2, 3], torch.arange(n, dtype=torch.float32))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,batch", [(7, 2), (10, 4), (9, 16), (3, 2)])
def test_chunked_gather_world2_gloo(n, batch):
    """Per-batch gather (the overlap path of BASELINE config 3): ragged last chunks, shards of unequal size, a shard
    that needs fewer chunks than its neighbour, more ranks' worth of chunks than frames."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_chunked, args=(2, port, n, batch), nprocs=2, join=True)


def test_chunked_gather_single_process():
    def produce(lo, hi, out):
        out[: hi - lo] = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1)

    full = gather_chunked(produce, 11, 4, (1,), torch.float32, "cpu")
    assert torch.equal(full[:, 0], torch.arange(11, dtype=torch.float32))
