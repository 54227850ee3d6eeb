"""Frame sharding + ordered gather (host logic), world_size 2 on gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from calipsync_b200.sharding import frame_shard, gather_frames, shard_sizes


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
