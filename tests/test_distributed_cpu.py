"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (ray sharding of a panorama, flat-gradient all-reduce
semantics of FlatAdam == DDP's mean of local means)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from panonerf_b200.parallel import shard_rows, all_reduce_flat_
    # 1. contiguous row blocks cover the image exactly once
    h = 37
    row0, nrows = shard_rows(h, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (row0, nrows))
    # 2. flat gradient all-reduce: sum over ranks, scale returned for the optimiser
    g = torch.full((1000,), float(rank + 1))
    scale = all_reduce_flat_(g)
    ret[rank] = (spans, g[:3].tolist(), scale)
    dist.destroy_process_group()


def test_ray_sharding_and_flat_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for rank in range(world):
        spans, g, scale = ret[rank]
        assert spans == [(0, 19), (19, 18)]
        assert g == [3.0, 3.0, 3.0] and scale == 0.5


def test_shard_rows_partitions():
    from panonerf_b200.parallel import shard_rows
    for h in (1, 7, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(h, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == h
            for (a, n), (b, _) in zip(spans, spans[1:]):
                assert a + n == b
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
