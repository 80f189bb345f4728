"""CPU: the N>1 sharding/gather plumbing with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uforecon_b200 import dist as ufodist


def test_shard_rows_partition():
    for H, W, world in [(1216, 1600, 8), (320, 416, 3), (7, 5, 4), (2, 9, 4)]:
        spans = [ufodist.shard_rows(H, W, world, r) for r in range(world)]
        pos = 0
        for b, n in spans:
            assert b == pos and n % W == 0
            pos += n
        assert pos == H * W
        rows = [n // W for _, n in spans]
        assert max(rows) - min(rows) <= 1


def test_shard_images_round_robin():
    got = sorted(sum((ufodist.shard_images(49, 8, r) for r in range(8)), []))
    assert got == list(range(49))
    assert len(ufodist.shard_images(49, 8, 0)) == 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, n = ufodist.shard_rows(H, W, world, rank)
    idx = torch.arange(begin, begin + n, dtype=torch.float32)
    depth = idx * 0.5                                   # stand-in for the rendered shard
    rgb = torch.stack([idx, idx + 1, idx + 2], 1)
    out = ufodist.gather_depth_rgb(depth, rgb, ufodist.shard_counts(H, W, world))
    if rank == 0:
        q.put((out[0], out[1]))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    H, W, world = 5, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    depth, rgb = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    idx = torch.arange(H * W, dtype=torch.float32)
    assert torch.equal(depth, idx * 0.5)
    assert torch.equal(rgb, torch.stack([idx, idx + 1, idx + 2], 1))
