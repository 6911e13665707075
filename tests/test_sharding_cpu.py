"""Multi-GPU partitioning logic on CPU: world_size-2 gloo processes shard
frames round-robin and gather screen bands; band-restricted flush descriptors
tile the frame exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rive_runtime_b200 import sharding, trace as T


def test_partition_helpers():
    for h in (1080, 2160, 16384, 17, 16):
        for n in (1, 2, 4, 8):
            bands = [sharding.band_for_rank(h, r, n) for r in range(n)]
            assert bands[0][0] == 0 and bands[-1][1] == h
            for a, b in zip(bands, bands[1:]):
                assert a[1] == b[0] and a[1] % 16 == 0
    assert sharding.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((sharding.frames_for_rank(1000, r, 8) for r in range(8)), [])) == list(range(1000))
    d = T.FlushDesc()
    d.update_bounds[:] = [0, 100, 640, 400]
    b = sharding.restrict_to_band(d, (256, 512))
    assert list(b.update_bounds) == [0, 256, 640, 400]
    b = sharding.restrict_to_band(d, (0, 96))
    assert b.update_bounds[3] <= b.update_bounds[1] or b.update_bounds[3] - b.update_bounds[1] == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, height, width, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # Each rank "renders" its band of a synthetic frame whose pixel value encodes (row, col).
        rows = torch.arange(height).view(-1, 1, 1).expand(height, width, 4)
        cols = torch.arange(width).view(1, -1, 1).expand(height, width, 4)
        full = ((rows * 7 + cols * 3) % 251).to(torch.uint8)
        r0, r1 = sharding.band_for_rank(height, rank, world)
        frame = sharding.gather_bands(full[r0:r1].contiguous(), height, width, dst_rank=0)
        ok_band = rank != 0 or bool(torch.equal(frame, full))
        # Frame sharding: rank r renders frames r, r+N, ...; gather puts them back in order.
        n_frames = 7
        mine = [np.full((2, 2, 4), i, np.uint8) for i in sharding.frames_for_rank(n_frames, rank, world)]
        frames = sharding.gather_frames(mine, n_frames, dst_rank=0)
        ok_frames = rank != 0 or all(int(f[0, 0, 0]) == i for i, f in enumerate(frames))
        q.put((rank, ok_band, ok_frames))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 100, 48, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok_band and ok_frames for _, ok_band, ok_frames in results), results
