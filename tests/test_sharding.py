"""Host-side sharding logic and the frame reduce, world_size 2 over gloo (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracingoneweekendapplication_b200 import sharding


@pytest.mark.parametrize("w,h,spp,count", [(400, 225, 10, 2), (3840, 2160, 64, 8), (600, 600, 200, 4), (33, 17, 5, 3)])
def test_plans_partition_the_frame(w, h, spp, count):
    for mode in (sharding.RT_SHARD_TILES, sharding.RT_SHARD_SAMPLES, sharding.RT_SHARD_AUTO):
        plans = [sharding.plan(w, h, spp, r, count, mode) for r in range(count)]
        assert sum(p.samples(w, h) for p in plans) == w * h * spp
        if plans[0].mode == sharding.RT_SHARD_TILES:
            tiles = sorted(t for p in plans for t in p.local_tiles)
            assert tiles == list(range(plans[0].tiles_x * plans[0].tiles_y))
        else:
            samples = sorted(p.sample_offset + k * p.sample_stride for p in plans for k in range(p.local_samples))
            assert samples == list(range(spp))


def test_auto_mode_picks_tiles_for_4k_and_samples_for_small_frames():
    assert sharding.plan(3840, 2160, 1024, 0, 8).mode == sharding.RT_SHARD_TILES
    assert sharding.plan(400, 225, 10, 0, 8).mode == sharding.RT_SHARD_SAMPLES
    assert sharding.plan(400, 225, 10, 0, 1).mode == sharding.RT_SHARD_TILES


def _synthetic_partial(w, h, spp, p):
    """What a rank's accumulation buffer holds: a deterministic integer per (pixel, sample)."""
    acc = np.zeros((h, w, 4), dtype=np.int64)
    ys, xs = np.mgrid[0:h, 0:w]
    pix = ys * w + xs
    owned = np.zeros((h, w), dtype=bool)
    for t in p.local_tiles:
        tx, ty = t % p.tiles_x, t // p.tiles_x
        owned[ty * p.tile_size:(ty + 1) * p.tile_size, tx * p.tile_size:(tx + 1) * p.tile_size] = True
    for k in range(p.local_samples):
        s = p.sample_offset + k * p.sample_stride
        val = (pix * 2654435761 + s * 40503) % (1 << 40)
        for c in range(3):
            acc[..., c] += np.where(owned, val + c, 0)
    return acc


def _worker(rank, world, port, w, h, spp, mode, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = sharding.plan(w, h, spp, rank, world, mode)
    acc = torch.from_numpy(_synthetic_partial(w, h, spp, p))
    sharding.reduce_frame(acc, dst=0)
    if rank == 0:
        np.save(out, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", [sharding.RT_SHARD_TILES, sharding.RT_SHARD_SAMPLES])
def test_two_rank_reduce_equals_single_rank(tmp_path, mode):
    w, h, spp = 70, 45, 6
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, port, w, h, spp, mode, out), nprocs=2, join=True)
    whole = _synthetic_partial(w, h, spp, sharding.plan(w, h, spp, 0, 1))
    assert np.array_equal(np.load(out), whole)
