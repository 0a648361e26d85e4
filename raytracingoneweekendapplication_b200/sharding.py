"""Frame sharding across GPUs — the host-side mirror of the plan rt_render applies
(csrc/rt_b200.cu) plus the NVLink exchange.

The reference splits a frame into contiguous row bands, one per CPU thread, with no
communication (Camera.txt:59-61, 96-100).  Here the frame is cut into tile x tile pixel tiles
and tile t goes to rank t mod G (interleaved, so that expensive regions are spread over all
GPUs); frames with too few tiles shard the sample index instead (sample s -> rank s mod G).
Every rank renders into a full-frame accumulation buffer of 64-bit fixed-point sums (zero
where it owns nothing), and ONE collective finishes the frame: an int64 SUM reduce to rank 0
over NCCL/NVLink.  Integer addition is associative, so the result is bit-identical for any
number of ranks and either sharding mode.
"""
from __future__ import annotations

from dataclasses import dataclass

RT_SHARD_AUTO, RT_SHARD_TILES, RT_SHARD_SAMPLES = 0, 1, 2


@dataclass
class ShardPlan:
    mode: int
    rank: int
    count: int
    tile_size: int
    tiles_x: int
    tiles_y: int
    local_tiles: list       # global tile indices this rank renders
    local_samples: int      # samples per pixel this rank traces
    sample_offset: int
    sample_stride: int

    def pixels(self, width: int, height: int) -> int:
        n = 0
        for t in self.local_tiles:
            tx, ty = t % self.tiles_x, t // self.tiles_x
            n += min(self.tile_size, width - tx * self.tile_size) * min(self.tile_size, height - ty * self.tile_size)
        return n

    def samples(self, width: int, height: int) -> int:
        return self.pixels(width, height) * self.local_samples


def plan(width: int, height: int, spp: int, rank: int, count: int, mode: int = RT_SHARD_AUTO, tile_size: int = 16) -> ShardPlan:
    count = max(1, count)
    if not 0 <= rank < count:
        raise ValueError("rank outside [0, count)")
    tiles_x, tiles_y = -(-width // tile_size), -(-height // tile_size)
    n_tiles = tiles_x * tiles_y
    if mode == RT_SHARD_AUTO:
        mode = RT_SHARD_TILES if n_tiles // count >= 256 else RT_SHARD_SAMPLES
    if count == 1:
        mode = RT_SHARD_TILES
    if mode == RT_SHARD_TILES:
        return ShardPlan(mode, rank, count, tile_size, tiles_x, tiles_y, list(range(rank, n_tiles, count)), spp, 0, 1)
    if mode == RT_SHARD_SAMPLES:
        return ShardPlan(mode, rank, count, tile_size, tiles_x, tiles_y, list(range(n_tiles)), (spp - rank + count - 1) // count,
                         rank, count)
    raise ValueError("unknown shard mode")


def reduce_frame(accum, dst: int = 0):
    """Sum the per-rank accumulation buffers (torch int64 tensor, any device) onto `dst`.
    The only collective of the path; NCCL on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
