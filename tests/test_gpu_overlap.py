"""RT_FLAG_OVERLAP: consecutive accumulate passes on two internal streams with their own work counters, so that the drain
of one pass overlaps the start of the next.  The sums are order-independent integers: whatever runs beside whatever, the
frame must equal the serially rendered one bit for bit -- through rt_download, rt_download_begin (stream-ordered join),
rt_resolve_tiles on a compact shard, on a multi-device context, and after rt_join on the caller's stream."""
import numpy as np
import pytest
import torch

from raytracingoneweekendapplication_b200 import capi

pytestmark = pytest.mark.gpu


def _serial(ctx, sc, w, h, spp, passes, seed, **kw):
    for i in range(passes):
        ctx.render(w, h, spp, max_depth=sc.depth, seed=seed, spp_begin=i * spp, accumulate=i > 0, **kw)
    return ctx.accum_download()


def _overlapped(ctx, sc, w, h, spp, passes, seed, **kw):
    ctx.render(w, h, spp, max_depth=sc.depth, seed=seed, spp_begin=0, blocking=False, **kw)
    for i in range(1, passes):
        ctx.render(w, h, spp, max_depth=sc.depth, seed=seed, spp_begin=i * spp, accumulate=True, blocking=False, overlap=True, **kw)


@pytest.mark.parametrize("name,w,h,spp,passes", [("final", 320, 180, 3, 7), ("cornell_smoke", 200, 200, 2, 5), ("mesh", 256, 144, 1, 6)])
def test_overlapped_passes_render_the_serial_frame(ctx, scene_of, name, w, h, spp, passes):
    sc = scene_of(name)
    ctx.upload(sc)
    ref = _serial(ctx, sc, w, h, spp, passes, seed=4)
    _overlapped(ctx, sc, w, h, spp, passes, seed=4)
    got = ctx.accum_download()                      # waits for both lanes
    assert np.array_equal(got, ref)
    assert ctx.stats()["render_ms"] > 0.0
    # the same frame once more, handed out through the stream-ordered path (no host wait before the resolve)
    img_ref = ctx.download(spp * passes, linear=False, rgb8=True)
    _overlapped(ctx, sc, w, h, spp, passes, seed=4)
    ctx.download_begin(spp * passes, linear=False, rgb8=True)
    img = ctx.frame_end()[1].copy()
    assert np.array_equal(img, img_ref)
    # a plain pass after overlapped ones waits for them and clears the frame
    ctx.render(w, h, spp, max_depth=sc.depth, seed=4)
    one = ctx.accum_download()
    ctx.render(w, h, spp, max_depth=sc.depth, seed=4)
    assert np.array_equal(one, ctx.accum_download())


def test_join_orders_the_callers_stream_after_the_passes(ctx, scene_of):
    """What bench.py relies on: an event recorded on the caller's stream after rt_join covers the overlapped passes,
    and work enqueued there afterwards sees the finished frame."""
    sc = scene_of("final")
    ctx.upload(sc)
    w, h, spp, passes = 512, 288, 4, 6
    ref = torch.from_numpy(_serial(ctx, sc, w, h, spp, passes, seed=2).astype(np.int64).reshape(-1))
    buf = torch.zeros(w * h * 4, dtype=torch.int64, device="cuda")
    ctx.bind_accum(buf.data_ptr(), buf.numel() * 8, w, h)
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            _overlapped(ctx, sc, w, h, spp, passes, seed=2, stream=s.cuda_stream)
            ctx.join(s.cuda_stream)
            copy = buf.clone()                       # enqueued on s, after the join
        s.synchronize()
        assert torch.equal(copy.cpu(), ref)
        ctx.sync()
    finally:
        ctx.bind_accum(None, 0, 0, 0)


def test_overlap_on_compact_shards_and_multi_device_contexts(scene_of, monkeypatch):
    sc = scene_of("final")
    w, h, spp, passes = 480, 270, 2, 5
    c = capi.Context(0)
    try:
        c.upload(sc)
        full = _serial(c, sc, w, h, spp, passes, seed=7)
        # one shard of three, compact buffer
        kw = dict(shard_rank=1, shard_count=3, shard_mode=capi.RT_SHARD_TILES, compact=True, tile_size=16)
        ref = _serial(c, sc, w, h, spp, passes, seed=7, **kw)
        _overlapped(c, sc, w, h, spp, passes, seed=7, **kw)
        assert np.array_equal(c.accum_download(), ref)
    finally:
        c.close()
    monkeypatch.setenv("RT_B200_ALLOW_DUPLICATE_DEVICES", "1")
    m = capi.Context([0, 0])
    try:
        m.upload(sc)
        _overlapped(m, sc, w, h, spp, passes, seed=7)
        m.join()
        assert np.array_equal(m.accum_download(), full)
    finally:
        m.close()


def test_overlap_needs_the_frame_it_accumulates_onto(ctx, scene_of):
    sc = scene_of("quads")
    ctx.upload(sc)
    ctx.render(64, 64, 1, max_depth=sc.depth, seed=1)
    with pytest.raises(capi.RtError):
        ctx.render(96, 64, 1, max_depth=sc.depth, seed=1, spp_begin=1, accumulate=True, blocking=False, overlap=True)
    # without ACCUMULATE / ASYNC the flag is ignored: a plain blocking pass
    ctx.render(64, 64, 1, max_depth=sc.depth, seed=1, overlap=True)
    a = ctx.accum_download()
    ctx.render(64, 64, 1, max_depth=sc.depth, seed=1)
    assert np.array_equal(a, ctx.accum_download())
