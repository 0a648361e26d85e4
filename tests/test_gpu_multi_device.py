"""Multi-GPU inside the library (replaces the reference's row bands over CPU threads,
Camera.txt:59-61, 96-100): compact per-shard tile buffers, gather, reassembly.  The frame is a sum of integers
and a sample is a pure function of its Philox counter, so every comparison here is for EQUALITY with the
one-device, one-shard render.

A context over several devices needs several GPUs; on a one-GPU box the same code runs with the GPU listed
twice or four times (RT_B200_ALLOW_DUPLICATE_DEVICES, peer-copy gather); with >= 2 GPUs visible the real
thing runs too (NCCL send/recv when libnccl loads)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
from raytracingoneweekendapplication_b200 import capi  # noqa: E402


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def _reference(scene, w, h, spp, depth, seed=7, **kw):
    c = capi.Context(0)
    try:
        c.upload(scene)
        c.render(w, h, spp, max_depth=depth, seed=seed, **kw)
        lin, b8 = c.download(spp, linear=True, rgb8=True)
        return c.accum_download(), lin, b8
    finally:
        c.close()


@pytest.mark.parametrize("name,w,h,count,tile", [("cornell", 200, 136, 3, 16), ("final", 256, 144, 4, 8), ("mesh", 131, 77, 2, 32)])
def test_compact_shards_reassemble_to_the_full_frame(scene_of, name, w, h, count, tile):
    """One process per GPU (torchrun): every shard renders its tiles into a compact buffer, resolves them on the device,
    the shards are gathered (here: torch tensors on one GPU) and rt_untile rebuilds the frame."""
    import torch

    sc = scene_of(name)
    sums_ref, lin_ref, b8_ref = _reference(sc, w, h, 6, 12)
    cap = max(capi.load().rt_shard_pixels(w, h, tile, r, count) for r in range(count))
    assert cap * count >= w * h
    g8 = torch.zeros((count, cap * 3), dtype=torch.uint8, device="cuda:0")
    glin = torch.zeros((count, cap * 3), dtype=torch.float32, device="cuda:0")
    total = np.zeros_like(sums_ref)
    ctxs = []
    for r in range(count):
        c = capi.Context(0)
        ctxs.append(c)
        c.upload(sc)
        c.render(w, h, 6, max_depth=12, seed=7, shard_rank=r, shard_count=count, shard_mode=capi.RT_SHARD_TILES, tile_size=tile, compact=True)
        c.resolve_tiles(6, g8[r].data_ptr(), cap, dev_linear=glin[r].data_ptr())
        part = c.accum_download()          # this shard's tiles in place, zero elsewhere
        assert not (total.astype(bool) & part.astype(bool)).any(), "two shards wrote the same pixel"
        total += part
        one_lin, one_b8 = c.download(6, linear=True, rgb8=True)
        own = part[..., :3].any(axis=-1) | (one_b8 == b8_ref).all(axis=-1)
        assert (one_b8[own] == b8_ref[own]).all()
    torch.cuda.synchronize()
    assert np.array_equal(total, sums_ref)
    b8 = ctxs[0].untile(g8.data_ptr(), cap * 3, 3, count, w, h, tile)
    lin = ctxs[0].untile(glin.data_ptr(), cap * 12, 12, count, w, h, tile)
    assert np.array_equal(b8, b8_ref)
    assert np.array_equal(lin, lin_ref)
    for c in ctxs:
        c.close()


def test_compact_needs_tile_sharding(ctx, scene_of):
    ctx.upload(scene_of("quads"))
    with pytest.raises(capi.RtError):
        ctx.render(64, 64, 2, shard_rank=0, shard_count=2, shard_mode=capi.RT_SHARD_SAMPLES, compact=True)
    ctx.render(64, 64, 2, shard_rank=0, shard_count=2, shard_mode=capi.RT_SHARD_TILES, compact=True)
    with pytest.raises(capi.RtError):   # accumulate onto another layout
        ctx.render(64, 64, 2, spp_begin=2, accumulate=True)
    ctx.render(64, 64, 2)


def _multi(devices, scene, w, h, spp, depth, seed=7, passes=1, resume_from=None, **kw):
    c = capi.Context(devices)
    try:
        assert c.device_count() == len(devices)
        c.upload(scene)
        done = 0
        if resume_from is not None:
            c.accum_upload(resume_from[0])
            done = resume_from[1]
        per = (spp - done) // passes
        for i in range(passes):
            n = per if i < passes - 1 else spp - done
            c.render(w, h, n, max_depth=depth, seed=seed, spp_begin=done, accumulate=(done > 0), **kw)
            done += n
        lin, b8 = c.download(spp, linear=True, rgb8=True)
        return c.accum_download(), lin, b8, c.stats()
    finally:
        c.close()


def _device_lists():
    n = _gpu_count()
    lists = []
    if n >= 2:
        lists.append(list(range(min(n, 8))))
        lists.append([0, 1])
    return lists


@pytest.mark.parametrize("name,w,h", [("cornell_smoke", 320, 200), ("final", 480, 270), ("book1", 96, 54)])
def test_multi_device_context_renders_the_same_frame(scene_of, name, w, h, monkeypatch):
    """480x270 and 320x200 are tile-sharded (compact buffers, resolved tiles gathered), 96x54 has too few tiles and
    is sample-sharded (full-frame sums added on device 0)."""
    sc = scene_of(name)
    sums_ref, lin_ref, b8_ref = _reference(sc, w, h, 8, 10)
    cases = [(d, False) for d in _device_lists()]
    cases += [([0, 0], True), ([0, 0, 0, 0], True)]
    for devices, dup in cases:
        if dup:
            monkeypatch.setenv("RT_B200_ALLOW_DUPLICATE_DEVICES", "1")
        else:
            monkeypatch.delenv("RT_B200_ALLOW_DUPLICATE_DEVICES", raising=False)
        sums, lin, b8, st = _multi(devices, sc, w, h, 8, 10)
        assert st["devices"] == len(devices) and st["gather_mode"] in (1, 2)
        assert np.array_equal(b8, b8_ref), (name, devices)
        assert np.array_equal(lin, lin_ref), (name, devices)
        assert np.array_equal(sums, sums_ref), (name, devices)
        assert st["samples"] == w * h * 8
        # progressive passes and a restored checkpoint on a multi-device context
        sums2, _, b82, _ = _multi(devices, sc, w, h, 8, 10, passes=2)
        assert np.array_equal(sums2, sums_ref) and np.array_equal(b82, b8_ref)
        half = _reference(sc, w, h, 3, 10)[0]
        sums3, _, b83, _ = _multi(devices, sc, w, h, 8, 10, resume_from=(half, 3))
        assert np.array_equal(sums3, sums_ref) and np.array_equal(b83, b8_ref)


def test_duplicate_devices_are_refused_without_the_test_hook(built, monkeypatch):
    monkeypatch.delenv("RT_B200_ALLOW_DUPLICATE_DEVICES", raising=False)
    with pytest.raises(capi.RtError):
        capi.Context([0, 0])


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_nccl_and_peer_copy_gathers_agree(scene_of, monkeypatch):
    sc = scene_of("final")
    out = {}
    for mode in ("nccl", "p2p"):
        monkeypatch.setenv("RT_B200_GATHER", mode)
        try:
            out[mode] = _multi([0, 1], sc, 512, 288, 4, 10)
        except capi.RtError as e:      # libnccl.so.2 not loadable on this box
            if mode == "nccl":
                pytest.skip(str(e))
            raise
    assert out["nccl"][3]["gather_mode"] == 2 and out["p2p"][3]["gather_mode"] == 1
    assert np.array_equal(out["nccl"][2], out["p2p"][2]) and np.array_equal(out["nccl"][0], out["p2p"][0])


def test_asynchronous_hand_out_equals_the_blocking_download(scene_of):
    """rt_download_begin / rt_frame_end: the copy runs behind the next pass; frames come back in order, equal to
    rt_download's, and an unchanged scene is recognised on re-upload (only its staged arena is copied again)."""
    sc = scene_of("cornell")
    c = capi.Context(0)
    try:
        c.upload(sc)
        assert c.stats()["scene_reused"] == 0 and c.stats()["upload_bytes"] > 0
        want = []
        for i in range(3):
            c.render(160, 120, 2, max_depth=8, seed=3, spp_begin=2 * i)
            want.append(c.download(2, linear=True, rgb8=True))
        got = []
        for i in range(3):
            c.upload(sc)
            assert c.stats()["scene_reused"] == 1
            c.render(160, 120, 2, max_depth=8, seed=3, spp_begin=2 * i, blocking=False)
            c.download_begin(2, linear=True, rgb8=True)
            if i > 0:
                lin, b8 = c.frame_end()
                got.append((lin.copy(), b8.copy()))
        lin, b8 = c.frame_end()
        got.append((lin.copy(), b8.copy()))
        for (wl, w8), (gl, g8) in zip(want, got):
            assert np.array_equal(wl, gl) and np.array_equal(w8, g8)
        with pytest.raises(capi.RtError):
            c.frame_end()                      # nothing outstanding
        # a changed scene is uploaded in full again
        other = scene_of("quads")
        c.upload(other)
        assert c.stats()["scene_reused"] == 0
    finally:
        c.close()


def test_multi_device_asynchronous_hand_out(scene_of, monkeypatch):
    monkeypatch.setenv("RT_B200_ALLOW_DUPLICATE_DEVICES", "1")
    sc = scene_of("final")
    _, _, b8_ref = _reference(sc, 640, 360, 4, 10)
    c = capi.Context([0, 0])
    try:
        c.upload(sc)
        c.render(640, 360, 4, max_depth=10, seed=7)     # 920 tiles: tile-sharded
        c.download_begin(4, linear=False, rgb8=True)
        c.render(640, 360, 4, max_depth=10, seed=8)     # the next pass, while the first frame travels
        _, b8 = c.frame_end()
        assert np.array_equal(b8, b8_ref)
    finally:
        c.close()


def test_cpp_host_program_splits_the_frame_over_devices(built, tmp_path, monkeypatch):
    """apps/rtow_b200 ... --devices 0,0,0: main() -> camera::render(world, lights) with camera::devices set -> ONE context
    over three "devices" -> the same PNG as on one device (a resumed render included)."""
    import subprocess

    import helpers
    from raytracingoneweekendapplication_b200.assets import ensure_assets

    exe = os.path.join(helpers.ROOT, "apps", "rtow_b200")
    env = dict(os.environ, RT_B200_ALLOW_DUPLICATE_DEVICES="1")
    one, many, resumed = (str(tmp_path / n) for n in ("one.png", "many.png", "resumed.png"))
    ck = str(tmp_path / "ck.bin")
    base = ["final", None, ensure_assets(), "640", "360", "6"]
    for out, extra in ((one, []), (many, ["--devices", "0,0,0"])):
        cmd = [exe] + [out if a is None else a for a in base] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr
    assert open(one, "rb").read() == open(many, "rb").read()
    for extra in (["--stop-after", "2"], []):
        cmd = [exe] + [resumed if a is None else a for a in base] + ["--devices", "0,0", "--checkpoint", ck] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr
    assert "Resuming" in r.stderr
    assert open(one, "rb").read() == open(resumed, "rb").read()
