"""Download-side additions: linear-radiance writers and resumable renders.

CPU part: the EXR / PFM writers of host/frame_io.h are parsed back with an independent reader
written here.  GPU part: a render interrupted at any sample count and resumed from its
checkpoint is BIT-IDENTICAL to the uninterrupted render (the frame is a sum of integers, and
every sample is a pure function of its Philox counter)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from raytracingoneweekendapplication_b200 import capi


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline() == b"PF\n"
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        assert scale < 0  # little endian
        data = np.frombuffer(f.read(), dtype="<f4").reshape(h, w, 3)
    return data[::-1]  # stored bottom-up


def read_exr(path):
    """Minimal reader of what write_exr emits: scanline, uncompressed, FLOAT channels."""
    b = open(path, "rb").read()
    magic, version = struct.unpack_from("<II", b, 0)
    assert magic == 20000630 and version == 2
    pos = 8
    attrs = {}
    while b[pos] != 0:
        end = b.index(b"\0", pos)
        name = b[pos:end].decode()
        pos = end + 1
        end = b.index(b"\0", pos)
        typ = b[pos:end].decode()
        pos = end + 1
        (size,) = struct.unpack_from("<I", b, pos)
        pos += 4
        attrs[name] = (typ, b[pos:pos + size])
        pos += size
    pos += 1
    for required in ("channels", "compression", "dataWindow", "displayWindow", "lineOrder", "pixelAspectRatio",
                     "screenWindowCenter", "screenWindowWidth"):
        assert required in attrs, required
    assert attrs["compression"] == ("compression", b"\0")
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans = []
    cb = attrs["channels"][1]
    p = 0
    while cb[p] != 0:
        e = cb.index(b"\0", p)
        chans.append(cb[p:e].decode())
        ptype, = struct.unpack_from("<i", cb, e + 1)
        assert ptype == 2
        p = e + 1 + 16
    assert chans == sorted(chans) == ["B", "G", "R"]
    offsets = struct.unpack_from("<%dQ" % h, b, pos)
    img = np.empty((h, w, 3), dtype=np.float32)
    for y in range(h):
        yy, nbytes = struct.unpack_from("<ii", b, offsets[y])
        assert yy == y and nbytes == w * 12
        line = np.frombuffer(b, dtype="<f4", count=3 * w, offset=offsets[y] + 8).reshape(3, w)
        img[y, :, 2], img[y, :, 1], img[y, :, 0] = line[0], line[1], line[2]
    assert offsets[-1] + 8 + w * 12 == len(b)
    return img


@pytest.mark.parametrize("w,h", [(1, 1), (7, 3), (64, 48)])
def test_writers_round_trip(built, tmp_path, w, h):
    s = capi.load_scenes()
    rng = np.random.default_rng(w * 100 + h)
    img = (rng.random((h, w, 3), dtype=np.float32) * 50.0).astype(np.float32)
    img[0, 0] = [0.0, 1e-30, 3.0e38]
    pe, pp = str(tmp_path / "a.exr"), str(tmp_path / "a.pfm")
    assert s.rtsc_write_exr(pe.encode(), w, h, img.ctypes.data) == 0
    assert s.rtsc_write_pfm(pp.encode(), w, h, img.ctypes.data) == 0
    assert np.array_equal(read_exr(pe), img)
    assert np.array_equal(read_pfm(pp), img)


def test_writers_report_unwritable_paths(built, tmp_path):
    s = capi.load_scenes()
    img = np.zeros((2, 2, 3), dtype=np.float32)
    bad = str(tmp_path / "no_such_dir" / "a.exr").encode()
    assert s.rtsc_write_exr(bad, 2, 2, img.ctypes.data) != 0
    assert s.rtsc_write_pfm(bad, 2, 2, img.ctypes.data) != 0


def test_scene_hash_tells_scenes_and_seeds_apart(built):
    s = capi.load_scenes()
    a, a2, b, c = capi.Scene("book1", 1), capi.Scene("book1", 1), capi.Scene("book1", 2), capi.Scene("cornell", 1)
    ha, ha2, hb, hc = (s.rtsc_scene_hash(x._h) for x in (a, a2, b, c))
    assert ha == ha2 and len({ha, hb, hc}) == 3 and 0 not in (ha, hb, hc)


# ---------------------------------------------------------------------------------------
# GPU: resume == uninterrupted, through the C ABI and through camera::render
# ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("scene,split", [("quads", 1), ("cornell_smoke", 5), ("final", 3)])
def test_accum_download_upload_resumes_bit_identically(ctx, scene_of, scene, split):
    sc = scene_of(scene)
    w, h, total = 96, 64, 8
    ctx.upload(sc)
    ctx.render(w, h, total, seed=7)
    whole = ctx.accum_download()
    whole8 = ctx.download(total, linear=False, rgb8=True)
    # first part in this context, the rest in a NEW context restored from the host copy
    ctx.render(w, h, split, seed=7)
    part = ctx.accum_download()
    assert part.shape == (h, w, 4) and part.dtype == np.uint64
    other = capi.Context(0)
    try:
        other.upload(sc)
        other.accum_upload(part)
        other.render(w, h, total - split, seed=7, spp_begin=split, accumulate=True)
        assert np.array_equal(other.accum_download(), whole)
        assert np.array_equal(other.download(total, linear=False, rgb8=True), whole8)
    finally:
        other.close()


@pytest.mark.gpu
def test_accum_upload_rejects_a_wrong_size(ctx, scene_of):
    ctx.upload(scene_of("quads"))
    sums = np.zeros((8, 8, 4), dtype=np.uint64)
    rc = ctx.lib.rt_accum_upload(ctx._h, sums.ctypes.data, sums.nbytes - 8, 8, 8)
    assert rc == capi.RT_ERR_INVALID
    rc = ctx.lib.rt_accum_upload(ctx._h, sums.ctypes.data, sums.nbytes, 0, 8)
    assert rc == capi.RT_ERR_INVALID


def _render(tmp_path, name, spp, ckpt=None, every=0, stop=0, linear=None, scene="cornell", depth=50, w=80, h=80):
    from raytracingoneweekendapplication_b200.assets import ensure_assets

    s = capi.load_scenes()
    out = str(tmp_path / name)
    done, resumed = C.c_int(), C.c_int()
    rc = s.rtsc_render_resumable(scene.encode(), 1, ensure_assets().encode(), w, h, spp, depth, out.encode(),
                                 ckpt.encode() if ckpt else None, every, stop, linear.encode() if linear else None,
                                 C.byref(done), C.byref(resumed))
    assert rc == 0
    return open(out, "rb").read(), done.value, resumed.value


@pytest.mark.gpu
def test_camera_render_resumes_from_its_checkpoint(built, tmp_path):
    """camera::render (Camera.txt:54) stopped after 5 of 12 samples, then called again: it picks
    the checkpoint up, renders the missing 7, writes the same PNG as one uninterrupted call and
    removes the checkpoint."""
    ck = str(tmp_path / "cornell.ckpt")
    ref_png, done, resumed = _render(tmp_path, "ref.png", 12)
    assert (done, resumed) == (12, 0)
    _, done, resumed = _render(tmp_path, "part.png", 12, ckpt=ck, stop=5)
    assert (done, resumed) == (5, 0) and os.path.exists(ck)
    assert os.path.getsize(ck) == 48 + 80 * 80 * 32
    png, done, resumed = _render(tmp_path, "full.png", 12, ckpt=ck)
    assert (done, resumed) == (12, 5)
    assert png == ref_png
    assert not os.path.exists(ck)


@pytest.mark.gpu
def test_checkpoint_of_another_render_is_ignored(built, tmp_path):
    ck = str(tmp_path / "x.ckpt")
    _render(tmp_path, "a.png", 6, ckpt=ck, stop=3, scene="cornell")
    assert os.path.exists(ck)
    ref_png, _, _ = _render(tmp_path, "ref.png", 6, scene="quads")
    png, done, resumed = _render(tmp_path, "b.png", 6, ckpt=ck, scene="quads")     # different scene: start over
    assert (done, resumed) == (6, 0) and png == ref_png
    _render(tmp_path, "a2.png", 6, ckpt=ck, stop=3, scene="cornell")
    png, done, resumed = _render(tmp_path, "c.png", 6, ckpt=ck, scene="cornell", depth=7)   # different depth
    assert resumed == 0
    open(ck, "wb").write(b"RTB2CKPT garbage")                                     # truncated file
    png, done, resumed = _render(tmp_path, "d.png", 6, ckpt=ck, scene="quads")
    assert (done, resumed) == (6, 0) and png == ref_png


@pytest.mark.gpu
def test_periodic_checkpoints_and_linear_output(built, tmp_path):
    ck = str(tmp_path / "p.ckpt")
    exr, pfm = str(tmp_path / "o.exr"), str(tmp_path / "o.pfm")
    _, done, _ = _render(tmp_path, "p1.png", 10, ckpt=ck, every=4, stop=9, linear=exr)
    assert done == 9
    hd = struct.unpack("<8sIiiiiIQQ", open(ck, "rb").read(48))
    assert hd[0] == b"RTB2CKPT" and hd[2:6] == (80, 80, 9, 50)
    _, done, resumed = _render(tmp_path, "p2.png", 10, ckpt=ck, linear=pfm)
    assert (done, resumed) == (10, 9)
    # the linear outputs are the library's float frame: 9-spp EXR vs 10-spp PFM of the same render
    c = capi.Context(0)
    sc = capi.Scene("cornell")
    c.upload(sc)
    c.render(80, 80, 9, seed=1)
    lin9 = c.download(9)
    c.render(80, 80, 10, seed=1)
    lin10 = c.download(10)
    c.close()
    assert np.array_equal(read_exr(exr), lin9)
    assert np.array_equal(read_pfm(pfm), lin10)
