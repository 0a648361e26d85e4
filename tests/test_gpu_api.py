"""Error behaviour of the C ABI and the reference-facing C++ path (camera::render)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
from raytracingoneweekendapplication_b200 import capi  # noqa: E402


def test_render_before_upload_is_a_state_error(built):
    c = capi.Context(0)
    with pytest.raises(capi.RtError) as e:
        c.render(16, 16, 1)
    assert e.value.code == capi.RT_ERR_STATE
    c.close()


def test_bad_arguments_are_rejected(ctx, scene_of):
    ctx.upload(scene_of("quads"))
    for kwargs in ({"width": 0, "height": 16, "spp": 1}, {"width": 16, "height": 16, "spp": 0},
                   {"width": 16, "height": 16, "spp": 1, "shard_rank": 3, "shard_count": 2},
                   {"width": 16, "height": 16, "spp": 1, "tile_size": 12}):
        with pytest.raises(capi.RtError) as e:
            ctx.render(**kwargs)
        assert e.value.code == capi.RT_ERR_INVALID


def test_malformed_scene_is_rejected(ctx, scene_of):
    import copy

    sc = scene_of("quads")
    d = capi.rt_scene_desc()
    C.memmove(C.byref(d), C.byref(sc.desc), C.sizeof(d))
    d.abi_version = 999
    with pytest.raises(capi.RtError) as e:
        ctx.upload(d)
    assert e.value.code == capi.RT_ERR_INVALID
    C.memmove(C.byref(d), C.byref(sc.desc), C.sizeof(d))
    bad = (capi.rt_prim_ref * d.n_world)(*[capi.rt_prim_ref(1, 10 ** 6) for _ in range(d.n_world)])
    d.world = bad
    with pytest.raises(capi.RtError):
        ctx.upload(d)
    ctx.upload(sc)  # the context is still usable afterwards


def test_bound_accumulation_buffer_is_used(ctx, scene_of):
    """rt_bind_accum: the frame lands in a caller-owned device buffer (a torch tensor here),
    which is what bench.py reduces over NCCL."""
    import torch

    ctx.upload(scene_of("quads"))
    w, h = 64, 64
    ctx.render(w, h, 4, seed=1)
    own = ctx.download(4).copy()
    buf = torch.zeros(w * h * 4, dtype=torch.int64, device="cuda")
    ctx.bind_accum(buf.data_ptr(), buf.numel() * 8, w, h)
    ctx.render(w, h, 4, seed=1, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    lin = (buf.view(h, w, 4)[..., :3].double() / (2.0 ** capi.RT_ACCUM_FRAC_BITS) / 4).cpu().numpy()
    assert np.allclose(lin, own, atol=1e-6)
    ctx.bind_accum(None, 0, 0, 0)


def test_cpp_host_program_renders_through_camera_render(built, tmp_path):
    """apps/rtow_b200 is the reference's main() with the B200 path underneath: scenes.h ->
    host mirror -> camera::render(world, lights) -> C ABI -> PNG."""
    exe = os.path.join(helpers.ROOT, "apps", "rtow_b200")
    out = str(tmp_path / "quads.png")
    from raytracingoneweekendapplication_b200.assets import ensure_assets

    r = subprocess.run([exe, "quads", out, ensure_assets(), "96", "96", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Rendering Image" in r.stdout and "Done rendering" in r.stdout
    data = open(out, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    import struct
    import zlib

    w, h = struct.unpack(">II", data[16:24])
    assert (w, h) == (96, 96)
    idat = data[data.index(b"IDAT") + 4: data.index(b"IEND") - 8]
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)
    # same pixels as the same render through the Python binding (seed 1, 8 spp)
    c = capi.Context(0)
    sc = capi.Scene("quads")
    c.upload(sc)
    c.render(96, 96, 8, max_depth=50, seed=1)
    b8 = c.download(8, linear=False, rgb8=True)
    c.close()
    assert np.array_equal(raw, b8)
