"""Opt-in next-event estimation (RT_FLAG_NEE).

The estimator is a different one from the reference's (individual samples differ), but it is built
to have the SAME expectation: direct light from the quad emitters is sampled on the emitters and
weighted with the density of the reference's own direction sampler, and the scattered ray does
not collect a listed emitter's emission a second time.  So the parity gate is gate (b) of the
converged images — against the images rendered by the UNMODIFIED reference (tests/golden/
image_*.npz) — reached with far fewer samples, plus a direct NEE-vs-plain comparison."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
from raytracingoneweekendapplication_b200 import capi  # noqa: E402

NEE_SCENES = ["cornell", "cornell_smoke", "final", "emissive", "kitchen_sink"]


def _halves(ctx, w, h, spp, depth, nee):
    ctx.render(w, h, spp // 2, max_depth=depth, seed=11, nee=nee)
    a = ctx.download(spp // 2).astype(np.float64)
    ctx.render(w, h, spp // 2, max_depth=depth, seed=22, nee=nee)
    b = ctx.download(spp // 2).astype(np.float64)
    return a, b


@pytest.mark.parametrize("name", NEE_SCENES)
def test_nee_converges_to_the_reference_image(ctx, scene_of, name):
    sc = scene_of(name)
    ctx.upload(sc)
    g = helpers.golden("image", name)
    ref, ref_var, ref_spp, depth = g["image"].astype(np.float64), g["var"].astype(np.float64), int(g["spp"]), int(g["depth"])
    h, w, _ = ref.shape
    a, b = _halves(ctx, w, h, 8192, depth, nee=True)
    dev = 0.5 * (a + b)
    n = h * w
    sigma_ref = np.sqrt((ref_var / ref_spp).sum(axis=(0, 1))) / n
    sigma_dev = 0.5 * np.sqrt(((a - b) ** 2).sum(axis=(0, 1))) / n
    sigma = np.sqrt(sigma_ref ** 2 + sigma_dev ** 2)
    diff = np.abs(dev.mean(axis=(0, 1)) - ref.mean(axis=(0, 1)))
    assert np.all(diff <= 3 * sigma + 1e-7), (name, diff, sigma, diff / sigma)
    psnr = helpers.psnr_after_gamma(dev, ref)
    assert psnr >= 40.0, (name, psnr)


@pytest.mark.parametrize("name,w,h", [("cornell", 96, 96), ("cornell_smoke", 96, 96)])
def test_nee_has_the_same_mean_and_less_noise_than_plain_sampling(ctx, scene_of, name, w, h):
    sc = scene_of(name)
    ctx.upload(sc)
    spp = 2048
    pa, pb = _halves(ctx, w, h, spp, 50, nee=False)
    na, nb = _halves(ctx, w, h, spp, 50, nee=True)
    n = w * h
    s_plain = 0.5 * np.sqrt(((pa - pb) ** 2).sum(axis=(0, 1))) / n
    s_nee = 0.5 * np.sqrt(((na - nb) ** 2).sum(axis=(0, 1))) / n
    diff = np.abs((0.5 * (pa + pb)).mean(axis=(0, 1)) - (0.5 * (na + nb)).mean(axis=(0, 1)))
    assert np.all(diff <= 3 * np.sqrt(s_plain ** 2 + s_nee ** 2) + 1e-7), (diff, s_plain, s_nee)
    # per-pixel noise: the two halves of the NEE render agree much better than the plain ones
    noise_plain, noise_nee = np.mean((pa - pb) ** 2), np.mean((na - nb) ** 2)
    assert noise_nee < 0.5 * noise_plain, (noise_plain, noise_nee)


def test_nee_without_listed_emitters_is_the_plain_estimator(ctx, scene_of):
    """book1 has no quad emitter: the flag changes nothing, bit for bit."""
    ctx.upload(scene_of("book1"))
    ctx.render(120, 68, 4, seed=3)
    plain = ctx.accum_download()
    ctx.render(120, 68, 4, seed=3, nee=True)
    assert np.array_equal(plain, ctx.accum_download())


def test_nee_is_deterministic_and_shardable(ctx, scene_of):
    ctx.upload(scene_of("cornell"))
    ctx.render(64, 64, 6, seed=5, nee=True)
    whole = ctx.accum_download()
    ctx.render(64, 64, 6, seed=5, nee=True, shard_rank=0, shard_count=2, shard_mode=2)
    ctx.render(64, 64, 6, seed=5, nee=True, shard_rank=1, shard_count=2, shard_mode=2, accumulate=True)
    assert np.array_equal(whole, ctx.accum_download())


def test_camera_render_takes_the_flag(built, tmp_path):
    """The reference-facing call: `cam.next_event_estimation = true; cam.render(world, lights)`
    (apps/rtow_b200 --nee 1) writes the PNG of the same frame the C ABI renders with RT_FLAG_NEE."""
    import os
    import struct
    import subprocess
    import zlib

    from raytracingoneweekendapplication_b200.assets import ensure_assets

    exe = os.path.join(helpers.ROOT, "apps", "rtow_b200")
    out = str(tmp_path / "cornell_nee.png")
    r = subprocess.run([exe, "cornell", out, ensure_assets(), "72", "72", "6", "--nee", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    data = open(out, "rb").read()
    w, h = struct.unpack(">II", data[16:24])
    idat = data[data.index(b"IDAT") + 4: data.index(b"IEND") - 8]
    png = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)
    sc = capi.Scene("cornell")
    c = capi.Context(0)
    c.upload(sc)
    c.render(72, 72, 6, max_depth=sc.depth, seed=1, nee=True)
    with_nee = c.download(6, linear=False, rgb8=True)
    c.render(72, 72, 6, max_depth=sc.depth, seed=1)
    plain = c.download(6, linear=False, rgb8=True)
    c.close()
    assert np.array_equal(png, with_nee)
    assert not np.array_equal(png, plain)


def test_shadowed_point_lights(built):
    """RT_FLAG_SHADOWED_POINT_LIGHTS: a floor, a sphere above it, a point light above the sphere and a
    black sky.  The reference's unshadowed term (Camera.txt:240-272) lights the floor under the
    sphere as if the sphere were not there; with the flag that patch goes dark and nothing else changes."""
    import ctypes as C

    import fuzz_scenes

    d3 = fuzz_scenes.d3
    tex = capi.rt_texture()
    tex.type, tex.even, tex.odd, tex.image, tex.perlin = capi.RT_TEX_SOLID, -1, -1, -1, -1
    tex.color = d3((0.8, 0.8, 0.8))
    mat = capi.rt_material()
    mat.type, mat.texture = capi.RT_MAT_LAMBERTIAN, 0
    floor = capi.rt_quad()
    floor.Q, floor.u, floor.v = d3((-20, 0, -20)), d3((40, 0, 0)), d3((0, 0, 40))
    floor.material, floor.xform = 0, -1
    ball = capi.rt_sphere()
    ball.center0, ball.center_vec, ball.radius = d3((0, 3, 0)), d3((0, 0, 0)), 1.5
    ball.material, ball.xform = 0, -1
    light = capi.rt_point_light()
    light.position, light.intensity, light.size = d3((0, 6, 0)), d3((40, 40, 40)), 0.1
    cam = capi.rt_camera()
    cam.lookfrom, cam.lookat, cam.vup = d3((0, 14, 0.01)), d3((0, 0, 0)), d3((0, 0, -1))
    cam.vfov, cam.defocus_angle, cam.focus_dist = 60, 0, 10
    cam.background = d3((0, 0, 0))
    sc = fuzz_scenes.PyScene([ball], [floor], [], [], [mat], [tex], camera=cam, lights=[light])
    c = capi.Context(0)
    c.upload(sc)
    w = h = 96
    c.render(w, h, 64, max_depth=1, seed=1)                       # depth 1: the point-light term of the first hit only
    plain = c.download(64).mean(axis=2)
    c.render(w, h, 64, max_depth=1, seed=1, shadowed_point_lights=True)
    shadowed = c.download(64).mean(axis=2)
    aov = c.aov(w, h)["prim_id"].reshape(h, w)
    c.close()
    floor_px = aov == [r.type for r in sc.desc.world[:2]].index(1)   # the quad's canonical id
    # the shadow is a disc of radius 1.5 * 6 / sqrt(3^2 - 1.5^2) = 3.46 (20 pixels) around the centre, the ball's
    # silhouette covers 11 pixels of it: the ring of floor between 13 and 18 pixels is in shadow
    yy, xx = np.mgrid[0:h, 0:w]
    r_px = np.hypot(xx - (w - 1) / 2, yy - (h - 1) / 2)
    ring = floor_px & (r_px > 13) & (r_px < 18)
    assert ring.sum() > 100
    assert plain[ring].min() > 0.05 and shadowed[ring].max() < 1e-6
    far = floor_px & (r_px > 40)
    assert np.array_equal(plain[far], shadowed[far])
    ball_px = ~floor_px & (r_px < 9)                                 # well inside the silhouette (samples are jittered)
    assert ball_px.sum() > 100 and np.array_equal(plain[ball_px], shadowed[ball_px])   # the ball itself is lit in both


def test_nee_is_the_same_with_the_device_built_scene(scene_of):
    """The device upload path builds the emitter list from the source quad array; same records, same
    light numbers, so an NEE frame does not depend on who built the scene."""
    sc = scene_of("final")
    frames = []
    for mode in ("host", "device"):
        c = capi.Context(0)
        c.set_bvh_builder(mode)
        c.upload(sc)
        c.render(160, 90, 4, max_depth=50, seed=6, nee=True)
        frames.append(c.accum_download())
        c.render(160, 90, 4, max_depth=50, seed=6)
        frames.append(c.accum_download())
        c.close()
    assert np.array_equal(frames[0], frames[2])
    assert np.array_equal(frames[1], frames[3])
    assert not np.array_equal(frames[0], frames[1])


def test_nee_lists_only_emitters_the_world_references():
    """An emissive quad that is not in the world list (the boundary of a medium, or referenced by nothing) can never be
    hit by a path, so it must not be sampled as an emitter either -- on the host-built and on the device-built scene
    alike (the device path reads the source quad array, the host path the world order)."""
    import ctypes as C

    import fuzz_scenes

    base = fuzz_scenes.random_scene(3, 12, 8, 0, with_xforms=False)
    d = base.desc
    # two extra quads: the world light, and a huge bright emissive quad that nothing references
    nq = d.n_quads
    quads = (capi.rt_quad * (nq + 2))()
    for i in range(nq):
        quads[i] = d.quads[i]
    mats = (capi.rt_material * (d.n_materials + 1))()
    for i in range(d.n_materials):
        mats[i] = d.materials[i]
    texs = (capi.rt_texture * (d.n_textures + 1))()
    for i in range(d.n_textures):
        texs[i] = d.textures[i]
    texs[d.n_textures].type = capi.RT_TEX_SOLID
    texs[d.n_textures].color[:] = [30.0, 30.0, 30.0]
    mats[d.n_materials].type, mats[d.n_materials].texture = capi.RT_MAT_DIFFUSE_LIGHT, d.n_textures
    for k, (q0, u, v) in enumerate((((-3, 12, -3), (6, 0, 0), (0, 0, 6)), ((-40, -12, -40), (80, 0, 0), (0, 0, 80)))):
        q = quads[nq + k]
        q.Q[:], q.u[:], q.v[:] = q0, u, v
        q.material, q.xform = d.n_materials, -1
    world = (capi.rt_prim_ref * (d.n_world + 1))()
    for i in range(d.n_world):
        world[i] = d.world[i]
    world[d.n_world].type, world[d.n_world].index = capi.RT_PRIM_QUAD, nq      # only the first of the two
    d2 = capi.rt_scene_desc()
    C.memmove(C.byref(d2), C.byref(d), C.sizeof(d))
    d2.quads, d2.n_quads = quads, nq + 2
    d2.materials, d2.n_materials = mats, d.n_materials + 1
    d2.textures, d2.n_textures = texs, d.n_textures + 1
    d2.world, d2.n_world = world, d.n_world + 1
    frames = {}
    for mode in ("host", "device"):
        c = capi.Context(0)
        try:
            c.set_bvh_builder(mode)
            c.upload(d2)
            assert c.stats()["bvh_on_device"] == (1 if mode == "device" else 0)
            c.render(96, 64, 8, max_depth=6, seed=4, nee=True)
            frames[mode] = c.accum_download()
        finally:
            c.close()
    assert np.array_equal(frames["host"], frames["device"])
