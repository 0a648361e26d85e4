"""Device-side scene upload (csrc/bvh_device.cuh): LBVH build and record
emission in CUDA kernels must give the same answers as the host path (binned SAH on the CPU).
The device records are bit-identical by construction (csrc/prim_derive.h) and the closest hit
does not depend on the tree, so primary hits and whole images are compared for EQUALITY; the
only legitimate differences are rays that hit two primitives at exactly the same t (shared
edges of the Cornell box), which the two trees may resolve differently."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
from raytracingoneweekendapplication_b200 import capi  # noqa: E402

sys.path.insert(0, os.path.join(helpers.ROOT, "tools"))


def _both(scene, fn, modes=("host", "device")):
    out = []
    for mode in modes:
        c = capi.Context(0)
        try:
            c.set_bvh_builder(mode)
            c.upload(scene)
            out.append((fn(c), c.stats()))
        finally:
            c.close()
    return out


@pytest.mark.parametrize("name", ["book1", "final", "mesh", "kitchen_sink", "quads", "cornell", "cornell_smoke", "mixed"])
def test_primary_hits_and_images_match_the_host_path(scene_of, name):
    sc = scene_of(name)
    w, h = 160, 120

    def run(c):
        a = c.aov(w, h)
        c.render(w, h, 4, max_depth=sc.depth, seed=3)
        return a, c.accum_download()

    (host, hs), (dev, ds) = _both(sc, run)
    assert hs["bvh_on_device"] == 0
    if sc.desc.n_world < 8:
        assert ds["bvh_on_device"] == 0   # tiny scenes stay on the host path
        return
    assert ds["bvh_on_device"] == 1 and ds["bvh_nodes"] > 0 and 0 < ds["bvh_depth"] < 62
    tie = host[0]["prim_id"] != dev[0]["prim_id"]
    # the ids may differ ONLY where two primitives are hit at exactly the same t (kitchen_sink has
    # coincident faces, the Cornell box shared edges): t is equal everywhere, bit for bit
    assert np.array_equal(host[0]["t"], dev[0]["t"]), f"{name}: primary t differs"
    assert tie.mean() <= 0.01, f"{name}: primary ids differ on {tie.sum()} pixels"
    if name not in ("kitchen_sink", "cornell"):
        assert not tie.any()
    assert np.array_equal(host[0]["normal"][~tie], dev[0]["normal"][~tie])
    same_px = (host[1] == dev[1]).all(axis=-1)
    assert same_px.mean() >= (0.99 if tie.any() else 1.0), f"{name}: {(~same_px).sum()} pixels differ"


@pytest.mark.parametrize("name", ["final", "mesh", "book1", "cornell"])
def test_device_built_tree_against_the_reference_fixture(scene_of, name):
    """The device-built tree checked DIRECTLY against the fixture dumped from the unmodified reference
    (not only against the host-built tree): the gate of test_gpu_primary.py with set_bvh_builder('device')."""
    import test_gpu_primary

    sc = scene_of(name)
    gold = helpers.golden("primary", name)
    h, w = gold["ids"].shape
    c = capi.Context(0)
    try:
        c.set_bvh_builder("device")
        c.upload(sc)
        st = c.stats()
        assert st["bvh_on_device"] == 1 and 0 < st["bvh_depth"] < 62   # the depth the unchecked device stack relies on
        aov = c.aov(w, h)
    finally:
        c.close()
    r = helpers.compare_primary(gold, aov, helpers.flat_leaf_keys(sc.desc), helpers.undecidable_pixels(sc, w, h))
    test_gpu_primary.record(f"device_built/{name}/{w}x{h}", r)
    test_gpu_primary.check(r)


def test_large_triangle_soup(built):
    import upload_scale

    tris = upload_scale.terrain(200_000)
    d, keep = upload_scale.scene_desc(tris)
    (host, hs), (dev, ds) = _both(d, lambda c: c.aov(640, 360))
    assert ds["bvh_on_device"] == 1
    assert np.array_equal(host["prim_id"], dev["prim_id"])
    assert np.array_equal(host["t"], dev["t"])
    assert np.array_equal(host["uv"], dev["uv"])
    assert (host["prim_id"] >= 0).mean() > 0.5
    # the device path exists to make the upload cheap
    assert ds["device_build_ms"] > 0 and ds["device_build_ms"] < 50


@pytest.mark.parametrize("name", ["final", "mesh"])
def test_pure_lbvh_and_sah_top_levels_agree(scene_of, name):
    """'device' rebuilds the radix tree's small subtrees (one warp per cluster) and its top levels
    (host, over a cut of a few thousand clusters) with SAH; 'lbvh' keeps the radix tree as it is.
    Same records, same closest hits."""
    sc = scene_of(name)

    def run(c):
        c.render(160, 120, 3, max_depth=sc.depth, seed=11)
        return c.accum_download()

    (hyb, hs), (pure, ps) = _both(sc, run, modes=("device", "lbvh"))
    assert hs["bvh_on_device"] == ps["bvh_on_device"] == 1
    assert hs["bvh_nodes"] != ps["bvh_nodes"]     # different trees ...
    assert np.array_equal(hyb, pure)


def test_auto_mode_keeps_small_scenes_on_the_host(ctx, scene_of):
    ctx.set_bvh_builder("auto")
    ctx.upload(scene_of("final"))
    assert ctx.stats()["bvh_on_device"] == 0


def test_coincident_primitives_do_not_break_the_radix_tree(built):
    """All Morton codes equal: the tree is built from the position tie-break alone."""
    n = 300
    sph = (capi.rt_sphere * n)()
    refs = (capi.rt_prim_ref * n)()
    for i in range(n):
        sph[i].center0[:] = [0.0, 0.0, -5.0]
        sph[i].radius = 1.0 + 0.001 * (i % 7)
        sph[i].material, sph[i].xform = 0, -1
        refs[i].type, refs[i].index = 0, i
    mats = (capi.rt_material * 1)()
    texs = (capi.rt_texture * 1)()
    mats[0].type, mats[0].texture = 0, 0
    texs[0].type = 0
    texs[0].color[:] = [0.5, 0.5, 0.5]
    d = capi.rt_scene_desc()
    d.struct_size, d.abi_version = C.sizeof(d), capi.RT_B200_ABI_VERSION
    d.world, d.n_world = refs, n
    d.spheres, d.n_spheres = sph, n
    d.materials, d.n_materials = mats, 1
    d.textures, d.n_textures = texs, 1
    d.camera.lookat[:] = [0.0, 0.0, -1.0]
    d.camera.vup[:] = [0.0, 1.0, 0.0]
    d.camera.vfov, d.camera.focus_dist = 60.0, 1.0
    (host, hs), (dev, ds) = _both(d, lambda c: c.aov(64, 64))
    assert np.array_equal(host["t"], dev["t"])
    assert (dev["prim_id"] >= 0).any()


@pytest.mark.parametrize("seed,ns,nq,nt", [(1, 40, 30, 60), (2, 300, 200, 500), (3, 0, 0, 700), (4, 900, 0, 0), (5, 5, 700, 3),
                                           (6, 130, 0, 0), (7, 64, 64, 1)])
def test_random_mixed_scenes_probe_rays(built, seed, ns, nq, nt):
    """Random soups of static + moving spheres, quads and triangles under instance transforms: 20 000
    arbitrary rays must find the same closest hit whichever builder made the tree (the per-cluster
    SAH rebuild sees mixed types, coincident centres, ranges of every size up to its 128 limit)."""
    import fuzz_scenes

    sc = fuzz_scenes.random_scene(seed, ns, nq, nt)
    rays = fuzz_scenes.random_rays(50 + seed, 20000)
    res = {}
    for mode in ("host", "device", "lbvh"):
        c = capi.Context(0)
        try:
            c.set_bvh_builder(mode)
            c.upload(sc)
            res[mode] = (c.probe_hit(rays), c.stats())
        finally:
            c.close()
    assert res["device"][1]["bvh_on_device"] == 1 and res["host"][1]["bvh_on_device"] == 0
    for mode in ("device", "lbvh"):
        assert np.array_equal(res["host"][0]["t"], res[mode][0]["t"]), mode
        same = res["host"][0]["prim_id"] == res[mode][0]["prim_id"]
        assert same.mean() >= 0.9995, (mode, int((~same).sum()))     # exact-t ties between overlapping random primitives
        assert np.array_equal(res["host"][0]["normal"][same], res[mode][0]["normal"][same])
    assert (res["host"][0]["prim_id"] >= 0).mean() > 0.2
