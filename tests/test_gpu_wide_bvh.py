"""The wide quantised BVH (csrc/bvh_wide.h, csrc/rt_wide.cuh; replaces bvh.h:13-45,64-72 and
aabb.h:61-85) must give the same answers as the binary tree: the primitive records are the same
and the closest hit does not depend on the acceleration structure, so primary t, probe-ray t and
whole frames are compared for EQUALITY between widths 2, 4 and 8.  The only legitimate difference
is WHICH of two primitives hit at exactly the same t is reported (shared edges of the Cornell box,
coincident faces of kitchen_sink)."""
import ctypes as C

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
from raytracingoneweekendapplication_b200 import capi  # noqa: E402


def _per_width(scene, fn, widths=(2, 4, 8)):
    out = {}
    for w in widths:
        c = capi.Context(0)
        try:
            c.set_bvh_builder("host")
            c.set_bvh_width(w)
            c.upload(scene)
            out[w] = (fn(c), c.stats())
        finally:
            c.close()
    return out


@pytest.mark.parametrize("name", ["book1", "final", "mesh", "kitchen_sink", "quads", "cornell", "cornell_smoke", "mixed", "specular",
                                  "emissive"])
def test_primary_hits_and_frames_do_not_depend_on_the_width(scene_of, name):
    sc = scene_of(name)
    w, h = 160, 120

    def run(c):
        a = c.aov(w, h)
        c.render(w, h, 4, max_depth=sc.depth, seed=3)
        return a, c.accum_download()

    res = _per_width(sc, run)
    (ref, rs) = res[2]
    assert rs["bvh_width"] == 2 and rs["wide_nodes"] == 0
    for width in (4, 8):
        (got, st) = res[width]
        assert st["bvh_width"] == width and st["wide_nodes"] > 0 and 0 < st["wide_depth"] <= 24
        assert np.array_equal(ref[0]["t"], got[0]["t"]), f"{name} width {width}: primary t differs"
        tie = ref[0]["prim_id"] != got[0]["prim_id"]
        assert tie.mean() <= 0.01, f"{name} width {width}: primary ids differ on {tie.sum()} pixels"
        if name not in ("kitchen_sink", "cornell"):
            assert not tie.any(), f"{name} width {width}: {tie.sum()} ids differ"
        assert np.array_equal(ref[0]["normal"][~tie], got[0]["normal"][~tie])
        assert np.array_equal(ref[0]["uv"][~tie], got[0]["uv"][~tie])
        same_px = (ref[1] == got[1]).all(axis=-1)
        assert same_px.mean() >= (0.99 if tie.any() else 1.0), f"{name} width {width}: {(~same_px).sum()} pixels differ"


@pytest.mark.parametrize("name", ["final", "mesh", "cornell_smoke"])
def test_primary_hits_against_the_reference_fixture_with_the_wide_tree(scene_of, name):
    """The oracle gate of test_gpu_primary.py, through the width-8 traversal."""
    sc = scene_of(name)
    gold = helpers.golden("primary", name)
    assert gold is not None
    h, w = gold["ids"].shape
    c = capi.Context(0)
    try:
        c.set_bvh_width(8)
        c.upload(sc)
        assert c.stats()["bvh_width"] == 8
        aov = c.aov(w, h)
    finally:
        c.close()
    und = helpers.undecidable_pixels(sc, w, h)
    r = helpers.compare_primary(gold, aov, helpers.flat_leaf_keys(sc.desc), und)
    assert r["id_match"] >= 0.9999, r
    assert r["t_within_1e5"] >= 0.9999, r


@pytest.mark.parametrize("seed,ns,nq,nt", [(1, 40, 30, 60), (2, 300, 200, 500), (3, 0, 0, 700), (4, 900, 0, 0), (5, 5, 700, 3),
                                           (6, 1, 0, 0), (7, 2, 1, 0), (8, 3, 3, 3)])
def test_random_mixed_scenes_probe_rays(built, seed, ns, nq, nt):
    """Random soups of static + moving spheres, quads and triangles under instance transforms (and
    worlds of one, three and nine primitives: single-leaf and one-node trees): 20 000 arbitrary rays
    find the same closest hit through every width."""
    import fuzz_scenes

    sc = fuzz_scenes.random_scene(seed, ns, nq, nt)
    rays = fuzz_scenes.random_rays(50 + seed, 20000)
    res = _per_width(sc, lambda c: c.probe_hit(rays))
    for width in (4, 8):
        assert res[width][1]["bvh_width"] == width
        assert np.array_equal(res[2][0]["t"], res[width][0]["t"]), width
        same = res[2][0]["prim_id"] == res[width][0]["prim_id"]
        assert same.mean() >= 0.9995, (width, int((~same).sum()))
        assert np.array_equal(res[2][0]["normal"][same], res[width][0]["normal"][same])
    if ns + nq + nt > 20:
        assert (res[2][0]["prim_id"] >= 0).mean() > 0.2


def test_nee_and_shadowed_lights_through_the_wide_tree(scene_of):
    """Shadow rays (RT_FLAG_NEE, RT_FLAG_SHADOWED_POINT_LIGHTS) go through the same structure."""
    for name, kw in (("cornell", dict(nee=True)), ("mesh", dict(shadowed_point_lights=True))):
        sc = scene_of(name)

        def run(c):
            c.render(96, 96, 4, max_depth=12, seed=9, **kw)
            return c.accum_download()

        res = _per_width(sc, run, widths=(2, 8))
        same_px = (res[2][0] == res[8][0]).all(axis=-1)
        assert same_px.mean() >= 0.99, (name, int((~same_px).sum()))


def test_empty_world_and_stats(built):
    d = capi.rt_scene_desc()
    d.struct_size, d.abi_version = C.sizeof(d), capi.RT_B200_ABI_VERSION
    d.camera.lookat[:] = [0.0, 0.0, -1.0]
    d.camera.vup[:] = [0.0, 1.0, 0.0]
    d.camera.vfov, d.camera.focus_dist = 60.0, 1.0
    d.camera.background[:] = [0.25, 0.5, 1.0]
    for width in (4, 8):
        c = capi.Context(0)
        try:
            c.set_bvh_width(width)
            c.upload(d)
            c.render(32, 16, 2, max_depth=5, seed=1)
            lin = c.download(2)
            assert np.allclose(lin, [0.25, 0.5, 1.0], atol=1e-6)
            assert (c.aov(32, 16)["prim_id"] == -1).all()
        finally:
            c.close()


def test_invalid_width_is_refused(ctx):
    with pytest.raises(capi.RtError):
        ctx.set_bvh_width(3)


def test_device_built_scenes_keep_the_binary_tree(scene_of):
    c = capi.Context(0)
    try:
        c.set_bvh_builder("device")
        c.set_bvh_width(8)
        c.upload(scene_of("final"))
        st = c.stats()
        assert st["bvh_on_device"] == 1 and st["bvh_width"] == 2
        c.render(64, 36, 1, max_depth=5)
    finally:
        c.close()
