"""Host mirror of the scene API: the flattened scenes are the reference's scenes (no GPU).

The golden leaf tables were dumped by oracle/ref_driver.cpp from the object graph the
UNMODIFIED reference headers built for the same scenes.h and seed."""
import json
import os

import numpy as np
import pytest

import helpers

SCENES = ["book1", "cornell", "cornell_smoke", "mesh", "final", "quads", "emissive", "specular", "mixed", "kitchen_sink", "monkey"]
SUMMARY = json.load(open(os.path.join(helpers.GOLDEN, "scenes.json")))


@pytest.mark.parametrize("name", SCENES)
def test_flattened_geometry_matches_reference_graph(scene_of, name):
    sc = scene_of(name)
    d = sc.desc
    ref = SUMMARY[name]
    assert d.n_world == ref["leaves"]
    assert (sc.width, sc.height, sc.spp, sc.depth) == (ref["width"], ref["height"], ref["spp"], ref["depth"])
    gold = helpers.golden("primary", name)
    keys = helpers.flat_leaf_keys(d)
    mapping, dist = helpers.match_leaves(gold["leaves"], keys)
    # same multiset of primitives: every reference leaf has a twin and vice versa
    assert len(mapping) == d.n_world
    canon = helpers.equivalent_ids(keys)
    assert sorted(canon[mapping].tolist()) == sorted(canon.tolist())
    assert dist < 1e-7


@pytest.mark.parametrize("name", SCENES)
def test_medium_density_and_bvh_leaf_multiplicity(scene_of, name):
    """a constant_medium that lands in a 1-object span of the reference's BVH is
    hit() twice per visit; the mirror must reproduce which media those are."""
    d = scene_of(name).desc
    ref = SUMMARY[name]
    mine = sorted((round(d.media[i].density, 12), d.media[i].multiplicity) for i in range(d.n_media))
    theirs = sorted(zip([round(x, 12) for x in ref["medium_density"]], ref["medium_multiplicity"]))
    assert mine == theirs
    if name == "kitchen_sink":
        assert [m for _, m in mine] == [2]


def test_perlin_tables_follow_the_reference_shuffle(scene_of):
    """perlin.h:64-71 swaps with random_int(0,1): the tables are weakly permuted but still
    permutations, and the gradient vectors are unit length."""
    d = scene_of("final").desc
    assert d.n_perlins == 1
    p = d.perlins[0]
    for perm in (p.perm_x, p.perm_y, p.perm_z):
        assert sorted(perm) == list(range(256))
        # "weak": swapping only with slots 0/1 leaves every entry within a few places of home
        assert np.median([abs(perm[i] - i) for i in range(256)]) <= 2
    v = np.array([[p.randvec[i][k] for k in range(3)] for i in range(256)])
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-12)


def test_image_texture_bytes_are_gamma_linearised(scene_of):
    """rtw_stb_image.h:99-105 on top of stb_image.h:1869: byte -> pow(b/255, 2.2) -> byte."""
    d = scene_of("mixed").desc
    assert d.n_images == 1
    im = d.images[0]
    assert (im.width, im.height) == (1024, 512)
    row0 = np.ctypeslib.as_array(im.rgb, shape=(512, 1024, 3))[0, :256, 0]  # the 0..255 ramp strip of earth.ppm
    lin = np.power(np.arange(256, dtype=np.float32) / np.float32(255.0), np.float32(2.2)).astype(np.float32)
    expect = np.where(lin <= 0, 0, np.where(lin >= 1, 255, np.floor(256.0 * lin.astype(np.float64)))).astype(np.uint8)
    assert np.abs(row0.astype(int) - expect.astype(int)).max() <= 1
    assert (row0 == expect).mean() > 0.98


def test_scene_is_reproducible_and_seeded(built):
    from raytracingoneweekendapplication_b200 import capi

    a, b, c = capi.Scene("book1", 1), capi.Scene("book1", 1), capi.Scene("book1", 2)
    ka, kb, kc = (helpers.flat_leaf_keys(s.desc) for s in (a, b, c))
    assert np.array_equal(ka, kb)
    assert ka.shape != kc.shape or not np.array_equal(ka, kc)


def test_unknown_scene_raises(built):
    from raytracingoneweekendapplication_b200 import capi

    with pytest.raises(ValueError):
        capi.Scene("no_such_scene")
