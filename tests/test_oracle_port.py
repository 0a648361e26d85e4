"""Pins the CPU restatement (oracle/rt_oracle.cpp) against the REAL reference: the golden
fixtures were produced by the unmodified reference headers (tests/golden/make_golden.py).
No GPU needed."""
import numpy as np
import pytest

import helpers

PRIMARY_SCENES = ["book1", "cornell", "cornell_smoke", "mesh", "final", "quads", "emissive", "specular", "mixed", "kitchen_sink", "monkey"]


@pytest.mark.parametrize("name", PRIMARY_SCENES)
def test_port_primary_hits_equal_reference(scene_of, name):
    from oracle import port

    sc = scene_of(name)
    g = helpers.golden("primary", name)
    h, w = g["ids"].shape
    r = helpers.compare_primary(g, port.primary(sc, w, h), helpers.flat_leaf_keys(sc.desc))
    assert r["mismatches"] == 0, r
    assert r["t_rel_max"] <= 1e-12, r           # same formulas, same doubles
    assert r["n_err_max"] <= 1e-7 and r["uv_err_max"] <= 1e-7, r  # fixture stores float32


def _kat_inputs(kat):
    rec = np.zeros((kat.shape[0], 16))
    rec[:, 0:3] = kat[:, 1:4]
    rec[:, 3:6] = kat[:, 4:7]
    rec[:, 6] = kat[:, 7]
    rec[:, 7:10] = kat[:, 9:12]
    rec[:, 10:13] = kat[:, 12:15]
    rec[:, 13] = kat[:, 15]
    rec[:, 14] = kat[:, 16]
    rec[:, 15] = kat[:, 17]
    return rec, kat[:, 18:22]


def kat_cases(sc, g):
    """Yields (material index, rows) of a KAT fixture grouped by the flattened material."""
    d = sc.desc
    kat = g["kat"]
    mapping, _ = helpers.match_leaves(g["leaves"], helpers.flat_leaf_keys(d))
    mats = []
    for leaf in kat[:, 0].astype(int):
        ref = d.world[int(mapping[leaf])]
        arr = (d.spheres, d.quads, d.triangles)[ref.type]
        mats.append(arr[ref.index].material)
    mats = np.array(mats)
    for m in np.unique(mats):
        yield int(m), np.nonzero(mats == m)[0]


@pytest.mark.parametrize("name", ["kitchen_sink", "mixed", "final", "book1", "specular", "mesh", "monkey"])
def test_port_scatter_and_textures_equal_reference(scene_of, name):
    from oracle import port

    sc = scene_of(name)
    g = helpers.golden("kat", name)
    kat = g["kat"]
    rec, uni = _kat_inputs(kat)
    seen_types = set()
    for m, rows in kat_cases(sc, g):
        out = port.scatter(sc, m, rec[rows], uni[rows])
        ref = kat[rows]
        seen_types.add(sc.desc.materials[m].type)
        assert np.array_equal(out[:, 0], ref[:, 22]), f"scatter flag, material {m}"
        ok = ref[:, 22] != 0
        np.testing.assert_allclose(out[ok, 1:4], ref[ok, 23:26], rtol=0, atol=1e-12, err_msg=f"attenuation, material {m}")
        np.testing.assert_allclose(out[ok, 4:7], ref[ok, 26:29], rtol=0, atol=1e-9, err_msg="scattered origin")
        np.testing.assert_allclose(out[ok, 7:10], ref[ok, 29:32], rtol=0, atol=1e-12, err_msg=f"scattered direction, material {m}")
        np.testing.assert_allclose(out[:, 10:13], ref[:, 33:36], rtol=0, atol=1e-12, err_msg="emitted")
    assert len(seen_types) >= 2


@pytest.mark.parametrize("name,spp", [("quads", 64), ("emissive", 128), ("mixed", 48), ("kitchen_sink", 48)])
def test_port_image_statistics_match_reference(scene_of, name, spp):
    """The restatement's Monte Carlo estimate agrees with the reference's converged render:
    per-channel image mean within 4 sigma of the combined noise."""
    from oracle import port

    sc = scene_of(name)
    g = helpers.golden("image", name)
    ref, ref_var, ref_spp = g["image"].astype(np.float64), g["var"].astype(np.float64), int(g["spp"])
    h, w, _ = ref.shape
    img, var = port.render(sc, w, h, spp, depth=int(g["depth"]), seed=11, want_var=True)
    n = h * w
    sigma = np.sqrt((ref_var / ref_spp).sum(axis=(0, 1)) + (var / spp).sum(axis=(0, 1))) / n
    diff = np.abs(img.mean(axis=(0, 1)) - ref.mean(axis=(0, 1)))
    assert np.all(diff <= 4 * sigma + 1e-6), (diff, sigma)
