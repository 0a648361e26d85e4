"""constant_medium boundaries made of six quads: the upload recognises ONE box() (the faces of a parallelepiped, under
any rigid transform) and intersects it with a three-slab test (boundary_pair_box); anything else -- a face moved, a face
too small, five faces -- keeps the reference's two passes over the boundary's primitives (constant_medium.h:23-27).
Observable through the counters: one boundary test per (ray, medium) for a box, 2 x 6 otherwise."""
import ctypes as C

import numpy as np
import pytest

import fuzz_scenes
from raytracingoneweekendapplication_b200 import capi

pytestmark = pytest.mark.gpu


def _box_quads(lo, hi, xform):
    """quad.h:76-97 box(a, b): six quads."""
    lo, hi = np.minimum(lo, hi).astype(float), np.maximum(lo, hi).astype(float)
    dx, dy, dz = np.array([hi[0] - lo[0], 0, 0.0]), np.array([0, hi[1] - lo[1], 0.0]), np.array([0, 0, hi[2] - lo[2]])
    faces = [((lo[0], lo[1], hi[2]), dx, dy), ((hi[0], lo[1], hi[2]), -dz, dy), ((hi[0], lo[1], lo[2]), -dx, dy),
             ((lo[0], lo[1], lo[2]), dz, dy), ((lo[0], hi[1], hi[2]), dx, -dz), ((lo[0], lo[1], lo[2]), dx, dz)]
    out = []
    for Q, u, v in faces:
        q = capi.rt_quad()
        q.Q, q.u, q.v = fuzz_scenes.d3(Q), fuzz_scenes.d3(u), fuzz_scenes.d3(v)
        q.material, q.xform = 1, xform
        out.append(q)
    return out


def _scene(mutate=None, n_faces=6):
    tex = capi.rt_texture()
    tex.type, tex.even, tex.odd, tex.image, tex.perlin = capi.RT_TEX_SOLID, -1, -1, -1, -1
    tex.color = fuzz_scenes.d3((0.7, 0.7, 0.7))
    lam, iso = capi.rt_material(), capi.rt_material()
    lam.type, lam.texture = capi.RT_MAT_LAMBERTIAN, 0
    iso.type, iso.texture = capi.RT_MAT_ISOTROPIC, 0
    ang = 0.4
    x = capi.rt_xform()
    x.r = (C.c_double * 9)(np.cos(ang), 0, np.sin(ang), 0, 1, 0, -np.sin(ang), 0, np.cos(ang))
    x.t = fuzz_scenes.d3((0.5, -1.0, 2.0))
    floor = capi.rt_quad()
    floor.Q, floor.u, floor.v = fuzz_scenes.d3((-40, -8, -40)), fuzz_scenes.d3((80, 0, 0)), fuzz_scenes.d3((0, 0, 80))
    floor.material, floor.xform = 0, -1
    faces = _box_quads(np.array([-4.0, -3.0, -2.0]), np.array([5.0, 4.0, 3.0]), 0)
    if mutate:
        mutate(faces)
    faces = faces[:n_faces]
    sc = fuzz_scenes.PyScene([], [floor] + faces, [], [x], [lam, iso], [tex])
    d = sc.desc
    world = (capi.rt_prim_ref * 1)(capi.rt_prim_ref(1, 0))
    refs = (capi.rt_prim_ref * len(faces))(*[capi.rt_prim_ref(1, 1 + i) for i in range(len(faces))])
    med = (capi.rt_medium * 1)()
    med[0].boundary_first, med[0].boundary_count = 0, len(faces)
    med[0].density, med[0].multiplicity, med[0].material, med[0].xform = 0.15, 1, 1, 0
    sc._keep += [world, refs, med]
    d.world, d.n_world = world, 1
    d.boundary_refs, d.n_boundary_refs = refs, len(faces)
    d.media, d.n_media = med, 1
    return sc


def _render(sc, monkeypatch, quads_only=False):
    if quads_only:
        monkeypatch.setenv("RT_B200_NO_BOX_MEDIA", "1")
    else:
        monkeypatch.delenv("RT_B200_NO_BOX_MEDIA", raising=False)
    c = capi.Context(0)
    try:
        c.upload(sc)
        c.render(160, 120, 48, max_depth=12, seed=5, stats=True)
        st = c.stats()
        return c.download(48).astype(np.float64), st["medium_queries"], st["boundary_tests"]
    finally:
        c.close()


def test_a_rotated_box_is_one_slab_test_and_renders_the_same_image(built, monkeypatch):
    sc = _scene()
    a, qa, ba = _render(sc, monkeypatch)
    b, qb, bb = _render(sc, monkeypatch, quads_only=True)
    assert qa > 0 and ba == qa and bb == 12 * qb
    assert a.std() > 0.01                                   # the medium is in view and scatters
    same = np.isclose(a, b, rtol=1e-4, atol=1e-6).all(axis=2).mean()
    assert same >= 0.98, same
    assert np.allclose(a.mean(axis=(0, 1)), b.mean(axis=(0, 1)), rtol=3e-3)


def _move_a_face(faces):
    faces[2].Q = fuzz_scenes.d3((5.0, -3.0, -2.5))          # the back face, half a unit further out: no longer a closed box


def _shrink_a_face(faces):
    faces[4].u = fuzz_scenes.d3((8.0, 0.0, 0.0))            # the top face one unit short


def _shear_a_face(faces):
    faces[0].v = fuzz_scenes.d3((0.5, 7.0, 0.0))            # front face no longer a rectangle of the box


@pytest.mark.parametrize("mutate,n_faces", [(_move_a_face, 6), (_shrink_a_face, 6), (_shear_a_face, 6), (None, 5)])
def test_anything_else_keeps_the_two_passes_over_the_quads(built, monkeypatch, mutate, n_faces):
    sc = _scene(mutate, n_faces)
    a, qa, ba = _render(sc, monkeypatch)
    b, qb, bb = _render(sc, monkeypatch, quads_only=True)
    assert qa == qb and ba == bb == 2 * n_faces * qa        # the generic path both times
    assert np.array_equal(a, b)
