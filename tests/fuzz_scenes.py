"""Small random scenes built directly as rt_scene_desc from Python (ctypes), for fuzzing the
device intersection code against the CPU restatement."""
import ctypes as C

import numpy as np

from raytracingoneweekendapplication_b200 import capi


class PyScene:
    """Keeps the ctypes arrays alive and exposes .desc_ptr like capi.Scene."""

    def __init__(self, spheres, quads, tris, xforms, materials, textures, camera=None, lights=()):
        d = capi.rt_scene_desc()
        d.struct_size = C.sizeof(capi.rt_scene_desc)
        d.abi_version = capi.RT_B200_ABI_VERSION
        self._keep = []

        def arr(ctype, items):
            a = (ctype * max(1, len(items)))(*items)
            self._keep.append(a)
            return a

        world = [capi.rt_prim_ref(0, i) for i in range(len(spheres))] + [capi.rt_prim_ref(1, i) for i in range(len(quads))] + \
                [capi.rt_prim_ref(2, i) for i in range(len(tris))]
        rng = np.random.RandomState(len(world))
        rng.shuffle(world)
        d.world, d.n_world = arr(capi.rt_prim_ref, world), len(world)
        d.boundary_refs, d.n_boundary_refs = arr(capi.rt_prim_ref, []), 0
        d.spheres, d.n_spheres = arr(capi.rt_sphere, spheres), len(spheres)
        d.quads, d.n_quads = arr(capi.rt_quad, quads), len(quads)
        d.triangles, d.n_triangles = arr(capi.rt_triangle, tris), len(tris)
        d.media, d.n_media = arr(capi.rt_medium, []), 0
        d.xforms, d.n_xforms = arr(capi.rt_xform, xforms), len(xforms)
        d.materials, d.n_materials = arr(capi.rt_material, materials), len(materials)
        d.textures, d.n_textures = arr(capi.rt_texture, textures), len(textures)
        d.images, d.n_images = arr(capi.rt_image, []), 0
        d.perlins, d.n_perlins = arr(capi.rt_perlin, []), 0
        d.lights, d.n_lights = arr(capi.rt_point_light, list(lights)), len(lights)
        cam = camera or capi.rt_camera()
        if camera is None:
            cam.lookfrom = (C.c_double * 3)(0, 0, -30)
            cam.lookat = (C.c_double * 3)(0, 0, 0)
            cam.vup = (C.c_double * 3)(0, 1, 0)
            cam.vfov, cam.defocus_angle, cam.focus_dist = 40, 0, 10
            cam.background = (C.c_double * 3)(0.5, 0.7, 1.0)
        d.camera = cam
        self.desc = d
        self.desc_ptr = C.pointer(d)


def d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def random_scene(seed, n_spheres=40, n_quads=30, n_tris=60, extent=10.0, with_xforms=True):
    rng = np.random.RandomState(seed)
    tex = capi.rt_texture()
    tex.type, tex.even, tex.odd, tex.image, tex.perlin = capi.RT_TEX_SOLID, -1, -1, -1, -1
    tex.color = d3((0.6, 0.6, 0.6))
    mat = capi.rt_material()
    mat.type, mat.texture = capi.RT_MAT_LAMBERTIAN, 0
    xforms = []
    if with_xforms:
        for _ in range(4):
            ang = rng.uniform(-np.pi, np.pi)
            c, s = np.cos(ang), np.sin(ang)
            x = capi.rt_xform()
            x.r = (C.c_double * 9)(c, 0, s, 0, 1, 0, -s, 0, c)
            x.t = d3(rng.uniform(-3, 3, 3))
            xforms.append(x)

    def pick_xf():
        return int(rng.randint(-1, len(xforms))) if xforms else -1

    spheres, quads, tris = [], [], []
    for i in range(n_spheres):
        s = capi.rt_sphere()
        s.center0 = d3(rng.uniform(-extent, extent, 3))
        s.center_vec = d3(rng.uniform(-1, 1, 3) if i % 5 == 0 else (0, 0, 0))
        s.radius = float(rng.uniform(0.2, 2.5))
        s.material, s.xform = 0, pick_xf()
        spheres.append(s)
    for _ in range(n_quads):
        q = capi.rt_quad()
        q.Q = d3(rng.uniform(-extent, extent, 3))
        u = rng.uniform(-3, 3, 3)
        v = np.cross(u, rng.uniform(-1, 1, 3))
        v *= rng.uniform(0.5, 3) / max(np.linalg.norm(v), 1e-9)
        q.u, q.v = d3(u), d3(v)
        q.material, q.xform = 0, pick_xf()
        quads.append(q)
    for _ in range(n_tris):
        t = capi.rt_triangle()
        p0 = rng.uniform(-extent, extent, 3)
        t.p0, t.p1, t.p2 = d3(p0), d3(p0 + rng.uniform(-2, 2, 3)), d3(p0 + rng.uniform(-2, 2, 3))
        t.uv0 = (C.c_float * 2)(*rng.uniform(0, 1, 2))
        t.uv1 = (C.c_float * 2)(*rng.uniform(0, 1, 2))
        t.uv2 = (C.c_float * 2)(*rng.uniform(0, 1, 2))
        t.material, t.xform = 0, pick_xf()
        tris.append(t)
    return PyScene(spheres, quads, tris, xforms, [mat], [tex])


def random_rays(seed, n, extent=10.0):
    """Rays whose components are exactly representable in FP32 (so both sides see the same input)."""
    rng = np.random.RandomState(seed)
    o = rng.uniform(-2 * extent, 2 * extent, (n, 3))
    target = rng.uniform(-extent, extent, (n, 3))
    d = (target - o) * rng.uniform(0.2, 3.0, (n, 1))
    rays = np.zeros((n, 9))
    rays[:, 0:3], rays[:, 3:6] = o, d
    rays[:, 6] = rng.uniform(0, 1, n)
    rays[:, 7] = 0.001
    rays[:, 8] = np.inf
    return rays.astype(np.float32).astype(np.float64)
