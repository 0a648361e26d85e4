"""How rt_upload_scene and the render kernel scale with the triangle count:
a synthetic soup of N small triangles on a bumpy terrain + one ground sphere, built straight
into the C ABI structs with numpy.  Prints one JSON line per N.

Usage: upload_scale.py [N ...]   (default 10000 100000 1000000)"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi  # noqa: E402

TRI = np.dtype([("p0", "<f8", 3), ("p1", "<f8", 3), ("p2", "<f8", 3), ("uv0", "<f4", 2), ("uv1", "<f4", 2), ("uv2", "<f4", 2),
                ("material", "<i4"), ("xform", "<i4")])
REF = np.dtype([("type", "<i4"), ("index", "<i4")])
assert TRI.itemsize == C.sizeof(capi.rt_triangle) and REF.itemsize == C.sizeof(capi.rt_prim_ref)


def terrain(n_tris: int, seed: int = 1):
    """A (k x k) height-field grid, two triangles per cell, ~n_tris triangles in [-50,50]^2."""
    k = max(2, int(np.sqrt(n_tris / 2)) + 1)
    xs = np.linspace(-50.0, 50.0, k)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    rng = np.random.default_rng(seed)
    Y = 3.0 * np.sin(0.21 * X) * np.cos(0.17 * Z) + 0.3 * rng.standard_normal(X.shape)
    P = np.stack([X, Y, Z], axis=-1)
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    t = np.zeros(2 * (k - 1) * (k - 1), dtype=TRI)
    t["p0"][0::2], t["p1"][0::2], t["p2"][0::2] = a.reshape(-1, 3), b.reshape(-1, 3), c.reshape(-1, 3)
    t["p0"][1::2], t["p1"][1::2], t["p2"][1::2] = a.reshape(-1, 3), c.reshape(-1, 3), d.reshape(-1, 3)
    t["uv1"] = [1, 0]
    t["uv2"] = [0, 1]
    t["material"] = 0
    t["xform"] = -1
    return t[:n_tris] if len(t) > n_tris else t


def scene_desc(tris: np.ndarray):
    keep = {}
    d = capi.rt_scene_desc()
    d.struct_size = C.sizeof(d)
    d.abi_version = capi.RT_B200_ABI_VERSION
    n = len(tris)
    refs = np.zeros(n + 1, dtype=REF)
    refs["type"][:n] = 2
    refs["index"][:n] = np.arange(n)
    refs["type"][n], refs["index"][n] = 0, 0
    sph = (capi.rt_sphere * 1)()
    sph[0].center0[:] = [0.0, -1010.0, 0.0]
    sph[0].radius = 1000.0
    sph[0].material, sph[0].xform = 1, -1
    mats = (capi.rt_material * 2)()
    texs = (capi.rt_texture * 2)()
    for i, col in enumerate([(0.6, 0.5, 0.3), (0.3, 0.6, 0.3)]):
        mats[i].type, mats[i].texture = 0, i
        texs[i].type = 0
        texs[i].color[:] = col
    d.world = refs.ctypes.data_as(C.POINTER(capi.rt_prim_ref)); d.n_world = n + 1
    d.triangles = tris.ctypes.data_as(C.POINTER(capi.rt_triangle)); d.n_triangles = n
    d.spheres = sph; d.n_spheres = 1
    d.materials = mats; d.n_materials = 2
    d.textures = texs; d.n_textures = 2
    d.camera.lookfrom[:] = [0.0, 40.0, 90.0]
    d.camera.lookat[:] = [0.0, 0.0, 0.0]
    d.camera.vup[:] = [0.0, 1.0, 0.0]
    d.camera.vfov = 45.0
    d.camera.focus_dist = 10.0
    d.camera.background[:] = [0.7, 0.8, 1.0]
    keep.update(refs=refs, sph=sph, mats=mats, texs=texs, tris=tris)
    return d, keep


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [10_000, 100_000, 1_000_000]
    for n in sizes:
        t0 = time.perf_counter()
        tris = terrain(n)
        d, keep = scene_desc(tris)
        t_gen = time.perf_counter() - t0
        ctx = capi.Context(0)
        mode = os.environ.get("RT_B200_BVH", "host")
        ctx.set_bvh_builder(mode)
        ups = []
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.upload(d)
            ups.append(time.perf_counter() - t0)
        w, h, spp = 1920, 1080, 4
        ctx.render(w, h, spp, max_depth=8, seed=1)
        best = 1e30
        for _ in range(2):
            ctx.render(w, h, spp, max_depth=8, seed=1)
            best = min(best, ctx.stats()["render_ms"])
        st = ctx.stats()
        aov = ctx.aov(w, h)
        print(json.dumps({"triangles": len(tris), "gen_s": round(t_gen, 3), "upload_ms_first": round(1e3 * ups[0], 1),
                          "upload_ms_best": round(1e3 * min(ups), 1), "bvh_nodes": st["bvh_nodes"], "bvh_depth": st["bvh_depth"],
                          "render_ms": round(best, 2), "msamples_s": round(w * h * spp / best / 1e3, 1),
                          "primary_hit_fraction": round(float((aov["prim_id"] >= 0).mean()), 4),
                          "builder": mode, "on_device": st["bvh_on_device"], "device_build_ms": round(st["device_build_ms"], 3),
                          "device_copy_in_ms": round(st["device_copy_in_ms"], 3), "device_top_ms": round(st["device_top_ms"], 3), "bvh_leaves": st["bvh_leaves"]}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
