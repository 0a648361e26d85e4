"""OBJ file -> first frame, end to end, for a generated terrain of N triangles (SURVEY 8f ranks 1+2):
load (mesh::loadObj: reference structure vs fast path), flatten, rt_upload_scene (host SAH vs
device LBVH), one 1080p 4-spp frame.  Usage: obj_e2e.py [N ...]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi  # noqa: E402


def write_terrain(path, n_tris):
    k = int(np.sqrt(n_tris / 2)) + 1
    xs = np.linspace(-5, 5, k)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    Y = 0.4 * np.sin(2.1 * X) * np.cos(1.7 * Z)
    with open(path, "w") as f:
        f.write("".join(f"v {x:.6f} {y:.6f} {z:.6f}\n" for x, y, z in zip(X.ravel(), Y.ravel(), Z.ravel())))
        f.write("".join(f"vt {i / (k - 1):.6f} {j / (k - 1):.6f}\n" for i in range(k) for j in range(k)))
        idx = (np.arange(k - 1)[:, None] * k + np.arange(k - 1)[None, :] + 1).ravel()
        f.write("".join(f"f {a}/{a} {a + k}/{a + k} {a + k + 1}/{a + k + 1} {a + 1}/{a + 1}\n" for a in idx))
    return 2 * (k - 1) ** 2


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100_000, 1_000_000]
    for n in sizes:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "terrain.obj")
            tris = write_terrain(path, n)
            size_mb = os.path.getsize(path) / 2 ** 20
            for per_triangle, builder in ((True, "host"), (False, "host"), (False, "device")):
                t0 = time.perf_counter()
                sc = capi.ObjScene(path, per_triangle=per_triangle)
                t_load = time.perf_counter() - t0
                ctx = capi.Context(0)
                ctx.set_bvh_builder(builder)
                t0 = time.perf_counter()
                ctx.upload(sc)
                t_up = time.perf_counter() - t0
                t0 = time.perf_counter()
                ctx.render(1920, 1080, 4, max_depth=10, seed=1)
                img = ctx.download(4, linear=False, rgb8=True)
                t_frame = time.perf_counter() - t0
                print(json.dumps({"triangles": tris, "obj_mb": round(size_mb, 1), "loader": "reference structure" if per_triangle else "fast path",
                                  "builder": builder, "parse_ms": round(sc.load_ms, 1), "flatten_ms": round(sc.flatten_ms, 1),
                                  "upload_ms": round(1e3 * t_up, 1), "frame_ms": round(1e3 * t_frame, 1),
                                  "file_to_frame_ms": round(1e3 * (t_load + t_up + t_frame), 1), "mean_rgb8": round(float(img.mean()), 2)}), flush=True)
                ctx.close()
                sc.close()


if __name__ == "__main__":
    main()
