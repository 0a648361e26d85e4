"""Traversal work per ray with the host SAH tree, the device tree and the pure LBVH (RT_FLAG_STATS).  Usage: tree_quality.py scene:w:h:spp ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi

for a in sys.argv[1:] or ["mesh:1920:1080:2", "final:1920:1080:2"]:
    name, w, h, spp = a.split(":")
    sc = capi.Scene(name)
    for mode in ("host", "device", "lbvh"):
        ctx = capi.Context(0)
        ctx.set_bvh_builder(mode)
        ctx.upload(sc)
        ctx.render(int(w), int(h), int(spp), max_depth=sc.depth, seed=1, stats=True)
        st = ctx.stats()
        r = st["rays"]
        print(json.dumps({"scene": name, "builder": mode, "nodes": st["bvh_nodes"], "leaves": st["bvh_leaves"], "depth": st["bvh_depth"],
                          "node_visits_per_ray": round(st["node_visits"] / r, 2), "leaf_visits_per_ray": round(st["leaf_lanes"] / r, 2),
                          "sphere_tests_per_ray": round(st["sphere_tests"] / r, 2), "quad_tests_per_ray": round(st["quad_tests"] / r, 2),
                          "tri_tests_per_ray": round(st["triangle_tests"] / r, 2), "fp64_sphere_per_ray": round(st["fp64_sphere_tests"] / r, 3)}))
        ctx.close()
