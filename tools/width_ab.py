#!/usr/bin/env python
"""A/B of the acceleration-structure width (2 binary, 4 / 8 wide quantised) on the BASELINE scenes:
device time of a plain pass, and the per-ray traversal counters of a STATS pass.
usage: width_ab.py [scene ...]  (env AB_W, AB_H, AB_SPP, AB_WIDTHS)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi  # noqa: E402

scenes = sys.argv[1:] or ["final", "mesh", "cornell", "cornell_smoke", "book1"]
W, H, SPP = int(os.environ.get("AB_W", 1920)), int(os.environ.get("AB_H", 1080)), int(os.environ.get("AB_SPP", 32))
widths = [int(x) for x in os.environ.get("AB_WIDTHS", "2,4,8").split(",")]
for name in scenes:
    sc = capi.Scene(name)
    for width in widths:
        c = capi.Context(0)
        c.set_bvh_builder("host")
        c.set_bvh_width(width)
        c.upload(sc)
        for _ in range(2):
            c.render(W, H, SPP, max_depth=sc.depth, seed=1)
        best = 1e30
        for _ in range(3):
            c.render(W, H, SPP, max_depth=sc.depth, seed=1)
            best = min(best, c.stats()["render_ms"])
        c.render(W, H, max(1, SPP // 8), max_depth=sc.depth, seed=1, stats=True)
        st = c.stats()
        rays = max(st["rays"], 1)
        row = {"scene": name, "width": st["bvh_width"], "ms": round(best, 3), "msamples_s": round(W * H * SPP / best / 1e3, 1),
               "nodes": st["wide_nodes"] or st["bvh_nodes"], "depth": st["wide_depth"] or st["bvh_depth"],
               "node_steps_per_ray": round(st["node_visits"] / rays, 2), "box_tests_per_ray": round(st["box_tests"] / rays, 2),
               "empty_steps_per_ray": round(st["empty_node_steps"] / rays, 2),
               "prim_tests_per_ray": round((st["sphere_tests"] + st["quad_tests"] + st["triangle_tests"]) / rays, 2),
               "desc_lanes": round(st["desc_lanes"] / max(st["desc_iters"], 1), 2),
               "desc_trav_lanes": round(st["desc_trav_lanes"] / max(st["desc_iters"], 1), 2),
               "leaf_lanes": round(st["leaf_lanes"] / max(st["leaf_iters"], 1), 2),
               "leaf_visits_per_ray": round(st["leaf_iters"] and st["leaf_lanes"] / rays, 2),
               "shade_lanes": round(st["shade_lanes"] / max(st["shade_iters"], 1), 2),
               "desc_iters_per_ray": round(st["desc_iters"] / rays, 3), "regs": st["regs_per_thread"], "blocks": st["blocks"]}
        print(json.dumps(row), flush=True)
        c.close()
