#!/usr/bin/env python
"""The command ncu captures: N plain passes of one scene.  usage: prof_case.py scene W H spp [bvh_width] [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi  # noqa: E402

name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
width = int(sys.argv[5]) if len(sys.argv) > 5 else 2
passes = int(sys.argv[6]) if len(sys.argv) > 6 else 3
sc = capi.Scene(name)
c = capi.Context(0)
c.set_bvh_width(width)
c.upload(sc)
for i in range(passes):
    c.render(w, h, spp, max_depth=sc.depth, seed=1)
    print(name, "width", c.stats()["bvh_width"], "pass", i, round(c.stats()["render_ms"], 3), "ms",
          round(w * h * spp / c.stats()["render_ms"] / 1e3, 1), "Msamples/s", flush=True)
c.close()
