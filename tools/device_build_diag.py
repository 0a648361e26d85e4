import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from raytracingoneweekendapplication_b200 import capi
for name in ["book1", "final", "mesh", "kitchen_sink", "quads", "cornell", "cornell_smoke", "mixed", "emissive", "specular"]:
    sc = capi.Scene(name)
    res = []
    for mode in ("host", "device"):
        c = capi.Context(0); c.set_bvh_builder(mode); c.upload(sc)
        a = c.aov(160, 120); c.render(160, 120, 4, max_depth=sc.depth, seed=3); acc = c.accum_download(); st = c.stats(); c.close()
        res.append((a, acc, st))
    (ha, hacc, hs), (da, dacc, ds) = res
    diff = ha["prim_id"] != da["prim_id"]
    trel = np.abs(ha["t"][diff] - da["t"][diff]) / np.maximum(1e-30, np.abs(ha["t"][diff]))
    same_px = (hacc == dacc).all(axis=-1)
    print(name, "n_world", sc.desc.n_world, "on_dev", ds["bvh_on_device"], "nodes", hs["bvh_nodes"], ds["bvh_nodes"], "depth", hs["bvh_depth"], ds["bvh_depth"],
          "id diff", int(diff.sum()), "max t rel on diff", float(trel.max()) if diff.any() else 0.0,
          "t equal elsewhere", bool(np.array_equal(ha["t"][~diff], da["t"][~diff])), "img px differ", int((~same_px).sum()))
