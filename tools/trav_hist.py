"""Dumps the traversal-shape histograms (RT_FLAG_STATS) of a scene as JSON.  Usage: trav_hist.py scene:w:h:spp ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi

for a in sys.argv[1:] or ["final:1920:1080:4"]:
    name, w, h, spp = a.split(":")
    sc = capi.Scene(name)
    ctx = capi.Context(0)
    ctx.upload(sc)
    ctx.render(int(w), int(h), int(spp), max_depth=sc.depth, seed=1, stats=True)
    st = ctx.stats()
    print(json.dumps({"scene": name, **{k: st[k] for k in st if k not in ("trav_hist",)}, "first": st["trav_hist"][0:64],
                      "between": st["trav_hist"][64:128], "tail": st["trav_hist"][128:192], "leaf_visits": st["trav_hist"][192:208]}))
    ctx.close()
