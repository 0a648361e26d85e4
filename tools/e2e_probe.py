import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from raytracingoneweekendapplication_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "book1"
sc = capi.Scene(name)
ctx = capi.Context(0)
ctx.upload(sc)
w, h, spp = sc.width, sc.height, min(sc.spp, 50)
frame = np.empty((h, w, 3), np.uint8)
for rep in range(8):
    t0 = time.perf_counter(); ctx.upload(sc)
    t1 = time.perf_counter(); ctx.render(w, h, spp, max_depth=sc.depth, seed=1)
    t2 = time.perf_counter(); ctx.lib.rt_download(ctx._h, spp, None, frame.ctypes.data)
    t3 = time.perf_counter()
    print(name, "upload %.2f render %.2f (device %.2f) download %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1), ctx.stats()["render_ms"], 1e3*(t3-t2)), flush=True)
