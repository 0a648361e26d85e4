"""Small renders of every scene for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi
ctx = capi.Context(0)
for name in capi.scene_names():
    sc = capi.Scene(name)
    ctx.upload(sc)
    ctx.aov(64, 36)
    ctx.render(64, 36, 2, max_depth=min(sc.depth, 12), seed=3, stats=True)
    ctx.render(64, 36, 2, max_depth=min(sc.depth, 12), seed=3, shard_rank=1, shard_count=2, shard_mode=2, accumulate=True)
    img = ctx.download(2, linear=True, rgb8=True)
    print(name, "ok", float(img[0].mean()), flush=True)
ctx.close()
print("done")
