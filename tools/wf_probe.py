import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RT_B200_KERNEL"] = sys.argv[1] if len(sys.argv) > 1 else "wf"
from raytracingoneweekendapplication_b200 import capi
sc = capi.Scene("final")
ctx = capi.Context(0)
ctx.upload(sc)
ctx.render(1920, 1080, 16, max_depth=50, seed=1)
ctx.render(1920, 1080, 16, max_depth=50, seed=1)
print(ctx.stats()["render_ms"], ctx.stats()["kernel_launches"])
