"""PSNR against the reference's converged image as a function of spp, with and without RT_FLAG_NEE."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from raytracingoneweekendapplication_b200 import capi

for name in sys.argv[1:] or ["cornell", "cornell_smoke", "final"]:
    sc = capi.Scene(name)
    g = helpers.golden("image", name)
    ref, depth = g["image"].astype(np.float64), int(g["depth"])
    h, w, _ = ref.shape
    ctx = capi.Context(0)
    ctx.upload(sc)
    for nee in (False, True):
        row = {"scene": name, "nee": nee, "frame": [w, h]}
        for spp in (16, 64, 256, 1024, 4096):
            ctx.render(w, h, spp, max_depth=depth, seed=9, nee=nee)
            img = ctx.download(spp).astype(np.float64)
            row[f"psnr_{spp}"] = round(float(helpers.psnr_after_gamma(img, ref)), 2)
            row[f"ms_{spp}"] = round(ctx.stats()["render_ms"], 2)
            row[f"mean_{spp}"] = [round(float(x), 5) for x in img.mean(axis=(0, 1))]
        row["ref_mean"] = [round(float(x), 5) for x in ref.mean(axis=(0, 1))]
        print(json.dumps(row), flush=True)
    ctx.close()
