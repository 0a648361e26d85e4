"""First-light check on a GPU box: primary-hit parity, converged-image statistics against the
golden fixtures, PNGs into gpurun_out/, and a first throughput figure per scene."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from raytracingoneweekendapplication_b200 import capi
import helpers
from raytracingoneweekendapplication_b200.host_png import write_png

out_dir = os.path.join(ROOT, "gpurun_out"); os.makedirs(out_dir, exist_ok=True)
ctx = capi.Context(0)
print("fp32 peak TFLOP/s:", ctx.measure_fp32_peak(), flush=True)
report = {}
scenes = sys.argv[1:] or capi.scene_names()
for name in scenes:
    sc = capi.Scene(name)
    t0 = time.time(); ctx.upload(sc); up = time.time() - t0
    r = {"upload_s": up}
    g = helpers.golden("primary", name)
    if g is not None:
        h, w = g["ids"].shape
        aov = ctx.aov(w, h)
        keys = helpers.flat_leaf_keys(sc.desc)
        cmp_ = helpers.compare_primary(g, aov, keys)
        r["primary"] = cmp_
    gi = helpers.golden("image", name)
    if gi is not None:
        img_ref = gi["image"]; h, w, _ = img_ref.shape
        spp = 4096
        ctx.render(w, h, spp, max_depth=int(gi["depth"]), seed=7)
        lin = ctx.download(spp)
        r["image"] = {"psnr": helpers.psnr_after_gamma(lin, img_ref), "mean_dev": lin.mean(axis=(0, 1)).tolist(),
                      "mean_ref": img_ref.mean(axis=(0, 1)).tolist(), "ms": ctx.stats()["render_ms"]}
    # throughput at the scene's own frame, reduced spp
    spp = max(1, min(sc.spp, 16))
    ctx.render(sc.width, sc.height, spp, max_depth=sc.depth, seed=1)
    ctx.render(sc.width, sc.height, spp, max_depth=sc.depth, seed=1, stats=True)
    st_stats = ctx.stats()
    ctx.render(sc.width, sc.height, spp, max_depth=sc.depth, seed=1)
    st = ctx.stats()
    lin, b8 = ctx.download(spp, linear=True, rgb8=True)
    write_png(os.path.join(out_dir, f"{name}.png"), b8)
    samples = sc.width * sc.height * spp
    r["throughput"] = {"w": sc.width, "h": sc.height, "spp": spp, "ms": st["render_ms"], "msamples_s": samples / st["render_ms"] / 1e3,
                       "rays": st_stats["rays"], "mrays_s": st_stats["rays"] / st["render_ms"] / 1e3,
                       "rays_per_sample": st_stats["rays"] / samples, "nodes_per_ray": st_stats["node_visits"] / max(1, st_stats["rays"]),
                       "stats_ms": st_stats["render_ms"], "regs": st["regs_per_thread"], "blocks": st["blocks"],
                       "nonfinite": st_stats["nonfinite_samples"], "fp64_sphere": st_stats["fp64_sphere_tests"],
                       "bvh_nodes": st["bvh_nodes"], "bvh_depth": st["bvh_depth"]}
    report[name] = r
    print(name, json.dumps(r), flush=True)
json.dump(report, open(os.path.join(out_dir, "first_light.json"), "w"), indent=1)
