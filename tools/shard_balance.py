#!/usr/bin/env python
"""What limits 1 -> N scaling, measured on ONE GPU: device time of every shard of an N-way tile-sharded frame
(compact buffers, as the multi-GPU paths render them) against 1/N of the unsharded frame.  The slowest shard is
what an N-GPU step costs (plus the gather).  usage: shard_balance.py [N ...]   (env AB_SCENE, AB_W, AB_H, AB_SPP)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracingoneweekendapplication_b200 import capi  # noqa: E402

W, H, SPP = int(os.environ.get("AB_W", 3840)), int(os.environ.get("AB_H", 2160)), int(os.environ.get("AB_SPP", 64))
sc = capi.Scene(os.environ.get("AB_SCENE", "final"))
c = capi.Context(0)
c.upload(sc)


def timed(**kw):
    best = 1e30
    for _ in range(3):
        c.render(W, H, SPP, max_depth=sc.depth, seed=1, **kw)
        best = min(best, c.stats()["render_ms"])
    return best


c.render(W, H, SPP, max_depth=sc.depth, seed=1)
whole = timed()
print(json.dumps({"shards": 1, "ms": round(whole, 3)}), flush=True)
for n in [int(a) for a in sys.argv[1:]] or [2, 4, 8]:
    ms = [timed(shard_rank=r, shard_count=n, shard_mode=capi.RT_SHARD_TILES, compact=True) for r in range(n)]
    print(json.dumps({"shards": n, "ideal_ms": round(whole / n, 3), "slowest_ms": round(max(ms), 3), "fastest_ms": round(min(ms), 3),
                      "mean_ms": round(sum(ms) / n, 3), "efficiency_bound": round(whole / n / max(ms), 4),
                      "imbalance": round(max(ms) / (sum(ms) / n) - 1, 4), "per_shard_overhead": round(sum(ms) / whole - 1, 4)}), flush=True)
c.close()
