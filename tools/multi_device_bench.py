#!/usr/bin/env python
"""One process, n GPUs through ONE context (rt_create with n device ids): C5 at 4K, device time of a pass (slowest
device), and the same pass with the RGB8 frame gathered and downloaded.  usage: multi_device_bench.py [n ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from raytracingoneweekendapplication_b200 import capi  # noqa: E402

W, H, SPP = int(os.environ.get("AB_W", 3840)), int(os.environ.get("AB_H", 2160)), int(os.environ.get("AB_SPP", 64))
sc = capi.Scene(os.environ.get("AB_SCENE", "final"))
ref = None
for n in [int(a) for a in sys.argv[1:]] or [1, 2]:
    c = capi.Context(list(range(n)))
    c.upload(sc)
    for _ in range(2):
        c.render(W, H, SPP, max_depth=sc.depth, seed=1)
    ms = []
    for _ in range(3):
        c.render(W, H, SPP, max_depth=sc.depth, seed=1)
        ms.append(c.stats()["render_ms"])
    steps = 6
    img = c.download(SPP, linear=False, rgb8=True)      # first gather: NCCL sets its channels up here
    t0 = time.perf_counter()
    for i in range(steps):
        c.upload(sc)
        c.render(W, H, SPP, max_depth=sc.depth, seed=1)
        img = c.download(SPP, linear=False, rgb8=True)
    e2e_blocking = (time.perf_counter() - t0) / steps * 1e3
    c.render(W, H, SPP, max_depth=sc.depth, seed=1)
    c.download_begin(SPP, linear=False, rgb8=True)
    c.frame_end()
    t0 = time.perf_counter()
    for i in range(steps):
        c.upload(sc)
        c.render(W, H, SPP, max_depth=sc.depth, seed=1, blocking=False)
        c.download_begin(SPP, linear=False, rgb8=True)
        if i > 0:
            img2 = c.frame_end()[1]
    img2 = c.frame_end()[1]
    e2e = (time.perf_counter() - t0) / steps * 1e3
    assert np.array_equal(img, img2)
    if ref is None:
        ref = img
    st = c.stats()
    print(json.dumps({"devices": n, "gather": {0: "-", 1: "peer copies", 2: "nccl"}[st["gather_mode"]], "render_ms": round(min(ms), 3),
                      "msamples_s": round(W * H * SPP / min(ms) / 1e3, 1), "e2e_ms_blocking_download": round(e2e_blocking, 3),
                      "e2e_ms_overlapped_hand_out": round(e2e, 3), "e2e_msamples_s": round(W * H * SPP / e2e / 1e3, 1), "bit_identical_to_first": bool(np.array_equal(img, ref))}), flush=True)
    c.close()
