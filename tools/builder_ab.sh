#!/bin/bash
# render throughput with the host SAH tree, the device tree (LBVH + SAH top levels) and the pure LBVH on the five BASELINE configs
for m in host device lbvh; do echo "== RT_B200_BVH=$m"; RT_B200_BVH=$m python tools/perf_sweep.py v2 final:1920:1080:16 cornell:600:600:32 book1:800:450:16 mesh:1920:1080:8 cornell_smoke:600:600:32 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'], 'Msamples/s', r['v2']['ms'], 'ms')
"; done
