#!/bin/bash
# how much of the LBVH's render penalty sits below the cut: SAH top levels over clusters of at most N primitives
for c in 64 32 16 8 4; do
echo "== clusters of <= $c"
RT_B200_CLUSTER=$c RT_B200_BVH=device python tools/perf_sweep.py v2 final:1920:1080:16 mesh:1920:1080:8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'])
"
done
