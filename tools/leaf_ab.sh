#!/bin/bash
# device-built tree with leaves of at most 1/2/3 primitives (build/variants/librt_leaf*.so) vs the default 4
for lib in "" build/variants/librt_leaf1.so build/variants/librt_leaf2.so build/variants/librt_leaf3.so; do
echo "== ${lib:-default (4)}"
RT_B200_LIB=${lib:+$PWD/$lib} RT_B200_BVH=device python tools/perf_sweep.py v2 final:1920:1080:16 mesh:1920:1080:8 book1:800:450:16 cornell:600:600:32 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'])
"
done
