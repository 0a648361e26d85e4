#!/bin/bash
# host builder: SAH bins per axis (build/variants/librt_bins*.so; default 16)
for lib in "" build/variants/librt_bins8.so build/variants/librt_bins32.so build/variants/librt_bins64.so; do
echo "== ${lib:-default (16)}"
RT_B200_LIB=${lib:+$PWD/$lib} python tools/perf_sweep.py v2 final:1920:1080:16 mesh:1920:1080:8 book1:800:450:16 cornell:600:600:32 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'])
"
done
