#!/bin/bash
# device builder: node-visit cost of the bottom-up SAH leaf decision (build/variants/librt_tc*.so; default 1.0)
for lib in "" build/variants/librt_tc2.0f.so build/variants/librt_tc3.0f.so build/variants/librt_tc5.0f.so build/variants/librt_tc100.0f.so; do
echo "== ${lib:-default (1.0)}"
RT_B200_LIB=${lib:+$PWD/$lib} RT_B200_BVH=device python tools/perf_sweep.py v2 final:1920:1080:16 mesh:1920:1080:8 book1:800:450:16 cornell:600:600:32 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'])
"
RT_B200_LIB=${lib:+$PWD/$lib} RT_B200_BVH=device python tools/upload_scale.py 1000000 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('terrain 1M', r['msamples_s'], 'upload', r['upload_ms_best'])
"
done
