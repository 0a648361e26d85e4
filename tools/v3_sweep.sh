#!/bin/bash
# v3 tuning variants (build/variants/librt_v3_l<leaf trigger>_t<threshold>.so): throughput and lane occupancy on C5
fmt='
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print("  node iters/ray %.2f working %.1f holding %.1f | leaf iters/ray %.3f lanes %.1f | shade iters/ray %.3f lanes %.1f" % (r["desc_iters"]/r["rays"], r["desc_lanes"]/max(1,r["desc_iters"]), r["desc_trav_lanes"]/max(1,r["desc_iters"]), r["leaf_iters"]/r["rays"], r["leaf_lanes"]/max(1,r["leaf_iters"]), r["shade_iters"]/r["rays"], r["shade_lanes"]/max(1,r["shade_iters"])))
'
for lib in build/variants/librt_v3_*.so; do
  echo "== $lib"
  RT_B200_LIB=$PWD/$lib RT_B200_KERNEL=v3 timeout 120 python tools/perf_sweep.py v3 final:1920:1080:16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v3']['msamples_s'], 'Msamples/s')
"
  RT_B200_LIB=$PWD/$lib RT_B200_KERNEL=v3 timeout 120 python tools/trav_hist.py final:1920:1080:4 2>&1 | python -c "$fmt"
done
echo "== v2 reference"
timeout 120 python tools/trav_hist.py final:1920:1080:4 2>&1 | python -c "$fmt"
