#!/bin/bash
# host builder: largest range that still gets the full 3-axis SAH search (default 256)
for m in 256 4096 100000000; do
echo "== RT_B200_ALL_AXES=$m"
RT_B200_ALL_AXES=$m python tools/perf_sweep.py v2 final:1920:1080:16 mesh:1920:1080:8 book1:800:450:16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'])
"
done
