#!/bin/bash
# host builder: from how many primitives on the build-time shortcuts apply (default 1024)
for m in 1024 4096 100000; do
echo "== RT_B200_SHORTCUT_MIN=$m"
RT_B200_SHORTCUT_MIN=$m python tools/perf_sweep.py v2 final:1920:1080:16 mesh:1920:1080:8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'])
"
done
RT_B200_SHORTCUT_MIN=100000 python tools/tree_quality.py final:1920:1080:2 mesh:1920:1080:2 2>&1 | grep '"host"'
