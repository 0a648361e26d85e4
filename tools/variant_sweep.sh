#!/bin/bash
# runs tools/perf_sweep.py once per library variant in build/variants (tuning experiments)
for lib in build/variants/*.so; do
  echo "== $lib"
  RT_B200_LIB=$PWD/$lib timeout 120 python tools/perf_sweep.py v2 "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'], 'Msamples/s regs', r['v2']['regs'], 'blocks', r['v2']['blocks'])
"
done
