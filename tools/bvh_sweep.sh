#!/bin/bash
for ml in 1 2 4 8; do for tc in 0.5 1 2; do
  echo "== max_leaf $ml trav_cost $tc"
  RT_B200_MAX_LEAF=$ml RT_B200_TRAV_COST=$tc timeout 120 python tools/perf_sweep.py v2 "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(' ', r['scene'], r['v2']['msamples_s'])
"
done; done
