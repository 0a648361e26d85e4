for d in 0 12288 24576 36864; do echo "== dummy smem $d"; RT_B200_DUMMY_SMEM=$d python tools/perf_sweep.py v2 final:1920:1080:16 cornell:600:600:32 mesh:1920:1080:8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['scene'], r['v2']['msamples_s'], 'Msamples/s', r['v2']['ms'], 'ms')
"; done
