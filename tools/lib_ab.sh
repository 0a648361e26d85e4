#!/bin/bash
# A/B of build/variants/*.so against the in-tree library on the cases given (scene:w:h:spp ...), twice
for i in 1 2; do
for lib in "" build/variants/*.so; do
echo "== ${lib:-in-tree}"
RT_B200_LIB=${lib:+$PWD/$lib} python tools/perf_sweep.py v2 "$@" 2>&1 | python -c "
import sys, json
print(' '.join('%s %.1f' % (r['scene'], r['v2']['msamples_s']) for r in (json.loads(l) for l in sys.stdin if l.startswith('{'))))
"
done
done
