#!/bin/bash
# bench.py (device-timed value only) with the in-tree library and every build/variants/*.so, on the full frame and on
# a 1/8-size frame (what one GPU of eight renders per pass: the pass drain weighs 8x as much there)
for lib in "" build/variants/*.so; do
  for geo in "" "--width 1360 --height 764"; do
    RT_B200_LIB=${lib:+$PWD/$lib} python bench.py --steps 16 --warmup 3 --no-e2e --no-cpu-baseline $geo 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('${lib:-in-tree}', d['config']['width'], d['config']['height'], round(d['value'], 1), 'Msamples/s', round(d['ms_per_step'], 3), 'ms/step', d['device_stats']['blocks'], 'blocks')"
  done
done
