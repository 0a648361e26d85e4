#!/bin/bash
# sweep of the wide-BVH collapse heuristic and of the register budget of the wide kernel instances
# (build/variants/wide_mb2.so = -DRT_WIDE_MIN_BLOCKS=2); prints one JSON row per (library, setting, scene)
export AB_W=${AB_W:-1920} AB_H=${AB_H:-1080} AB_SPP=${AB_SPP:-16}
SCENES=${SCENES:-"final mesh book1"}
for lib in "" build/variants/wide_mb2.so; do
  for setting in "SA=0" "SA=1 NC=2" "SA=1 NC=4" "SA=1 NC=8" "SA=1 NC=16"; do
    eval $setting
    echo "== lib=${lib:-default} $setting"
    env ${lib:+RT_B200_LIB=$PWD/$lib} RT_B200_WIDE_SCALE_AWARE=$SA RT_B200_WIDE_NODE_COST=${NC:-4} AB_WIDTHS=${AB_WIDTHS:-8} python tools/width_ab.py $SCENES
  done
done
