#!/bin/bash
# width_ab.py once with the default library and once per library variant in build/variants (tuning experiments)
export AB_W=${AB_W:-1920} AB_H=${AB_H:-1080} AB_SPP=${AB_SPP:-16} AB_WIDTHS=${AB_WIDTHS:-2}
SCENES=${SCENES:-"final mesh book1 cornell cornell_smoke"}
echo "== lib=default"; python tools/width_ab.py $SCENES
for lib in build/variants/*.so; do echo "== lib=$lib"; RT_B200_LIB=$PWD/$lib python tools/width_ab.py $SCENES; done
