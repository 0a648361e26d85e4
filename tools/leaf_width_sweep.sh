for ml in 1 2 4; do for w in 2 8; do echo "== max_leaf=$ml width=$w"; RT_B200_MAX_LEAF=$ml AB_WIDTHS=$w AB_SPP=16 python tools/width_ab.py final mesh book1; done; done
