#!/bin/bash
# developer A/B: time every build_ab/lib_*.so (or the default build) at the three wave shapes
for f in ${LIBS:-pyhelp_b200/libhelp_b200.so}; do
  for cfg in ${CFGS:-"94720 1" "100000 2" "125000 0"}; do
    set -- $cfg
    echo "== $f n=$1 shape=$2: $(HELP_B200_LIB=$PWD/$f HELP_B200_SHAPE=$2 timeout 300 python scripts/dev_bench.py $1 ${NY:-2} 2>&1 | grep -E "^setup" | tail -1)"
  done
done
