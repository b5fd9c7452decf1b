#!/bin/bash
# developer A/B: time every build_ab/lib_*.so (or the default build) at the three wave shapes
for f in ${LIBS:-pyhelp_b200/libhelp_b200.so}; do
  for cfg in "94720 5" "100000 6" "125000 4"; do
    set -- $cfg
    echo "== $f n=$1 minb=$2: $(HELP_B200_LIB=$PWD/$f HELP_B200_MINB=$2 timeout 300 python scripts/dev_bench.py $1 ${NY:-2} 2>&1 | grep -E "^setup" | tail -1)"
  done
done
