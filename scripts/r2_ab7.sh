#!/bin/bash
# Round-2 A/B (gpurun): grouped loads in the bottom-up pass, early J+1 loads in the sweep
mkdir -p gpurun_out
L=gpurun_out/r2_ab7.log; : > $L
for f in base bu4 bu8 tde bu4tde base; do
  echo "== nat 100000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
cat $L
