#!/bin/bash
# Round-2 A/B (gpurun): one barrier per sub-step (plain profiles) instead of / next to the one per day
mkdir -p gpurun_out
L=gpurun_out/r2_ab13.log; : > $L
for f in base subsync subsync_noday base; do
  for B in 128 256; do
    echo "== nat 100000 x 5yr $f block $B" >> $L
    HELP_B200_BLOCK=$B HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
  done
done
echo "== example size subsync" >> $L
HELP_B200_LIB=$PWD/build_ab/lib_subsync.so timeout 300 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep "^setup" | tail -1 >> $L
cat $L
