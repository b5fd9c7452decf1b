#!/bin/bash
# Round-2 A/B (gpurun): relaxed day barrier (mbarrier, one day of slack) against the strict one
mkdir -p gpurun_out
L=gpurun_out/r2_ab16.log; : > $L
for f in base mbar base mbar; do
  echo "== nat 100000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 200 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
for f in base mbar; do
  echo "== example size 17178 x 11yr $f" >> $L
  HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 200 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep "^setup" | tail -1 >> $L
  echo "== nat 100000 x 5yr block 256 $f" >> $L
  HELP_B200_BLOCK=256 HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 200 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup" | tail -1 >> $L
done
cat $L
