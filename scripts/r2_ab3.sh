#!/bin/bash
# Round-2 A/B (gpurun): block size of the simulation kernel (threads that share the day barrier; blocks per SM), reversed block order
mkdir -p gpurun_out
L=gpurun_out/r2_ab3.log; : > $L
export HELP_B200_LIB=$PWD/build_ab/lib_rev2.so
for B in 128 64 96 256 352 384 704 128; do
  echo "== nat 100000 x 5yr block $B" >> $L
  HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup" | tail -1 >> $L
done
for B in 352 512; do
  echo "== nat 100000 x 5yr block $B shape 3 (64 regs)" >> $L
  HELP_B200_SHAPE=3 HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup" | tail -1 >> $L
done
for B in 128 256 512; do
  echo "== mixed 125000 x 5yr block $B" >> $L
  HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup" | tail -1 >> $L
  echo "== landfill 31250 x 5yr block $B" >> $L
  HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup" | tail -1 >> $L
done
for B in 128 64 256; do
  echo "== example size 17178 x 11yr block $B" >> $L
  HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep "^setup" | tail -1 >> $L
done
cat $L
