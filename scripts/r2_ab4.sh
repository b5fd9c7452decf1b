#!/bin/bash
# Round-2 A/B (gpurun): block size for batches with liner profiles
mkdir -p gpurun_out
L=gpurun_out/r2_ab4.log; : > $L
export HELP_B200_LIB=$PWD/build_ab/lib_rev2.so
for B in 160 192 224 256; do
  echo "== landfill 31250 x 5yr block $B" >> $L
  HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup" | tail -1 >> $L
done
for B in 384 512; do
  echo "== mixed 125000 x 5yr block $B" >> $L
  HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup" | tail -1 >> $L
done
for B in 128 256 512; do
  echo "== classes alone, 75000 cells x 3 yr, block $B" >> $L
  HELP_B200_BLOCK=$B timeout 600 python scripts/class_cost.py 75000 3 2>&1 | grep "^class" >> $L
done
echo "== classes alone, 37500 cells x 3 yr, block 256" >> $L
HELP_B200_BLOCK=256 timeout 600 python scripts/class_cost.py 37500 3 2>&1 | grep "^class" >> $L
cat $L
