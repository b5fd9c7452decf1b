#!/bin/bash
# Round-2 A/B (gpurun): 255-register instantiations for batches that leave most warp slots empty
mkdir -p gpurun_out
L=gpurun_out/r2_ab9.log; : > $L
for S in 0 4; do
  echo "== example size 17178 x 11yr shape $S" >> $L
  HELP_B200_SHAPE=$S timeout 300 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
echo "== landfill 31250 x 5yr 128 regs" >> $L
timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
echo "== landfill 31250 x 5yr 255 regs" >> $L
HELP_B200_BIGREG=1 timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
echo "== landfill 31250 x 5yr 255 regs block 256" >> $L
HELP_B200_BLOCK=256 HELP_B200_BIGREG=1 timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
cat $L
