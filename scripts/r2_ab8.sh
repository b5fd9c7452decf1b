#!/bin/bash
# Round-2 A/B (gpurun): cells per warp for batches that do not fill the SMs
mkdir -p gpurun_out
L=gpurun_out/r2_ab8.log; : > $L
for LN in 32 16 8 4; do
  echo "== example size 17178 x 11yr lanes $LN" >> $L
  HELP_B200_LANES=$LN timeout 300 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
for LN in 32 16 8; do
  echo "== landfill 31250 x 5yr lanes $LN" >> $L
  HELP_B200_LANES=$LN timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
for LN in 32 16; do
  echo "== nat 50000 x 5yr lanes $LN" >> $L
  HELP_B200_LANES=$LN timeout 300 python scripts/dev_bench.py 50000 5 2>&1 | grep "^setup" | tail -1 >> $L
done
for LN in 32 24; do
  echo "== nat 100000 x 5yr lanes $LN" >> $L
  HELP_B200_LANES=$LN timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup" | tail -1 >> $L
done
cat $L
