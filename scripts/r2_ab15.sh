#!/bin/bash
# Round-2 A/B (gpurun): block placement that evens out the SMs of a single-wave launch
mkdir -p gpurun_out
L=gpurun_out/r2_ab15.log; : > $L
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 ) >> $L
for BAL in 0 1 0 1; do
  echo "== nat 100000 x 5yr balance $BAL" >> $L
  HELP_B200_BALANCE=$BAL timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
for BAL in 0 1; do
  echo "== nat 142000 x 5yr (64-register shape) balance $BAL" >> $L
  HELP_B200_BALANCE=$BAL timeout 300 python scripts/dev_bench.py 142000 5 2>&1 | grep "^setup" | tail -1 >> $L
  echo "== nat 60000 x 5yr balance $BAL" >> $L
  HELP_B200_BALANCE=$BAL timeout 300 python scripts/dev_bench.py 60000 5 2>&1 | grep "^setup" | tail -1 >> $L
  echo "== class 2 alone 75000 (liner, one block per SM: no effect expected) balance $BAL" >> $L
done
cat $L
