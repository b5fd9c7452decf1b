#!/bin/bash
# Round-2 (gpurun): new launch policy (reversed block order, block size by class) -- parity, timings, shard balance at world 8 / 2
mkdir -p gpurun_out
L=gpurun_out/r2_ab5.log; : > $L
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py -m gpu -x -q 2>&1 | tail -3 ) >> $L
echo "== nat 100000 x 5yr" >> $L; timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup" | tail -1 >> $L
echo "== mixed 125000 x 5yr" >> $L; timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup" | tail -1 >> $L
echo "== mixed 125000 x 5yr block 512" >> $L; HELP_B200_BLOCK=512 timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup" | tail -1 >> $L
echo "== landfill 31250 x 5yr" >> $L; timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup" | tail -1 >> $L
for F in 2.1 1.7; do
  echo "== shard balance world 8, liner factor $F" >> $L
  HELP_B200_COST_LINER=$F timeout 900 python scripts/shard_balance.py 8 5 2>&1 | grep "^rank\|^world" >> $L
done
echo "== shard balance world 2, liner factor 1.7" >> $L
HELP_B200_COST_LINER=1.7 timeout 900 python scripts/shard_balance.py 2 5 2>&1 | grep "^rank\|^world" >> $L
cat $L
