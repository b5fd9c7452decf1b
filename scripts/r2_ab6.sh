#!/bin/bash
# Round-2 (gpurun): launch policy by class + dealt partition -- GPU tests, wide-tier register budgets, shard balance
mkdir -p gpurun_out
L=gpurun_out/r2_ab6.log; : > $L
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) >> $L
for f in "" build_ab/lib_w96.so build_ab/lib_w80.so; do
  [ -n "$f" ] && export HELP_B200_LIB=$PWD/$f
  echo "== mixed 125000 x 5yr lib=$f" >> $L; timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup" | tail -1 >> $L
  echo "== landfill 31250 x 5yr lib=$f" >> $L; timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup" | tail -1 >> $L
done
unset HELP_B200_LIB
echo "== shard balance world 8, dealt" >> $L
timeout 900 python scripts/shard_balance.py 8 5 2>&1 | grep "^rank\|^world" >> $L
echo "== shard balance world 2, dealt" >> $L
timeout 900 python scripts/shard_balance.py 2 5 2>&1 | grep "^rank\|^world" >> $L
cat $L
