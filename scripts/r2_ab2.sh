#!/bin/bash
# Round-2 A/B pass (gpurun -- 'bash scripts/r2_ab2.sh'): GPU tests, prefetch / block-order / day-loop variants in build_ab/, per-class cost calibration
mkdir -p gpurun_out
L=gpurun_out/r2_ab2.log; : > $L
( timeout 600 python -m pytest tests/test_gpu_pack.py -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2_pytest_pack.log
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2_pytest5.log
for f in base pftd1 pftd2 pfbu pfboth rolled nosync rev base; do
  echo "== nat 100000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup" | tail -2 >> $L
done
for f in base rev pfboth; do
  echo "== mixed 125000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup" | tail -2 >> $L
  echo "== landfill 31250 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/build_ab/lib_$f.so timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup" | tail -2 >> $L
done
echo "== class cost" >> $L
HELP_B200_LIB=$PWD/build_ab/lib_base.so timeout 600 python scripts/class_cost.py 100000 3 2>&1 | grep "^class" >> $L
cat gpurun_out/r2_pytest_pack.log | tail -12; cat gpurun_out/r2_pytest5.log | tail -3; cat $L
