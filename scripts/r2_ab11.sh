#!/bin/bash
# Round-2 A/B (gpurun): paired powers in the sweep (capillary limit next to HIGH / LOW) and re-use of LOW at the lower limit
mkdir -p gpurun_out
L=gpurun_out/r2_ab12.log; : > $L
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py tests/test_daily.py tests/test_abi.py tests/test_help_math.py -m gpu -x -q 2>&1 | tail -3 ) >> $L
for f in build_ab/lib_base.so pyhelp_b200/libhelp_b200.so build_ab/lib_base.so pyhelp_b200/libhelp_b200.so; do
  echo "== nat 100000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
for f in build_ab/lib_base.so pyhelp_b200/libhelp_b200.so; do
  echo "== mixed 125000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
  echo "== landfill 31250 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
  echo "== example size 17178 x 11yr $f" >> $L
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
cat $L
