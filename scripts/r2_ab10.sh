#!/bin/bash
# Round-2 A/B (gpurun): shared routing-record slots for bit-identical consecutive segments
mkdir -p gpurun_out
L=gpurun_out/r2_ab10.log; : > $L
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_configs.py tests/test_daily.py -m gpu -x -q 2>&1 | tail -3 ) >> $L
for f in build_ab/lib_base.so pyhelp_b200/libhelp_b200.so build_ab/lib_base.so pyhelp_b200/libhelp_b200.so; do
  echo "== nat 100000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
for f in build_ab/lib_base.so pyhelp_b200/libhelp_b200.so; do
  echo "== mixed 125000 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
  echo "== landfill 31250 x 5yr $f" >> $L
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep "^setup\|yearly mean" | tail -2 >> $L
done
cat $L
