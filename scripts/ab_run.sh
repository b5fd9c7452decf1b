#!/bin/bash
# developer A/B: time every build_ab/lib_*.so on the same workload (args: n_cells n_years)
N=${1:-94720}; NY=${2:-2}
mkdir -p gpurun_out
for f in build_ab/lib_*.so; do
  echo "== $f" 
  HELP_B200_LIB=$PWD/$f timeout 300 python scripts/dev_bench.py $N $NY 2>&1 | grep -E "simulate|yearly mean|Error|error" | tail -3
done
