#!/bin/bash
# End-of-round check on a GPU box (gpurun -- 'bash scripts/final_check.sh'): GPU test suite, smoke, both bench arms,
# ncu launch list of the bench command.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/final_pytest_gpu.log
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 ) > gpurun_out/final_smoke.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/final_ncu_launches.log 2>&1
# full capture of the bench launch: ~40 replay passes of a 3.9 s kernel, each restoring 1.5 GB -- needs ~20 min; opt-in
if [ -n "$FULL_NCU" ]; then timeout 1800 ncu --set full --clock-control none --import-source on -k regex:help_sim_kernel -c 1 -f -o gpurun_out/prof_r1_bench_final python bench.py --steps 1 --warmup 0 > gpurun_out/final_ncu_full.log 2>&1; fi
cat gpurun_out/final_pytest_gpu.log gpurun_out/final_smoke.log gpurun_out/final_bench.json
