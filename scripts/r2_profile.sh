#!/bin/bash
# Round-2 measurement pass on a GPU box (gpurun -- 'bash scripts/r2_profile.sh'): GPU tests, bench line, launch-shape A/B,
# ncu launch list of the bench command, ncu --set full captures (3-simulated-year launches of the same kernels) summarised
# on the box, DRAM bytes of the real bench launch.
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_pytest_final.log
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1b.json 2> gpurun_out/r2_bench1b.err
for S in 0 1 2 3; do echo "== shape $S"; HELP_B200_SHAPE=$S timeout 200 python scripts/dev_bench.py 100000 5 2>&1 | grep simulate | tail -1; done > gpurun_out/r2_shape_ab.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_ncu_launches.log 2>&1
for W in "nat 100000 2 1000 natural3" "mixed 125000 2 1250 mixed" "landfill 31250 2 312 landfill"; do
  set -- $W
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:help_sim_kernel -c 1 -f -o gpurun_out/prof_r2_$1 python scripts/dev_bench.py $2 $3 $4 $5 > gpurun_out/r2_ncu_full_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_r2_$1.ncu-rep r2_$1 $( [ $1 = nat ] && echo --limits ) > /dev/null 2>&1
  ls -la gpurun_out/prof_r2_$1.ncu-rep
done
cp profiles/r2_*_metrics.txt profiles/r2_*_hot_lines.txt profiles/r2_*_limits.json gpurun_out/ 2>/dev/null
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:help_sim_kernel -c 1 --csv --log-file gpurun_out/r2_bench_kernel_dram.csv python scripts/dev_bench.py 100000 30 > gpurun_out/r2_ncu_dram.log 2>&1
du -sh gpurun_out; rm -f gpurun_out/prof_r2_mixed.ncu-rep gpurun_out/prof_r2_landfill.ncu-rep
tail -3 gpurun_out/r2_pytest_final.log; cat gpurun_out/r2_shape_ab.log; head -c 600 gpurun_out/r2_bench1b.json
