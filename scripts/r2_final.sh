#!/bin/bash
# Round-2 final measurement pass on a GPU box (gpurun -- 'bash scripts/r2_final.sh'): GPU tests, both bench arms, ncu launch list
# of the bench command, ncu --set full captures (3-simulated-year launches of the same kernels) summarised on the box, DRAM
# bytes of the real bench launch.  Every ncu pass runs after its command has exited 0 without ncu.
set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2f_pytest.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2f_ncu_launches.log 2>&1
for W in "nat 100000 2 1000 natural3" "mixed 125000 2 1250 mixed" "landfill 31250 2 312 landfill"; do
  set -- $W
  timeout 300 python scripts/dev_bench.py $2 $3 $4 $5 > gpurun_out/r2f_dev_$1.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:help_sim_kernel -c 1 -f -o gpurun_out/prof_r2f_$1 python scripts/dev_bench.py $2 $3 $4 $5 > gpurun_out/r2f_ncu_full_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_r2f_$1.ncu-rep r2f_$1 --limits > /dev/null 2>&1
  ls -la gpurun_out/prof_r2f_$1.ncu-rep
done
cp profiles/r2f_*_metrics.txt profiles/r2f_*_hot_lines.txt profiles/r2f_*_limits.json gpurun_out/ 2>/dev/null
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:help_sim_kernel -c 1 --csv --log-file gpurun_out/r2f_bench_kernel_dram.csv python scripts/dev_bench.py 100000 30 > gpurun_out/r2f_ncu_dram.log 2>&1
du -sh gpurun_out; rm -f gpurun_out/prof_r2f_landfill.ncu-rep
tail -3 gpurun_out/r2f_pytest.log; head -c 400 gpurun_out/r2f_bench.json; echo; cat gpurun_out/r2f_bench_reference.json | head -c 600
