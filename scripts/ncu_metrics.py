"""Key metrics of an ncu report: python scripts/ncu_metrics.py rep.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_global_loads", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in want or ("smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and float(v or 0) > 0.05):
        print("%-78s %-12s %s" % (h, u, v))
