"""profiles/traffic.json (what bench.py reports as roofline.traffic / operative_limits) from the committed captures:
    python scripts/make_traffic_json.py profiles/r2f_bench_kernel_dram.csv profiles/r2f_nat_limits.json
the first is the metrics-only ncu CSV of the real bench launch (dram__bytes_read/write.sum), the second the counters of the
`--set full` capture of the same kernel (scripts/ncu_summary.py --limits)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dram_csv, limits = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(dram_csv, errors="replace")) if len(r) > 10 and "help_sim_kernel" in r[4]]
val = {r[12]: float(r[14]) for r in rows}
d = {"workload": "BASELINE configs[1]: synthetic 100k cells x 30 yr, 3-layer natural-soil profiles, 1000 weather nodes, per GPU",
     "kernel": rows[0][4].replace("void ", "").replace("(SimArgs)", ""),
     "report": "%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... of the bench launch: 100 000 cells x 31 simulated years)" % os.path.relpath(dram_csv, ROOT),
     "dram_bytes_read": val["dram__bytes_read.sum"], "dram_bytes_write": val["dram__bytes_write.sum"],
     "dram_bytes_per_launch": val["dram__bytes_read.sum"] + val["dram__bytes_write.sum"],
     "kernel_ms_under_ncu": val.get("gpu__time_duration.sum", 0.0) / 1e6,
     "operative_limits": json.load(open(limits))}
json.dump(d, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(d, indent=1))
