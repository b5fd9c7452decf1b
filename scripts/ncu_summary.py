"""Summarise an ncu report of the simulation kernel into profiles/ (tracked, judged).

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep r1_bench [--traffic]

writes profiles/<tag>_metrics.txt (launch shape, duration, pipe utilisation, stall reasons,
memory traffic), profiles/<tag>_hot_lines.txt (top source lines by stall samples) and, with
--traffic, profiles/traffic.json = measured DRAM bytes per launch of help_sim_kernel on the
bench workload (bench.py reads it into roofline.traffic)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
sim = [r for r in body if "help_sim_kernel" in r[hdr.index("Kernel Name")]] or body
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.avg",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_average_branch_targets_threads_uniform.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "sass__inst_executed_global_loads", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
out = ["# ncu --set full --clock-control none, report %s" % os.path.basename(rep)]
for k, r in enumerate(sim):
    out.append("## launch %d" % k)
    for h, u, v in zip(hdr, units, r):
        if h in WANT or ("smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.05):
            out.append("%-84s %-14s %s" % (h, u, v))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", tag + "_metrics.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
if "--limits" in sys.argv:            # the counters that bound the kernel, from this capture alone (no DRAM-byte claim)
    r = sim[-1]
    def num(name):
        try: return float(r[hdr.index(name)])
        except (ValueError, IndexError): return None
    lim = {"source": "profiles/%s_metrics.txt (ncu --set full of %s)" % (tag, os.path.basename(rep)),
           "fp64_pipe_pct_of_peak": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "issue_slots_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "lanes_active_of_32": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "warps_active_per_sm": num("sm__warps_active.avg.per_cycle_active"),
           "l2_hit_pct": num("lts__t_sector_hit_rate.pct"), "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
           "top_stalls_per_issue": {k: num("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k)
                                    for k in ("long_scoreboard", "wait", "no_instruction", "not_selected", "barrier")}}
    json.dump(lim, open(os.path.join(ROOT, "profiles", tag + "_limits.json"), "w"), indent=1)
if "--traffic" in sys.argv:
    r = sim[-1]
    def val(name):
        v, u = float(r[hdr.index(name)]), units[hdr.index(name)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    d = {"workload": "BASELINE configs[1]: synthetic 100k cells x 30 yr, 3-layer natural-soil profiles, 1000 weather nodes, per GPU",
         "kernel": "help_sim_kernel", "report": os.path.basename(rep),
         "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum")}
    d["dram_bytes_per_launch"] = d["dram_bytes_read"] + d["dram_bytes_write"]
    def num(name):
        try: return float(r[hdr.index(name)])
        except (ValueError, IndexError): return None
    # the counters that bound this kernel (bench.py reports them next to the HBM fraction)
    d["operative_limits"] = {
        "source": "%s (ncu --set full)" % os.path.basename(rep),
        "fp64_pipe_pct_of_peak": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_slots_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_active_of_32": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warps_active_per_sm": num("sm__warps_active.avg.per_cycle_active"),
        "l2_hit_pct": num("lts__t_sector_hit_rate.pct"), "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
        "top_stalls_per_issue": {k: num("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k)
                                 for k in ("long_scoreboard", "wait", "no_instruction", "not_selected", "barrier")}}
    old = os.path.join(ROOT, "profiles", "traffic.json")
    json.dump(d, open(old, "w"), indent=1)
    print(d)
# hot source lines
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True).stdout.decode("utf-8", "replace")
cur = None; h2 = None; acc = []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h2 = r; continue
    if h2 and len(r) == len(h2) and r[0] not in ("", "Line No"):
        d = dict(zip(h2, r))
        try:
            acc.append((cur, int(r[0]), r[1].strip()[:80], int(d["# Samples"]), int(d["Instructions Executed"]),
                        int(d["Thread Instructions Executed"]), int(d["stall_long_sb"]), int(d["stall_wait"]), int(d["stall_barrier"])))
        except (ValueError, KeyError):
            pass
def repo_line(fname, line):
    """the text of that line in the committed source (exposes a report whose line numbers belong to another revision)"""
    for d in ("pyhelp_b200/csrc", "include"):
        p = os.path.join(ROOT, d, fname or "")
        if os.path.exists(p):
            L = open(p, errors="replace").read().splitlines()
            return L[line - 1].strip()[:80] if 0 < line <= len(L) else "<beyond end of file>"
    return None
acc = [(o[0], o[1], repo_line(o[0], o[1]) or o[2]) + o[3:] for o in acc]
if acc:
    ts = sum(o[3] for o in acc) or 1; ti = sum(o[4] for o in acc) or 1
    L = ["# per source line: %% of stall samples, %% of warp instructions, active lanes, of which long-scoreboard / fixed-latency wait / barrier (%% of samples)"]
    byf = {}
    for o in acc:
        b = byf.setdefault(o[0], [0, 0, 0]); b[0] += o[3]; b[1] += o[4]; b[2] += o[6]
    for f, b in byf.items():
        L.append("file %-22s samples %5.1f%%  instructions %5.1f%%  long_scoreboard %5.1f%%" % (f, 100 * b[0] / ts, 100 * b[1] / ti, 100 * b[2] / ts))
    for o in sorted(acc, key=lambda o: -o[3])[:40]:
        L.append("%5.2f %5.2f %4.1f | %4.1f %4.1f %4.1f  %s:%d  %s" % (100 * o[3] / ts, 100 * o[4] / ti, o[5] / max(o[4], 1), 100 * o[6] / ts,
                                                                  100 * o[7] / ts, 100 * o[8] / ts, o[0], o[1], o[2]))
    open(os.path.join(ROOT, "profiles", tag + "_hot_lines.txt"), "w").write("\n".join(L) + "\n")
