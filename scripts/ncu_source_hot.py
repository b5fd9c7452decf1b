"""Top source lines of an ncu report by stall samples / instructions executed.
usage: ncu -i rep --page source --print-source cuda,sass --csv > src.csv; python scripts/ncu_source_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; out = []; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        try:
            out.append((cur, int(r[0]), r[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])))
        except ValueError:
            pass
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print("total samples", ts, "total warp insts", ti)
byfile = {}
for o in out: byfile.setdefault(o[0], [0, 0]); byfile[o[0]][0] += o[3]; byfile[o[0]][1] += o[4]
for f, (s, i) in byfile.items(): print("  %-22s samples %5.1f%%  insts %5.1f%%" % (f, 100 * s / ts, 100 * i / ti))
print("-- by samples")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print("%5.2f%% smp %5.2f%% inst lanes %4.1f  %s:%d  %s" % (100 * o[3] / ts, 100 * o[4] / ti, o[5] / max(o[4], 1), o[0], o[1], o[2]))
