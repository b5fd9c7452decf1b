"""Top SASS instructions of an ncu report by a stall reason, with the instructions before them.
usage: ncu -i rep --page source --print-source sass --csv > sass.csv; python scripts/ncu_sass_hot.py sass.csv [stall_long_sb] [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
key = sys.argv[2] if len(sys.argv) > 2 else "stall_long_sb"
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
hdr = rows[1]
sass = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(d["# Samples"]) for d in sass)
print("instructions", len(sass), "samples", tot, "share of", key, "%.1f%%" % (100 * sum(int(d[key]) for d in sass) / tot))
for i in sorted(range(len(sass)), key=lambda i: -int(sass[i][key]))[:top_n]:
    print("---- %.2f%% of samples in %s" % (100 * int(sass[i][key]) / tot, key))
    for j in range(max(0, i - 8), i + 1):
        print("   %s  %-64s smp %5.2f%% lanes %s" % (sass[j]["Address"][-5:], sass[j]["Source"].strip()[:64],
                                                 100 * int(sass[j]["# Samples"]) / tot, sass[j]["Avg. Threads Executed"]))
