"""aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/launch_summary.py x.csv [out.txt] [what]"""
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, total = {}, 0.0
for r in rows[1:]:
    ns = float(r[vi].replace(",", ""))
    name = r[ki].split("(")[0]
    total += ns
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
out.write("# ncu --metrics gpu__time_duration.sum --clock-control none  %s\n" % (sys.argv[3] if len(sys.argv) > 3 else ""))
out.write("# (cold-cache, serialised times: compare SHARES)\n")
out.write("%-70s %6s %12s %9s %7s\n" % ("kernel", "count", "total_us", "avg_us", "share"))
for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.write("%-70s %6d %12.1f %9.2f %6.1f%%\n" % (name[:70], c, ns / 1e3, ns / 1e3 / c, 100 * ns / total))
out.write("total %.1f us\n" % (total / 1e3))
