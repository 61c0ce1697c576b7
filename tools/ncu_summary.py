"""Summarise ncu outputs into small text/JSON files for profiles/ (dev aid).
usage: ncu_summary.py <launches.csv> <report.ncu-rep> <out_prefix>"""
import csv
import json
import subprocess
import sys

launch_csv, rep, out = sys.argv[1:4]
what = sys.argv[4] if len(sys.argv) > 4 else "python tools/prof_target.py 16384 6 (one fused FISTA call: C3 sampling, 16384 LOS, 6 iterations, bench slab)"
rows = list(csv.reader(l for l in open(launch_csv) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
lines, total, ours = [], 0.0, 0.0
agg = {}
for r in rows[1:]:
    ns = float(r[vi].replace(",", ""))
    name = r[ki].split("(")[0]
    total += ns
    if "csr::" in r[ki]:
        ours += ns
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
with open(out + "_launches.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  %s\n" % what)
    f.write("# (cold-cache, serialised times: compare SHARES)\n")
    f.write("%-60s %6s %12s %7s\n" % ("kernel", "count", "total_us", "share"))
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-60s %6d %12.1f %6.1f%%\n" % (name[:60], c, ns / 1e3, 100 * ns / total))
    f.write("total %.1f us; kernels of libcsromer_b200.so: %.1f%%\n" % (total / 1e3, 100 * ours / total))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "gpc__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
res = []
for r in rr[2:]:
    d = {}
    for w in want:
        if w in h:
            d[w] = r[h.index(w)] + (" " + rr[1][h.index(w)] if rr[1][h.index(w)] else "")
    res.append(d)
json.dump(res, open(out + "_gemm_metrics.json", "w"), indent=1)
for d in res:
    print(d)
