"""Reduce an `ncu --set full` report to per-kernel DRAM bytes per launch and a few headline metrics.
usage: python tools/ncu_traffic.py <report.ncu-rep> <out.json> [what]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1:3]
what = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = {"time_us": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l2_hit_pct": "lts__t_sector_hit_rate.pct", "xbar2l1tex_read": "l1tex__m_xbar2l1tex_read_bytes.sum",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "registers": "launch__registers_per_thread",
        "grid": "launch__grid_size", "block": "launch__block_size"}
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
res = []
for r in rows[2:]:
    d = {"kernel": r[h.index("Kernel Name")].split("(")[0]}
    for key, metric in want.items():
        if metric in h:
            i = h.index(metric)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            d[key] = v * mult.get(units[i], 1.0)
    if "dram_read" in d:
        d["dram_bytes"] = d["dram_read"] + d.get("dram_write", 0.0)
    res.append(d)
json.dump({"what": what, "launches": res}, open(out, "w"), indent=1)
for d in res:
    print(d)
