"""print the essentials of a bench.py JSON line: python tools/show_bench.py gpurun_out/x.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def kernels(rec):
    for k, v in (rec.get("roofline") or {}).get("per_kernel", {}).items():
        extra = ""
        if "executed_frac" in v:
            extra += " exec %.0f TF (%.2f)" % (v["executed"], v["executed_frac"])
        if "hbm" in v:
            extra += " hbm_alg %.0f GB/s (%.3f)" % (v["hbm"]["algorithmic"], v["hbm"]["algorithmic_frac"])
        if v.get("bound") == "hbm" and "algorithmic" in v:
            extra += " alg %.0f GB/s (%.2f)" % (v["algorithmic"], v["algorithmic_frac"])
        if "executed_gbs" in v:
            extra += " ncu %.0f GB/s (%.2f)" % (v["executed_gbs"], v["executed_hbm_frac"])
        print("    %-28s n=%-5d avg %.4f ms  share %.3f%s" % (k, v["launches"], v["avg_ms"], v["share_of_step"], extra))


print("c3: value %.1f LOS/s  ms/step %.2f  ms/iter %.4f  launches %d  clocks %s" % (
    d["value"], d["ms_per_step"], d["ms_per_iteration"], d["gpu_launches"], d["clocks"]))
if d.get("e2e"):
    print("    e2e %.1f LOS/s  h2d %.1f MB d2h %.1f MB" % (d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"] / 1e6, d["e2e"]["d2h_bytes_per_step"] / 1e6))
print("    roofline: %s frac %.3f (%s)" % (d["roofline"]["kind"], d["roofline"]["frac"], d["roofline"]["unit"]))
kernels(d)
for name, sub in d.get("configs", {}).items():
    print("%s: value %.1f LOS/s  ms/step %.2f  ms/iter %.4f  clocks %s" % (name, sub["value"], sub["ms_per_step"], sub["ms_per_iteration"], sub["clocks"]))
    kernels(sub)
for name, p in d.get("parity", {}).items():
    if p:
        o = p.get("oracle", {})
        sd = p.get("shortcuts_vs_dense", {})
        print("parity %s: bit_equal %s dense %.0f LOS/s oracle %s ok %s %s" % (name, sd.get("bit_equal"), sd.get("dense_los_per_s", 0),
              o.get("parity_max_rel"), p["ok"], {k: v for k, v in p.items() if k.startswith("nufft")}))
if d.get("cpu_baseline"):
    print("cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
print("peaks:", d["peaks"]["measured_in_run"])
