"""print the interesting numbers of bench JSON lines (dev aid)"""
import json
import sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e)
        continue
    r = d.get("roofline") or {}
    pk = r.get("per_kind", {})
    print("%s: value=%.0f e2e=%.0f LOS/s ms/step=%.1f  gemm: ach=%.1f TF (frac %.3f, exec %.3f) share=%.3f | %s | clocks %s" % (
        path, d["value"], d["e2e"]["value"] or 0.0, d["ms_per_step"], r.get("achieved", 0), r.get("frac", 0), r.get("executed_frac", 0),
        r.get("gemm_share_of_step") or 0,
        " ".join("%s=%.2fms x%d" % (k, v["avg_ms"], v["launches"]) for k, v in pk.items()) + " exec_share=%s" % r.get("executed_share_of_dense"), d.get("clocks")))
