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
    pa = r.get("path") or {}
    de = r.get("dense") or {}
    print("%s: value=%.0f e2e=%.0f LOS/s ms/step=%.1f | dominant: ach=%.0f TF frac %.3f | path: %.0f TF alg (%.2f of bf16) gemm share %.3f | dense: %.0f LOS/s frac %.3f exec %.3f | %s | clocks %s" % (
        path, d["value"], d["e2e"]["value"] or 0.0, d["ms_per_step"], r.get("achieved") or 0, r.get("frac") or 0,
        pa.get("algorithmic_tflops") or r.get("achieved") or 0, pa.get("frac_of_bf16_peak") or 0, pa.get("gemm_share_of_step") or r.get("gemm_share_of_step") or 0,
        de.get("value") or 0, de.get("frac") or 0, de.get("executed_frac") or 0,
        " ".join("%s=%.2fms x%d" % (k, v["avg_ms"], v["launches"]) for k, v in pk.items()), d.get("clocks")))
