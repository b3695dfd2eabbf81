import json, sys
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4g e2e %.4g ms/step %.1f iters %.2f p50lat %.3f flopfrac %.4f cpu %.0f status %s" % (
            d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("mean_admm_iters_per_step", 0), d.get("p50_step_latency_ms") or 0,
            d.get("roofline_flop", {}).get("frac", 0), d.get("cpu_baseline", {}).get("value", 0), list(d.get("status_counts", {}).values())))
    elif l.strip():
        print(l.rstrip()[:300])
