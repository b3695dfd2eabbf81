"""Small driver for ncu: one fused closed-loop launch (config-2 workload, B x T reduced) + one batched step."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import trajectory_generation_b200 as tg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
N = int(sys.argv[3]) if len(sys.argv) > 3 else 20
kw = dict(bench.GEN_KW); kw["N"] = N
gen = tg.ClosedLoopGenerator(**kw)
x0, u0, sc = bench.make_workload(gen, B)
for _ in range(2):
    res = gen.generate(x0, u0, sc, T)
print("iters/step", res["iters_total"].sum() / (B * T), "status", res["status_counts"].sum(0))
pr, vr = gen.ref_window(x0, sc)
out = gen.step(x0, u0, pr, vr, want_trajectory=False)
print("step iters", out["iters"].mean())
