"""Statistics of the benchmarked workload (bench.make_workload): how often the applied input sits on a bound, ADMM iterations
per step, status counts.  usage: python tools/workload_stats.py [B] [T]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import trajectory_generation_b200 as tg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
gen = tg.ClosedLoopGenerator(**bench.GEN_KW)
x0, u0, sc = bench.make_workload(gen, B)
gen.generate(x0[:8], u0[:8], sc.slice(0, 8), 10)
t0 = time.perf_counter(); res = gen.generate(x0, u0, sc, T); dt = time.perf_counter() - t0
U = res["U"]
dU = np.diff(np.concatenate([u0[:, None, :], U], axis=1), axis=1)
print(f"B {B} T {T}: {B*T/dt:.3e} steps/s through generate(); mean iters/step {res['iters_total'].sum()/(B*T):.3f}")
print("status counts", dict(zip(tg.STATUS_STRINGS, res["status_counts"].sum(0).tolist())))
print("per-trajectory iters/step percentiles (50,90,99,max):", np.percentile(res["iters_total"] / T, [50, 90, 99, 100]).round(2))
print("u0 on box: d %.4f delta %.4f of steps" % ((np.abs(U[..., 0]) >= 1 - 1e-7).mean(), (np.abs(U[..., 1]) >= 0.6 - 1e-7).mean()))
print("du0 on rate limit: d %.4f delta %.4f of steps" % ((np.abs(dU[..., 0]) >= 0.5 - 1e-7).mean(), (np.abs(dU[..., 1]) >= 0.3 - 1e-7).mean()))
first = (np.abs(U[..., 1]) >= 0.6 - 1e-7) | (np.abs(dU[..., 1]) >= 0.3 - 1e-7) | (np.abs(U[..., 0]) >= 1 - 1e-7) | (np.abs(dU[..., 0]) >= 0.5 - 1e-7)
print("steps with the applied input on some bound: %.4f; by step index (first 10 / 10-100 / rest): %.3f %.3f %.4f" % (first.mean(), first[:, :10].mean(), first[:, 10:100].mean(), first[:, 100:].mean()))
