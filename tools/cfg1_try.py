"""dev: config 1 latency (B = 1, N = 40, 600 steps) and mid-horizon throughput without state rows"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import trajectory_generation_b200 as tg
gen = tg.ClosedLoopGenerator(N=40, Ts=0.02)
x0 = np.array([[0, 0.5, 0, 1.0, 0, 0.0]]); u0 = np.array([[tg.d_steady_state(1.0), 0.0]])
gen.generate(x0, u0, tg.Scenarios(1), 5)
best = 1e9
for _ in range(5):
    t = time.perf_counter(); res = gen.generate(x0, u0, tg.Scenarios(1), 600); best = min(best, time.perf_counter() - t)
print(f"config 1: {best*1e3:.1f} ms = {best/600*1e6:.1f} us/step, iters/step {res['iters_total'][0]/600:.2f}, d mean {res['U'][0,:,0].mean():.4f}")
for N, B in ((40, 1184), (30, 2368), (50, 592)):
    kw = dict(bench.GEN_KW); kw["N"] = N
    g = tg.ClosedLoopGenerator(**kw)
    x0b, u0b, sc = bench.make_workload(g, B)
    buf = g.alloc_result(B, 300)                      # page-locked once; best of 4 launches
    g.generate(x0b, u0b, sc, 300, out=buf)
    dt = 1e9
    for _ in range(4):
        t = time.perf_counter(); r = g.generate(x0b, u0b, sc, 300, out=buf); dt = min(dt, time.perf_counter() - t)
    print(f"N={N} B={B} T=300 (bench workload): {B*300/dt:.3e} steps/s, iters/step {r['iters_total'].sum()/(B*300):.2f}, {g.info()}")
