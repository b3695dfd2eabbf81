import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import trajectory_generation_b200 as tg
g = np.load(os.path.join(bench.ROOT, "tests", "golden", "oracle_bench_config.npz"))
i0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
x0, u0, sc = bench.make_workload(i0 + 1)
gen = tg.ClosedLoopGenerator(**bench.GEN_KW)
prev = 0
for T in range(1, 14):
    r = gen.generate(x0[i0:], u0[i0:], sc.slice(i0, i0 + 1), T)
    tot = int(r["iters_total"][0])
    print(f"step {T-1}: iters {tot - prev:4d}  U {r['U'][0, T-1]}  oracle {g['U_ipm'][i0, T-1]}  err {np.abs(r['U'][0, T-1] - g['U_ipm'][i0, T-1]).max():.2e}")
    prev = tot
