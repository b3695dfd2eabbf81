"""Where the benchmarked configuration deviates from the golden oracle loops (per trajectory / time)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import trajectory_generation_b200 as tg
g = np.load(os.path.join(bench.ROOT, "tests", "golden", "oracle_bench_config.npz"))
n, T = int(g["n_traj"]), int(g["T"])
x0, u0, sc = bench.workload_from_golden(g)
kw = dict(bench.GEN_KW)
for a in sys.argv[1:]:
    k, v = a.split("="); kw.setdefault("solver_opts", {})[k] = float(v)
gen = tg.ClosedLoopGenerator(**kw)
res = gen.generate(x0, u0, sc, T)
dX, dU = np.abs(res["clean"] - g["X_ipm"]), np.abs(res["U"] - g["U_ipm"])
jump = g["fd_jump"]
print("iters/step", res["iters_total"].sum() / (n * T))
for i in np.argsort(-dX.max((1, 2)))[:6]:
    t = int(dX[i].max(1).argmax()); tu = int(dU[i].max(1).argmax())
    print(f"traj {i}: max|dX| {dX[i].max():.2e} at row {t} (state {int(dX[i, t].argmax())}), max|dU| {dU[i].max():.2e} at step {tu}; fd_jump steps {np.nonzero(jump[i])[0].tolist()}; "
          f"U_gpu {res['U'][i, tu]} U_ipm {g['U_ipm'][i, tu]} U_osqp {g['U_osqp'][i, tu]}")
    lo = max(tu - 3, 0)
    print("   dU around:", np.round(dU[i, lo:tu + 4].max(1), 6), " d_cmd gpu", np.round(res["U"][i, lo:tu + 4, 0], 5), "ipm", np.round(g["U_ipm"][i, lo:tu + 4, 0], 5))
