import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
import bench, trajectory_generation_b200 as tg
g = np.load("tests/golden/oracle_bench_config.npz")
n, T = int(g["n_traj"]), int(g["T"])
x0, u0, sc = bench.workload_from_golden(g)
for tight in (False, True):
    kw = dict(bench.GEN_KW); kw["jacobian"] = tg.JAC_FD if hasattr(tg, "JAC_FD") else 1
    if tight: kw["solver_opts"] = dict(eps_abs=1e-7, eps_rel=1e-7)
    gen = tg.ClosedLoopGenerator(**kw)
    res = gen.generate(x0, u0, sc, T)
    dX = np.abs(res["clean"] - g["X_ipm"]).max(axis=(1, 2)); dU = np.abs(res["U"] - g["U_ipm"]).max(axis=(1, 2))
    print("FD mode tight", tight, "max dX", dX.max(), "dU", dU.max(), "iters", res["iters_total"].sum() / (n * T))
    print(np.round(np.log10(dX + 1e-300), 1)); print(np.round(np.log10(dU + 1e-300), 1))
    print(res["status_counts"].sum(0))
    j = g["fd_jump"]
    for b in np.nonzero(j.any(1))[0]:
        t = np.nonzero(j[b])[0]
        print(b, t, np.abs(res["U"][b, t[0]:t[-1] + 3] - g["U_ipm"][b, t[0]:t[-1] + 3]).max(1))
    gen.close()
