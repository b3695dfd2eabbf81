"""Parity of the benchmarked configuration against the golden oracle closed loops for several solver settings.
usage: python tools/parity_sweep.py"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
golden = np.load(os.path.join(bench.ROOT, "tests", "golden", "oracle_bench_config.npz"))
for so in ({}, {"eps_abs": 1e-6, "eps_rel": 1e-6}, {"eps_abs": 1e-7, "eps_rel": 1e-7}, {"eps_abs": 1e-8, "eps_rel": 1e-8},
           {"eps_abs": 1e-8, "eps_rel": 1e-8, "rho": 1e-4}, {"eps_abs": 1e-8, "eps_rel": 1e-8, "rho": 1e-6}, {"rho": 1e-5}, {"rho": 1e-6, "check_every": 1}):
    kw = dict(bench.GEN_KW); kw["solver_opts"] = so
    p = bench.parity_vs_golden(golden, kw)
    print(json.dumps(so), "| X %.2e U %.2e | vs osqp X %.2e U %.2e | iters %.2f ok %s" % (p["max_abs_err_X"], p["max_abs_err_U"], p["max_abs_err_X_vs_osqp"], p["max_abs_err_U_vs_osqp"], p["mean_admm_iters"], p["all_steps_accepted"]), flush=True)
# where does the default-settings error sit?
import trajectory_generation_b200 as tg
n, T = int(golden["n_traj"]), int(golden["T"])
x0, u0, sc = bench.workload_from_golden(golden)
gen = tg.ClosedLoopGenerator(**bench.GEN_KW)
res = gen.generate(x0, u0, sc, T)
eX = np.abs(res["clean"] - golden["X_ipm"]); eU = np.abs(res["U"] - golden["U_ipm"])
print("per state max err", eX.max((0, 1)).round(6), "per input", eU.max((0, 1)).round(6))
print("max err over time (t=10,100,300,600,1200):", [float(eX[:, :t + 1].max().round(6)) for t in (10, 100, 300, 600, 1200)])
print("worst trajectories:", np.argsort(-eX.max((1, 2)))[:5], np.sort(eX.max((1, 2)))[::-1][:5].round(5))
