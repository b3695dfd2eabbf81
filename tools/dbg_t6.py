import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import trajectory_generation_b200 as tg
g = np.load(os.path.join(bench.ROOT, "tests", "golden", "oracle_bench_config.npz"))
x0, u0, sc = bench.make_workload(1)
gen = tg.ClosedLoopGenerator(**bench.GEN_KW)
ctl = tg.BatchedMPC(N=20, Ts=0.01)
X, U = g["X_ipm"][0], g["U_ipm"][0]
out = {}
for t in (5, 6, 7):
    pr, vr = gen.ref_window(X[t:t + 1], sc, t_index=t)
    r = ctl.step(X[t:t + 1], U[t - 1:t], pr, vr)
    a = ctl.assemble(X[t:t + 1], U[t - 1:t], pr, vr)
    A, Bm, gg, xbar = ctl.linearize(X[t:t + 1], U[t - 1:t])
    print(t, "iters", r["iters"], "status", r["status"], "u_cmd", r["u_cmd"][0], "oracle", U[t], "err", np.abs(r["u_cmd"][0] - U[t]).max())
    out.update({f"H{t}": a["H"][0], f"q{t}": a["q"][0], f"U{t}": r["U_opt"][0], f"pr{t}": pr[0], f"vr{t}": vr[0], f"A{t}": A[0], f"B{t}": Bm[0], f"xbar{t}": xbar[0]})
np.savez("gpurun_out/dbg_t6.npz", **out)
