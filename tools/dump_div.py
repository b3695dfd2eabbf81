import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import trajectory_generation_b200 as tg
golden = np.load(os.path.join(bench.ROOT, "tests", "golden", "oracle_bench_config.npz"))
n, T = 32, 40
x0, u0, sc = bench.make_workload(n)
out = {}
for name, kw in (("analytic", {}), ("fd", {"jacobian": tg.JAC_FD}), ("cold", {"warm_start": False})):
    g = dict(bench.GEN_KW); g.update(kw)
    gen = tg.ClosedLoopGenerator(**g)
    res = gen.generate(x0, u0, sc, T)
    out[name + "_X"] = res["clean"]; out[name + "_U"] = res["U"]
    eX = np.abs(res["clean"] - golden["X_ipm"][:, :T + 1]); eU = np.abs(res["U"] - golden["U_ipm"][:, :T])
    print(name, "max err X %.3e U %.3e" % (eX.max(), eU.max()), "first step with U err > 1e-4 (traj 2, 23):",
          [int(np.argmax(eU[i].max(1) > 1e-4)) for i in (2, 23)], [float(eU[i].max()) for i in (2, 23)])
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/div.npz", **out)
