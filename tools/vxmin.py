import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import trajectory_generation_b200 as tg
x0, u0, sc = bench.make_workload(1024)
gen = tg.ClosedLoopGenerator(**bench.GEN_KW)
res = gen.generate(x0, u0, sc, 1200)
vx = res["clean"][:, :, 3]
print("min vx per trajectory: count < 0.31:", (vx.min(1) < 0.31).sum(), " count <1e-3:", (vx.min(1) < 1e-3).sum(), "steps with vx<0.31:", (vx < 0.31).sum(), " steps vx<=1e-5:", (vx <= 1e-5).sum())
i = np.argsort(vx.min(1))[:8]; print(i, vx.min(1)[i], (vx[i] < 0.31).sum(1), res["iters_total"][i])
print("d at those:", res["U"][i, :, 0].min(1), res["U"][i, :, 0].mean(1))
