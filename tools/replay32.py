import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import trajectory_generation_b200 as tg
i0 = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x0, u0, sc = bench.make_workload(i0 + 1)
gen = tg.ClosedLoopGenerator(**bench.GEN_KW)
T = 80
res = gen.generate(x0[i0:], u0[i0:], sc.slice(i0, i0 + 1), T)
X, U = res["clean"][0], res["U"][0]
ctl = tg.BatchedMPC(N=20, Ts=0.01)
up = np.concatenate([u0[i0:i0 + 1], U[:-1]], 0)
out = []
for t in range(T):
    pr, vr = gen.ref_window(X[t:t + 1], sc.slice(i0, i0 + 1), t_index=t)
    r = ctl.step(X[t:t + 1], up[t:t + 1], pr, vr)
    out.append((t, int(r["iters"][0]), int(r["status"][0]), X[t, 3], up[t, 0], float(np.abs(r["u_cmd"][0] - U[t]).max())))
    if r["iters"][0] > 100:
        np.savez(f"gpurun_out/qp32_{t}.npz", x=X[t], up=up[t], pr=pr[0], vr=vr[0], U_opt=r["U_opt"][0], iters=r["iters"][0])
for o in out:
    if o[1] > 30 or o[0] < 3: print("t %d iters %d status %d vx %.4f d_prev %.3f |u_step-u_loop| %.2e" % o)
print("total iters closed loop", res["iters_total"])
