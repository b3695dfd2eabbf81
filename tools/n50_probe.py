"""dev: ADMM iteration counts of rate-limited N = 50 problems: first step (step API, cold) and closed loops with / without warm start"""
import sys
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
from oracle import refgen as R
N, Ts, B = int(sys.argv[1]) if len(sys.argv) > 1 else 50, 0.02, 256
RATE = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)))
rng = np.random.default_rng(4)
x0 = np.zeros((B, 6)); x0[:, 1] = rng.uniform(-1.5, 1.5, B); x0[:, 3] = rng.uniform(0.8, 1.2, B)
u0 = np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1)
sc = tg.Scenarios(B); sc.set_sine(slice(0, B), 0.5, 0.5, 0.0, 0.0)
for warm in (True, False):
    gen = tg.ClosedLoopGenerator(N=N, Ts=Ts, warm_start=warm, **RATE)
    for T in (1, 5, 20, 60):
        res = gen.generate(x0, u0, sc, T)
        it = res["iters_total"] / T
        print(f"N={N} warm={warm} T={T}: iters/step mean {it.mean():.0f} p50 {np.median(it):.0f} max {it.max():.0f}; statuses {res['status_counts'].sum(0).tolist()}")
ctl = tg.BatchedMPC(N=N, Ts=Ts, **RATE)
pr, vr = gen.ref_window(x0, sc)
out = ctl.step(x0, u0, pr, vr)
print("step API cold first step: iters mean", out["iters"].mean(), "max", out["iters"].max(), "status", np.bincount(out["status"], minlength=6))
i = int(np.argmax(out["iters"]))
print("worst", i, x0[i], out["iters"][i])
