"""dev: which rate-limited N = 50 closed loops need thousands of ADMM iterations"""
import sys
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
N, Ts, B = 50, 0.02, 256
RATE = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)))
rng = np.random.default_rng(4)
x0 = np.zeros((B, 6)); x0[:, 1] = rng.uniform(-1.5, 1.5, B); x0[:, 3] = rng.uniform(0.8, 1.2, B)
u0 = np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1)
sc = tg.Scenarios(B); sc.set_sine(slice(0, B), 0.5, 0.5, 0.0, 0.0)
gen = tg.ClosedLoopGenerator(N=N, Ts=Ts, **RATE)
prev = np.zeros(B)
for T in (1, 2, 3, 4, 5, 8):
    res = gen.generate(x0, u0, sc, T)
    it = res["iters_total"].astype(float)
    step_it = it - prev; prev = it
    print(f"T={T}: iterations of the last step(s): mean {step_it.mean():.0f} p50 {np.median(step_it):.0f} p90 {np.percentile(step_it, 90):.0f}; corr with vx0 {np.corrcoef(step_it, x0[:, 3])[0, 1]:.2f}, |y0| {np.corrcoef(step_it, np.abs(x0[:, 1]))[0, 1]:.2f}")
    w = np.argsort(step_it)[-3:]
    for i in w:
        print(f"    traj {i}: y0 {x0[i,1]:.3f} vx0 {x0[i,3]:.3f} iters {step_it[i]:.0f} state {res['clean'][i, T]} u {res['U'][i, T-1]} statuses {res['status_counts'][i]}")
# single trajectory of the step-by-step debug (tools/n50_step_debug.py)
x1 = np.array([[0, 1.5, 0, 1.0, 0, 0.]]); u1 = np.array([[tg.d_steady_state(1.0), 0.0]])
r = gen.generate(x1, u1, tg.Scenarios(1).slice(0, 1) if False else sc.slice(0, 1), 6)
print("single (1.5, 1.0) closed loop of 6 steps: iterations", r["iters_total"], r["status_counts"])
