import sys
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
from oracle import mpc as ompc, refgen as R, dynamics as dyn
HARD = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)), x_lo=[-1e20] * 4 + [-0.15, -2.0], x_hi=[1e20] * 4 + [0.15, 2.0])
for N in (10, 20, 50):
    T, Ts, B = 12, 0.02, 48
    rng = np.random.default_rng(N)
    x0 = np.zeros((B, 6)); x0[:, 1] = rng.uniform(-1.5, 1.5, B); x0[:, 3] = rng.uniform(0.8, 1.2, B); x0[:2, 1] = (1.5, -1.2)
    u0 = np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1)
    sc = tg.Scenarios(B); sc.set_sine(slice(0, B), 0.5, 0.5, 0.0, 0.0)
    gen = tg.ClosedLoopGenerator(N=N, Ts=Ts, solver_opts={"eps_abs": 1e-6, "eps_rel": 1e-6}, **HARD)
    res = gen.generate(x0, u0, sc, T)
    print("N", N, "status totals", res["status_counts"].sum(0), "iters/step mean", res["iters_total"].mean() / T, "max", res["iters_total"].max() / T)
    bad = np.where(res["status_counts"][:, 3:].sum(1) > 0)[0]
    print("  trajectories with user_limit/NaN:", bad, res["status_counts"][bad])
    if N == 20 and len(bad):
        ctl = tg.BatchedMPC(N=N, Ts=Ts, solver_opts={"eps_abs": 1e-6, "eps_rel": 1e-6}, **HARD)
        i = int(bad[0]); x = x0[i].copy(); up = u0[i].copy()
        for t in range(T):
            v = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, Ts); pr = R.ref_window(x[0], N, Ts, v, R.PATH_SINE, (0.5, 0.5, 0, 0))
            out = ctl.step(x[None], up[None], pr[None], v[None])
            u1, s1, i1 = ompc.mpc_step(x, up, pr, Ts=Ts, N=N, vref=v, solver="ipm", **HARD)
            print("  ", i, t, "gpu(cold)", tg.STATUS_STRINGS[out["status"][0]], out["iters"][0], "| ipm", s1, "| vy,om %.4f %.4f" % (x[4], x[5]))
            ucmd = u1 if s1 == "optimal" else up
            x = dyn.plant_step(x, ucmd, Ts); up = ucmd
