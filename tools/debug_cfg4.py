import sys
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
from oracle import mpc as ompc, refgen as R, dynamics as dyn
HARD = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)), x_lo=[-1e20] * 4 + [-0.15, -2.0], x_hi=[1e20] * 4 + [0.15, 2.0])
N, T, Ts, B = 20, 12, 0.02, 48
rng = np.random.default_rng(N)
x0 = np.zeros((B, 6)); x0[:, 1] = rng.uniform(-1.5, 1.5, B); x0[:, 3] = rng.uniform(0.8, 1.2, B); x0[:2, 1] = (1.5, -1.2)
u0 = np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1)
sc = tg.Scenarios(B); sc.set_sine(slice(0, B), 0.5, 0.5, 0.0, 0.0)
gen = tg.ClosedLoopGenerator(N=N, Ts=Ts, solver_opts={"eps_abs": 1e-6, "eps_rel": 1e-6}, **HARD)
res = gen.generate(x0, u0, sc, T)
bad = np.where(res["status_counts"][:, 4] > 0)[0]
print("user_limit trajectories", bad, res["status_counts"][bad])
ctl = tg.BatchedMPC(N=N, Ts=Ts, solver_opts={"eps_abs": 1e-6, "eps_rel": 1e-6}, **HARD)
for i in bad[:2]:
    x = x0[i].copy(); up = u0[i].copy()
    for t in range(T):
        v = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, Ts); pr = R.ref_window(x[0], N, Ts, v, R.PATH_SINE, (0.5, 0.5, 0, 0))
        out = ctl.step(x[None], up[None], pr[None], v[None])
        u1, s1, i1 = ompc.mpc_step(x, up, pr, Ts=Ts, N=N, vref=v, solver="ipm", **HARD)
        u2, s2, i2 = ompc.mpc_step(x, up, pr, Ts=Ts, N=N, vref=v, solver="osqp", **HARD)
        print(i, t, "gpu(cold)", tg.STATUS_STRINGS[out["status"][0]], out["iters"][0], "| ipm", s1, "| osqp", s2, i2.get("iters") if i2 else None,
              "| vy,om %.4f %.4f" % (x[4], x[5]), "| du %.2e" % (np.abs(out["u_cmd"][0] - u1).max()))
        ucmd = u1 if s1 == "optimal" else up
        x = dyn.plant_step(x, ucmd, Ts); up = ucmd
print("---- inspect the disputed solve")
i = 9
x = x0[i].copy(); up = u0[i].copy()
for t in range(7):
    v = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, Ts); pr = R.ref_window(x[0], N, Ts, v, R.PATH_SINE, (0.5, 0.5, 0, 0))
    if t == 6:
        out = ctl.step(x[None], up[None], pr[None], v[None])
        asm = ctl.assemble(x[None], up[None], pr[None], v[None])
        dU = (out["U_opt"][0] - up[None, :]).reshape(-1)
        n = 2 * N
        Ac = np.vstack([np.eye(n), np.eye(n) - np.eye(n, k=-2), asm["Gs"][0]])
        lhs = Ac @ dU
        viol = np.maximum(asm["l"][0] - lhs, 0) + np.maximum(lhs - asm["u"][0], 0)
        print("status", out["status"], "iters", out["iters"], "|dU|max", np.abs(dU).max(), "|y|max", np.abs(out["y_opt"]).max(), "max viol", viol.max(), "at row", viol.argmax(), "m", len(viol))
        print("l,u,lhs at worst", asm["l"][0][viol.argmax()], asm["u"][0][viol.argmax()], lhs[viol.argmax()])
        print("state rows l:", asm["l"][0][2*n:2*n+6], "u:", asm["u"][0][2*n:2*n+6])
        print("obj", out["objective"])
    u1, s1, i1 = ompc.mpc_step(x, up, pr, Ts=Ts, N=N, vref=v, solver="ipm", **HARD)
    ucmd = u1 if s1 == "optimal" else up
    x = dyn.plant_step(x, ucmd, Ts); up = ucmd
