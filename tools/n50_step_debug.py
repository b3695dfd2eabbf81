"""dev: the rate-limited N = 50 problems of a closed loop, one by one through the step API with solver variants, beside the NumPy model"""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import trajectory_generation_b200 as tg
import condensed_model as cm
from oracle import dynamics as dyn, refgen as R, mpc as M
N, Ts = 50, 0.02
kw = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)))
x = np.array([0, 1.5, 0, 1.0, 0, 0.]); u = np.array([R.d_steady_state(1.0), 0.])
variants = {"default": None, "no free mode": {"solver_flags": 1}, "no adaptive rho": {"adaptive_rho": 0}, "rho 1": {"rho": 1.0},
            "alpha_warm=alpha": {"alpha_warm": 1.6}, "check 25": {"check_every": 25}}
ctls = {k: tg.BatchedMPC(N=N, Ts=Ts, solver_opts=v, **kw) for k, v in variants.items()}
for t in range(6):
    v = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, Ts); pr = R.ref_window(x[0], N, Ts, v, R.PATH_SINE, (0.5, 0.5, 0, 0))
    A, B, g, _ = dyn.linearize_horizon(x, u, Ts, N)
    H, q, c0, Ac, l, uu, c, G = cm.condense(x, u, A, B, g, pr, v, **kw)
    xs, ys, it, nf, st = cm.admm(H, q, Ac, l, uu, eps_abs=1e-6, eps_rel=1e-6, check=5, max_iter=10000)
    ucmd, status, info = M.mpc_step(x, u, pr, Ts=Ts, N=N, vref=v, solver="ipm", **kw)
    line = f"t={t} model {it} ({nf} factorisations)"
    for k, ctl in ctls.items():
        o = ctl.step(x[None], u[None], pr[None], v[None])
        line += f" | {k}: {int(o['iters'][0])} st {int(o['status'][0])} |dU| {np.abs(o['u_cmd'][0] - ucmd).max():.1e}"
    print(line)
    x = dyn.plant_step(x, ucmd, Ts); u = ucmd
