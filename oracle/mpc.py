"""Oracle: ``mpc_step`` and the closed loop of ``MPC/main.py`` (fp64, NumPy).

Test infrastructure -- see ``oracle/__init__.py``.  ``mpc_step`` follows
MPC/mpc_6stati.py:120-275 (same signature minus ``solver``/``verbose``, same return tuple,
same fallback); the QP half is ``oracle/qp.py`` (PARITY UNPINNED, see there).
``closed_loop`` follows MPC/main.py:72-101.
"""
import numpy as np

from . import dynamics as dyn
from . import qp as oqp
from . import refgen

ACCEPTED = ("optimal", "optimal_inaccurate")   # MPC/mpc_6stati.py:261


def mpc_step(x0, u_prev, path_ref, Ts=0.02, N=20, params=None,
             q_c=6.0, q_phi=0.5, q_vx=0.5, R=np.diag([0.02, 2.0]), Rd=np.diag([0.01, 5.0]),
             vref=None, u_bounds=((-1.0, 1.0), (-0.6, 0.6)), du_bounds=((-0.5, 0.5), (-0.3, 0.3)),
             x_lo=None, x_hi=None, solver="ipm", variant=dyn.VARIANT_MPC, solver_opts=None):
    """-> (u_cmd[2], status, info).  ``solver``: "ipm" (exact optimum) or "osqp" (restated
    OSQP at CVXPY's settings, cold start: a fresh cp.Problem has no solver cache, so the
    reference's warm_start=True at :256 is a no-op)."""
    p = dict(dyn.PARAMS)                                   # :144-146
    if params is not None:
        p.update(params)
    x0 = np.asarray(x0, float).reshape(6)                  # :148-151
    u_pr = np.asarray(u_prev, float).reshape(2)
    path_ref = np.asarray(path_ref, float)
    assert path_ref.shape[0] == N + 1 and path_ref.shape[1] == 3
    if vref is None:                                       # :158-163
        vref = np.full(N + 1, x0[3])
    elif np.isscalar(vref):
        vref = np.full(N + 1, float(vref))
    else:
        vref = np.asarray(vref, float).reshape(N + 1)

    A, B, g, _ = dyn.linearize_horizon(x0, u_pr, Ts, N, p, variant)    # :165-178
    prob = oqp.build_sparse_qp(x0, u_pr, A, B, g, path_ref, vref, q_c, q_phi, q_vx, R, Rd,
                               u_bounds, du_bounds, x_lo, x_hi)        # :180-252
    opts = dict(solver_opts or {})
    iters = 0
    try:                                                               # :255-259
        if oqp.is_trivially_infeasible(prob):
            z, status = None, "infeasible"
        elif solver == "ipm":
            z, y, status = oqp.solve_ipm(prob, **opts)
        else:
            out = oqp.solve_osqp(prob, **opts)
            z, y, status, iters = out["z"], out["y"][prob.n_eq:], out["status"], out["iters"]
    except Exception as e:  # noqa: BLE001
        return u_pr, f"Solver Error: {type(e).__name__}", {}
    if status not in ACCEPTED:                                         # :261-262
        return u_pr, status, {}
    X, U = prob.split(z)
    u_cmd = np.array([U[0, 0], U[1, 0]])                               # :265
    info = {"status": status, "objective": prob.objective(z), "X_opt": X, "U_opt": U,
            "path_ref": path_ref, "vref": vref, "y_ineq": y, "iters": iters}
    return u_cmd, status, info


def closed_loop(x0, u_prev, T, Ts, N, path_kind=refgen.PATH_PARABOLA, path_prm=(0.1, 0.0, 0.0, 0.0),
                spline=None, vref_kind=refgen.VREF_RAMP, vref_prm=(0.8, 2.0, 2.0), vref_advance=False,
                plant=dyn.PLANT_MPC, solver="ipm", solver_opts=None, arc=None, **mpc_kwargs):
    """MPC/main.py:85-101: for t in range(T): vref window (:87), path window anchored at x[0]
    (:90), mpc_step (:94), Euler plant step (:97), u_prev <- u_cmd (:101).
    Returns X[T+1,6] (row 0 = x0), U[T,2], statuses[T], iters[T]."""
    x = np.asarray(x0, float).copy()
    u_prev = np.asarray(u_prev, float).copy()
    Xh = np.zeros((T + 1, 6)); Uh = np.zeros((T, 2)); st = []; its = np.zeros(T, int)
    Xh[0] = x
    for t in range(T):
        t0 = t * Ts if vref_advance else 0.0
        vref_seq = refgen.vref_profile(vref_kind, vref_prm, N, Ts, t0, x[3])
        if path_kind == refgen.PATH_ARC:     # arclength-parameterised path: the anchor s is tracked from step to step
            path_ref, s_anchor = refgen.ref_window_arc(x, path_prm[0] if t == 0 else s_anchor, N, Ts, vref_seq, arc)
        else:
            path_ref = refgen.ref_window(x[0], N, Ts, vref_seq, path_kind, path_prm, spline)
        u_cmd, status, info = mpc_step(x, u_prev, path_ref, Ts=Ts, N=N, vref=vref_seq,
                                       solver=solver, solver_opts=solver_opts, **mpc_kwargs)
        x = dyn.plant_step(x, u_cmd, Ts, plant=plant)
        Xh[t + 1] = x; Uh[t] = u_cmd; st.append(status); its[t] = info.get("iters", 0) if info else 0
        u_prev = u_cmd
    return Xh, Uh, st, its
