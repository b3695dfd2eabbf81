"""CPU: independent third-party cross-check of the oracle's QP half (oracle/qp.py, PARITY UNPINNED -- cvxpy / osqp are
absent).  Every fixture of tests/golden/oracle_qp.npz is rebuilt in the reference's sparse form (MPC/mpc_6stati.py:180-250
restated by build_sparse_qp) and handed to SciPy: `scipy.optimize.minimize(method="trust-constr")` must land on the
oracle's interior-point solution, and HiGHS (`scipy.optimize.linprog`, zero objective) must agree with the oracle's
feasible / infeasible verdicts.  This does not pin the oracle to the reference's CVXPY -> OSQP answer (only the real
solver could); it rules out an error shared between the oracle's problem builder and its own solvers."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, linprog, minimize

from oracle import dynamics as dyn, mpc as ompc, qp as oqp, refgen as R
from conftest import HARD


def _problem(c):
    kw = HARD if bool(c["hard"]) else {}
    N, Ts = int(c["N"]), float(c["Ts"])
    A, B, g, _ = dyn.linearize_horizon(c["x0"], c["u_prev"], Ts, N)
    return oqp.build_sparse_qp(c["x0"], c["u_prev"], A, B, g, c["path_ref"], c["vref"], **kw)


def _scipy_solve(prob, z0):
    P, A = sp.csr_matrix(prob.P), sp.csr_matrix(prob.A)       # sparse: 20x faster than dense in trust-constr
    l = np.where(prob.l <= -oqp.INF, -np.inf, prob.l)
    u = np.where(prob.u >= oqp.INF, np.inf, prob.u)
    fun = lambda z: 0.5 * z @ (P @ z) + prob.q @ z
    jac = lambda z: P @ z + prob.q
    hess = lambda z: P
    res = minimize(fun, z0, jac=jac, hess=hess, method="trust-constr", constraints=[LinearConstraint(A, l, u)],
                   options={"gtol": 1e-10, "xtol": 1e-12, "barrier_tol": 1e-12, "maxiter": 3000})
    return res


def test_scipy_trust_constr_agrees_with_oracle_ipm_on_every_fixture(golden_qp):
    worst_u, worst_obj = 0.0, 0.0
    for c in golden_qp:
        prob = _problem(c)
        # start from the nominal rollout (a feasible point of the equalities), not from the oracle's answer
        N = int(c["N"])
        xbar = dyn.linearize_horizon(c["x0"], c["u_prev"], float(c["Ts"]), N)[3]
        z0 = np.concatenate([np.asarray(xbar).reshape(-1), np.tile(c["u_prev"], N)])
        res = _scipy_solve(prob, z0)
        X, U = prob.split(res.x)
        du = np.abs(U - c["U_opt"]).max()
        obj = prob.objective(res.x)
        o_ref = float(c["objective"])
        worst_u = max(worst_u, du)
        worst_obj = max(worst_obj, abs(obj - o_ref) / (1.0 + abs(o_ref)))
        Az = prob.A @ res.x
        assert (Az >= prob.l - 1e-7).all() and (Az <= prob.u + 1e-7).all()      # SciPy's point is feasible ...
        # ... so its objective bounds the optimum from above: the oracle's (feasible, KKT-checked) point must not be worse.
        # This one-sided test is tight; SciPy's barrier method itself stops up to ~1e-5 above the optimum when 30+ rows
        # are active (it is then the less accurate of the two), hence the looser two-sided bounds.
        assert o_ref <= obj + 1e-8 * (1.0 + abs(o_ref)), (N, bool(c["hard"]), o_ref - obj)
        assert abs(obj - o_ref) <= 2e-5 * (1.0 + abs(o_ref)), (N, bool(c["hard"]), obj - o_ref)
        # inputs: within 5e-4 = half the 1e-3 parity bar (measured: <= 2e-7 on the lightly constrained fixtures, up to
        # 1.8e-4 where 30+ rows are active -- vy / omega are unpenalised, so the optimum is flat: objective differences
        # of 1e-9 move U by 1e-5)
        assert du <= 5e-4, (N, bool(c["hard"]), du)
    print("scipy trust-constr vs oracle IPM: max |U - U*| %.2e, max rel objective error %.2e" % (worst_u, worst_obj))


def _highs_feasible(prob):
    A = prob.A.toarray() if hasattr(prob.A, "toarray") else np.asarray(prob.A)
    l = np.where(prob.l <= -oqp.INF, -np.inf, prob.l)
    u = np.where(prob.u >= oqp.INF, np.inf, prob.u)
    eq = np.isclose(l, u)
    fin_u, fin_l = np.isfinite(u) & ~eq, np.isfinite(l) & ~eq
    A_ub = np.vstack([A[fin_u], -A[fin_l]])
    b_ub = np.concatenate([u[fin_u], -l[fin_l]])
    res = linprog(np.zeros(A.shape[1]), A_ub=A_ub, b_ub=b_ub, A_eq=A[eq], b_eq=l[eq], bounds=(None, None), method="highs")
    return res.status == 0


def test_highs_agrees_with_oracle_feasibility_verdicts(golden_qp):
    # every stored fixture is feasible ...
    for c in golden_qp:
        assert _highs_feasible(_problem(c))
    # ... and the infeasible family of BASELINE config 4 (tight rate limits + vy / omega box, large offsets) is flagged by
    # both: sweep lateral offsets / yaw rates until the oracle reports infeasible steps and compare verdict by verdict
    rng = np.random.default_rng(11)
    n_inf = 0
    for trial in range(40):
        N, Ts = 20, 0.02
        vx = rng.uniform(0.5, 1.5)
        x = np.array([rng.uniform(-1, 1), 0.0, rng.uniform(-0.6, 0.6), vx, rng.uniform(-0.2, 0.2), rng.uniform(-2.5, 2.5)])
        prm = (rng.uniform(0.2, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0, 2 * np.pi), 0.0)
        x[1] = R.path_eval(R.PATH_SINE, prm, x[0:1])[0][0] + rng.uniform(-1.5, 1.5)
        up = np.array([R.d_steady_state(vx), rng.uniform(-0.3, 0.3)])
        v = R.vref_profile(R.VREF_RAMP, (0.8, rng.uniform(0.8, 2.0), 2.0), N, Ts)
        pr = R.ref_window(x[0], N, Ts, v, R.PATH_SINE, prm)
        A, B, g, _ = dyn.linearize_horizon(x, up, Ts, N)
        prob = oqp.build_sparse_qp(x, up, A, B, g, pr, v, **HARD)
        _, st, _ = ompc.mpc_step(x, up, pr, Ts=Ts, N=N, vref=v, solver="ipm", **HARD)
        feas = _highs_feasible(prob)
        assert (st == "infeasible") == (not feas), (trial, st, feas)
        n_inf += st == "infeasible"
    assert n_inf >= 5, n_inf     # the sweep does exercise the infeasible branch
