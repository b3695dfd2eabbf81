"""Oracle: the time-varying QP of ``mpc_step`` in the reference's own sparse form (fp64).

Test infrastructure -- see ``oracle/__init__.py``.  PARITY UNPINNED for this file: the
reference hands the problem to CVXPY -> OSQP (MPC/mpc_6stati.py:252-256), un-pinned
third-party packages that are absent here; no reference test or stored output exists.

``build_sparse_qp`` restates MPC/mpc_6stati.py:180-250 row by row:

  variables   z = [x_0 .. x_N (6 each) ; u_0 .. u_{N-1} (2 each)]        (:181-182)
  equalities  x_0 = x0                                                    (:186)
              x_{k+1} = A_k x_k + B_k u_k + g_k                           (:189-192)
  inequal.    u_lo <= u_k <= u_hi ; du_lo <= u_k - u_{k-1} <= du_hi       (:198-213)
              optional x_lo <= x_k <= x_hi for k = 0..N (k = 0 included)  (:216-221)
  cost        sum_{k<=N} q_c e_c^2 + q_phi (phi-phi*)^2 + q_vx (vx-vref)^2
              + sum_{k<N} u'Ru + du'Rd du         (no 1/2 factors)        (:224-250)

Two solvers for it:

* ``solve_ipm``   Mehrotra primal-dual interior point to ~1e-10: "the exact optimum".
* ``solve_osqp``  restatement of the published OSQP algorithm (Stellato et al., "OSQP: an
  operator splitting solver for quadratic programs", Math. Prog. Comp. 2020) with the
  settings CVXPY passes (eps_abs = eps_rel = 1e-5, max_iter 10000) and OSQP's defaults
  (rho 0.1, x1e3 on equality rows, sigma 1e-6, alpha 1.6, Ruiz scaling 10 iterations,
  adaptive rho, check_termination 25, polish off).  This is the CPU baseline's solver.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

INF = 1e20   # OSQP_INFTY


class SparseQP:
    """min 1/2 z'Pz + q'z + const  s.t.  l <= A z <= u (equalities have l == u)."""

    def __init__(self, P, q, const, A, l, u, N, n_eq):
        self.P, self.q, self.const, self.A, self.l, self.u = P, q, const, A, l, u
        self.N, self.n_eq = N, n_eq

    def split(self, z):
        N = self.N
        return z[: 6 * (N + 1)].reshape(N + 1, 6).T.copy(), z[6 * (N + 1):].reshape(N, 2).T.copy()

    def objective(self, z):
        return 0.5 * z @ (self.P @ z) + self.q @ z + self.const


def build_sparse_qp(x0, u_prev, A_list, B_list, g_list, path_ref, vref,
                    q_c=6.0, q_phi=0.5, q_vx=0.5, R=None, Rd=None,
                    u_bounds=((-1.0, 1.0), (-0.6, 0.6)), du_bounds=((-0.5, 0.5), (-0.3, 0.3)),
                    x_lo=None, x_hi=None):
    """MPC/mpc_6stati.py:180-250 as matrices (see module docstring)."""
    R = np.diag([0.02, 2.0]) if R is None else np.asarray(R, float)
    Rd = np.diag([0.01, 5.0]) if Rd is None else np.asarray(Rd, float)
    N = len(A_list)
    nx, nu = 6 * (N + 1), 2 * N
    nz = nx + nu
    xi = lambda k: slice(6 * k, 6 * k + 6)
    ui = lambda k: slice(nx + 2 * k, nx + 2 * k + 2)

    P = np.zeros((nz, nz))
    q = np.zeros(nz)
    const = 0.0
    Rs, Rds = 0.5 * (R + R.T), 0.5 * (Rd + Rd.T)   # quad_form uses the symmetric part
    for k in range(N + 1):
        Xr, Yr, Pr = path_ref[k]
        s, c = np.sin(Pr), np.cos(Pr)
        a = np.zeros(6)
        a[0], a[1] = s, -c                       # lateral_error :111-117
        e0 = -(s * Xr - c * Yr)                  # e_c = a.x + e0
        Qk = q_c * np.outer(a, a)
        Qk[2, 2] += q_phi
        Qk[3, 3] += q_vx
        P[xi(k), xi(k)] += 2.0 * Qk
        qk = 2.0 * q_c * e0 * a
        qk[2] += -2.0 * q_phi * Pr
        qk[3] += -2.0 * q_vx * vref[k]
        q[xi(k)] += qk
        const += q_c * e0 ** 2 + q_phi * Pr ** 2 + q_vx * vref[k] ** 2
    for k in range(N):
        P[ui(k), ui(k)] += 2.0 * Rs               # u'Ru            :238
        P[ui(k), ui(k)] += 2.0 * Rds              # du'Rd du        :241-245
        if k == 0:
            q[ui(0)] += -2.0 * Rds @ u_prev
            const += u_prev @ Rds @ u_prev
        else:
            P[ui(k - 1), ui(k - 1)] += 2.0 * Rds
            P[ui(k), ui(k - 1)] += -2.0 * Rds
            P[ui(k - 1), ui(k)] += -2.0 * Rds

    rows, lo, hi = [], [], []
    # equalities :186-192
    for i in range(6):
        r = np.zeros(nz); r[i] = 1.0
        rows.append(r); lo.append(x0[i]); hi.append(x0[i])
    for k in range(N):
        blk = np.zeros((6, nz))
        blk[:, xi(k + 1)] = np.eye(6)
        blk[:, xi(k)] = -A_list[k]
        blk[:, ui(k)] = -B_list[k]
        for i in range(6):
            rows.append(blk[i]); lo.append(g_list[k][i]); hi.append(g_list[k][i])
    n_eq = len(rows)
    # input boxes and rates :198-213
    for k in range(N):
        for j in range(2):
            r = np.zeros(nz); r[nx + 2 * k + j] = 1.0
            rows.append(r); lo.append(u_bounds[j][0]); hi.append(u_bounds[j][1])
        for j in range(2):
            r = np.zeros(nz); r[nx + 2 * k + j] = 1.0
            if k == 0:
                rows.append(r); lo.append(du_bounds[j][0] + u_prev[j]); hi.append(du_bounds[j][1] + u_prev[j])
            else:
                r[nx + 2 * (k - 1) + j] = -1.0
                rows.append(r); lo.append(du_bounds[j][0]); hi.append(du_bounds[j][1])
    # optional state boxes, k = 0..N :216-221
    if x_lo is not None or x_hi is not None:
        xl = np.full(6, -INF) if x_lo is None else np.asarray(x_lo, float).reshape(6)
        xh = np.full(6, INF) if x_hi is None else np.asarray(x_hi, float).reshape(6)
        for k in range(N + 1):
            for i in range(6):
                if xl[i] <= -INF and xh[i] >= INF:
                    continue
                r = np.zeros(nz); r[6 * k + i] = 1.0
                rows.append(r); lo.append(max(xl[i], -INF)); hi.append(min(xh[i], INF))
    return SparseQP(P, q, const, np.array(rows), np.array(lo, float), np.array(hi, float), N, n_eq)


# --------------------------------------------------------------------------- exact optimum
def solve_ipm(qp, tol=1e-10, max_iter=80):
    """Mehrotra predictor-corrector on  min 1/2 z'Pz+q'z, Ez=b, Cz<=d.
    Returns (z, lam_ineq[m_in] signed like OSQP's y (>0 at upper, <0 at lower), status)."""
    P, q, A, l, u, n_eq = qp.P, qp.q, qp.A, qp.l, qp.u, qp.n_eq
    nz = P.shape[0]
    E, b = A[:n_eq], l[:n_eq]
    Ain, lin, uin = A[n_eq:], l[n_eq:], u[n_eq:]
    up = np.where(uin < INF)[0]
    dn = np.where(lin > -INF)[0]
    C = np.vstack([Ain[up], -Ain[dn]]) if (len(up) + len(dn)) else np.zeros((0, nz))
    d = np.concatenate([uin[up], -lin[dn]])
    mi = C.shape[0]
    Ps, Es, Cs = sp.csc_matrix(P), sp.csc_matrix(E), sp.csc_matrix(C)

    def kkt_solve(w, r1, r2):
        # [P + C'WC  E'; E 0] [dz; dnu] = [r1; r2]
        K = sp.bmat([[Ps + Cs.T @ sp.diags(w) @ Cs + 1e-13 * sp.eye(nz), Es.T],
                     [Es, -1e-13 * sp.eye(n_eq)]], format="csc")
        sol = spla.splu(K).solve(np.concatenate([r1, r2]))
        return sol[:nz], sol[nz:]

    z = np.zeros(nz)
    nu = np.zeros(n_eq)
    s = np.ones(mi)
    lam = np.ones(mi)
    # a feasible-ish start: solve the equality-constrained problem, then push slacks positive
    z, nu = kkt_solve(np.zeros(mi), -q, b)
    if mi:
        viol = d - C @ z
        s = np.maximum(viol, 1.0)
        lam = np.ones(mi)
    status = "max_iter"
    for it in range(max_iter):
        rd = P @ z + q + E.T @ nu + (C.T @ lam if mi else 0.0)
        rp = E @ z - b
        rc = (C @ z + s - d) if mi else np.zeros(0)
        mu = (s @ lam / mi) if mi else 0.0
        scale = 1.0 + max(np.abs(q).max(), np.abs(d).max() if mi else 0.0)
        if max(np.abs(rd).max(), np.abs(rp).max(), np.abs(rc).max() if mi else 0.0) <= tol * scale and mu <= tol:
            status = "optimal"
            break
        if not mi:
            dz, dnu = kkt_solve(np.zeros(0), -rd, -rp)
            z, nu = z + dz, nu + dnu
            continue
        w = lam / s

        def direction(rs):
            # rs = target for  s*dlam + lam*ds = -rs
            r1 = -rd - C.T @ (w * rc - rs / s)
            dz, dnu = kkt_solve(w, r1, -rp)
            ds = -rc - C @ dz
            dlam = -(rs + lam * ds) / s
            return dz, dnu, ds, dlam

        def step_len(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

        dz_a, dnu_a, ds_a, dl_a = direction(s * lam)
        ap, ad = step_len(s, ds_a), step_len(lam, dl_a)
        mu_aff = (s + ap * ds_a) @ (lam + ad * dl_a) / mi
        sig = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        dz, dnu, ds, dl = direction(s * lam + ds_a * dl_a - sig * mu)
        ap = min(1.0, 0.995 * step_len(s, ds) if step_len(s, ds) < 1.0 else 1.0)
        ad = min(1.0, 0.995 * step_len(lam, dl) if step_len(lam, dl) < 1.0 else 1.0)
        a = min(ap, ad)
        z, nu, s, lam = z + a * dz, nu + a * dnu, s + a * ds, lam + a * dl
        if not np.all(np.isfinite(z)) or np.abs(z).max() > 1e12:
            status = "infeasible"
            break
    y = np.zeros(Ain.shape[0])
    if mi:
        y[up] += lam[: len(up)]
        y[dn] -= lam[len(up):]
    if status == "max_iter" and mi and (np.abs(lam).max() > 1e8 or mu > 1e-4):
        status = "infeasible"
    return z, y, status


def is_trivially_infeasible(qp):
    """Cheap exact checks used to label fixtures: crossed bounds / x0 outside its box."""
    return bool(np.any(qp.l > qp.u + 1e-12))


# --------------------------------------------------------------------------- OSQP restatement
def _ruiz(P, q, A, iters=10):
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, q, A = P.copy(), q.copy(), A.copy()
    MIN_S, MAX_S = 1e-4, 1e4
    for _ in range(iters):
        cn = np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0) if m else 0.0)
        rn = np.abs(A).max(axis=1) if m else np.zeros(0)
        cn = np.where(cn < MIN_S, 1.0, np.minimum(cn, MAX_S))
        rn = np.where(rn < MIN_S, 1.0, np.minimum(rn, MAX_S))
        dD, dE = 1.0 / np.sqrt(cn), 1.0 / np.sqrt(rn)
        P = dD[:, None] * P * dD[None, :]
        A = dE[:, None] * A * dD[None, :]
        q = dD * q
        D, E = D * dD, E * dE
        pn = np.abs(P).max(axis=0).mean()
        ct = max(pn, np.abs(q).max())
        ct = 1.0 if ct < MIN_S else min(ct, MAX_S)
        P, q, c = P / ct, q / ct, c / ct
    return P, q, A, D, E, c


def solve_osqp(qp, eps_abs=1e-5, eps_rel=1e-5, max_iter=10000, rho=0.1, sigma=1e-6, alpha=1.6,
               check_termination=25, adaptive_rho=True, adaptive_rho_interval=25,
               adaptive_rho_tolerance=5.0, scaling=10, eps_prim_inf=1e-4, warm=None):
    """OSQP algorithm 1 (+ Ruiz scaling, per-row rho, adaptive rho, infeasibility certificate).
    Returns dict(z, y, status, iters, rho_updates).  Status strings are the ones CVXPY
    reports and mpc_step tests at MPC/mpc_6stati.py:261."""
    P0, q0, A0, l0, u0 = qp.P, qp.q, qp.A, qp.l, qp.u
    n, m = P0.shape[0], A0.shape[0]
    if scaling:
        P, q, A, D, E, c = _ruiz(P0, q0, A0, scaling)
    else:
        P, q, A, D, E, c = P0.copy(), q0.copy(), A0.copy(), np.ones(n), np.ones(m), 1.0
    l, u = E * l0, E * u0
    l = np.where(l0 <= -INF, -INF, l)
    u = np.where(u0 >= INF, INF, u)
    is_eq = np.abs(l0 - u0) < 1e-4 * 1.0     # RHO_TOL
    is_free = (l0 <= -INF) & (u0 >= INF)
    Ps, As = sp.csc_matrix(P), sp.csc_matrix(A)

    def rho_vec(r):
        v = np.full(m, r)
        v[is_eq] = 1e3 * r
        v[is_free] = 1e-6
        return v

    def factor(r):
        rv = rho_vec(r)
        K = sp.bmat([[Ps + sigma * sp.eye(n), As.T], [As, -sp.diags(1.0 / rv)]], format="csc")
        return rv, spla.splu(K)

    rv, lu = factor(rho)
    x = np.zeros(n); z = np.zeros(m); y = np.zeros(m)
    if warm is not None:
        x = warm[0] / D
        z = A @ x
        y = warm[1] / E * c
    status, rho_updates = "user_limit", 0
    Einv, Dinv = 1.0 / E, 1.0 / D
    it = 0
    for it in range(1, max_iter + 1):
        rhs = np.concatenate([sigma * x - q, z - y / rv])
        sol = lu.solve(rhs)
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rv
        x_new = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        z_new = np.clip(zr + y / rv, l, u)
        y_new = y + rv * (zr - z_new)
        dy = y_new - y
        x, z, y = x_new, z_new, y_new
        if it % check_termination == 0 or it == max_iter:
            Ax, Px, Aty = A @ x, P @ x, A.T @ y
            r_prim = np.abs(Einv * (Ax - z)).max() if m else 0.0
            r_dual = np.abs(Dinv * (Px + q + Aty)).max() / c
            n_prim = max(np.abs(Einv * Ax).max(), np.abs(Einv * z).max()) if m else 0.0
            n_dual = max(np.abs(Dinv * Px).max(), np.abs(Dinv * Aty).max(), np.abs(Dinv * q).max()) / c
            if r_prim <= eps_abs + eps_rel * n_prim and r_dual <= eps_abs + eps_rel * n_dual:
                status = "optimal"
                break
            # primal infeasibility certificate
            ndy = np.abs(E * dy).max() if m else 0.0
            if ndy > 1e-30:
                lhs = np.sum(np.where(u < INF, u, 0.0) * np.maximum(dy, 0)
                             + np.where(l > -INF, l, 0.0) * np.minimum(dy, 0))
                unb = np.any((u >= INF) & (dy > eps_prim_inf * ndy)) or np.any((l <= -INF) & (dy < -eps_prim_inf * ndy))
                if (not unb) and lhs < -eps_prim_inf * ndy and np.abs(Dinv * (A.T @ dy)).max() <= eps_prim_inf * ndy:
                    status = "infeasible"
                    break
            if adaptive_rho and it % adaptive_rho_interval == 0:
                sc_p = max(np.abs(Ax).max(), np.abs(z).max())
                sc_d = max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(q).max())
                rp_s = np.abs(Ax - z).max() / (sc_p + 1e-10)
                rd_s = np.abs(Px + q + Aty).max() / (sc_d + 1e-10)
                new = float(np.clip(rho * np.sqrt(rp_s / (rd_s + 1e-10)), 1e-6, 1e6))
                if new > rho * adaptive_rho_tolerance or new < rho / adaptive_rho_tolerance:
                    rho = new
                    rv, lu = factor(rho)
                    rho_updates += 1
    return {"z": D * x, "y": E * y / c, "status": status, "iters": it, "rho_updates": rho_updates}
