"""Development model (NumPy, fp64) of the algorithm the CUDA kernels implement: condensed QP +
per-row-rho ADMM.  Used to choose solver settings and to debug kernels; not shipped, not an
oracle (the oracle is oracle/, which follows the reference's sparse statement)."""
import numpy as np

from oracle import dynamics as dyn

INF = 1e20


def condense(x0, u_prev, A, B, g, path_ref, vref, q_c=6.0, q_phi=0.5, q_vx=0.5,
             R=np.diag([0.02, 2.0]), Rd=np.diag([0.01, 5.0]),
             u_bounds=((-1.0, 1.0), (-0.6, 0.6)), du_bounds=((-0.5, 0.5), (-0.3, 0.3)),
             x_lo=None, x_hi=None):
    N = len(A)
    n = 2 * N
    c = np.zeros((N + 1, 6)); c[0] = x0
    G = np.zeros((N + 1, 6, n))
    for k in range(N):
        c[k + 1] = A[k] @ c[k] + g[k]
        G[k + 1] = A[k] @ G[k]
        G[k + 1][:, 2 * k:2 * k + 2] = B[k]
    H = np.zeros((n, n)); q = np.zeros(n); const = 0.0
    for k in range(N + 1):
        Xr, Yr, Pr = path_ref[k]
        s, co = np.sin(Pr), np.cos(Pr)
        W = np.stack([s * G[k][0] - co * G[k][1], G[k][2], G[k][3]])
        r = np.array([s * (c[k][0] - Xr) - co * (c[k][1] - Yr), c[k][2] - Pr, c[k][3] - vref[k]])
        L = np.array([q_c, q_phi, q_vx])
        H += 2 * W.T @ (L[:, None] * W)
        q += 2 * W.T @ (L * r)
        const += L @ (r * r)
    Rs, Rds = 0.5 * (R + R.T), 0.5 * (Rd + Rd.T)
    for k in range(N):
        s = slice(2 * k, 2 * k + 2)
        H[s, s] += 2 * Rs + 2 * Rds
        if k > 0:
            sp = slice(2 * k - 2, 2 * k)
            H[sp, sp] += 2 * Rds; H[s, sp] -= 2 * Rds; H[sp, s] -= 2 * Rds
    q[0:2] -= 2 * Rds @ u_prev
    const += u_prev @ Rds @ u_prev
    rows, lo, hi = [], [], []
    for k in range(N):
        for j in range(2):
            r = np.zeros(n); r[2 * k + j] = 1; rows.append(r); lo.append(u_bounds[j][0]); hi.append(u_bounds[j][1])
    for k in range(N):
        for j in range(2):
            r = np.zeros(n); r[2 * k + j] = 1
            if k == 0:
                rows.append(r); lo.append(du_bounds[j][0] + u_prev[j]); hi.append(du_bounds[j][1] + u_prev[j])
            else:
                r[2 * k - 2 + j] = -1; rows.append(r); lo.append(du_bounds[j][0]); hi.append(du_bounds[j][1])
    if x_lo is not None or x_hi is not None:
        xl = np.full(6, -INF) if x_lo is None else np.asarray(x_lo, float)
        xh = np.full(6, INF) if x_hi is None else np.asarray(x_hi, float)
        for k in range(1, N + 1):
            for i in range(6):
                if xl[i] <= -INF and xh[i] >= INF:
                    continue
                rows.append(G[k][i].copy()); lo.append(max(xl[i], -INF) - c[k][i] if xl[i] > -INF else -INF)
                hi.append(xh[i] - c[k][i] if xh[i] < INF else INF)
    return H, q, const, np.array(rows), np.array(lo), np.array(hi), c, G


def row_rho(H, A, rho, mode="hscaled"):
    if mode == "uniform":
        return np.full(A.shape[0], rho)
    dH = np.diag(H)
    # rho_i = rho / max_j (a_ij^2 / H_jj)
    return rho / np.max(A ** 2 / dH[None, :], axis=1)


def admm(H, q, A, l, u, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-5, eps_rel=1e-5, max_iter=4000,
         check=25, mode="hscaled", adaptive=True, warm=None, adapt_tol=5.0):
    n, m = H.shape[0], A.shape[0]
    rv = row_rho(H, A, rho, mode)
    base = rv / rho

    def fac(r):
        rv = base * r
        return rv, np.linalg.inv(H + sigma * np.eye(n) + A.T @ (rv[:, None] * A))

    rv, Kinv = fac(rho)
    x = np.zeros(n); z = np.zeros(m); y = np.zeros(m)
    if warm is not None:
        x, y = warm[0].copy(), warm[1].copy(); z = np.clip(A @ x, l, u)
    nfac = 1
    for it in range(1, max_iter + 1):
        xt = Kinv @ (sigma * x - q + A.T @ (rv * z - y))
        zt = A @ xt
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, l, u)
        y = y + rv * (zr - zn)
        z = zn
        if it % check == 0:
            Ax, Hx, Aty = A @ x, H @ x, A.T @ y
            rp, rd = np.abs(Ax - z).max(), np.abs(Hx + q + Aty).max()
            sp, sd = max(np.abs(Ax).max(), np.abs(z).max()), max(np.abs(Hx).max(), np.abs(Aty).max(), np.abs(q).max())
            if rp <= eps_abs + eps_rel * sp and rd <= eps_abs + eps_rel * sd:
                return x, y, it, nfac, "optimal"
            if adaptive:
                new = rho * np.sqrt((rp / (sp + 1e-10)) / (rd / (sd + 1e-10) + 1e-10))
                new = float(np.clip(new, 1e-6, 1e6))
                if new > adapt_tol * rho or new < rho / adapt_tol:
                    rho = new; rv, Kinv = fac(rho); nfac += 1
    return x, y, max_iter, nfac, "user_limit"
