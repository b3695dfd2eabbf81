"""Oracle: 6-state dynamic-bicycle / Pacejka model and its FD linearisation (fp64, NumPy).

Test infrastructure -- see ``oracle/__init__.py``.  Restates

* ``Params``                 MPC/mpc_6stati.py:9-19 (identical copies generation_type1.py:11-16,
                             generation_type2.py:14-19)
* ``tire_forces``/``f_cont`` MPC/mpc_6stati.py:25-71, generation_type1.py:38-68,
                             generation_type2.py:52-86 (three variants)
* ``numerical_jacobian``     MPC/mpc_6stati.py:73-97
* ``linearize_discretize``   MPC/mpc_6stati.py:99-109
* plant steps                MPC/main.py:97 (no clipping), generation_type1.py:70-84 /
                             generation_type2.py:180-187 (vx>=0, |omega|<=6)
"""
import math

import numpy as np

# MPC/mpc_6stati.py:9-19
PARAMS = {
    "Cm1": 0.287, "Cm2": 0.0545, "Cr0": 0.0518, "Cr2": 0.00035,
    "Br": 3.3852, "Cr": 1.2691, "Dr": 0.1737,
    "Bf": 2.579, "Cf": 1.2, "Df": 0.192,
    "m": 0.041, "Iz": 27.8e-6, "lf": 0.029, "lr": 0.033, "g": 9.81,
    "maxAlpha": 0.6, "vx_zero": 0.3,
}

VARIANT_MPC = 0    # MPC/mpc_6stati.py
VARIANT_GEN1 = 1   # generation_traj/generation_type1.py
VARIANT_GEN2 = 2   # generation_traj/generation_type2.py


def _clamp(v, lo, hi):
    # MPC/mpc_6stati.py:21-23  np.minimum(np.maximum(x, lo), hi)
    return min(max(v, lo), hi)


def tire_forces(x, u, p=PARAMS, variant=VARIANT_MPC):
    """Slip angles -> Pacejka lateral forces + rear drive force.

    MPC variant        MPC/mpc_6stati.py:25-53  (vx_eff keeps the sign of vx, sign(0)=0; both
                       slip angles clamped; Frx uses the raw vx)
    gen type1 variant  generation_type1.py:38-54 (vx_eff=max(|vx|,vx_zero); only alpha_f clamped;
                       Frx uses vx_eff)
    gen type2 variant  generation_type2.py:52-71 (as type1 but alpha_r clamped too)
    """
    vx, vy, omega = float(x[3]), float(x[4]), float(x[5])
    d, delta = float(u[0]), float(u[1])
    if variant == VARIANT_MPC:
        sgn = (vx > 0) - (vx < 0)
        vx_eff = sgn * max(abs(vx), p["vx_zero"])
    else:
        vx_eff = max(abs(vx), p["vx_zero"])
    alpha_f = -math.atan2(omega * p["lf"] + vy, vx_eff) + delta
    alpha_r = math.atan2(omega * p["lr"] - vy, vx_eff)
    alpha_f = _clamp(alpha_f, -p["maxAlpha"], p["maxAlpha"])
    if variant != VARIANT_GEN1:
        alpha_r = _clamp(alpha_r, -p["maxAlpha"], p["maxAlpha"])
    Fy_f = p["Df"] * math.sin(p["Cf"] * math.atan(p["Bf"] * alpha_f))
    Fy_r = p["Dr"] * math.sin(p["Cr"] * math.atan(p["Br"] * alpha_r))
    v_long = vx if variant == VARIANT_MPC else vx_eff
    Frx = (p["Cm1"] - p["Cm2"] * v_long) * d - p["Cr0"] - p["Cr2"] * (v_long ** 2)
    return Fy_f, Fy_r, Frx


def f_cont(x, u, p=PARAMS, variant=VARIANT_MPC):
    """Continuous-time dynamics.  MPC/mpc_6stati.py:55-71 (generation_type1.py:56-68,
    generation_type2.py:73-86 differ only in how the 1/m, 1/Iz factors are written)."""
    phi, vx, vy, omega = float(x[2]), float(x[3]), float(x[4]), float(x[5])
    delta = float(u[1])
    m, Iz, lf, lr = p["m"], p["Iz"], p["lf"], p["lr"]
    Fy_f, Fy_r, Frx = tire_forces(x, u, p, variant)
    Xdot = vx * math.cos(phi) - vy * math.sin(phi)
    Ydot = vx * math.sin(phi) + vy * math.cos(phi)
    if variant == VARIANT_MPC:
        vxdot = (1.0 / m) * (Frx - Fy_f * math.sin(delta) + m * vy * omega)
        vydot = (1.0 / m) * (Fy_r + Fy_f * math.cos(delta) - m * vx * omega)
        omdot = (1.0 / Iz) * (Fy_f * lf * math.cos(delta) - Fy_r * lr)
    else:
        vxdot = (Frx - Fy_f * math.sin(delta) + m * vy * omega) / m
        vydot = (Fy_r + Fy_f * math.cos(delta) - m * vx * omega) / m
        omdot = (Fy_f * lf * math.cos(delta) - Fy_r * lr) / Iz
    return np.array([Xdot, Ydot, omega, vxdot, vydot, omdot])


def numerical_jacobian(x, u, p=PARAMS, variant=VARIANT_MPC, eps_x=1e-5, eps_u=1e-5):
    """Central differences, 12 + 4 + 1 evaluations.  MPC/mpc_6stati.py:73-97."""
    x = np.asarray(x, dtype=float)
    u = np.asarray(u, dtype=float)
    n, m = x.size, u.size
    Jx = np.zeros((n, n))
    Ju = np.zeros((n, m))
    for i in range(n):
        dx = np.zeros(n)
        dx[i] = eps_x
        Jx[:, i] = (f_cont(x + dx, u, p, variant) - f_cont(x - dx, u, p, variant)) / (2.0 * eps_x)
    for j in range(m):
        du = np.zeros(m)
        du[j] = eps_u
        Ju[:, j] = (f_cont(x, u + du, p, variant) - f_cont(x, u - du, p, variant)) / (2.0 * eps_u)
    return Jx, Ju, f_cont(x, u, p, variant)


def linearize_discretize(x_bar, u_bar, Ts, p=PARAMS, variant=VARIANT_MPC):
    """Ad = I + Ts Jx, Bd = Ts Ju, g = xbar + Ts f - Ad xbar - Bd ubar.  MPC/mpc_6stati.py:99-109."""
    x_bar = np.asarray(x_bar, dtype=float)
    u_bar = np.asarray(u_bar, dtype=float)
    Jx, Ju, fval = numerical_jacobian(x_bar, u_bar, p, variant)
    Ad = np.eye(x_bar.size) + Ts * Jx
    Bd = Ts * Ju
    g = x_bar + Ts * fval - Ad @ x_bar - Bd @ u_bar
    return Ad, Bd, g


def nominal_rollout(x0, u_prev, Ts, N, p=PARAMS, variant=VARIANT_MPC):
    """xbar_{k+1} = xbar_k + Ts f(xbar_k, u_prev), ubar_k = u_prev.  MPC/mpc_6stati.py:165-172."""
    xbar = np.zeros((6, N + 1))
    ubar = np.zeros((2, N))
    xbar[:, 0] = x0
    for k in range(N):
        ubar[:, k] = u_prev
        xbar[:, k + 1] = xbar[:, k] + Ts * f_cont(xbar[:, k], ubar[:, k], p, variant)
    return xbar, ubar


def linearize_horizon(x0, u_prev, Ts, N, p=PARAMS, variant=VARIANT_MPC):
    """Rollout + N linearisations.  MPC/mpc_6stati.py:165-178.  Returns A[N,6,6], B[N,6,2], g[N,6]."""
    xbar, ubar = nominal_rollout(x0, u_prev, Ts, N, p, variant)
    A = np.zeros((N, 6, 6))
    B = np.zeros((N, 6, 2))
    g = np.zeros((N, 6))
    for k in range(N):
        A[k], B[k], g[k] = linearize_discretize(xbar[:, k], ubar[:, k], Ts, p, variant)
    return A, B, g, xbar


PLANT_MPC = 0      # MPC/main.py:97  x <- x + Ts f(x,u), MPC variant, no clipping
PLANT_GEN1 = 1     # generation_type1.py:70-84  gen1 variant + clipping
PLANT_GEN2 = 2     # generation_type2.py:180-187 gen2 variant + clipping


def plant_step(x, u, Ts, p=PARAMS, plant=PLANT_MPC):
    """One explicit-Euler plant step with the generator's stability clipping where it applies."""
    x = np.asarray(x, dtype=float)
    xn = x + Ts * f_cont(x, u, p, plant)
    if plant != PLANT_MPC:
        # generation_type1.py:81-82 / generation_type2.py:186-187
        xn[3] = max(xn[3], 0.0)
        xn[5] = min(max(xn[5], -6.0), 6.0)
    return xn
