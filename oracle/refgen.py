"""Oracle: reference-window and velocity-profile generators of the closed-loop driver (fp64).

Test infrastructure -- see ``oracle/__init__.py``.  Restates MPC/main.py:9-18 (``d_steady_state``),
:28-47 (``vref_profile_*``), :51-68 (``ref_window_from_x_with_vref``) and the sinusoidal path
documented in MPC/README.md:73-76.  The reference hard-codes the parabola y = 0.1 x^2; here the
path is a small spec (kind + coefficients) so the same window code serves the parabola, the
README's sinusoid and a natural cubic spline y(x) (BASELINE.json configs 2-3).
"""
import math

import numpy as np

from .dynamics import PARAMS

PATH_PARABOLA = 0   # y = c2 x^2 + c1 x + c0           (MPC/main.py:64-66: c2=0.1)
PATH_SINE = 1       # y = A sin(k x + psi) + c0        (MPC/README.md:75-76: A=0.5,k=0.5)
PATH_SPLINE = 2     # piecewise cubic y(x), scipy PPoly layout, end pieces extrapolate
PATH_ARC = 3        # parametric path (x(s), y(s)): two piecewise cubics over one set of breaks
ARC_PROJECT_ITERS = 4

VREF_HOLD = 0       # vref=None in mpc_step -> x0[3]   (MPC/mpc_6stati.py:158-159)
VREF_CONST = 1      # scalar vref                      (MPC/mpc_6stati.py:160-161)
VREF_RAMP = 2       # vref_profile_ramp_cruise         (MPC/main.py:28-32)
VREF_TRAPEZOID = 3  # vref_profile_trapezoid           (MPC/main.py:34-42)
VREF_SINE = 4       # vref_profile_sine                (MPC/main.py:44-47)


def d_steady_state(v, p=PARAMS):
    """(Cr0 + Cr2 v^2) / (Cm1 - Cm2 v).  MPC/main.py:9-18."""
    return (p["Cr0"] + p["Cr2"] * v ** 2) / (p["Cm1"] - p["Cm2"] * v)


def vref_profile(kind, prm, N, Ts, t0=0.0, vx0=0.0):
    """Velocity reference over the horizon, length N+1, t = t0 + k Ts.

    The reference calls its profile with no time offset on every closed-loop step
    (MPC/main.py:87), i.e. t0 = 0 always; t0 != 0 is the time-advancing extension.
    """
    t = t0 + np.arange(N + 1) * Ts
    if kind == VREF_HOLD:
        return np.full(N + 1, float(vx0))
    if kind == VREF_CONST:
        return np.full(N + 1, float(prm[0]))
    if kind == VREF_RAMP:
        v0, v_cruise, tramp = prm[0], prm[1], prm[2]
        return v0 + (v_cruise - v0) * np.clip(t / tramp, 0.0, 1.0)
    if kind == VREF_TRAPEZOID:
        v0, vmax, t_acc, t_flat, t_dec = prm[0], prm[1], prm[2], prm[3], prm[4]
        v = np.where(t <= t_acc, v0 + (vmax - v0) * (t / t_acc), vmax)
        v = np.where(t > t_acc + t_flat, vmax - (vmax - v0) * ((t - (t_acc + t_flat)) / t_dec), v)
        return np.clip(v, v0, vmax)
    if kind == VREF_SINE:
        v_mean, v_amp, period = prm[0], prm[1], prm[2]
        return v_mean + v_amp * np.sin(2 * np.pi * t / period)
    raise ValueError(f"unknown vref kind {kind}")


def path_eval(kind, prm, xs, spline=None):
    """-> (y(xs), dy/dx(xs)).  ``spline`` = (breaks[K+1], coef[K,4]) with
    y = c0 (x-b)^3 + c1 (x-b)^2 + c2 (x-b) + c3 on [b_i, b_{i+1}) (scipy PPoly order)."""
    xs = np.asarray(xs, dtype=float)
    if kind == PATH_PARABOLA:
        c2, c1, c0 = prm[0], prm[1], prm[2]
        return c2 * xs ** 2 + c1 * xs + c0, 2.0 * c2 * xs + c1
    if kind == PATH_SINE:
        A, k, psi, c0 = prm[0], prm[1], prm[2], prm[3]
        return A * np.sin(k * xs + psi) + c0, A * k * np.cos(k * xs + psi)
    if kind == PATH_SPLINE:
        breaks, coef = spline
        K = coef.shape[0]
        idx = np.clip(np.searchsorted(breaks, xs, side="right") - 1, 0, K - 1)
        dx = xs - breaks[idx]
        c = coef[idx]
        y = ((c[:, 0] * dx + c[:, 1]) * dx + c[:, 2]) * dx + c[:, 3]
        dy = (3.0 * c[:, 0] * dx + 2.0 * c[:, 1]) * dx + c[:, 2]
        return y, dy
    raise ValueError(f"unknown path kind {kind}")


def ref_window(x_start, N, Ts, vref_seq, kind=PATH_PARABOLA, prm=(0.1, 0.0, 0.0, 0.0), spline=None):
    """Reference poses (N+1, 3) re-anchored at the vehicle's current X.  MPC/main.py:51-68:
    xs[k+1] = xs[k] + vref[k] Ts; ys = path(xs); phi* = atan(path'(xs))."""
    vref_seq = np.asarray(vref_seq, dtype=float).reshape(N + 1)
    xs = np.zeros(N + 1)
    xs[0] = x_start
    for k in range(N):
        xs[k + 1] = xs[k] + vref_seq[k] * Ts
    ys, dydx = path_eval(kind, prm, xs, spline)
    return np.stack([xs, ys, np.arctan(dydx)], axis=1)


def arc_eval(arc, s):
    """(x, y, x', y') of the parametric path at parameter s.  ``arc`` = (breaks[K] piece starts, coef_x[K,4], coef_y[K,4]),
    scipy PPoly coefficient order; both ends extrapolate their end piece."""
    breaks, cx, cy = arc
    K = len(cx)
    lo = 0
    while lo + 1 < K and s >= breaks[lo + 1]:
        lo += 1
    d = s - breaks[lo]
    a, b = cx[lo], cy[lo]
    x = ((a[0] * d + a[1]) * d + a[2]) * d + a[3]
    y = ((b[0] * d + b[1]) * d + b[2]) * d + b[3]
    dx = (3.0 * a[0] * d + 2.0 * a[1]) * d + a[2]
    dy = (3.0 * b[0] * d + 2.0 * b[1]) * d + b[2]
    return x, y, dx, dy


def arc_project(arc, s_guess, X, Y, iters=ARC_PROJECT_ITERS):
    """Parameter of the path point closest to (X, Y): ``iters`` Gauss-Newton steps s <- s + (P - p(s)).p'(s) / |p'(s)|^2
    from ``s_guess`` (a fixed count, so that the CUDA path and this restatement do the same arithmetic)."""
    s = float(s_guess)
    for _ in range(iters):
        x, y, dx, dy = arc_eval(arc, s)
        s = s + ((X - x) * dx + (Y - y) * dy) / (dx * dx + dy * dy)
    return s


def ref_window_arc(x_state, s_guess, N, Ts, vref_seq, arc):
    """The window of MPC/main.py:51-68 for a path that is NOT a graph over X.  The reference anchors the window at the
    vehicle's X and advances X by vref Ts (:59-61); here the anchor is the path parameter s0 of the point closest to the
    vehicle (tracked from step to step through ``s_guess``), the window advances along the path by the arclength vref Ts,
    ds = vref Ts / |p'(s)|, and phi* = atan2(y', x') is unwrapped so that it is continuous along the window and within pi of
    the vehicle's (unwrapped) heading -- the cost (phi - phi*)^2 of mpc_6stati.py:233 has no wrap.
    -> (path_ref[N+1,3], s0)"""
    vref_seq = np.asarray(vref_seq, dtype=float).reshape(N + 1)
    s0 = arc_project(arc, s_guess, float(x_state[0]), float(x_state[1]))
    out = np.zeros((N + 1, 3))
    s = s0
    prev = float(x_state[2])
    two_pi = 2.0 * math.pi
    for k in range(N + 1):
        x, y, dx, dy = arc_eval(arc, s)
        raw = math.atan2(dy, dx)
        ph = raw + two_pi * np.rint((prev - raw) / two_pi)
        out[k] = (x, y, ph)
        prev = ph
        if k < N:
            s = s + vref_seq[k] * Ts / math.sqrt(dx * dx + dy * dy)
    return out, s0


def arc_spline_tables(px, py):
    """Natural cubic splines x(s), y(s) through the way-points (px, py), s = cumulative chord length.
    -> (breaks[K] piece starts, coef_x[K,4], coef_y[K,4])"""
    from scipy.interpolate import CubicSpline
    px = np.asarray(px, float); py = np.asarray(py, float)
    s = np.concatenate([[0.0], np.cumsum(np.hypot(np.diff(px), np.diff(py)))])
    sx, sy = CubicSpline(s, px, bc_type="natural"), CubicSpline(s, py, bc_type="natural")
    return np.ascontiguousarray(s[:-1]), np.ascontiguousarray(sx.c.T), np.ascontiguousarray(sy.c.T)


def natural_spline_ppoly(knots_x, knots_y):
    """Natural cubic spline through the knots as (breaks, coef[K,4]) -- the same object
    scipy's ``CubicSpline(bc_type='natural')`` builds (generation_type1.py:97 uses it for
    control profiles; here it describes a path y(x))."""
    from scipy.interpolate import CubicSpline
    cs = CubicSpline(np.asarray(knots_x, float), np.asarray(knots_y, float), bc_type="natural")
    return np.ascontiguousarray(cs.x), np.ascontiguousarray(cs.c.T)
