"""Oracle: scenario + initial-state generation of the B200 path (csrc/tg_scenarios.cuh), restated in NumPy.

Test infrastructure -- see ``oracle/__init__.py``.  The generator draws what SURVEY.md section 8(d) asks of configs 2 / 3 / 5:
x0 from the ranges of generation_type1.py:260-265 (generation_type2.py:171-174 for the type-2 ranges) with Y and phi placed
relative to the reference path, one reference path per trajectory (natural cubic spline through random knots -- the
construction of generation_type1.py:86-100, here as a path y(x) --, sinusoid MPC/README.md:75, parabola MPC/main.py:64) and the
ramp-cruise speed profile of MPC/main.py:28-32.  Random numbers: Philox4x32-10 keyed by seed_base + trajectory id
(oracle/philox.py), uniforms (r + 0.5) 2^-32, normals by the shared Box-Muller.  Plain NumPy fp64 arithmetic in the order the
kernel uses with non-contracted operations, so everything but the libm-dependent entries (sin / cos / atan) is bit-identical.
"""
import numpy as np

from . import dynamics as dyn
from . import philox as oph
from . import refgen as R

STREAM = 7

DEFAULT_RULES = dict(
    x0_lo=(-2.0, 0.0, 0.0, 0.4, -0.05, -1.0), x0_hi=(2.0, 0.0, 0.0, 1.5, 0.05, 1.0),     # generation_type1.py:260-265
    lat_off=(-0.2, 0.2), head_off=(-0.2, 0.2), vref0=0.8, vcruise=(0.8, 2.0), t_ramp=2.0,
    sine_A=(0.2, 1.0), sine_k=(0.3, 1.0), sine_psi=(0.0, 6.283185307179586), parab_c=(-0.2, 0.2),
    spl_x0=-6.0, spl_dx=(1.0, 3.0), spl_sigma=0.3, spl_knots=27, cycle=(R.PATH_SPLINE, R.PATH_SINE), seed_base=2025)


def _uniform(r, lo, hi):
    u = (np.asarray(r, np.float64) + 0.5) * 2.0 ** -32
    return lo + (hi - lo) * u


def _block(seed, blk):
    """the 4 words of philox(counter = (blk, STREAM, 0, 0), key = seed)"""
    return oph.philox_stream(seed, blk, STREAM, 1)[0]


def natural_spline(kx, ky):
    """second-derivative form + Thomas algorithm (the arithmetic of trajectory_generation_b200.Scenarios.set_splines)
    -> coef[K-1, 4] in scipy's PPoly order"""
    K = len(kx)
    h = kx[1:] - kx[:-1]
    d = (ky[1:] - ky[:-1]) / h
    n = K - 2
    m = np.zeros(K)
    if n > 0:
        b = 2.0 * (h[:-1] + h[1:]); r = 6.0 * (d[1:] - d[:-1])
        for i in range(1, n):
            w = h[i] / b[i - 1]
            b[i] = b[i] - w * h[i]
            r[i] = r[i] - w * r[i - 1]
        m[n] = r[n - 1] / b[n - 1]
        for i in range(n - 2, -1, -1):
            m[i + 1] = (r[i] - h[i + 1] * m[i + 2]) / b[i]
    c0 = (m[1:] - m[:-1]) / (6.0 * h)
    c1 = m[:-1] / 2.0
    c2 = d - h * (2.0 * m[:-1] + m[1:]) / 6.0
    return np.stack([c0, c1, c2, ky[:-1]], axis=-1)


def make_scenarios(B, traj_id0=0, params=dyn.PARAMS, **overrides):
    """-> dict(x0[B,6], u0[B,2], path_kind[B], path[B,4], vref[B,6], breaks[B,K-1], coef[B,K-1,4])"""
    ru = dict(DEFAULT_RULES); ru.update(overrides)
    K = int(ru["spl_knots"]); P = K - 1
    out = {"x0": np.zeros((B, 6)), "u0": np.zeros((B, 2)), "path_kind": np.zeros(B, np.int32), "path": np.zeros((B, 4)),
           "vref": np.zeros((B, 6)), "breaks": np.zeros((B, P)), "coef": np.zeros((B, P, 4))}
    for b in range(B):
        i = traj_id0 + b
        seed = int(ru["seed_base"]) + i
        r0, r1, r2 = _block(seed, 0), _block(seed, 1), _block(seed, 2)
        X = _uniform(r0[0], ru["x0_lo"][0], ru["x0_hi"][0])
        lat = _uniform(r0[1], *ru["lat_off"]); head = _uniform(r0[2], *ru["head_off"])
        vx = _uniform(r0[3], ru["x0_lo"][3], ru["x0_hi"][3])
        vy = _uniform(r1[0], ru["x0_lo"][4], ru["x0_hi"][4]); om = _uniform(r1[1], ru["x0_lo"][5], ru["x0_hi"][5])
        vcr = _uniform(r1[2], *ru["vcruise"])
        kind = int(ru["cycle"][i % len(ru["cycle"])])
        out["path_kind"][b] = kind
        out["vref"][b, :3] = (ru["vref0"], vcr, ru["t_ramp"])
        if kind == R.PATH_SPLINE:
            kx = np.zeros(K); ky = np.zeros(K)
            kx[0] = ru["spl_x0"]
            for j in range(P):
                q = _block(seed, 3 + j // 4)
                kx[j + 1] = kx[j] + _uniform(q[j % 4], *ru["spl_dx"])
            for pj in range((K + 1) // 2):
                q = _block(seed, 16 + pj // 2)
                n0, n1 = oph.box_muller([q[(pj % 2) * 2]], [q[(pj % 2) * 2 + 1]])
                ky[2 * pj] = ru["spl_sigma"] * n0[0]
                if 2 * pj + 1 < K:
                    ky[2 * pj + 1] = ru["spl_sigma"] * n1[0]
            coef = natural_spline(kx, ky)
            out["breaks"][b] = kx[:-1]; out["coef"][b] = coef
            piece = int(np.clip((kx[:-1] <= X).sum() - 1, 0, P - 1))
            dx = X - kx[piece]
            c0, c1, c2, c3 = coef[piece]
            y = ((c0 * dx + c1) * dx + c2) * dx + c3
            dy = (3.0 * c0 * dx + 2.0 * c1) * dx + c2
        elif kind == R.PATH_SINE:
            A = _uniform(r2[0], *ru["sine_A"]); k = _uniform(r2[1], *ru["sine_k"]); psi = _uniform(r2[2], *ru["sine_psi"])
            out["path"][b, :3] = (A, k, psi)
            y = A * np.sin(k * X + psi); dy = (A * k) * np.cos(k * X + psi)
        else:
            cc = _uniform(r2[0], *ru["parab_c"])
            out["path"][b, 0] = cc
            y = cc * (X * X); dy = (2.0 * cc) * X
        out["x0"][b] = (X, y + lat, np.arctan(dy) + head, vx, vy, om)
        out["u0"][b, 0] = (params["Cr0"] + params["Cr2"] * (vx * vx)) / (params["Cm1"] - params["Cm2"] * vx)   # MPC/main.py:9-18
    return out
