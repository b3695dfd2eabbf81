"""Oracle: the estimator-side physics -- KalmanNet's prior step.

Test infrastructure -- see ``oracle/__init__.py``.  Restates, with plain torch tensor operations on the CPU,

* KalmanNet/vehicle_model.py:19-40  ``pt_tire_forces`` (gen-type-1 tyre model: only the front slip angle is clamped)
* KalmanNet/vehicle_model.py:43-79  ``pt_f_cont`` (phi, vx, vy, omega clamped to the data-set limits before use)
* KalmanNet/vehicle_model.py:109-134 ``VehicleModel.f`` (explicit Euler from the UNclamped state, then all six states clamped)
* KalmanNet/test_prediction.py:68-87 ``rollout_open_loop``

PINNED: tests/golden/reference_estimator.npz holds outputs and autograd gradients of the reference's own
``VehicleModel.f`` / ``rollout_open_loop`` (imported from /root/reference by tests/golden/make_golden.py, fp32 and fp64);
tests/test_oracle_vs_golden.py checks this file against them.  Gradients are whatever torch.autograd gives for these
operations (clamp: pass-through on the closed interval; max(|vx|, vx_zero): the active branch) -- the CUDA VJP must match.
"""
import torch

from .dynamics import PARAMS

LIMIT_KEYS = (("x_min", "x_max"), ("y_min", "y_max"), ("phi_min", "phi_max"), ("vx_min", "vx_max"), ("vy_min", "vy_max"),
              ("omega_min", "omega_max"))        # KalmanNet/training.py:60-90, test_vehicle.py:86-93


def with_limits(lo, hi, p=PARAMS):
    q = dict(p)
    for i, (a, b) in enumerate(LIMIT_KEYS):
        q[a], q[b] = float(lo[i]), float(hi[i])
    return q


def f_cont(x, u, p):
    """pt_f_cont: x[B,6], u[B,2] -> xdot[B,6]."""
    phi = x[:, 2].clamp(p["phi_min"], p["phi_max"])
    vx = x[:, 3].clamp(p["vx_min"], p["vx_max"])
    vy = x[:, 4].clamp(p["vy_min"], p["vy_max"])
    om = x[:, 5].clamp(p["omega_min"], p["omega_max"])
    d, delta = u[:, 0], u[:, 1]
    # :26 builds the threshold with torch.tensor(p["vx_zero"]) -- a float32 scalar whatever the state dtype, so in fp64
    # the reference compares against float32(0.3) = 0.30000001192...; kept, it is the reference's arithmetic
    ve = torch.max(vx.abs(), torch.tensor(p["vx_zero"]))
    af = (-torch.atan2(om * p["lf"] + vy, ve) + delta).clamp(-p["maxAlpha"], p["maxAlpha"])
    ar = torch.atan2(om * p["lr"] - vy, ve)
    Fyf = p["Df"] * torch.sin(p["Cf"] * torch.atan(p["Bf"] * af))
    Fyr = p["Dr"] * torch.sin(p["Cr"] * torch.atan(p["Br"] * ar))
    Frx = (p["Cm1"] - p["Cm2"] * ve) * d - p["Cr0"] - p["Cr2"] * (ve ** 2)
    m, Iz, lf, lr = p["m"], p["Iz"], p["lf"], p["lr"]
    return torch.stack([vx * torch.cos(phi) - vy * torch.sin(phi),
                        vx * torch.sin(phi) + vy * torch.cos(phi),
                        om,
                        (Frx - Fyf * torch.sin(delta) + m * vy * om) / m,
                        (Fyr + Fyf * torch.cos(delta) - m * vx * om) / m,
                        (Fyf * lf * torch.cos(delta) - Fyr * lr) / Iz], dim=1)


def step(x, u, Ts, p):
    """VehicleModel.f on [B,6] / [B,2] tensors -> [B,6]."""
    nxt = x + Ts * f_cont(x, u, p)
    cols = [nxt[:, i].clamp(p[a], p[b]) for i, (a, b) in enumerate(LIMIT_KEYS)]
    return torch.stack(cols, dim=1)


def rollout(x0, U, t_start, H, Ts, p):
    """rollout_open_loop: x0[B,6], U[B,2,T] -> preds[B,6,Hn], Hn = min(H, T - t_start) (x0 itself when that is <= 0)."""
    x, out = x0, []
    for k in range(H):
        t = t_start + k
        if t >= U.shape[2]:
            break
        x = step(x, U[:, :, t], Ts, p)
        out.append(x)
    return torch.stack(out, dim=2) if out else x0.unsqueeze(2)
