"""Oracle: the open-loop control synthesis of the two generator scripts.

Test infrastructure -- see ``oracle/__init__.py``.  Restates, in fp64 NumPy/SciPy,

* generation_type1.py:86-137  ``create_spline_signal`` / ``generate_smooth_profiles`` / ``apply_du_bounds`` and the
  per-trajectory body of its ``__main__`` loop (:279-292): transient natural-cubic-spline segment, steady segment
  ("straight" or "sinusoid"), high-frequency control noise, slew limiter, clipping, plant integration;
* generation_type2.py:95-157  ``sample_controls_piecewise``: the accelerate/cruise/turn finite-state machine whose
  decisions read a shadow simulation of the vehicle (which is the ground-truth integration of :180-187 itself:
  same x0, same controls, same Euler step, same clipping).

Every random number is taken from a *draw source*, so that one restatement serves two purposes:

  PINNED:  with the NumPy sources (the legacy global ``np.random`` stream of generation_type1.py:17, resp.
           ``default_rng(seed + i)`` of generation_type2.py:166) the functions consume the streams in the
           reference's own call order and reproduce the reference's outputs bit for bit
           (tests/test_oracle_vs_golden.py against tests/golden/reference_openloop.npz, made by running the
           reference's code).
  B200:    with the Philox sources the same functions define what the CUDA kernels must produce.  A sequential
           generator cannot be evaluated per (trajectory, step) on a GPU, so -- exactly as for the sensor noise
           (philox_ref.c) -- the contract is kept (distributions, ranges, seed = base + trajectory id) and the
           bit stream is replaced by counter-based Philox4x32-10:
             key = ctrl_seed_base + trajectory_id, counter = (index, block, 0, 0), u = (r + 0.5) 2^-32,
             normals = Box-Muller on word pairs (0,1) and (2,3).
           type 1 blocks: 0x10 scalars (index 0: mode, transient, checkpoint, period; index 1: amplitude, phase),
                          0x11 knot k: (z_d, z_delta);  0x12 step t: (z_d_steady, z_delta_steady, z_noise_d, z_noise_delta)
           type 2 blocks: 0x20 segment s: (choice, seg_len, d, magnitude);  0x21 segment s: BM(0,1) = (z_delta, z_stall_delta), word 2 = stall d
"""
from dataclasses import dataclass, field

import numpy as np
from scipy.interpolate import CubicSpline

from . import dynamics as dyn, philox

MODE_STRAIGHT, MODE_SINUSOID = 0, 1                               # generation_type1.py:108
MODE_ACCELERATE, MODE_CRUISE, MODE_TURN_LEFT, MODE_TURN_RIGHT = 0, 1, 2, 3   # generation_type2.py:108-112
TYPE2_MODE_NAMES = ("accelerate", "cruise", "turn_left", "turn_right")
TYPE1_MODE_NAMES = ("straight", "sinusoid")
CTRL_SEED_BASE = 42                                               # generation_type1.py:17, generation_type2.py:292


@dataclass
class Type1Rules:
    """Constants of generation_type1.py (:90, :108-128, :250-251, :283-289)."""
    d_mean: float = 0.2161
    d_std: float = 0.1314
    delta_mean: float = 0.0035
    delta_std: float = 0.0338
    du_lo: tuple = (-0.1, -0.04)
    du_hi: tuple = (0.1, 0.04)
    u_lo: tuple = (-1.0, -0.6)
    u_hi: tuple = (1.0, 0.6)
    transient_s: tuple = (1.5, 3.0)
    checkpoint_s: tuple = (3.0, 5.0)
    period_s: tuple = (4.0, 8.0)
    amp_frac: tuple = (0.5, 1.5)
    p_straight: float = 0.5
    tr_d_frac: float = 0.2
    tr_delta_frac: float = 0.3
    st_d_frac: float = 0.05
    sin_noise_frac: float = 0.1
    straight_frac: float = 0.01
    ctrl_noise_frac: float = 0.1
    mode: int = -1            # -1 = 'random' (:108), else MODE_STRAIGHT / MODE_SINUSOID


@dataclass
class Type2Rules:
    """ControlRules + the literals of sample_controls_piecewise (generation_type2.py:21-30, :97-131)."""
    v_turn_max: float = 1.2
    v_high: float = 4.0
    d_range: tuple = (0.0, 0.33)
    delta_turn_range: tuple = (0.015, 0.04)
    delta_straight_noise: float = 0.004
    delta_rate_max: float = 0.30
    v_floor: float = 0.35
    d_boost_min: float = 0.15
    seg_s: tuple = (0.4, 1.5)
    p_modes: tuple = (0.35, 0.35, 0.15, 0.15)
    p_after_turn: tuple = (0.5, 0.5)
    acc_d_lo: float = 0.3
    cruise_d: tuple = (-0.05, 0.2)
    turn_d_fast: tuple = (0.0, 0.15)
    turn_d_slow: tuple = (0.05, 0.25)
    stall_v: float = 0.5
    stall_d: tuple = (0.5, 1.0)
    stall_min_s: float = 0.3
    delta_clip: float = 0.6


# ------------------------------------------------------------------------------------------------ draw sources
class LegacyNumpySource:
    """generation_type1.py's stream: the legacy global np.random (MT19937), consumed in call order."""

    def __init__(self, random_state):
        self.rs = random_state

    def choice(self, name, p):
        return int(self.rs.choice(len(p), p=list(p)))

    def uniform(self, name, lo, hi):
        return self.rs.uniform(lo, hi)

    def normal(self, name, mu, sd, n, first=0):
        return self.rs.normal(mu, sd, n)


class GeneratorSource:
    """generation_type2.py's stream: np.random.default_rng(seed + i) (PCG64), consumed in call order."""

    def __init__(self, rng):
        self.rng = rng

    def segment(self, s):
        pass

    def choice(self, name, p):
        return int(self.rng.choice(len(p), p=list(p)))

    def uniform(self, name, lo, hi):
        return self.rng.uniform(lo, hi)

    def normal(self, name, mu, sd):
        return self.rng.normal(mu, sd)


def _choice_from_u(u, p):
    """Generator.choice / RandomState.choice with p: cdf = cumsum(p) / cdf[-1]; searchsorted(u, 'right')."""
    cdf = np.cumsum(np.asarray(p, dtype=float))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


class PhiloxType1Source:
    """The B200 path's stream for type-1 control synthesis (layout in the module docstring)."""
    SCALARS = {"mode": (0, 0), "transient": (0, 1), "checkpoint": (0, 2), "period": (0, 3), "amplitude": (1, 0), "phase": (1, 1)}
    VECTORS = {"d_chk": (0x11, 0), "delta_chk": (0x11, 1), "d_st": (0x12, 0), "delta_st": (0x12, 1),
               "noise_d": (0x12, 2), "noise_delta": (0x12, 3)}

    def __init__(self, seed):
        self.seed = int(seed)
        self.sc = philox.uniform01(philox.philox_stream(self.seed, 0, 0x10, 2))

    def choice(self, name, p):
        return _choice_from_u(self.sc[self.SCALARS[name]], p)

    def uniform(self, name, lo, hi):
        return lo + (hi - lo) * self.sc[self.SCALARS[name]]

    def normal(self, name, mu, sd, n, first=0):
        block, word = self.VECTORS[name]
        raw = philox.philox_stream(self.seed, first, block, n)
        pair = 2 * (word // 2)
        z = philox.box_muller(raw[:, pair], raw[:, pair + 1])[word % 2]
        return mu + sd * z


class PhiloxType2Source:
    """The B200 path's stream for the type-2 state machine (one counter pair per segment)."""
    UNIFORMS = {"seglen": ("a", 1), "d": ("a", 2), "mag": ("a", 3), "stall_d": ("b", 2)}
    NORMALS = {"delta": 0, "stall_delta": 1}

    def __init__(self, seed):
        self.seed = int(seed)

    def segment(self, s):
        self.a = philox.philox_stream(self.seed, s, 0x20, 1)[0]
        self.b = philox.philox_stream(self.seed, s, 0x21, 1)[0]
        z0, z1 = philox.box_muller(self.b[0:1], self.b[1:2])
        self.z = (float(z0[0]), float(z1[0]))

    def choice(self, name, p):
        return _choice_from_u(float(philox.uniform01(self.a[0])), p)

    def uniform(self, name, lo, hi):
        which, w = self.UNIFORMS[name]
        return lo + (hi - lo) * float(philox.uniform01((self.a if which == "a" else self.b)[w]))

    def normal(self, name, mu, sd):
        return mu + sd * self.z[self.NORMALS[name]]


# ------------------------------------------------------------------------------------------------ type 1
def slew_limit(u, du_lo, du_hi):
    """apply_du_bounds, generation_type1.py:131-137: out[k] = out[k-1] + clip(u[k] - out[k-1], lo, hi)."""
    out = np.empty_like(u)
    out[0] = u[0]
    for k in range(1, len(u)):
        out[k] = out[k - 1] + min(max(u[k] - out[k - 1], du_lo), du_hi)
    return out


def type1_transient(src, n_steps, Ts, r):
    """create_spline_signal, generation_type1.py:86-102 -> (d[n_steps], delta[n_steps], knot indices)."""
    if n_steps <= 1:
        return np.array([r.d_mean]), np.array([r.delta_mean]), np.zeros(1, dtype=int)
    every = max(1, int(round(src.uniform("checkpoint", *r.checkpoint_s) / Ts)))
    knots = np.arange(0, n_steps, every)
    if knots[-1] != n_steps - 1:
        knots = np.append(knots, n_steps - 1)
    d_k = src.normal("d_chk", r.d_mean, r.d_std * r.tr_d_frac, len(knots))
    delta_k = src.normal("delta_chk", r.delta_mean, r.delta_std * r.tr_delta_frac, len(knots))
    t = np.arange(n_steps)
    return CubicSpline(knots, d_k, bc_type="natural")(t), CubicSpline(knots, delta_k, bc_type="natural")(t), knots


def type1_profiles(src, T, Ts, r=Type1Rules()):
    """generate_smooth_profiles, generation_type1.py:104-129 -> (d[T], delta[T], mode code)."""
    mode = r.mode if r.mode >= 0 else src.choice("mode", (r.p_straight, 1.0 - r.p_straight))
    n_tr = min(int(src.uniform("transient", *r.transient_s) / Ts), T)
    n_st = T - n_tr
    d_tr, delta_tr, _ = type1_transient(src, n_tr, Ts, r)
    if n_st <= 0:
        return d_tr, delta_tr, mode
    d_st = src.normal("d_st", r.d_mean, r.d_std * r.st_d_frac, n_st, first=n_tr)
    if mode == MODE_SINUSOID:
        t_st = np.arange(n_st) * Ts
        period = src.uniform("period", *r.period_s)
        amplitude = src.uniform("amplitude", r.delta_std * r.amp_frac[0], r.delta_std * r.amp_frac[1])
        phase = src.uniform("phase", 0, 2 * np.pi)
        w = 2 * np.pi / period
        wave = amplitude * np.sin(w * t_st + phase)
        jitter = src.normal("delta_st", 0, r.delta_std * r.sin_noise_frac, n_st, first=n_tr)
        delta_st = r.delta_mean + wave + jitter
    else:
        delta_st = src.normal("delta_st", r.delta_mean, r.delta_std * r.straight_frac, n_st, first=n_tr)
    return np.concatenate([d_tr, d_st]), np.concatenate([delta_tr, delta_st]), mode


def type1_controls(src, T, Ts, r=Type1Rules()):
    """The per-trajectory control block of generation_type1.py:279-290 -> (U[T,2], mode code)."""
    d_clean, delta_clean, mode = type1_profiles(src, T, Ts, r)
    noise_d = src.normal("noise_d", 0, r.d_std * r.ctrl_noise_frac, T)
    noise_delta = src.normal("noise_delta", 0, r.delta_std * r.ctrl_noise_frac, T)
    d = np.clip(slew_limit(d_clean + noise_d, r.du_lo[0], r.du_hi[0]), r.u_lo[0], r.u_hi[0])
    delta = np.clip(slew_limit(delta_clean + noise_delta, r.du_lo[1], r.du_hi[1]), r.u_lo[1], r.u_hi[1])
    return np.stack([d, delta], axis=1), mode


def plant_rollout(x0, U, Ts, plant, p=dyn.PARAMS):
    """simulate_trajectory, generation_type1.py:70-84 / generation_type2.py:180-187 -> X[T+1,6]."""
    X = np.empty((len(U) + 1, 6))
    X[0] = x0
    for k in range(len(U)):
        X[k + 1] = dyn.plant_step(X[k], U[k], Ts, p, plant)
    return X


def type1_trajectory(src, x0, T, Ts, r=Type1Rules(), plant=dyn.PLANT_GEN1, p=dyn.PARAMS):
    """-> (U[T,2], X[T+1,6], mode): controls + ground-truth integration (generation_type1.py:279-292)."""
    U, mode = type1_controls(src, T, Ts, r)
    return U, plant_rollout(x0, U, Ts, plant, p), mode


# ------------------------------------------------------------------------------------------------ type 2
def type2_trajectory(src, x0, T, Ts, r=Type2Rules(), plant=dyn.PLANT_GEN2, p=dyn.PARAMS):
    """sample_controls_piecewise, generation_type2.py:95-157 -> (U[T,2], X[T+1,6], modes[T] int8).
    X is the shadow state history, which equals the ground truth of :180-187."""
    U = np.zeros((T, 2))
    modes = np.zeros(T, dtype=np.int8)
    X = np.empty((T + 1, 6))
    X[0] = np.asarray(x0, dtype=float)
    xs = X[0].copy()
    prev_delta, prev_mode, i, seg = 0.0, -1, 0, 0
    step_max = r.delta_rate_max * Ts
    while i < T:
        src.segment(seg)
        seg += 1
        if prev_mode in (MODE_TURN_LEFT, MODE_TURN_RIGHT):
            mode = src.choice("mode", r.p_after_turn)          # accelerate / cruise only
        else:
            mode = src.choice("mode", r.p_modes)
        seg_len = max(1, int(np.round(src.uniform("seglen", *r.seg_s) / Ts)))
        v = np.hypot(xs[3], xs[4])
        if mode == MODE_ACCELERATE:
            if v >= r.v_high:
                mode = MODE_CRUISE
            d = src.uniform("d", r.acc_d_lo, r.d_range[1])
            delta = src.normal("delta", 0.0, r.delta_straight_noise)
        elif mode == MODE_CRUISE:
            d = src.uniform("d", *r.cruise_d)
            delta = src.normal("delta", 0.0, r.delta_straight_noise)
        else:
            d = src.uniform("d", *r.turn_d_fast) if v > r.v_turn_max else src.uniform("d", *r.turn_d_slow)
            mag = src.uniform("mag", *r.delta_turn_range)
            scale = min(1.0, r.v_turn_max / max(v, 1e-3))
            delta = (mag if mode == MODE_TURN_LEFT else -mag) * scale
        if v < r.stall_v:
            mode = MODE_ACCELERATE
            d = src.uniform("stall_d", *r.stall_d)
            delta = src.normal("stall_delta", 0.0, r.delta_straight_noise)
            seg_len = max(seg_len, int(round(r.stall_min_s / Ts)))
        for _ in range(seg_len):
            if i >= T:
                break
            if np.hypot(xs[3], xs[4]) < r.v_floor:
                d = max(d, r.d_boost_min)
            d_k = float(min(max(d, r.d_range[0]), r.d_range[1]))
            delta_k = float(min(max(delta, -r.delta_clip), r.delta_clip))
            delta_k = float(min(max(delta_k, prev_delta - step_max), prev_delta + step_max))
            prev_delta = delta_k
            U[i] = (d_k, delta_k)
            modes[i] = mode
            xs = dyn.plant_step(xs, U[i], Ts, p, plant)
            X[i + 1] = xs
            i += 1
        prev_mode = mode
    return U, X, modes
