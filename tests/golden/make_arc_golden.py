"""Generates tests/golden/oracle_arc_uturn.npz: the oracle's closed loop (exact optimum per step, oracle/mpc.py) on an
arclength-parameterised U-turn path (a reference that is not a graph over X), plus reference windows
at a few poses.  NumPy / SciPy only (no GPU, no /root/reference).  About half a minute.

    python tests/golden/make_arc_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mpc as ompc, refgen as R   # noqa: E402

N, TS, T = 20, 0.02, 450


def uturn_waypoints(L=2.0, r=1.2, ds=0.3):
    """straight east, half circle to the left, straight back west"""
    s1 = np.arange(0, L, ds); a = np.arange(0, np.pi, ds / r); s2 = np.arange(0, L + 1.0, ds)
    return np.concatenate([s1, L + r * np.sin(a), L - s2]), np.concatenate([0 * s1, r - r * np.cos(a), 2 * r + 0 * s2])


def main():
    px, py = uturn_waypoints()
    arc = R.arc_spline_tables(px, py)
    x0 = np.array([0.1, 0.15, 0.05, 1.0, 0.0, 0.0]); u0 = np.array([R.d_steady_state(1.0), 0.0])
    X, U, st, _ = ompc.closed_loop(x0, u0, T, TS, N, path_kind=R.PATH_ARC, path_prm=(0.0, 0, 0, 0), arc=arc,
                                   vref_kind=R.VREF_CONST, vref_prm=(1.0,), solver="ipm")
    assert set(st) == {"optimal"}
    # windows: poses off the path all along it, each searched from a rough guess of s
    rng = np.random.default_rng(7)
    s_true = np.linspace(0.2, 8.5, 12)
    poses, guesses, wins, s0s = [], [], [], []
    v = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, TS)
    for s in s_true:
        x, y, dx, dy = R.arc_eval(arc, s)
        th = np.arctan2(dy, dx)
        off = rng.uniform(-0.25, 0.25)
        pose = np.array([x - off * np.sin(th), y + off * np.cos(th), th + rng.uniform(-0.3, 0.3) + 2 * np.pi * rng.integers(-1, 2), 1.0, 0, 0])
        g = s + rng.uniform(-0.3, 0.3)
        w, s0 = R.ref_window_arc(pose, g, N, TS, v, arc)
        poses.append(pose); guesses.append(g); wins.append(w); s0s.append(s0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_arc_uturn.npz"), px=px, py=py, breaks=arc[0], coef_x=arc[1],
                        coef_y=arc[2], x0=x0, u0=u0, X=X, U=U, N=N, Ts=TS, T=T, poses=np.array(poses), guesses=np.array(guesses),
                        windows=np.array(wins), s0=np.array(s0s), vref=v)
    print("final state", X[-1])


if __name__ == "__main__":
    main()
