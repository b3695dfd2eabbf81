"""Arclength-parameterised reference paths (the window construction of MPC/main.py:51-68 carried over to
paths that are not graphs over X).  CPU: the oracle restatement's properties and its golden fixture
(tests/golden/make_arc_golden.py).  GPU (-m gpu): the CUDA window and the fused closed loop through a 180-degree turn
against the oracle."""
import os

import numpy as np
import pytest

from oracle import mpc as ompc, refgen as R
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "oracle_arc_uturn.npz"))


def _arc(g):
    return g["breaks"], g["coef_x"], g["coef_y"]


# ------------------------------------------------------------------------------------------------------------------ CPU
def test_oracle_arc_window_on_a_circle():
    """a circle of radius 2 sampled finely: window points stay on the circle, advance by vref Ts of arclength, and the
    heading reference keeps growing past pi (unwrapped) instead of jumping."""
    th = np.linspace(0, 2 * np.pi * 0.9, 200)
    arc = R.arc_spline_tables(2 * np.sin(th), 2 - 2 * np.cos(th))
    N, Ts = 40, 0.05
    v = np.full(N + 1, 2.0)
    pose = np.array([2 * np.sin(2.0) * 1.05, 2 - 2 * np.cos(2.0) * 1.05, 2.0 + 0.1, 2.0, 0, 0])   # near theta = 2 rad
    w, s0 = R.ref_window_arc(pose, 3.5, N, Ts, v, arc)
    assert abs(s0 - 4.0) < 2e-3                                   # closest point: arclength = radius * theta
    np.testing.assert_allclose(np.hypot(w[:, 0], w[:, 1] - 2.0), 2.0, atol=1e-4)
    np.testing.assert_allclose(np.hypot(np.diff(w[:, 0]), np.diff(w[:, 1])), 2.0 * Ts, rtol=2e-3)
    np.testing.assert_allclose(w[:, 2], 2.0 + np.arange(N + 1) * 2.0 * Ts / 2.0, atol=2e-3)   # theta grows through pi, no wrap
    assert w[-1, 2] > np.pi
    # the unwrap follows the vehicle's own (unwrapped) heading
    pose[2] += 4 * np.pi
    w2, _ = R.ref_window_arc(pose, 3.5, N, Ts, v, arc)
    np.testing.assert_allclose(w2[:, 2], w[:, 2] + 4 * np.pi, atol=1e-12)


def test_oracle_arc_golden_is_reproducible_and_tracks_the_path(g):
    arc = _arc(g)
    N, Ts = int(g["N"]), float(g["Ts"])
    X, U, st, _ = ompc.closed_loop(g["x0"], g["u0"], 25, Ts, N, path_kind=R.PATH_ARC, path_prm=(0.0, 0, 0, 0), arc=arc,
                                   vref_kind=R.VREF_CONST, vref_prm=(1.0,), solver="ipm")
    np.testing.assert_allclose(X, g["X"][:26], atol=1e-9)
    np.testing.assert_allclose(U, g["U"][:25], atol=1e-8)
    for i in range(len(g["poses"])):
        w, s0 = R.ref_window_arc(g["poses"][i], g["guesses"][i], N, Ts, g["vref"], arc)
        np.testing.assert_allclose(w, g["windows"][i], atol=1e-12)
    # the stored loop turned through 180 degrees and ended on the return leg, heading pi (unwrapped), within 2 cm of the path
    Xg = g["X"]
    assert abs(Xg[-1, 2] - np.pi) < 1e-3 and abs(Xg[-1, 1] - 2.4) < 1e-3 and Xg[-1, 0] < 0
    d = []
    for k in range(50, len(Xg), 10):
        s = R.arc_project(arc, 0.02 * k, Xg[k, 0], Xg[k, 1], iters=30)
        d.append(np.hypot(R.arc_eval(arc, s)[0] - Xg[k, 0], R.arc_eval(arc, s)[1] - Xg[k, 1]))
    assert max(d) < 0.03


def test_scenarios_set_arc_table_layout(g):
    import trajectory_generation_b200 as tg
    sc = tg.Scenarios(2)
    sc.set_spline(0, np.array([0.0, 1.0, 2.0, 3.0]), np.array([0.0, 0.2, -0.1, 0.0]))
    sc.set_arc(1, g["px"], g["py"], s_start=0.5)
    brk, coef = sc.tables()
    f, K = int(sc.spec["spline_first"][1]), int(sc.spec["spline_count"][1])
    assert f == 3 and len(brk) == len(coef) == 3 + 2 * K and sc.spec["path"][1, 0] == 0.5
    np.testing.assert_allclose(brk[f:f + K], g["breaks"]); np.testing.assert_allclose(brk[f + K:f + 2 * K], g["breaks"])
    np.testing.assert_allclose(coef[f:f + K], g["coef_x"]); np.testing.assert_allclose(coef[f + K:], g["coef_y"])


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_arc_window_matches_oracle(g):
    import trajectory_generation_b200 as tg
    N, Ts = int(g["N"]), float(g["Ts"])
    gen = tg.ClosedLoopGenerator(N=N, Ts=Ts)
    P = len(g["poses"])
    sc = tg.Scenarios(P)
    for i in range(P):
        sc.set_arc(i, g["px"], g["py"], s_start=g["guesses"][i])
    pr, vr = gen.ref_window(g["poses"], sc)
    np.testing.assert_allclose(vr, np.tile(g["vref"], (P, 1)), atol=1e-14)
    np.testing.assert_allclose(pr, g["windows"], atol=1e-11)


@pytest.mark.gpu
def test_closed_loop_through_a_u_turn_matches_oracle(g):
    """fused closed loop on the U-turn path: 450 steps, the heading reference passes through pi; vs the oracle's exact-optimum
    loop (tolerance 1e-4, the bar of the benchmarked-configuration test), and inside a mixed batch next to graph paths."""
    import trajectory_generation_b200 as tg
    N, Ts, T = int(g["N"]), float(g["Ts"]), int(g["T"])
    gen = tg.ClosedLoopGenerator(N=N, Ts=Ts)
    sc = tg.Scenarios(3)
    sc.set_arc(1, g["px"], g["py"], s_start=0.0)
    sc.set_spline(2, np.array([-1.0, 1.0, 3.0, 9.0, 30.0]), np.array([0.0, 0.3, -0.2, 0.1, 0.0]))
    for i in range(3):
        sc.set_vref(i, tg.VREF_CONST, 1.0)
    x0 = np.tile(g["x0"], (3, 1)); u0 = np.tile(g["u0"], (3, 1))
    res = gen.generate(x0, u0, sc, T)
    assert res["status_counts"][1, 0] == T
    assert np.abs(res["clean"][1] - g["X"]).max() <= 1e-4
    assert np.abs(res["U"][1] - g["U"]).max() <= 1e-4
    assert abs(res["clean"][1, -1, 2] - np.pi) < 1e-3
    alone = gen.generate(x0[1:2], u0[1:2], sc.slice(1, 2), T)
    assert np.array_equal(alone["clean"][0], res["clean"][1])
