"""CPU: the oracle against the committed golden vectors (outputs of the reference's own code,
tests/golden/make_golden.py) and against itself where the reference cannot run (the QP)."""
import io
import os

import numpy as np
import pandas as pd
import pytest

from oracle import dataset as ods, dynamics as dyn, mpc as ompc, philox as oph, qp as oqp, refgen as R
from conftest import GOLDEN, HARD


def test_f_cont_three_variants(golden_physics):
    g = golden_physics
    for v in range(3):
        for i in range(len(g["X"])):
            f = dyn.f_cont(g["X"][i], g["U"][i], variant=v)
            np.testing.assert_allclose(f, g["F"][v, i], rtol=1e-12, atol=1e-12)


def test_survey_appendix_a_known_answers():
    f = dyn.f_cont([0.3, -0.2, 0.4, 1.2, 0.08, -0.7], [0.35, -0.12])
    np.testing.assert_allclose(f, [1.07411972541877, 0.5409868902906114, -0.7, 0.29365939871874336,
                                   -2.8548086485435813, -22.151010408831763], rtol=1e-13)
    x, u = [0, 0, -1.0, 0.2, 0.25, 3.0], [0.8, 0.5]
    assert abs(dyn.f_cont(x, u, variant=0)[3] - 6.589295682725033) < 1e-12      # MPC variant
    assert abs(dyn.f_cont(x, u, variant=1)[3] - 6.482527390042107) < 1e-12      # generator variants
    assert abs(dyn.f_cont(x, u, variant=2)[3] - 6.482527390042107) < 1e-12


def test_linearize_discretize(golden_physics):
    g = golden_physics
    for i in range(len(g["XL"])):
        A, B, c = dyn.linearize_discretize(g["XL"][i], g["UL"][i], float(g["TsL"][i]))
        np.testing.assert_allclose(A, g["AL"][i], atol=1e-9)
        np.testing.assert_allclose(B, g["BL"][i], atol=1e-9)
        np.testing.assert_allclose(c, g["GL"][i], atol=1e-9)


def test_linearize_horizon(golden_physics):
    g = golden_physics
    for i in range(4):
        A, B, c, xbar = dyn.linearize_horizon(g["XL"][i], g["UL"][i], 0.02, 20)
        np.testing.assert_allclose(A, g["AH"][i], atol=1e-9)
        np.testing.assert_allclose(B, g["BH"][i], atol=1e-9)
        np.testing.assert_allclose(c, g["GH"][i], atol=1e-9)
        np.testing.assert_allclose(xbar.T, g["XB"][i], atol=1e-12)


def test_reference_generators(golden_physics):
    g = golden_physics
    np.testing.assert_array_equal(R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), 40, 0.02), g["vr40"])
    np.testing.assert_array_equal(R.vref_profile(R.VREF_TRAPEZOID, (0.8, 2.0, 1.0, 1.5, 1.0), 40, 0.1), g["vtr"])
    np.testing.assert_array_equal(R.vref_profile(R.VREF_SINE, (1.5, 0.5, 6.0), 40, 0.1), g["vsn"])
    np.testing.assert_array_equal(R.ref_window(0.37, 40, 0.02, g["vr40"]), g["win"])
    np.testing.assert_allclose([R.d_steady_state(v) for v in (0.5, 1.0, 2.0)], g["dss"], rtol=1e-15)


def test_plants_with_clipping(golden_physics):
    g = golden_physics
    for plant, key in ((dyn.PLANT_MPC, "Xm"), (dyn.PLANT_GEN1, "Xg1"), (dyn.PLANT_GEN2, "Xg2")):
        x = g["x0p"].copy()
        X = [x]
        for k in range(len(g["Usim"])):
            x = dyn.plant_step(x, g["Usim"][k], 0.01, plant=plant)
            X.append(x)
        np.testing.assert_allclose(np.array(X), g[key], rtol=1e-10, atol=1e-10)
    assert (g["Xg2"][:, 3] == 0.0).any() or (np.abs(g["Xg2"][:, 5]) == 6.0).any()   # a clip really engaged


def test_dataset_shell_reproduces_reference_csv(tmp_path):
    """x0 draw, PCG64 noise, frame layout and CSV text of generation_type2 (3 traj x 0.5 s)."""
    ref_clean = pd.read_csv(os.path.join(GOLDEN, "reference_gen2_clean.csv"), float_precision="round_trip")
    ref_noisy_txt = open(os.path.join(GOLDEN, "reference_gen2_noisy.csv")).read()
    ref_clean_txt = open(os.path.join(GOLDEN, "reference_gen2_clean.csv")).read()
    x0 = ods.sample_x0_type2(3, 42)
    cf, nf = [], []
    for i in range(3):
        rows = ref_clean[ref_clean["trajectory_id"] == i]
        X = rows[["X", "Y", "phi", "vx", "vy", "omega"]].values
        U = rows[["d", "delta"]].values[:-1]
        np.testing.assert_array_equal(X[0], x0[i])
        # the plant restatement regenerates the truth rows from the reference's controls
        x = X[0].copy()
        for k in range(len(U)):
            x = dyn.plant_step(x, U[k], 0.01, plant=dyn.PLANT_GEN2)
            np.testing.assert_allclose(x, X[k + 1], rtol=1e-10, atol=1e-12)
        c, n = ods.trajectory_frames(X, U, i, 0.01, ods.pcg64_noise(i, len(X)))
        cf.append(c); nf.append(n)
    ods.write_csv(cf, nf, tmp_path / "c.csv", tmp_path / "n.csv")
    assert open(tmp_path / "c.csv").read() == ref_clean_txt
    assert open(tmp_path / "n.csv").read() == ref_noisy_txt


def test_numpy_noise_facts():
    rng = np.random.default_rng(12345)
    np.testing.assert_allclose(rng.normal(0, 0.05, 3), [-0.07119125, 0.06318642, -0.04353309], atol=1e-8)


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, e in kat:
        assert " ".join("%08x" % v for v in oph.philox4x32_10_py(c, k)) == e
    s = oph.philox_stream((7 << 32) | 12345, 3, 1, 4)          # C restatement == pure-Python rounds
    for i in range(4):
        assert tuple(int(v) for v in s[i]) == oph.philox4x32_10_py((3 + i, 1, 0, 0), (12345, 7))


def test_philox_normals_are_standard_normal():
    n = oph.standard_normals(12345, 100000)
    assert np.abs(n.mean(0)).max() < 0.02 and np.abs(n.std(0) - 1).max() < 0.02
    from scipy import stats
    assert min(stats.kstest(n[:, c], "norm").pvalue for c in range(6)) > 1e-3
    import math
    u1, u2 = (5 + 0.5) / 2 ** 32, (77 + 0.5) / 2 ** 32
    import ctypes
    a, b = ctypes.c_double(), ctypes.c_double()
    oph.lib().tgo_box_muller(5, 77, ctypes.byref(a), ctypes.byref(b))
    assert abs(a.value - math.sqrt(-2 * math.log(u1)) * math.cos(2 * math.pi * u2)) < 1e-13


# ------------------------------------------------------------------ the QP (parity unpinned: self-consistency)
def _case(N=20, hard=False, seed=0, Ts=0.02):
    rng = np.random.default_rng(seed)
    vx = rng.uniform(0.5, 1.5)
    x = np.array([rng.uniform(-1, 1), rng.uniform(-1.0, 1.0) if hard else rng.uniform(-0.2, 0.2), rng.uniform(-0.3, 0.3), vx, 0.0, 0.0])
    up = np.array([R.d_steady_state(vx), 0.0])
    v = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, Ts)
    pr = R.ref_window(x[0], N, Ts, v)
    return x, up, pr, v


@pytest.mark.parametrize("hard", [False, True])
def test_ipm_and_osqp_restatement_agree(hard):
    for seed in range(3):
        x, up, pr, v = _case(hard=hard, seed=seed)
        kw = HARD if hard else {}
        u1, s1, i1 = ompc.mpc_step(x, up, pr, vref=v, solver="ipm", **kw)
        u2, s2, i2 = ompc.mpc_step(x, up, pr, vref=v, solver="osqp", **kw)
        assert s1 == "optimal" and s2 == "optimal"
        assert np.abs(i1["U_opt"] - i2["U_opt"]).max() < 1e-3
        assert abs(i1["objective"] - i2["objective"]) < 1e-4 * (1 + abs(i1["objective"]))


def test_ipm_solution_satisfies_kkt():
    x, up, pr, v = _case(hard=True, seed=5)
    A, B, g, _ = dyn.linearize_horizon(x, up, 0.02, 20)
    prob = oqp.build_sparse_qp(x, up, A, B, g, pr, v, **HARD)
    z, y, st = oqp.solve_ipm(prob)
    assert st == "optimal"
    Az = prob.A @ z
    assert (Az >= prob.l - 1e-8).all() and (Az <= prob.u + 1e-8).all()
    # stationarity in the null space of the equalities: P z + q + A_in' y is orthogonal to ker(E)
    E = prob.A[: prob.n_eq]
    r = prob.P @ z + prob.q + prob.A[prob.n_eq:].T @ y
    nu = np.linalg.lstsq(E.T, -r, rcond=None)[0]
    assert np.abs(r + E.T @ nu).max() < 1e-6
    Ain = Az[prob.n_eq:]
    assert (y[(Ain > prob.l[prob.n_eq:] + 1e-6) & (Ain < prob.u[prob.n_eq:] - 1e-6)] ** 2).sum() < 1e-12   # complementarity


def test_fallback_and_status_strings():
    x, up, pr, v = _case()
    for solver in ("ipm", "osqp"):
        u, st, info = ompc.mpc_step(x, [2.0, 0.0], pr, vref=v, solver=solver)   # u_prev outside what the rate allows
        assert st == "infeasible" and info == {} and np.array_equal(u, [2.0, 0.0])     # MPC/mpc_6stati.py:261-262
    u, st, info = ompc.mpc_step(x, up, pr, vref=v, x_lo=[-1e20, -1e20, -1e20, 2.0, -1e20, -1e20])   # x0 violates k=0 row
    assert st == "infeasible" and info == {}
    u, st, info = ompc.mpc_step(x, up, pr, vref=None)
    assert st == "optimal" and set(info) >= {"status", "objective", "X_opt", "U_opt", "path_ref", "vref"}
    assert np.array_equal(info["vref"], np.full(21, x[3]))                       # :158-159
    with pytest.raises(AssertionError):
        ompc.mpc_step(x, up, pr[:-1], vref=v)                                    # :151


def test_golden_qp_fixtures_are_consistent(golden_qp):
    for c in golden_qp[:6]:
        kw = HARD if bool(c["hard"]) else {}
        u, st, info = ompc.mpc_step(c["x0"], c["u_prev"], c["path_ref"], Ts=float(c["Ts"]), N=int(c["N"]), vref=c["vref"],
                                    solver="osqp", **kw)
        assert st == "optimal" and np.abs(info["U_opt"] - c["U_opt"]).max() < 1e-3


def test_closed_loop_golden_prefix(golden_loop):
    g = golden_loop
    X, U, st, _ = ompc.closed_loop(g["x0"], g["u0"], 5, 0.02, 40)
    np.testing.assert_allclose(X, g["X40"][:6], atol=1e-7)
    d = g["U40"][:, 0]
    assert abs(d.mean() - 0.2161) < 5e-3        # soft anchor: generation_type1.py:250 (statistics of an MPC run)


# ------------------------------------------------------------------ open-loop generators
def test_type2_state_machine_reproduces_reference(golden_openloop):
    """oracle/openloop.py with the reference's own PCG64 streams == generation_type2.generate_dataset."""
    from oracle import openloop as ol
    g = golden_openloop
    for c in range(int(g["n_t2"])):
        n, T, Ts, seed = g[f"t2_{c}_meta"]
        n, T, seed = int(n), int(T), int(seed)
        X0 = ods.sample_x0_type2(n, seed)
        for i in range(n):
            U, X, modes = ol.type2_trajectory(ol.GeneratorSource(np.random.default_rng(seed + i)), X0[i], T, float(Ts))
            np.testing.assert_array_equal(U, g[f"t2_{c}_U"][i])
            np.testing.assert_array_equal(modes, g[f"t2_{c}_modes"][i])
            np.testing.assert_allclose(X, g[f"t2_{c}_X"][i], rtol=1e-10, atol=1e-10)
    assert len(np.unique(g["t2_0_modes"])) == 4          # every mode of the machine occurs in the fixture


def test_type1_profiles_reproduce_reference(golden_openloop):
    """oracle/openloop.py with the legacy MT19937 stream == the per-trajectory body of generation_type1's main loop."""
    from oracle import openloop as ol
    g = golden_openloop
    seen = set()
    for c in range(int(g["n_t1"])):
        n, T, Ts, seed = g[f"t1_{c}_meta"]
        n, T, seed = int(n), int(T), int(seed)
        rs = np.random.RandomState(seed)
        for i in range(n):
            x0 = np.array([rs.uniform(lo, hi) for lo, hi in ods.X0_RANGES_TYPE1])
            np.testing.assert_array_equal(x0, g[f"t1_{c}_x0"][i])
            U, X, mode = ol.type1_trajectory(ol.LegacyNumpySource(rs), x0, T, float(Ts))
            np.testing.assert_array_equal(U, g[f"t1_{c}_U"][i])
            assert mode == g[f"t1_{c}_modes"][i]
            np.testing.assert_allclose(X, g[f"t1_{c}_X"][i], rtol=1e-10, atol=1e-10)
            seen.add(int(mode))
    assert seen == {0, 1}


def test_philox_sources_keep_the_generators_contract():
    """The B200 path's Philox streams: same distributions / ranges as the reference's draws."""
    from oracle import openloop as ol
    r1, r2 = ol.Type1Rules(), ol.Type2Rules()
    modes, Us = [], []
    for i in range(24):
        U, mode = ol.type1_controls(ol.PhiloxType1Source(42 + i), 1200, 0.01, r1)
        modes.append(mode)
        Us.append(U)
        assert np.all(np.abs(np.diff(U[:, 0])) <= 0.1 + 1e-15) and np.all(np.abs(np.diff(U[:, 1])) <= 0.04 + 1e-15)
    U = np.concatenate(Us)
    assert 4 <= sum(modes) <= 20                                               # p(sinusoid) = 0.5
    assert abs(U[:, 0].mean() - r1.d_mean) < 0.02 and abs(U[:, 1].mean() - r1.delta_mean) < 0.01
    # independent streams per trajectory and reproducibility
    a = ol.type1_controls(ol.PhiloxType1Source(42), 300, 0.01, r1)[0]
    b = ol.type1_controls(ol.PhiloxType1Source(42), 300, 0.01, r1)[0]
    c = ol.type1_controls(ol.PhiloxType1Source(43), 300, 0.01, r1)[0]
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # type 2: bounds of the machine hold, every mode is reachable, turns are followed by straights
    counts = np.zeros(4, dtype=int)
    for i in range(6):
        x0 = ods.sample_x0_type2(6, 42)[i]
        U, X, m = ol.type2_trajectory(ol.PhiloxType2Source(42 + i), x0, 1200, 0.01, r2)
        counts += np.bincount(m, minlength=4)
        assert U[:, 0].min() >= r2.d_range[0] and U[:, 0].max() <= r2.d_range[1]
        assert np.all(np.abs(np.diff(np.concatenate([[0.0], U[:, 1]]))) <= r2.delta_rate_max * 0.01 + 1e-15)
        assert X[:, 3].min() >= 0.0 and np.abs(X[:, 5]).max() <= 6.0
        change = np.flatnonzero(np.diff(m) != 0)
        for k in change:
            assert not (m[k] >= 2 and m[k + 1] >= 2)                          # generation_type2.py:107-109
    assert (counts > 0).all()


# ------------------------------------------------------------------ estimator physics
def test_estimator_step_and_rollout_reproduce_reference(golden_estimator):
    """oracle/estimator.py == KalmanNet's VehicleModel.f (values and autograd gradients) and rollout_open_loop, bit for bit."""
    import torch
    from oracle import estimator as oe
    g = golden_estimator
    p = oe.with_limits(g["lo"], g["hi"])
    Ts = float(g["Ts"])
    T, H, t0 = (int(v) for v in g["roll_meta"])
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        x = torch.tensor(g["X"], dtype=dt, requires_grad=True)
        u = torch.tensor(g["U"], dtype=dt, requires_grad=True)
        y = oe.step(x, u, Ts, p)
        y.backward(torch.tensor(g["G"], dtype=dt))
        np.testing.assert_array_equal(y.detach().numpy(), g[f"next_{name}"])
        np.testing.assert_array_equal(x.grad.numpy(), g[f"gx_{name}"])
        np.testing.assert_array_equal(u.grad.numpy(), g[f"gu_{name}"])
        r = oe.rollout(torch.tensor(g["X"][:8], dtype=dt), torch.tensor(g["Useq"], dtype=dt), t0, H, Ts, p)
        assert r.shape == (8, 6, T - t0)                                    # the reference stops at the end of u
        np.testing.assert_array_equal(r.numpy(), g[f"roll_{name}"])
        np.testing.assert_array_equal(g[f"h_{name}"], g["X"].astype(g[f"h_{name}"].dtype)[:, [0, 1, 3, 4, 5]])
    # the fixture exercises every clamp and the low-speed branch
    assert (((g["X"] < g["lo"]) | (g["X"] > g["hi"])).sum(0) > 0).all() and (np.abs(g["X"][:, 3]) < 0.3).any()
