"""GPU parity of the open-loop generator modes through the C ABI: the fused
control-synthesis + plant + sensor-noise kernels against oracle/openloop.py evaluated on the same Philox streams.
The oracle itself is pinned bit-exact against the reference's generators (tests/test_oracle_vs_golden.py)."""
import ctypes

import numpy as np
import pytest

import trajectory_generation_b200 as tg
from trajectory_generation_b200 import _lib
from oracle import dataset as ods, dynamics as dyn, openloop as ool, philox as oph

pytestmark = pytest.mark.gpu


def _rules_from_struct(cls, struct):
    kw = {}
    for name, _ in struct._fields_:
        if name == "reserved":
            continue
        v = getattr(struct, name)
        kw[name] = tuple(v) if isinstance(v, ctypes.Array) else v
    return cls(**kw)


def _noisy_expected(clean, traj_id0, std=oph.NOISE_STD):
    B, T1, _ = clean.shape
    return np.stack([clean[i] + oph.sensor_noise(traj_id0 + i, T1, std=std) for i in range(B)])


@pytest.mark.parametrize("Ts,T", [(0.01, 1200), (0.02, 301)])
def test_type1_matches_oracle(Ts, T):
    B, id0 = 37, 5
    gen = tg.OpenLoopGenerator("type1", Ts=Ts)
    x0 = gen.sample_x0(B, seed=3)
    res = gen.generate(x0, T, traj_id0=id0)
    rules = _rules_from_struct(ool.Type1Rules, gen.rules)
    seen = set()
    for i in range(B):
        U, X, mode = ool.type1_trajectory(ool.PhiloxType1Source(tg.CTRL_SEED_BASE + id0 + i), x0[i], T, Ts, rules)
        np.testing.assert_allclose(res["U"][i], U, rtol=0, atol=1e-12)
        assert res["modes"][i] == mode
        np.testing.assert_allclose(res["clean"][i], X, rtol=1e-8, atol=1e-8)
        seen.add(mode)
    assert seen == {0, 1}
    np.testing.assert_array_equal(res["clean"][:, 0], x0)
    np.testing.assert_allclose(res["noisy"], _noisy_expected(res["clean"], id0), rtol=5e-16, atol=1e-16)   # fma vs mul+add
    # the reference's statistics hold (generation_type1.py:250-251): slew limits and clip
    assert np.abs(np.diff(res["U"][:, :, 0], axis=1)).max() <= 0.1 + 1e-12
    assert np.abs(np.diff(res["U"][:, :, 1], axis=1)).max() <= 0.04 + 1e-12


def test_type1_many_knots_and_fixed_mode():
    """a transient with several spline knots (the natural-spline solve on the device) and mode forced."""
    Ts, T, B = 0.01, 700, 9
    kw = dict(transient_s=(4.0, 6.0), checkpoint_s=(0.5, 0.9), mode="sinusoid")
    gen = tg.OpenLoopGenerator("type1", Ts=Ts, ctrl_seed_base=1000, **kw)
    x0 = gen.sample_x0(B, seed=1)
    res = gen.generate(x0, T)
    rules = _rules_from_struct(ool.Type1Rules, gen.rules)
    assert rules.mode == 1
    for i in range(B):
        src = ool.PhiloxType1Source(1000 + i)
        U, mode = ool.type1_controls(src, T, Ts, rules)
        np.testing.assert_allclose(res["U"][i], U, rtol=0, atol=1e-11)
        assert res["modes"][i] == 1
    with pytest.raises(tg.TrajgenError):
        tg.OpenLoopGenerator("type1", Ts=Ts, transient_s=(4.0, 6.0), checkpoint_s=(0.05, 0.1)).generate(x0, T)   # > 16 knots


@pytest.mark.parametrize("Ts,T", [(0.01, 1200), (0.02, 300)])
def test_type2_matches_oracle(Ts, T):
    B, id0 = 41, 0
    gen = tg.OpenLoopGenerator("type2", Ts=Ts)
    x0 = gen.sample_x0(B, seed=42)
    np.testing.assert_array_equal(x0[:3], ods.sample_x0_type2(3, 42))
    res = gen.generate(x0, T, traj_id0=id0)
    rules = _rules_from_struct(ool.Type2Rules, gen.rules)
    counts = np.zeros(4, dtype=int)
    exact = 0
    for i in range(B):
        U, X, modes = ool.type2_trajectory(ool.PhiloxType2Source(tg.CTRL_SEED_BASE + id0 + i), x0[i], T, Ts, rules)
        # the machine's branches read the shadow state, whose last bits differ (libdevice vs libm); with the plant's slip angles
        # and tyre forces from the tables (1e-16 from libm) no threshold decision of these 41 trajectories flips: every one
        # must be the oracle's trajectory
        assert np.array_equal(res["modes"][i], modes), i
        exact += 1
        np.testing.assert_allclose(res["U"][i], U, rtol=0, atol=1e-9)
        np.testing.assert_allclose(res["clean"][i], X, rtol=1e-7, atol=1e-7)
        counts += np.bincount(res["modes"][i], minlength=4)
    assert exact == B
    assert (counts > 0).all()
    np.testing.assert_allclose(res["noisy"], _noisy_expected(res["clean"], id0), rtol=5e-16, atol=1e-16)   # fma vs mul+add
    U = res["U"]
    assert U[:, :, 0].min() >= 0.0 and U[:, :, 0].max() <= 0.33
    dd = np.diff(np.concatenate([np.zeros((B, 1)), U[:, :, 1]], axis=1), axis=1)
    assert np.abs(dd).max() <= 0.30 * Ts + 1e-12
    assert res["clean"][:, :, 3].min() >= 0.0 and np.abs(res["clean"][:, :, 5]).max() <= 6.0


def test_openloop_plant_consistency_and_shard_invariance():
    """clean == tg_plant_rollout(x0, U) exactly (same device code), and results do not depend on how the id range is cut."""
    T, B = 257, 70
    for kind in ("type1", "type2"):
        gen = tg.OpenLoopGenerator(kind, Ts=0.01)
        x0 = gen.sample_x0(B, seed=9)
        full = gen.generate(x0, T)
        plant = tg.ClosedLoopGenerator(N=5, Ts=0.01, plant=tg.PLANT_GEN1 if kind == "type1" else tg.PLANT_GEN2)
        np.testing.assert_array_equal(plant.plant_rollout(x0, full["U"]), full["clean"])
        a = gen.generate(x0[:33], T, traj_id0=0)
        b = gen.generate(x0[33:], T, traj_id0=33)
        for k in ("clean", "noisy", "U", "modes"):
            np.testing.assert_array_equal(np.concatenate([a[k], b[k]]), full[k])
        # optional outputs: asking for a subset gives the same numbers
        only = gen.generate(x0, T, want=("noisy",))
        assert set(only) == {"noisy"}
        np.testing.assert_array_equal(only["noisy"], full["noisy"])


def test_openloop_dataset_loads_through_the_reference_schema(tmp_path):
    """type-2 open-loop rows -> CSV (the reference's schema) -> the arrays KalmanNet's loader builds."""
    import pandas as pd
    gen = tg.OpenLoopGenerator("type2", Ts=0.01)
    res = gen.generate(gen.sample_x0(6), 120)
    c, n = tmp_path / "c.csv", tmp_path / "n.csv"
    tg.write_csv(res, 0.01, c, n)
    dc = pd.read_csv(c, float_precision="round_trip")
    dn = pd.read_csv(n, float_precision="round_trip")
    assert list(dc.columns) == tg.CLEAN_COLS and list(dn.columns) == tg.NOISY_COLS
    np.testing.assert_array_equal(dc[["X", "Y", "phi", "vx", "vy", "omega"]].to_numpy().reshape(6, 121, 6), res["clean"])
    np.testing.assert_array_equal(dc["d"].to_numpy().reshape(6, 121)[:, :-1], res["U"][:, :, 0])
    assert dc["d"].isna().sum() == 6
    y, u, x = tg.to_loader_tensors(res, 120)
    assert y.shape == (6, 5, 120) and u.shape == (6, 2, 120) and x.shape == (6, 6, 120)


def test_openloop_argument_errors():
    gen = tg.OpenLoopGenerator("type2", Ts=0.01)
    with pytest.raises(tg.TrajgenError):
        gen.generate(np.zeros((2, 6)), 0)                      # T >= 1
    with pytest.raises(TypeError):
        tg.OpenLoopGenerator("type2", no_such_rule=1.0)
    with pytest.raises(tg.TrajgenError):
        tg.OpenLoopGenerator("type2", p_modes=(0, 0, 0, 0)).generate(np.zeros((1, 6)), 4)
    with pytest.raises(tg.TrajgenError):
        tg.OpenLoopGenerator("type1", Ts=2.0).generate(np.zeros((1, 6)), 4)     # transient shorter than one step
    assert gen.generate(np.zeros((0, 6)), 5)["clean"].shape == (0, 6, 6)
