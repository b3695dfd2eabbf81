"""CPU: the C-ABI library loads, exports every symbol include/trajgen.h declares, the ctypes mirror of
tg_config has the C layout, and the product path fails loudly (no fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import trajectory_generation_b200 as tg
from trajectory_generation_b200 import _lib
from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "trajgen.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/trajgen.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert L.tg_version() == 100


def test_default_config_matches_mpc_step_defaults():
    c = _lib.default_config()            # MPC/mpc_6stati.py:124-140
    assert (c.N, c.Ts) == (20, 0.02)
    assert (c.q_c, c.q_phi, c.q_vx) == (6.0, 0.5, 0.5)
    assert list(c.R) == [0.02, 0, 0, 2.0] and list(c.Rd) == [0.01, 0, 0, 5.0]
    assert list(c.u_lo) == [-1.0, -0.6] and list(c.u_hi) == [1.0, 0.6]
    assert list(c.du_lo) == [-0.5, -0.3] and list(c.du_hi) == [0.5, 0.3]
    assert all(v <= -1e20 for v in c.x_lo) and all(v >= 1e20 for v in c.x_hi)
    assert dict(zip(_lib.PARAM_ORDER, c.params)) == tg.Params
    assert list(c.noise_std) == [0.05, 0.05, 0.003, 0.010, 0.003, 0.030] and c.noise_seed_base == 12345
    assert ctypes.sizeof(_lib.TgConfig) % 8 == 0


def test_make_config_mirrors_keyword_arguments():
    c = tg.make_config(Ts=0.01, N=40, params={"m": 0.05}, q_c=3.0, R=np.diag([0.1, 1.0]), du_bounds=((-0.1, 0.1), (-0.04, 0.04)),
                       x_lo=[-np.inf, -np.inf, -np.inf, 0.0, -0.15, -2.0], solver_opts={"eps_abs": 1e-6, "max_iter": 500})
    assert (c.N, c.Ts, c.q_c, c.max_iter, c.eps_abs) == (40, 0.01, 3.0, 500, 1e-6)
    assert c.params[_lib.PARAM_ORDER.index("m")] == 0.05 and c.params[0] == 0.287
    assert list(c.du_hi) == [0.1, 0.04] and c.x_lo[0] == -1e20 and c.x_lo[4] == -0.15
    with pytest.raises(ValueError):
        tg.make_config(solver_opts={"nope": 1})


def test_ref_spec_layout():
    assert _lib.REF_SPEC_DTYPE.itemsize == 96
    assert _lib.REF_SPEC_DTYPE.fields["path"][1] == 16 and _lib.REF_SPEC_DTYPE.fields["vref"][1] == 48


def test_no_gpu_fails_loudly():
    n = ctypes.c_int(0)
    rc = _lib.load().tg_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(tg.TrajgenError):
        tg.BatchedMPC()
    with pytest.raises(tg.TrajgenError):           # the shim does not fall back to any CPU path either
        tg.mpc_step(np.zeros(6), np.zeros(2), np.zeros((21, 3)))


def test_openloop_rule_defaults_are_the_generators_constants():
    """tg_default_type{1,2}_rules == the constants of generation_type{1,2}.py as oracle/openloop.py (pinned) states them."""
    from oracle import openloop as ool
    for struct, ref in ((tg.type1_rules(), ool.Type1Rules()), (tg.type2_rules(), ool.Type2Rules())):
        for name, _ in struct._fields_:
            if name == "reserved":
                continue
            v = getattr(struct, name)
            assert (tuple(v) if isinstance(v, ctypes.Array) else v) == getattr(ref, name), name
        assert ctypes.sizeof(struct) % 8 == 0
    r = tg.type1_rules(du_bounds=((-0.2, 0.3), (-0.01, 0.02)), mode="sinusoid", period_s=(2.0, 3.0))
    assert (list(r.du_lo), list(r.du_hi), r.mode, list(r.period_s)) == ([-0.2, -0.01], [0.3, 0.02], 1, [2.0, 3.0])
    with pytest.raises(TypeError):
        tg.type2_rules(bogus=1)
    with pytest.raises(ValueError):
        tg.type2_rules(p_modes=(1, 2))
    n = ctypes.c_int(0)
    if not (_lib.load().tg_device_count(ctypes.byref(n)) == 0 and n.value > 0):
        with pytest.raises(tg.TrajgenError):
            tg.OpenLoopGenerator("type1")
