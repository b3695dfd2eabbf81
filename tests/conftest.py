import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_physics():
    return np.load(os.path.join(GOLDEN, "reference_physics.npz"))


@pytest.fixture(scope="session")
def golden_qp():
    z = np.load(os.path.join(GOLDEN, "oracle_qp.npz"))
    n = int(z["n"])
    keys = ("N", "Ts", "hard", "x0", "u_prev", "path_ref", "vref", "u_cmd", "U_opt", "X_opt", "objective", "y")
    return [{k: z[f"{k}_{i}"] for k in keys} for i in range(n)]


@pytest.fixture(scope="session")
def golden_openloop():
    return np.load(os.path.join(GOLDEN, "reference_openloop.npz"))


@pytest.fixture(scope="session")
def golden_estimator():
    return np.load(os.path.join(GOLDEN, "reference_estimator.npz"))


@pytest.fixture(scope="session")
def golden_loop():
    return np.load(os.path.join(GOLDEN, "oracle_closed_loop.npz"))


HARD = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)), x_lo=[-1e20] * 4 + [-0.15, -2.0], x_hi=[1e20] * 4 + [0.15, 2.0])
