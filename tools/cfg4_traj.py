"""dev: one trajectory of the config-4 test (tests/test_gpu_configs.py) on the GPU and in the oracle, step by step"""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import trajectory_generation_b200 as tg
from oracle import mpc as ompc, refgen as R
from conftest import HARD
N, idx = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(N)
B, T, Ts = 48, 12, 0.02
x0 = np.zeros((B, 6)); x0[:, 1] = rng.uniform(-1.5, 1.5, B); x0[:, 3] = rng.uniform(0.8, 1.2, B)
x0[:2, 1] = (1.5, -1.2)
u0 = np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1)
sc = tg.Scenarios(B); sc.set_sine(slice(0, B), 0.5, 0.5, 0.0, 0.0)
for opts in (None, {"eps_abs": 1e-7, "eps_rel": 1e-7}):
    gen = tg.ClosedLoopGenerator(N=N, Ts=Ts, solver_opts=opts, **HARD) if opts else tg.ClosedLoopGenerator(N=N, Ts=Ts, **HARD)
    res = gen.generate(x0, u0, sc, T)
    print("opts", opts, "status counts", res["status_counts"][idx], "iters", res["iters_total"][idx])
Xo, Uo, st, _ = ompc.closed_loop(x0[idx], u0[idx], T, Ts, N, path_kind=R.PATH_SINE, path_prm=(0.5, 0.5, 0.0, 0.0), **HARD)
np.set_printoptions(precision=6, linewidth=200)
for t in range(T):
    print(t, st[t], "oracle U", Uo[t], "gpu U", res["U"][idx, t], "oracle x vy,om", Xo[t, 4:], "gpu x vy,om", res["clean"][idx, t, 4:])
