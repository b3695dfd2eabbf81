"""dev: phase breakdown of config 1 (B = 1, N = 40, Ts = 0.02, parabola; needs a -DTG_PHASE_TIMING build via TRAJGEN_LIB)"""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
from trajectory_generation_b200 import _lib
gen = tg.ClosedLoopGenerator(N=40, Ts=0.02)
x0 = np.array([[0, 0.5, 0, 1.0, 0, 0.0]]); u0 = np.array([[tg.d_steady_state(1.0), 0.0]])
gen.generate(x0, u0, tg.Scenarios(1), 5)
L = _lib.load(); L.tg_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.tg_debug_phases(None, 1)
T = 600
gen.generate(x0, u0, tg.Scenarios(1), T)
out = (ctypes.c_longlong * 16)(); L.tg_debug_phases(out, 0)
v = np.array(out[:16], dtype=float)
names = ["K1a rollout / window / noise", "K1b linearise + residuals", "-", "K2 condense + bounds + rho", "K3 build K + sweep", "ADMM tail", "exit", "ADMM init", "ADMM iterations", "ADMM checks"]
for n_, c in zip(names, v):
    if c > 0: print(f"  {n_:32s} {c/T:9.0f} cycles/step {100*c/v.sum():5.1f} %")
print(f"  body total {v.sum()/T:.0f} cycles/step")
