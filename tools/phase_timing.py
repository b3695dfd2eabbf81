"""dev: per-phase clock64 breakdown of the step body (needs a -DTG_PHASE_TIMING build via TRAJGEN_LIB)."""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import bench
import trajectory_generation_b200 as tg
from trajectory_generation_b200 import _lib
B, T = int(sys.argv[1]), int(sys.argv[2])
kw = dict(bench.GEN_KW)
if len(sys.argv) > 3:
    kw["N"] = int(sys.argv[3])
gen = tg.ClosedLoopGenerator(**kw)
x0, u0, sc = bench.make_workload(gen, B)
gen.generate(x0, u0, sc, 5)
L = _lib.load()
L.tg_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_longlong * 16)()
L.tg_debug_phases(None, 1)
import time; t = time.time(); res = gen.generate(x0, u0, sc, T); dt = time.time() - t
L.tg_debug_phases(out, 0)
v = np.array(out[:16], dtype=float)
names = ["rollout(+zero,sincos)", "linearise+resid", "K2 condense", "R-terms,bounds,rho", "buildK+sweep", "ADMM tail", "objective/exit", "ADMM init", "ADMM iterations", "ADMM check", "R-terms+dH", "bounds+rho", "H taps/copy", "K2 producer (G, W rows)", "K2 barrier wait", "K2 consumer (rank-3)"]
# trajectories that slot 0 of CTA 0 runs (N = 20 launch geometry: P problems per CTA, at most 2 CTAs of 4 per SM)
P = 4
while P > 1 and (B + P - 1) // P < gen.info()["num_sms"]:
    P //= 2
G = min((B + P - 1) // P, 2 * gen.info()["num_sms"] * (4 // P) if P > 1 else 10 ** 9)
ntraj0 = -(-B // (G * P))
steps = T * ntraj0
print(f"B={B} T={T} wall {dt*1e3:.1f} ms  -> {B*T/dt:.3e} steps/s; CTA0 ran {ntraj0} trajectories")
for nme, c in zip(names, v):
    print(f"  {nme:24s} {c/steps:10.0f} cycles/step  {100*c/v.sum():5.1f}%")
print(f"  body total {v.sum()/steps:.0f} cycles/step; wall per step {dt/T*1.965e9/ntraj0:.0f} cycles")
