"""dev: phase breakdown + workload statistics of BASELINE config 4 variants (needs a -DTG_PHASE_TIMING build via TRAJGEN_LIB for the
phases).  usage: python tools/cfg4_phase.py N B T [y0max vybox ombox]"""
import ctypes, sys, time
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
from trajectory_generation_b200 import _lib
N, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
y0max = float(sys.argv[4]) if len(sys.argv) > 4 else 1.5
vyb = float(sys.argv[5]) if len(sys.argv) > 5 else 0.15
omb = float(sys.argv[6]) if len(sys.argv) > 6 else 2.0
HARD = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)), x_lo=[-1e20] * 4 + [-vyb, -omb], x_hi=[1e20] * 4 + [vyb, omb])
rng = np.random.default_rng(4)
x0 = np.zeros((B, 6)); x0[:, 1] = rng.uniform(-y0max, y0max, B); x0[:, 3] = rng.uniform(0.8, 1.2, B)
u0 = np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1)
sc = tg.Scenarios(B); sc.set_sine(slice(0, B), 0.5, 0.5, 0.0, 0.0)
gen = tg.ClosedLoopGenerator(N=N, Ts=0.02, **HARD)
gen.generate(x0[:8], u0[:8], sc.slice(0, 8), 3)
L = _lib.load()
has_pt = hasattr(L, "tg_debug_phases")
if has_pt:
    L.tg_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.tg_debug_phases(None, 1)
t = time.perf_counter(); res = gen.generate(x0, u0, sc, T); dt = time.perf_counter() - t
st = res["status_counts"].sum(0); its = res["iters_total"] / T
print(f"N={N} B={B} T={T} y0max={y0max} vy={vyb} om={omb}: {dt:.2f} s = {B*T/dt:.3e} steps/s; statuses {dict(zip(tg.STATUS_STRINGS, st.tolist()))}; "
      f"iters/step mean {its.mean():.0f} p50 {np.median(its):.0f} p90 {np.percentile(its, 90):.0f} max {its.max():.0f}; {gen.info()}")
if has_pt:
    out = (ctypes.c_longlong * 16)(); L.tg_debug_phases(out, 0)
    v = np.array(out[:16], dtype=float)
    names = ["rollout", "linearise+resid", "-", "condense+bounds+rho (first pass)", "refactor: (re-condense) buildK+sweep", "ADMM tail", "exit", "ADMM init", "ADMM iterations", "ADMM check", "-", "-", "-", "-", "-", "-"]
    for nme, c in zip(names, v):
        if c > 0: print(f"  {nme:40s} {c/1e6:10.1f} Mcycles  {100*c/v.sum():5.1f}%")
