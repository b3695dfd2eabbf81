"""BASELINE.json configs 1, 3, 4, 5 at full size on one B200 (config 2 is bench.py): timing + sanity numbers.
Writes a short report to stdout; used to fill profiles/rNN_configs.txt."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import trajectory_generation_b200 as tg

def timed(gen, x0, u0, sc, T, reps=2):
    """best of `reps` launches into result buffers that are page-locked once (page-locking is not part of the path)"""
    buf = gen.alloc_result(len(x0), T)
    gen.generate(x0[:8], u0[:8], sc.slice(0, 8), 3)
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); res = gen.generate(x0, u0, sc, T, out=buf); best = min(best, time.perf_counter() - t)
    return res, best

ONLY = set(sys.argv[1:])          # e.g. "4" to run config 4 alone

# ---- config 1: MPC/main.py verbatim (B = 1, N = 40, Ts = 0.02, 600 steps)
def config1():
    gen = tg.ClosedLoopGenerator(N=40, Ts=0.02)
    x0 = np.array([[0, 0.5, 0, 1.0, 0, 0.0]]); u0 = np.array([[tg.d_steady_state(1.0), 0.0]])
    res, dt = timed(gen, x0, u0, tg.Scenarios(1), 600, reps=3)
    d, de = res["U"][0, :, 0], res["U"][0, :, 1]
    print(f"config 1  B=1 N=40 T=600: {dt*1e3:.1f} ms end to end = {dt/600*1e6:.0f} us per closed-loop step; statuses {res['status_counts'][0]}; "
          f"d mean {d.mean():.4f} std {d.std():.4f}, delta mean {de.mean():.4f} std {de.std():.4f} (generation_type1.py:250: 0.2161/0.1314, 0.0035/0.0338)")
    ctl = tg.BatchedMPC(N=40, Ts=0.02)
    from oracle import refgen as R
    v = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), 40, 0.02); pr = R.ref_window(0.0, 40, 0.02, v)
    ctl.step(x0, u0, pr[None], v[None])
    ts = []
    for _ in range(50):
        t = time.perf_counter(); ctl.step(x0, u0, pr[None], v[None]); ts.append(time.perf_counter() - t)
    print(f"          one mpc_step call through the host API (B=1, cold start): p50 {np.median(ts)*1e6:.0f} us")


# ---- config 3: 65536 trajectories, parabola + mixed references, generator plant, clean+noisy CSV (first 5000 ids)
def config3():
    B, T = 65536, 1200
    gen = tg.ClosedLoopGenerator(N=20, Ts=0.01, plant=tg.PLANT_GEN2, vref_advance=True)
    # parabola / sinusoid / spline by id mod 3, x0 from generation_type2.py:171-174's ranges with the vx floor of SURVEY.md 8(d)
    rules = tg.scenario_rules(cycle=(tg.PATH_PARABOLA, tg.PATH_SINE, tg.PATH_SPLINE), x0_lo=(-2, 0, 0, 0.4, -0.05, -1), x0_hi=(2, 0, 0, 0.6, 0.05, 1), seed_base=42)
    t_all = time.perf_counter()
    t = time.perf_counter(); x0, u0, sc = gen.make_scenarios(B, rules); t_setup = time.perf_counter() - t
    import os
    for rep in range(2):     # the second pass re-uses the generator's page-locked chunk buffers
        t = time.perf_counter()
        res = gen.generate_to_csv(x0, u0, sc, T, "/tmp/c3_clean.csv", "/tmp/c3_noisy.csv", csv_ids=5000)
        dt = time.perf_counter() - t
        if rep == 0:
            t_first = time.perf_counter() - t_all
    st = res["status_counts"].sum(0)
    print(f"config 3  B={B} N=20 T={T}: scenario generation on the device {t_setup:.2f} s + generate_to_csv (rows streamed through two page-locked "
          f"chunk buffers, CSV of the first 5000 ids = 6.0 M rows x 2 files, {os.path.getsize('/tmp/c3_clean.csv')/1e9 + os.path.getsize('/tmp/c3_noisy.csv')/1e9:.2f} GB, "
          f"written while later chunks compute): first call {t_first:.2f} s wall-clock in all = {B*T/t_first:.3e} MPC steps/s; "
          f"second call {dt:.2f} s = {B*T/dt:.3e} MPC steps/s; statuses {dict(zip(tg.STATUS_STRINGS, st.tolist()))}; "
          f"mean ADMM iterations/step {res['iters_total'].sum()/(B*T):.2f}")
    t = time.perf_counter(); full = gen.generate(x0, u0, sc, T); dt = time.perf_counter() - t
    print(f"          generate() with every row gathered into host arrays ({full['clean'].nbytes*2/1e9 + full['U'].nbytes/1e9:.1f} GB): {dt:.2f} s = {B*T/dt:.3e} MPC steps/s")
    del res, full


# ---- config 4: horizon sweep with active rate / state boxes, B = 16384, T = 200
#  (a) the SURVEY.md 8(d) stress scenario: tight rate limits + a box on vy / omega, lateral offsets up to 1.5 m.  Many of its steps
#      are INFEASIBLE by construction at long horizons (the box cannot be kept over 1 s of prediction), so (a) measures the
#      infeasibility certificate as much as the solver;
#  (b) the same offsets and rate limits without the state box: every step feasible, rate rows active -- the solver's own throughput.
def config4():
    HARD = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)), x_lo=[-1e20] * 4 + [-0.15, -2.0], x_hi=[1e20] * 4 + [0.15, 2.0])
    RATE = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)))
    B, T = 16384, 200
    rng = np.random.default_rng(4)
    x0 = np.zeros((B, 6)); x0[:, 1] = rng.uniform(-1.5, 1.5, B); x0[:, 3] = rng.uniform(0.8, 1.2, B)
    u0 = np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1)
    sc = tg.Scenarios(B); sc.set_sine(slice(0, B), 0.5, 0.5, 0.0, 0.0)
    for name, kw in (("(a) rate + state box", HARD), ("(b) rate limits only", RATE)):
        for N in (10, 20, 50):
            gen = tg.ClosedLoopGenerator(N=N, Ts=0.02, **kw)
            Bn = B if N < 50 else 4096          # N = 50: a quarter of the batch (the full one takes minutes at ~3000 ADMM iterations per step)
            res, dt = timed(gen, x0[:Bn], u0[:Bn], sc.slice(0, Bn), T, reps=2 if N < 50 else 1)
            st = res["status_counts"].sum(0)
            its = res["iters_total"] / T
            rate_active = np.mean(np.abs(np.abs(np.diff(res["U"][:, :, 1], axis=1)) - 0.04) < 1e-4)
            print(f"config 4 {name}  B={Bn} N={N} T={T}: {dt:.2f} s = {Bn*T/dt:.3e} MPC steps/s; statuses {dict(zip(tg.STATUS_STRINGS, st.tolist()))}; "
                  f"ADMM iterations/step mean {its.mean():.0f} p50 {np.median(its):.0f} p90 {np.percentile(its, 90):.0f} max {its.max():.0f}; "
                  f"steering-rate row active in {100*rate_active:.0f} % of steps; launch geometry {gen.info()}")


for name, f in (("1", config1), ("3", config3), ("4", config4)):
    if not ONLY or name in ONLY:
        f()
