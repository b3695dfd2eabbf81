import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
B, T = 65536, 1200
gen = tg.ClosedLoopGenerator(N=20, Ts=0.01, plant=tg.PLANT_GEN2, vref_advance=True)
rules = tg.scenario_rules(cycle=(tg.PATH_PARABOLA, tg.PATH_SINE, tg.PATH_SPLINE), x0_lo=(-2, 0, 0, 0.4, -0.05, -1), x0_hi=(2, 0, 0, 0.6, 0.05, 1), seed_base=42)
t_all = time.perf_counter()
x0, u0, sc = gen.make_scenarios(B, rules); print("setup", time.perf_counter() - t_all)
for rep in range(2):
    t = time.perf_counter()
    res = gen.generate_to_csv(x0, u0, sc, T, "/tmp/c3_clean.csv", "/tmp/c3_noisy.csv", csv_ids=5000)
    dt = time.perf_counter() - t
    print("to_csv", rep, dt, B * T / dt)
for rep in range(2):
    t = time.perf_counter(); full = gen.generate(x0, u0, sc, T); dt = time.perf_counter() - t
    print("generate", dt, B * T / dt)
