"""dev: where and why a type-2 open-loop trajectory of the GPU leaves the oracle's (tests/test_gpu_openloop.py::test_type2_matches_oracle)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
from oracle import openloop as ool
sys.path.insert(0, "tests")
from test_gpu_openloop import _rules_from_struct
for Ts, T in ((0.01, 1200), (0.02, 300)):
    B = 41
    gen = tg.OpenLoopGenerator("type2", Ts=Ts)
    x0 = gen.sample_x0(B, seed=42)
    res = gen.generate(x0, T)
    rules = _rules_from_struct(ool.Type2Rules, gen.rules)
    for i in range(B):
        U, X, modes = ool.type2_trajectory(ool.PhiloxType2Source(tg.CTRL_SEED_BASE + i), x0[i], T, Ts, rules)
        dX = np.abs(res["clean"][i] - X).max(1)
        if not np.array_equal(res["modes"][i], modes) or dX.max() > 1e-7:
            t = int(np.argmax(res["modes"][i] != modes)) if not np.array_equal(res["modes"][i], modes) else -1
            tU = int(np.argmax(np.abs(res["U"][i] - U).max(1) > 1e-9))
            print(f"Ts={Ts} traj {i}: first mode difference at step {t}, first U difference at {tU}; dX before: {dX[max(tU-5,0):tU+2]}")
            for tt in range(max(tU - 3, 0), tU + 2):
                v = np.hypot(X[tt, 3], X[tt, 4]); vg = np.hypot(res['clean'][i, tt, 3], res['clean'][i, tt, 4])
                print(f"   t={tt} oracle v={v!r} gpu v={vg!r} X_o={X[tt]} U_o={U[tt]} U_g={res['U'][i, tt]} mode {modes[tt]} / {res['modes'][i][tt]}")
            print("   rules: v_high", rules.v_high, "v_turn_max", rules.v_turn_max, "stall_v", rules.stall_v, "v_floor", rules.v_floor)
    print("done", Ts)
