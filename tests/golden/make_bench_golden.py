#!/usr/bin/env python
"""Oracle trajectories of the configuration bench.py times (BASELINE config 2), committed as
tests/golden/oracle_bench_config.npz.

    python tests/golden/make_bench_golden.py [n_traj=32] [T=1200]

For trajectory ids 0..n_traj-1 of the bench workload (oracle/scenarios.py = what bench.make_workload generates on the device; plant = generation_type1's clipped plant, spline / sinusoid
references, time-advancing ramp vref, N = 20, Ts = 0.01) this runs the oracle's closed loop (MPC/main.py:85-101
restated in oracle/mpc.py) twice:

  * solver="ipm"   -- every step solved to the exact optimum (the parity target), and
  * solver="osqp"  -- every step solved by the restated OSQP at CVXPY's settings (eps_abs = eps_rel = 1e-5, cold
                      start, check_termination 25): "the OSQP solution at matched eps".

The oracle is test infrastructure (PARITY UNPINNED for its QP half, see oracle/qp.py); this script only needs NumPy /
SciPy, no GPU and no /root/reference.  About 6 minutes on 8 cores for the default sizes.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _one(args):
    b, T, x0, u0, spec, brk, coef, solver = args
    import bench
    from oracle import dynamics as dyn, mpc as ompc, refgen as R
    kind = int(spec["path_kind"])
    spline = None
    if kind == R.PATH_SPLINE:
        f, K = int(spec["spline_first"]), int(spec["spline_count"])
        spline = (np.append(brk[f:f + K], np.inf), coef[f:f + K])
    X, U, st, its = ompc.closed_loop(x0, u0, T, bench.TS, bench.N_HORIZON, path_kind=kind, path_prm=tuple(spec["path"]),
                                     spline=spline, vref_kind=int(spec["vref_kind"]), vref_prm=tuple(spec["vref"]),
                                     vref_advance=True, plant=dyn.PLANT_GEN1, solver=solver)
    ok = np.array([s in ompc.ACCEPTED for s in st])
    jump = np.zeros(T, bool)
    if solver == "ipm":
        # steps at which the reference's central differences (MPC/mpc_6stati.py:73-97) straddle a JUMP of f_cont -- the
        # atan2 branch cut behind vx_eff < 0 or the sign flip at vx = 0, reached when the nominal rollout brakes through
        # standstill -- and return the jump divided by 2 eps (entries of ~2e5 in Ad) instead of a derivative
        up = np.vstack([u0, U[:-1]])
        for t in range(T):
            A = dyn.linearize_horizon(X[t], up[t], bench.TS, bench.N_HORIZON)[0]
            jump[t] = np.abs(A).max() > 1e3
    return b, solver, X, U, ok, its, jump


def main():
    n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
    import bench
    import trajectory_generation_b200 as tg
    from oracle import scenarios as oscn
    scn = oscn.make_scenarios(n_traj)                     # the workload bench.make_workload generates on the device
    x0, u0 = scn["x0"], scn["u0"]
    sc = tg.Scenarios.from_arrays(scn["path_kind"], scn["path"], scn["vref"], scn["breaks"], scn["coef"])
    brk, coef = sc.tables()
    jobs = [(b, T, x0[b], u0[b], sc.spec[b], brk, coef, s) for s in ("ipm", "osqp") for b in range(n_traj)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(_one, jobs, chunksize=1)
    out = {"x0": x0, "u0": u0, "n_traj": n_traj, "T": T, "path_kind": scn["path_kind"], "path": scn["path"], "vref": scn["vref"],
           "breaks": scn["breaks"], "coef": scn["coef"]}
    for s in ("ipm", "osqp"):
        rs = sorted([r for r in res if r[1] == s], key=lambda r: r[0])
        out[f"X_{s}"] = np.stack([r[2] for r in rs])
        out[f"U_{s}"] = np.stack([r[3] for r in rs])
        out[f"ok_{s}"] = np.stack([r[4] for r in rs])
        out[f"iters_{s}"] = np.stack([r[5] for r in rs]).astype(np.int32)
        if s == "ipm":
            out["fd_jump"] = np.stack([r[6] for r in rs])
    path = os.path.join(ROOT, "tests", "golden", "oracle_bench_config.npz")
    np.savez_compressed(path, **out)
    d = np.abs(out["X_ipm"] - out["X_osqp"]).max(), np.abs(out["U_ipm"] - out["U_osqp"]).max()
    print(f"wrote {path} in {time.time() - t0:.0f} s; |X_ipm - X_osqp| {d[0]:.2e}, |U_ipm - U_osqp| {d[1]:.2e}; "
          f"non-accepted steps ipm {int((~out['ok_ipm']).sum())}, osqp {int((~out['ok_osqp']).sum())}; "
          f"steps with a finite-difference jump artefact {int(out['fd_jump'].sum())} in trajectories {np.nonzero(out['fd_jump'].any(1))[0].tolist()}")


if __name__ == "__main__":
    main()
