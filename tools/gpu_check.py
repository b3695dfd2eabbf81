"""Development check on a B200: CUDA path vs oracle on a handful of cases (not a test; tests/ has those)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import trajectory_generation_b200 as tg
from oracle import dynamics as dyn, mpc as ompc, refgen as R, philox as oph
from tools import condensed_model as CM

rng = np.random.default_rng(0)
B, N, Ts = 16, 20, 0.02
x0 = np.stack([np.array([rng.uniform(-1, 1), rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3), rng.uniform(0.5, 1.5),
                         rng.uniform(-0.05, 0.05), rng.uniform(-1, 1)]) for _ in range(B)])
up = np.stack([np.array([R.d_steady_state(x0[i, 3]), rng.uniform(-0.1, 0.1)]) for i in range(B)])
for jac in (tg.JAC_ANALYTIC, tg.JAC_FD):
    ctl = tg.BatchedMPC(N=N, Ts=Ts, jacobian=jac)
    A, Bm, g, xb = ctl.linearize(x0, up)
    err = 0
    for i in range(B):
        Ao, Bo, go, xbo = dyn.linearize_horizon(x0[i], up[i], Ts, N)
        err = max(err, np.abs(A[i] - Ao).max(), np.abs(Bm[i] - Bo).max(), np.abs(g[i] - go).max(), np.abs(xb[i] - xbo.T).max())
    print("linearize jac", jac, "max abs err", err)
ctl = tg.BatchedMPC(N=N, Ts=Ts)
pr = np.zeros((B, N + 1, 3)); vr = np.zeros((B, N + 1))
for i in range(B):
    vr[i] = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, Ts)
    pr[i] = R.ref_window(x0[i, 0], N, Ts, vr[i])
asm = ctl.assemble(x0, up, pr, vr)
errH = errq = 0
for i in range(B):
    Ao, Bo, go, _ = dyn.linearize_horizon(x0[i], up[i], Ts, N)
    H, q, const, Ac, l, u, c, G = CM.condense(x0[i], up[i], Ao, Bo, go, pr[i], vr[i])
    # model is in U coordinates; kernel in dU = U - u_prev: H equal, q_dU = q + H ubar
    ub = np.tile(up[i], N)
    errH = max(errH, np.abs(asm["H"][i] - H).max() / np.abs(H).max())
    errq = max(errq, np.abs(asm["q"][i] - (q + H @ ub)).max() / (1 + np.abs(q).max()))
print("assemble rel err H", errH, "q", errq)
t = time.time(); out = ctl.step(x0, up, pr, vr); dt = time.time() - t
print("step status", out["status"], "iters", out["iters"], "time", dt)
err = 0; eo = 0
for i in range(B):
    u1, s1, i1 = ompc.mpc_step(x0[i], up[i], pr[i], Ts=Ts, N=N, vref=vr[i], solver="ipm")
    err = max(err, np.abs(out["U_opt"][i].T - i1["U_opt"]).max(), np.abs(out["X_opt"][i].T - i1["X_opt"]).max())
    eo = max(eo, abs(out["objective"][i] - i1["objective"]) / abs(i1["objective"]))
print("step vs ipm: max |U,X| err", err, "rel obj err", eo)
# shim
u_cmd, status, info = tg.mpc_step(x0[0], up[0], pr[0], Ts=Ts, N=N, vref=vr[0])
print("shim", u_cmd, status, sorted(info.keys()))
print("infeasible:", tg.mpc_step(x0[0], [2.0, 0.0], pr[0], Ts=Ts, N=N, vref=vr[0])[:2])
# hard case with state bounds
kw = dict(du_bounds=((-0.1, 0.1), (-0.04, 0.04)), x_lo=[-1e20] * 4 + [-0.15, -2], x_hi=[1e20] * 4 + [0.15, 2])
xh = np.array([0, 1.5, 0, 1.0, 0, 0.0]); uh = np.array([R.d_steady_state(1.0), 0.0])
vh = R.vref_profile(R.VREF_RAMP, (0.8, 2.0, 2.0), N, Ts); ph = R.ref_window(0.0, N, Ts, vh, R.PATH_SINE, (0.5, 0.5, 0, 0))
u2, s2, i2 = tg.mpc_step(xh, uh, ph, Ts=Ts, N=N, vref=vh, **kw)
u3, s3, i3 = ompc.mpc_step(xh, uh, ph, Ts=Ts, N=N, vref=vh, solver="ipm", **kw)
print("hard:", s2, i2.get("iters"), u2, u3, np.abs(i2["U_opt"] - i3["U_opt"]).max())
# noise
gen = tg.ClosedLoopGenerator(N=N, Ts=Ts)
nz = gen.sensor_noise_normals(0, 3, 50)
ref = np.stack([oph.standard_normals(12345 + i, 50) for i in range(3)])
print("noise bit-exact:", np.array_equal(nz, ref), "philox:", np.array_equal(gen.philox_u32(12345, 7, 1, 5), oph.philox_stream(12345, 7, 1, 5)))
# closed loop, MPC/main.py scenario N=20, 60 steps
sc = tg.Scenarios(2)
sc.set_sine(1, 0.5, 0.5, 0.0, 0.0)
x00 = np.array([[0, 0.5, 0, 1.0, 0, 0], [0, 0.3, 0, 1.0, 0, 0]]); u00 = np.tile(uh, (2, 1))
t = time.time(); res = gen.generate(x00, u00, sc, 60); dt = time.time() - t
Xo, Uo, st, its = ompc.closed_loop(x00[0], u00[0], 60, Ts, N)
print("closed loop 60 steps time", dt, "status", res["status_counts"], "iters", res["iters_total"])
print("  vs oracle: X err", np.abs(res["clean"][0] - Xo).max(), "U err", np.abs(res["U"][0] - Uo).max())
Xo, Uo, st, its = ompc.closed_loop(x00[1], u00[1], 60, Ts, N, path_kind=R.PATH_SINE, path_prm=(0.5, 0.5, 0, 0))
print("  sine: X err", np.abs(res["clean"][1] - Xo).max(), "U err", np.abs(res["U"][1] - Uo).max())
nzc = (res["noisy"][0] - res["clean"][0]) / np.array(oph.NOISE_STD)
print("  noise rows match oracle:", np.abs(nzc - oph.standard_normals(12345, 61)).max())
# throughput probe
Bb, T = 1024, 100
scb = tg.Scenarios(Bb); xb = np.tile(x00[0], (Bb, 1)); xb[:, 1] += rng.uniform(-0.2, 0.2, Bb); ubb = np.tile(uh, (Bb, 1))
gen.generate(xb[:8], ubb[:8], scb.slice(0, 8), 5)
t = time.time(); res = gen.generate(xb, ubb, scb, T); dt = time.time() - t
print(f"B={Bb} T={T}: {dt:.3f}s -> {Bb*T/dt:.3e} steps/s (host e2e), mean iters/step {res['iters_total'].mean()/T:.1f}, status {res['status_counts'].sum(0)}")
print("fp64 peak TFLOP/s", gen.fma_peak_tflops("f64"), "fp32", gen.fma_peak_tflops("f32"))
