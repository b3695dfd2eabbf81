"""Timing of the fused estimator kernels (tg_estimator_step / _vjp / _rollout) with CUDA events.

    python tools/bench_estimator.py [--batch 4194304] [--reps 20]

One JSON line per kernel and dtype: ms, elements/s and achieved HBM bandwidth on the algorithmic bytes
(step: 8 values in + 6 out per element; vjp: 14 in + 8 out) against MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import trajectory_generation_b200 as tg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1 << 22)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); src = "MEASURED_PEAKS.json"
    except Exception:
        peak, src = 6650.0, "fallback"
    B = a.batch
    m = tg.VehicleModel(0.01, 1, 1, None, None, None, None)
    lo = [-3, -3, -3.2, 0.05, -0.4, -5.0]; hi = [3, 3, 3.2, 2.2, 0.4, 5.0]
    for i, (k0, k1) in enumerate(tg.estimator.LIMIT_KEYS):
        m.Params[k0], m.Params[k1] = lo[i], hi[i]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for dt, sz in ((torch.float32, 4), (torch.float64, 8)):
        x = (torch.rand(B, 6, 1, device="cuda", dtype=dt) - 0.5) * 4
        u = (torch.rand(B, 2, 1, device="cuda", dtype=dt) - 0.5)
        g = torch.randn(B, 6, 1, device="cuda", dtype=dt)
        xg = x.clone().requires_grad_(True); ug = u.clone().requires_grad_(True)

        def timed(fn):
            ts = []
            for r in range(a.reps + 3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                if r >= 3:
                    ts.append(e0.elapsed_time(e1))
            return sum(ts) / len(ts)
        with torch.no_grad():
            ms = timed(lambda: m.f(x, u))
        res = [("tg_estimator_step_kernel", ms, B * 14 * sz)]
        y = m.f(xg, ug)
        ms = timed(lambda: torch.autograd.grad(y, (xg, ug), g, retain_graph=True))
        res.append(("tg_estimator_vjp_kernel", ms, B * 22 * sz))
        for name, ms, nbytes in res:
            gbps = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"kernel": name, "dtype": str(dt).split(".")[1], "batch": B, "ms": ms, "elements_per_s": B / (ms * 1e-3),
                              "roofline": {"bound": "hbm", "achieved": gbps, "peak": peak, "unit": "GB/s", "frac": gbps / peak, "peak_source": src},
                              "note": "timed through the torch.autograd.Function wrapper (includes the output allocation)"}))
        # rollout: the reference's use case is B = 1; report latency per predicted step
        for Br in (1, 4096):
            x0 = x[:Br].clone(); U = (torch.rand(Br, 2, 256, device="cuda", dtype=dt) - 0.5) * 0.4
            ms = timed(lambda: m.rollout_open_loop(x0, U, 0, 200))
            print(json.dumps({"kernel": "tg_estimator_rollout_kernel", "dtype": str(dt).split(".")[1], "batch": Br, "H": 200, "ms": ms,
                              "us_per_predicted_step": ms * 1e3 / 200}))


if __name__ == "__main__":
    main()
