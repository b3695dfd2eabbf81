"""small invocations of the widened-row kernels for compute-sanitizer (memcheck): odd sizes on purpose."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
import trajectory_generation_b200 as tg

for kind in ("type1", "type2"):
    g = tg.OpenLoopGenerator(kind, Ts=0.01)
    for B, T in ((1, 1), (33, 7), (70, 130)):
        r = g.generate(g.sample_x0(B), T)
        assert np.isfinite(r["clean"]).all() and np.isfinite(r["noisy"]).all() and np.isfinite(r["U"]).all()
    g.generate(g.sample_x0(5), 9, want=("U",))
    g.close()
m = tg.VehicleModel(0.01, 1, 1, None, None, None, None)
for i, (a, b) in enumerate(tg.estimator.LIMIT_KEYS):
    m.Params[a], m.Params[b] = -3.0 - i, 3.0 + i
for dt in (torch.float32, torch.float64):
    for B in (1, 129, 1000):
        x = torch.randn(B, 6, 1, device="cuda", dtype=dt, requires_grad=True); u = torch.randn(B, 2, 1, device="cuda", dtype=dt, requires_grad=True)
        y = m.f(x[1:] if B > 1 else x, u[1:] if B > 1 else u)          # offset views
        y.sum().backward()
        assert torch.isfinite(y).all() and torch.isfinite(x.grad).all()
    r = m.rollout_open_loop(torch.randn(3, 6, 1, device="cuda", dtype=dt), torch.randn(3, 2, 17, device="cuda", dtype=dt) * 0.1, 5, 30)
    assert r.shape == (3, 6, 12)
gen = tg.ClosedLoopGenerator(N=20, Ts=0.01, plant=tg.PLANT_GEN1)
B = 5
x0 = tg.sample_x0(B); x0[:, 1:3] = 0; x0[:, 3] += 0.4
res = gen.generate(x0, np.stack([tg.d_steady_state(x0[:, 3]), np.zeros(B)], 1), tg.Scenarios(B), 12)
assert res["status_counts"][:, :2].sum() == B * 12
print("sanitize workload ok")
