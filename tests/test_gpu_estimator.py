"""GPU parity of the fused estimator physics against the reference's own outputs
(tests/golden/reference_estimator.npz: VehicleModel.f values + autograd gradients, rollout_open_loop) and against
oracle/estimator.py on fresh inputs.  fp64: 1e-12; fp32: a few ulp of the state scale (tolerances below)."""
import numpy as np
import pytest
import torch

import trajectory_generation_b200 as tg
from oracle import estimator as oe

pytestmark = pytest.mark.gpu
TOL = {torch.float64: dict(rtol=1e-11, atol=1e-12), torch.float32: dict(rtol=2e-5, atol=2e-5)}


def _model(g, Ts=0.01):
    m = tg.VehicleModel(Ts, 10, 10, None, None, None, None)
    m.Params.update(oe.with_limits(g["lo"], g["hi"], {}))
    return m


@pytest.mark.parametrize("name,dt", [("f32", torch.float32), ("f64", torch.float64)])
def test_step_vjp_rollout_vs_reference_golden(golden_estimator, name, dt):
    g = golden_estimator
    m = _model(g)
    x = torch.tensor(g["X"], dtype=dt, device="cuda", requires_grad=True)
    u = torch.tensor(g["U"], dtype=dt, device="cuda", requires_grad=True)
    y = m.f(x.unsqueeze(2), u.unsqueeze(2))
    assert y.shape == (len(g["X"]), 6, 1) and y.dtype == dt
    y.squeeze(2).backward(torch.tensor(g["G"], dtype=dt, device="cuda"))
    np.testing.assert_allclose(y.squeeze(2).detach().cpu().numpy(), g[f"next_{name}"], **TOL[dt])
    # gradients carry the 1/Iz = 3.6e4 amplification of the yaw row: compare relative to each row's scale
    for got, ref in ((x.grad, g[f"gx_{name}"]), (u.grad, g[f"gu_{name}"])):
        got = got.cpu().numpy()
        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
        np.testing.assert_allclose(got / scale, ref / scale, rtol=0, atol=1e-11 if dt == torch.float64 else 2e-4)
    np.testing.assert_array_equal(m.h(x.detach().unsqueeze(2)).squeeze(2).cpu().numpy(), g[f"h_{name}"])
    T, H, t0 = (int(v) for v in g["roll_meta"])
    x0 = torch.tensor(g["X"][:8], dtype=dt, device="cuda").unsqueeze(2)
    U = torch.tensor(g["Useq"], dtype=dt, device="cuda")
    r = tg.rollout_open_loop(m, x0, U, t0, H)
    assert r.shape == (8, 6, T - t0)
    np.testing.assert_allclose(r.cpu().numpy(), g[f"roll_{name}"], rtol=1e-9 if dt == torch.float64 else 1e-3, atol=1e-9 if dt == torch.float64 else 1e-3)
    assert tg.rollout_open_loop(m, x0, U, T, H) is x0                       # nothing left to predict: the reference's fallback


def test_step_matches_oracle_on_fresh_inputs_and_in_a_filter_like_graph(golden_estimator):
    g = golden_estimator
    m = _model(g, Ts=0.02)
    p = oe.with_limits(g["lo"], g["hi"])
    gen = torch.Generator().manual_seed(3)
    B = 4099                                                   # not a multiple of the block size
    x = (torch.rand(B, 6, generator=gen, dtype=torch.float64) - 0.5) * torch.tensor([7, 7, 7, 5.0, 1.0, 12.0], dtype=torch.float64)
    u = (torch.rand(B, 2, generator=gen, dtype=torch.float64) - 0.5) * torch.tensor([2.0, 1.2], dtype=torch.float64)
    ref = oe.step(x, u, 0.02, p)
    got = m.f(x.cuda().unsqueeze(2), u.cuda().unsqueeze(2)).squeeze(2).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-11, atol=1e-12)
    # three chained steps with a learnable gain in between (the way KalmanNet back-propagates through f)
    def chain(step, x0, u0, w):
        xk = x0
        for _ in range(3):
            xk = step(xk * w, u0)
        return (xk ** 2).sum()
    w_ref = torch.full((6,), 0.9, dtype=torch.float64, requires_grad=True)
    chain(lambda a, b: oe.step(a, b, 0.02, p), x[:64], u[:64], w_ref).backward()
    w_gpu = torch.full((6,), 0.9, dtype=torch.float64, device="cuda", requires_grad=True)
    chain(lambda a, b: m.f(a.unsqueeze(2), b.unsqueeze(2)).squeeze(2), x[:64].cuda(), u[:64].cuda(), w_gpu).backward()
    np.testing.assert_allclose(w_gpu.grad.cpu().numpy(), w_ref.grad.numpy(), rtol=1e-9)


def test_estimator_errors():
    m = tg.VehicleModel(0.01, 1, 1, None, None, None, None)
    x = torch.zeros(2, 6, 1, device="cuda"); u = torch.zeros(2, 2, 1, device="cuda")
    with pytest.raises(KeyError):
        m.f(x, u)                                              # limits not set, as the reference would fail in pt_f_cont
    m.Params.update(oe.with_limits([-1] * 6, [1] * 6, {}))
    with pytest.raises(tg.TrajgenError):
        m.f(x.cpu(), u.cpu())                                  # no CPU fallback
    with pytest.raises(TypeError):
        m.f(x.half(), u.half())
    assert m.f(x[:0], u[:0]).shape == (0, 6, 1)
