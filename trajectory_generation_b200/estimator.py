"""Estimator-side physics on B200: drop-in for KalmanNet/vehicle_model.py.

``VehicleModel`` has the reference's constructor and attributes (``m, n, d, Ts, Params, T, T_test, m1x_0, prior_*``,
KalmanNet/vehicle_model.py:86-106) and its ``f`` / ``h`` signatures ([B,6,1], [B,2,1] -> [B,6,1]; [B,6,1] -> [B,5,1]), so
``KalmanNetNN.NNBuild(sys_model)`` (kalman_net.py:32) takes it unchanged.  ``f`` is ONE fused CUDA kernel
(tg_estimator_step) instead of ~40 elementwise torch kernels, and is differentiable: the backward pass is the analytic
vector-Jacobian product kernel (tg_estimator_step_vjp), which reproduces what torch.autograd gives for the reference's
graph.  ``rollout_open_loop`` is the H-step prediction of KalmanNet/test_prediction.py:68-87 in one launch.
torch is used for memory and streams only; CUDA tensors in, CUDA tensors out, no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from .mpc import Params as _Params, make_config

LIMIT_KEYS = (("x_min", "x_max"), ("y_min", "y_max"), ("phi_min", "phi_max"), ("vx_min", "vx_max"), ("vy_min", "vy_max"),
              ("omega_min", "omega_max"))


def _dtype_code(t):
    if t.dtype == torch.float64:
        return 0
    if t.dtype == torch.float32:
        return 1
    raise TypeError("float32 or float64 tensors required")


def _cuda_contig(t, name):
    if not t.is_cuda:
        raise _lib.TrajgenError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t          # the kernels move 2-element vectors


class _StepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, u, model):
        x, u = _cuda_contig(x, "x"), _cuda_contig(u, "u")
        model._check_device(x, u)
        out = torch.empty_like(x)
        model._call("tg_estimator_step", x, x.shape[0], _dtype_code(x), x.data_ptr(), u.data_ptr(), model._limits(), out.data_ptr())
        ctx.save_for_backward(x, u)
        ctx.model = model
        return out

    @staticmethod
    def backward(ctx, g):
        x, u = ctx.saved_tensors
        g = _cuda_contig(g, "grad")
        gx, gu = torch.empty_like(x), torch.empty_like(u)
        ctx.model._call("tg_estimator_step_vjp", x, x.shape[0], _dtype_code(x), x.data_ptr(), u.data_ptr(), ctx.model._limits(),
                        g.data_ptr(), gx.data_ptr(), gu.data_ptr())
        return gx, gu, None


class VehicleModel:
    def __init__(self, Ts, T_train, T_test, m1x_0_real, prior_Q, prior_Sigma, prior_S, device=None):
        self.m, self.n, self.d = 6, 5, 2                    # vehicle_model.py:88-90
        self.Ts = Ts
        self.Params = dict(_Params)                         # the caller adds the *_min / *_max limits (test_vehicle.py:86-93)
        self.T, self.T_test = T_train, T_test
        self.m1x_0 = m1x_0_real
        self.prior_Q, self.prior_Sigma, self.prior_S = prior_Q, prior_Sigma, prior_S
        self._device = torch.cuda.current_device() if device is None else int(device)
        self._h, self._key = None, None

    # -- library plumbing
    def _handle(self):
        key = (float(self.Ts), tuple(float(self.Params[k]) for k in _lib.PARAM_ORDER))
        if self._h is None or key != self._key:             # Ts / vehicle parameters changed: new handle
            self.close()
            cfg = make_config(N=1, Ts=float(self.Ts), params=self.Params)
            h = _lib.vp()
            _lib.check(_lib.load().tg_create(ctypes.byref(cfg), self._device, ctypes.byref(h)))
            self._h, self._key = h, key
        return self._h

    def _limits(self):
        lim = _lib.TgStateLimits()
        for i, (a, b) in enumerate(LIMIT_KEYS):
            if a not in self.Params or b not in self.Params:
                raise KeyError(f"Params[{a!r}] / Params[{b!r}] are not set (the reference reads them in pt_f_cont)")
            lim.lo[i], lim.hi[i] = float(self.Params[a]), float(self.Params[b])
        return ctypes.byref(lim)

    def _check_device(self, *tensors):
        for t in tensors:
            if t.device.index != self._device:
                raise _lib.TrajgenError(f"tensor on cuda:{t.device.index}, but this VehicleModel is bound to cuda:{self._device} "
                                        "(pass device= at construction)")

    def _call(self, name, like, *args):
        L, h = _lib.load(), self._handle()
        _lib.check(L.tg_set_stream(h, torch.cuda.current_stream(like.device).cuda_stream))
        _lib.check(getattr(L, name)(h, *args))

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().tg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's interface
    def f(self, x_batch_in, u_batch_in):
        """x_t = f(x_{t-1}, u_t): [B,6,1], [B,2,1] -> [B,6,1]   (vehicle_model.py:109-134)"""
        x = torch.squeeze(x_batch_in, 2)
        u = torch.squeeze(u_batch_in, 2)
        if u.dtype != x.dtype:                                # the reference's torch code type-promotes; the kernel takes one
            dt = torch.promote_types(x.dtype, u.dtype)        # dtype (the cast stays in the autograd graph)
            x, u = x.to(dt), u.to(dt)
        return torch.unsqueeze(_StepFn.apply(x, u, self), 2)

    def h(self, x_batch_in):
        """y_t = h(x_t): [B,6,1] -> [B,5,1], states (X, Y, vx, vy, omega)   (vehicle_model.py:137-153)"""
        return x_batch_in[:, [0, 1, 3, 4, 5]]

    @torch.no_grad()
    def rollout_open_loop(self, x0_real, u, t_start_state, H):
        """KalmanNet/test_prediction.py:68-87 (there a free function taking the model): x0[B,6,1], u[B,2,T] -> [B,6,Hn]."""
        x0 = _cuda_contig(torch.squeeze(x0_real, 2), "x0")
        u = _cuda_contig(u.to(x0.dtype), "u")
        self._check_device(x0, u)
        B, T_u = x0.shape[0], u.shape[2]
        Hn = max(0, min(int(H), T_u - int(t_start_state)))
        if Hn == 0:
            return x0_real                                    # :86 fallback
        preds = torch.empty((B, 6, Hn), dtype=x0.dtype, device=x0.device)
        got = _lib.i32()
        self._call("tg_estimator_rollout", x0, B, _dtype_code(x0), T_u, int(t_start_state), int(H), x0.data_ptr(), u.data_ptr(),
                   self._limits(), preds.data_ptr(), ctypes.byref(got))
        assert got.value == Hn
        return preds


def rollout_open_loop(sys_model, x0_real, u, t_start_state, H):
    """the reference's free-function form (KalmanNet/test_prediction.py:68)."""
    return sys_model.rollout_open_loop(x0_real, u, t_start_state, H)
