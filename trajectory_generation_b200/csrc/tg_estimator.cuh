// tg_estimator.cuh -- the estimator-side physics: KalmanNet's prior step
// x_t = f(x_{t-1}, u_t), its vector-Jacobian product (for autograd through the filter) and the open-loop rollout.
// sm_100a; fp32 or fp64 (the reference runs in torch's default fp32).
//
// Reference (paths relative to the reference root):
//   KalmanNet/vehicle_model.py:19-40    pt_tire_forces (front slip angle clamped, rear free; Frx on vx_eff)
//   KalmanNet/vehicle_model.py:43-79    pt_f_cont (phi, vx, vy, omega clamped to the data-set limits before use)
//   KalmanNet/vehicle_model.py:109-134  VehicleModel.f (Euler step from the unclamped state, then all six states clamped)
//   KalmanNet/test_prediction.py:68-87  rollout_open_loop
// In the reference one call of f is ~40 torch kernels of B elements each; here it is one fused kernel, one thread per
// batch element (the arithmetic intensity is ~15 flop/B: HBM-bound, 64 B in / 24..48 B out per element for fp32/fp64).
#pragma once
#include "tg_device.cuh"

template <typename T> struct EstMath;
template <> struct EstMath<double> {
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    static __device__ __forceinline__ double atan_(double x) { return atan(x); }
    static __device__ __forceinline__ void sincos_(double x, double &s, double &c) { sincos(x, &s, &c); }
};
template <> struct EstMath<float> {
    static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float atan_(float x) { return atanf(x); }
    static __device__ __forceinline__ void sincos_(float x, float &s, float &c) { sincosf(x, &s, &c); }
};

// 2-wide vector access: rows of 6 (state) and 2 (input) values are 8-byte (fp32) / 16-byte (fp64) aligned pairs
template <typename T> struct EstVec;
template <> struct EstVec<double> { using V2 = double2; };
template <> struct EstVec<float> { using V2 = float2; };
template <typename T> __device__ __forceinline__ void est_load6(const T *__restrict__ p, T v[6])
{
    using V2 = typename EstVec<T>::V2;
    const V2 a = __ldg(reinterpret_cast<const V2 *>(p)), b = __ldg(reinterpret_cast<const V2 *>(p) + 1), c = __ldg(reinterpret_cast<const V2 *>(p) + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
}
template <typename T> __device__ __forceinline__ void est_load2(const T *__restrict__ p, T &a, T &b)
{
    using V2 = typename EstVec<T>::V2;
    const V2 v = __ldg(reinterpret_cast<const V2 *>(p));
    a = v.x; b = v.y;
}
template <typename T> __device__ __forceinline__ void est_store2(T *__restrict__ p, T a, T b)
{
    using V2 = typename EstVec<T>::V2;
    V2 v; v.x = a; v.y = b;
    *reinterpret_cast<V2 *>(p) = v;
}

template <typename T>
struct EstCfg {
    T Ts, lo[6], hi[6];
    T Cm1, Cm2, Cr0, Cr2, Br, Cr, Dr, Bf, Cf, Df, m, Iz, lf, lr, maxAlpha, vx_zero;
};

template <typename T> __device__ __forceinline__ T est_clamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
template <typename T> __device__ __forceinline__ T est_inside(T v, T lo, T hi) { return (v >= lo && v <= hi) ? T(1) : T(0); }   // torch.clamp backward

// everything the forward pass computes that the backward pass needs again
template <typename T>
struct EstFwd {
    T phi, vx, vy, om, ve, nf, nr, af, ar, af_in;   // clamped inputs, slip angles, d clamp(af)/d af
    T sp, cp, sd, cd, Fyf, Fyr, Frx, thf, thr;      // thf = Cf atan(Bf af), thr = Cr atan(Br ar)
    T pre[6];                                       // x + Ts f before the output clamp
};

template <typename T>
__device__ __forceinline__ void est_forward(const EstCfg<T> &c, const T x[6], T d, T delta, EstFwd<T> &w)
{
    using M = EstMath<T>;
    w.phi = est_clamp(x[2], c.lo[2], c.hi[2]);                                  // vehicle_model.py:54-57
    w.vx = est_clamp(x[3], c.lo[3], c.hi[3]);
    w.vy = est_clamp(x[4], c.lo[4], c.hi[4]);
    w.om = est_clamp(x[5], c.lo[5], c.hi[5]);
    const T avx = fabs(w.vx);
    w.ve = avx > c.vx_zero ? avx : c.vx_zero;                                   // :26
    w.nf = w.om * c.lf + w.vy;
    w.nr = w.om * c.lr - w.vy;
    const T af_raw = -M::atan2_(w.nf, w.ve) + delta;                            // :29
    w.ar = M::atan2_(w.nr, w.ve);                                               // :30
    w.af = est_clamp(af_raw, -c.maxAlpha, c.maxAlpha);                          // :33
    w.af_in = est_inside(af_raw, -c.maxAlpha, c.maxAlpha);
    w.thf = c.Cf * M::atan_(c.Bf * w.af);
    w.thr = c.Cr * M::atan_(c.Br * w.ar);
    T sf, cf_, sr, cr_;
    M::sincos_(w.thf, sf, cf_);
    M::sincos_(w.thr, sr, cr_);
    w.Fyf = c.Df * sf;                                                          // :36-37
    w.Fyr = c.Dr * sr;
    w.thf = cf_;                                                                // keep cos(theta) for the backward pass
    w.thr = cr_;
    w.Frx = (c.Cm1 - c.Cm2 * w.ve) * d - c.Cr0 - c.Cr2 * (w.ve * w.ve);         // :38
    M::sincos_(w.phi, w.sp, w.cp);
    M::sincos_(delta, w.sd, w.cd);
    T f[6];
    f[0] = w.vx * w.cp - w.vy * w.sp;                                           // :68-75
    f[1] = w.vx * w.sp + w.vy * w.cp;
    f[2] = w.om;
    f[3] = (w.Frx - w.Fyf * w.sd + c.m * w.vy * w.om) / c.m;
    f[4] = (w.Fyr + w.Fyf * w.cd - c.m * w.vx * w.om) / c.m;
    f[5] = (w.Fyf * c.lf * w.cd - w.Fyr * c.lr) / c.Iz;
#pragma unroll
    for (int i = 0; i < 6; ++i) w.pre[i] = x[i] + c.Ts * f[i];                  // :121
}

template <typename T>
__global__ void tg_estimator_step_kernel(const __grid_constant__ EstCfg<T> c, int B, const T *__restrict__ x, const T *__restrict__ u,
                                         T *__restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T xs[6], d, delta;
    est_load6(x + 6 * (size_t)b, xs);
    est_load2(u + 2 * (size_t)b, d, delta);
    EstFwd<T> w;
    est_forward(c, xs, d, delta, w);
#pragma unroll
    for (int i = 0; i < 6; i += 2)                                                                  // :124-130
        est_store2(out + 6 * (size_t)b + i, est_clamp(w.pre[i], c.lo[i], c.hi[i]), est_clamp(w.pre[i + 1], c.lo[i + 1], c.hi[i + 1]));
}

// grad_x = J_x^T g, grad_u = J_u^T g of the map (x, u) -> VehicleModel.f(x, u), as torch.autograd differentiates it
template <typename T>
__global__ void tg_estimator_vjp_kernel(const __grid_constant__ EstCfg<T> c, int B, const T *__restrict__ x, const T *__restrict__ u,
                                        const T *__restrict__ g, T *__restrict__ gx, T *__restrict__ gu)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T xs[6], go[6], d, delta;
    est_load6(x + 6 * (size_t)b, xs);
    est_load6(g + 6 * (size_t)b, go);
    est_load2(u + 2 * (size_t)b, d, delta);
    EstFwd<T> w;
    est_forward(c, xs, d, delta, w);
#pragma unroll
    for (int i = 0; i < 6; ++i) go[i] *= est_inside(w.pre[i], c.lo[i], c.hi[i]);   // output clamps
    const T gf0 = c.Ts * go[0], gf1 = c.Ts * go[1], gf2 = c.Ts * go[2], gf3 = c.Ts * go[3], gf4 = c.Ts * go[4], gf5 = c.Ts * go[5];
    const T im = T(1) / c.m, iI = T(1) / c.Iz;
    T g_phi = gf0 * (-w.vx * w.sp - w.vy * w.cp) + gf1 * (w.vx * w.cp - w.vy * w.sp);
    T g_vx = gf0 * w.cp + gf1 * w.sp - gf4 * w.om;
    T g_vy = -gf0 * w.sp + gf1 * w.cp + gf3 * w.om;
    T g_om = gf2 + gf3 * w.vy - gf4 * w.vx;
    const T g_Frx = gf3 * im;
    const T g_Fyf = -gf3 * w.sd * im + gf4 * w.cd * im + gf5 * c.lf * w.cd * iI;
    const T g_Fyr = gf4 * im - gf5 * c.lr * iI;
    T g_delta = -gf3 * w.Fyf * w.cd * im - gf4 * w.Fyf * w.sd * im - gf5 * w.Fyf * c.lf * w.sd * iI;
    const T g_d = g_Frx * (c.Cm1 - c.Cm2 * w.ve);
    T g_ve = g_Frx * (-c.Cm2 * d - T(2) * c.Cr2 * w.ve);
    // F = D sin(C atan(B a)):  dF/da = D cos(theta) C B / (1 + (B a)^2)   (w.thf / w.thr hold cos(theta))
    const T g_af = g_Fyf * c.Df * w.thf * c.Cf * c.Bf / (T(1) + (c.Bf * w.af) * (c.Bf * w.af)) * w.af_in;
    const T g_ar = g_Fyr * c.Dr * w.thr * c.Cr * c.Br / (T(1) + (c.Br * w.ar) * (c.Br * w.ar));
    g_delta += g_af;
    const T rf = T(1) / (w.nf * w.nf + w.ve * w.ve), rr = T(1) / (w.nr * w.nr + w.ve * w.ve);
    const T g_nf = -g_af * w.ve * rf, g_nr = g_ar * w.ve * rr;          // d atan2(y, x) = (x dy - y dx) / (x^2 + y^2)
    g_ve += g_af * w.nf * rf - g_ar * w.nr * rr;
    g_om += g_nf * c.lf + g_nr * c.lr;
    g_vy += g_nf - g_nr;
    const T avx = fabs(w.vx);
    // torch.max(|vx|, vx_zero): gradient to the larger argument (split evenly on a tie); d|vx| = sign(vx)
    const T sel = avx > c.vx_zero ? T(1) : (avx == c.vx_zero ? T(0.5) : T(0));
    g_vx += g_ve * sel * (w.vx > T(0) ? T(1) : (w.vx < T(0) ? T(-1) : T(0)));
    if (gx) {
        est_store2(gx + 6 * (size_t)b, go[0], go[1]);
        est_store2(gx + 6 * (size_t)b + 2, go[2] + g_phi * est_inside(xs[2], c.lo[2], c.hi[2]),     // input clamps of pt_f_cont
                   go[3] + g_vx * est_inside(xs[3], c.lo[3], c.hi[3]));
        est_store2(gx + 6 * (size_t)b + 4, go[4] + g_vy * est_inside(xs[4], c.lo[4], c.hi[4]),
                   go[5] + g_om * est_inside(xs[5], c.lo[5], c.hi[5]));
    }
    if (gu) est_store2(gu + 2 * (size_t)b, g_d, g_delta);
}

// rollout_open_loop: preds[b][:, k] = f(preds[b][:, k-1], U[b][:, t_start + k]) for k < Hn; U is [B][2][T_u], preds [B][6][Hn]
template <typename T>
__global__ void tg_estimator_rollout_kernel(const __grid_constant__ EstCfg<T> c, int B, int T_u, int t_start, int Hn,
                                            const T *__restrict__ x0, const T *__restrict__ U, T *__restrict__ preds)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T xs[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) xs[i] = x0[6 * (size_t)b + i];
    const T *ub = U + (size_t)b * 2 * T_u;
    T *pb = preds + (size_t)b * 6 * Hn;
    for (int k = 0; k < Hn; ++k) {
        EstFwd<T> w;
        est_forward(c, xs, ub[t_start + k], ub[T_u + t_start + k], w);
#pragma unroll
        for (int i = 0; i < 6; ++i) { xs[i] = est_clamp(w.pre[i], c.lo[i], c.hi[i]); pb[(size_t)i * Hn + k] = xs[i]; }
    }
}
