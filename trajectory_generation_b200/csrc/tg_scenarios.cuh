// tg_scenarios.cuh -- scenario + initial-state generation on the device (the x0 ranges
// of generation_type1.py:260-265 / generation_type2.py:171-174, reference paths per trajectory, ramp-cruise speed profiles).
// One thread per trajectory; every number is a function of (seed_base + global trajectory id) through Philox4x32-10, so
// shards and batch sizes do not change a trajectory.  oracle/scenarios.py restates it in NumPy; arithmetic that must match
// bit for bit uses explicit non-contracted operations (__dadd_rn / __dmul_rn / __ddiv_rn), in NumPy's evaluation order.
//
// Draw layout: philox(counter = (block, 7, 0, 0), key = seed):
//   block 0: X, lateral offset, heading offset, vx      block 1: vy, omega, v_cruise, -
//   block 2: path parameters (sine: A, k, psi; parabola: c)
//   blocks 3..: knot spacings of the spline (4 per block)   blocks 16..: knot ordinates, Box-Muller pairs (2 pairs per block)
#pragma once
#include "tg_device.cuh"

#define TG_SCN_MAX_KNOTS 32
#define TG_SCN_STREAM 7u

__device__ __forceinline__ double tg_scn_uniform(uint32_t r, double lo, double hi)
{
    const double u = __dmul_rn(__dadd_rn((double)r, 0.5), 2.3283064365386963e-10);   // (r + 0.5) 2^-32, exact
    return __dadd_rn(lo, __dmul_rn(__dadd_rn(hi, -lo), u));
}

__global__ void tg_scenario_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ tg_scenario_rules R, int B,
                                   long long traj_id0, double *x0, double *u0, tg_ref_spec *spec, double *brk, double *coef)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned long long seed = R.seed_base + (unsigned long long)(traj_id0 + b);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t r0[4], r1[4], r2[4];
    tg_philox4x32_10(0u, TG_SCN_STREAM, 0u, 0u, k0, k1, r0);
    tg_philox4x32_10(1u, TG_SCN_STREAM, 0u, 0u, k0, k1, r1);
    tg_philox4x32_10(2u, TG_SCN_STREAM, 0u, 0u, k0, k1, r2);
    const double X = tg_scn_uniform(r0[0], R.x0_lo[0], R.x0_hi[0]);
    const double lat = tg_scn_uniform(r0[1], R.lat_off[0], R.lat_off[1]);
    const double head = tg_scn_uniform(r0[2], R.head_off[0], R.head_off[1]);
    const double vx = tg_scn_uniform(r0[3], R.x0_lo[3], R.x0_hi[3]);
    const double vy = tg_scn_uniform(r1[0], R.x0_lo[4], R.x0_hi[4]);
    const double om = tg_scn_uniform(r1[1], R.x0_lo[5], R.x0_hi[5]);
    const double vcr = tg_scn_uniform(r1[2], R.vcruise[0], R.vcruise[1]);
    const int K = R.spl_knots, P = K - 1;
    tg_ref_spec sp;
    sp.path_kind = R.cycle[(int)((traj_id0 + b) % R.n_cycle)];
    sp.vref_kind = TG_VREF_RAMP;
    sp.spline_first = b * P; sp.spline_count = P;
    for (int i = 0; i < 4; ++i) sp.path[i] = 0.0;
    for (int i = 0; i < 6; ++i) sp.vref[i] = 0.0;
    sp.vref[0] = R.vref0; sp.vref[1] = vcr; sp.vref[2] = R.t_ramp;
    double y = 0.0, dy = 0.0;
    double *bk = brk + (size_t)b * P, *cf = coef + 4 * (size_t)b * P;
    if (sp.path_kind == TG_PATH_SPLINE) {
        double kx[TG_SCN_MAX_KNOTS], ky[TG_SCN_MAX_KNOTS], hh[TG_SCN_MAX_KNOTS], dd[TG_SCN_MAX_KNOTS], bb[TG_SCN_MAX_KNOTS],
            rr[TG_SCN_MAX_KNOTS], mm[TG_SCN_MAX_KNOTS];
        kx[0] = R.spl_x0;
        {
            uint32_t q[4];
            for (int j = 0; j < P; ++j) {
                if ((j & 3) == 0) tg_philox4x32_10(3u + (uint32_t)(j >> 2), TG_SCN_STREAM, 0u, 0u, k0, k1, q);
                kx[j + 1] = __dadd_rn(kx[j], tg_scn_uniform(q[j & 3], R.spl_dx[0], R.spl_dx[1]));
            }
        }
        for (int pj = 0; 2 * pj < K; ++pj) {
            uint32_t q[4];
            tg_philox4x32_10(16u + (uint32_t)(pj >> 1), TG_SCN_STREAM, 0u, 0u, k0, k1, q);
            double n0, n1;
            tg_box_muller(q[(pj & 1) * 2], q[(pj & 1) * 2 + 1], n0, n1);
            ky[2 * pj] = __dmul_rn(R.spl_sigma, n0);
            if (2 * pj + 1 < K) ky[2 * pj + 1] = __dmul_rn(R.spl_sigma, n1);
        }
        // natural cubic spline, the arithmetic of Scenarios.set_splines (Thomas algorithm on the second derivatives)
        for (int i = 0; i < P; ++i) { hh[i] = __dadd_rn(kx[i + 1], -kx[i]); dd[i] = __ddiv_rn(__dadd_rn(ky[i + 1], -ky[i]), hh[i]); }
        const int n = K - 2;
        for (int i = 0; i < K; ++i) mm[i] = 0.0;
        if (n > 0) {
            for (int i = 0; i < n; ++i) { bb[i] = __dmul_rn(2.0, __dadd_rn(hh[i], hh[i + 1])); rr[i] = __dmul_rn(6.0, __dadd_rn(dd[i + 1], -dd[i])); }
            for (int i = 1; i < n; ++i) {   // a_i = h_i, c_i = h_{i+1}
                const double w = __ddiv_rn(hh[i], bb[i - 1]);
                bb[i] = __dadd_rn(bb[i], -__dmul_rn(w, hh[i]));
                rr[i] = __dadd_rn(rr[i], -__dmul_rn(w, rr[i - 1]));
            }
            mm[n] = __ddiv_rn(rr[n - 1], bb[n - 1]);
            for (int i = n - 2; i >= 0; --i) mm[i + 1] = __ddiv_rn(__dadd_rn(rr[i], -__dmul_rn(hh[i + 1], mm[i + 2])), bb[i]);
        }
        int piece = 0;
        for (int i = 0; i < P; ++i) {
            const double c0 = __ddiv_rn(__dadd_rn(mm[i + 1], -mm[i]), __dmul_rn(6.0, hh[i]));
            const double c1 = __ddiv_rn(mm[i], 2.0);
            const double c2 = __dadd_rn(dd[i], -__ddiv_rn(__dmul_rn(hh[i], __dadd_rn(__dmul_rn(2.0, mm[i]), mm[i + 1])), 6.0));
            bk[i] = kx[i];
            cf[4 * i] = c0; cf[4 * i + 1] = c1; cf[4 * i + 2] = c2; cf[4 * i + 3] = ky[i];
            if (kx[i] <= X) piece = i;
        }
        {   // y(X), y'(X) on the piece that contains X (the first piece extrapolates to the left)
            const double dx = __dadd_rn(X, -kx[piece]);
            const double c0 = cf[4 * piece], c1 = cf[4 * piece + 1], c2 = cf[4 * piece + 2], c3 = cf[4 * piece + 3];
            y = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c0, dx), c1), dx), c2), dx), c3);
            dy = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(3.0, c0), dx), __dmul_rn(2.0, c1)), dx), c2);
        }
    } else {
        for (int i = 0; i < P; ++i) { bk[i] = 0.0; cf[4 * i] = cf[4 * i + 1] = cf[4 * i + 2] = cf[4 * i + 3] = 0.0; }
        if (sp.path_kind == TG_PATH_SINE) {
            const double A = tg_scn_uniform(r2[0], R.sine_A[0], R.sine_A[1]), kk = tg_scn_uniform(r2[1], R.sine_k[0], R.sine_k[1]);
            const double psi = tg_scn_uniform(r2[2], R.sine_psi[0], R.sine_psi[1]);
            sp.path[0] = A; sp.path[1] = kk; sp.path[2] = psi;
            double sn, cs;
            sincos(__dadd_rn(__dmul_rn(kk, X), psi), &sn, &cs);
            y = __dmul_rn(A, sn); dy = __dmul_rn(__dmul_rn(A, kk), cs);
        } else {   // parabola y = c x^2
            const double cc = tg_scn_uniform(r2[0], R.parab_c[0], R.parab_c[1]);
            sp.path[0] = cc;
            y = __dmul_rn(cc, __dmul_rn(X, X)); dy = __dmul_rn(__dmul_rn(2.0, cc), X);
        }
    }
    double *xo = x0 + 6 * (size_t)b;
    xo[0] = X; xo[1] = __dadd_rn(y, lat); xo[2] = __dadd_rn(atan(dy), head); xo[3] = vx; xo[4] = vy; xo[5] = om;
    // steady-state duty cycle at vx (MPC/main.py:9-18)
    u0[2 * (size_t)b] = __ddiv_rn(__dadd_rn(c.p[P_Cr0], __dmul_rn(c.p[P_Cr2], __dmul_rn(vx, vx))), __dadd_rn(c.p[P_Cm1], -__dmul_rn(c.p[P_Cm2], vx)));
    u0[2 * (size_t)b + 1] = 0.0;
    spec[b] = sp;
}
