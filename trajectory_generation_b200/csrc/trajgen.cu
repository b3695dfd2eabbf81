// trajgen.cu -- kernels + C ABI (include/trajgen.h) of libtrajgen.so.  sm_100a only.
//
// Kernels (one CTA per problem / trajectory, persistent grid-stride over the batch):
//   tg_mpc_step_kernel     K1+K2+K3: replaces mpc_step, MPC/mpc_6stati.py:120-275
//   tg_closed_loop_kernel  fused K1..K4: replaces the loop MPC/main.py:85-101 + the dataset shell
//                          generation_type2.py:180-200 (plant clipping, sensor noise)
//   tg_ref_window_kernel   MPC/main.py:28-47,51-68
//   tg_plant_kernel        generation_type1.py:70-84 (open-loop integration with clipping)
//   tg_noise_kernel / tg_philox_kernel   Philox4x32-10 sensor-noise stream
//   tg_fma_peak_kernel     FMA micro-benchmark (roofline denominator)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <new>
#include <string>
#include <vector>

#include "tw_solver.cuh"
#include "tg_openloop.cuh"
#include "tg_scenarios.cuh"
#include "tg_estimator.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
void tg_internal_set_error(const char *msg) { g_err = msg; }   // for tg_csv.cpp
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(TG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct StepArgs {
    int B;
    const double *x0, *u_prev, *path_ref, *vref;
    double *u_cmd; int *status; int *iters; double *objective; double *U_opt, *X_opt, *y_opt;
    // taps
    double *A, *Bm, *g, *xbar, *H, *q, *c0, *l, *u, *Gs;
    int stop;
    double *ws_x, *ws_y; int *ws_valid;   // per-problem warm-start state (step API), may be null
    int ppc;                              // problems per CTA
};

struct LoopArgs {
    int B, T;
    const double *x0, *u0;
    const tg_ref_spec *spec;
    const double *brk, *coef;
    long long traj_id0;
    double *clean, *noisy, *U;
    int *status_counts;
    long long *iters_total;
    int ppc;   // problems per CTA
};

// ------------------------------------------------------------------------------------------------ warp-per-problem kernels
// tw_solver.cuh: W warps per problem (W = 1 for N <= 20), P problems per CTA side by side, each with its own shared-memory
// block (and, for W > 1, its own named barrier).  Problems of a CTA are re-aligned at every step boundary (barrier 15) so that
// they fetch the same instruction lines at the same time (the step body is ~100 KB of mostly straight-line code).
#define TW_CTA_THREADS 512
// register budget: S = 2 tiles take 64 registers -> 128-register kernels; S = 1 tiles take 32 -> TW_REGS_S1 registers
#ifndef TW_REGS_S1
#define TW_REGS_S1 128
#endif
#define TW_REGS(S) ((S) == 1 ? TW_REGS_S1 : 128)
// shapes that have a single-problem-per-CTA instance (see tw_mpc_step_kernel): the one- and two-warp kernels, where four CTAs
// per SM would leave most of the register file and shared memory unused
#define TW_HAS_P1(W, NC, HS) ((W) <= 2 && ((NC) == 0 || (HS)))

// P1 = true: the single-problem-per-CTA instance.  Its only barrier id is the constant 1, so the kernel is built with 2
// hardware barriers instead of 16 (a barrier id held in a register makes ptxas reserve all 16, and an SM has 64: four
// resident CTAs at most, whatever the registers and the shared memory would allow).
template <int W, int S, int NC, bool P1 = false, bool HS = (NC == 0)>
__global__ void __launch_bounds__(TW_CTA_THREADS) __maxnreg__(TW_REGS(S))
tw_mpc_step_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ WLayout L, const __grid_constant__ StepArgs a)
{
    extern __shared__ __align__(16) double sm_all[];
    constexpr int NT = 32 * W;
    const int P = P1 ? 1 : a.ppc;
    int prob = P1 ? 0 : threadIdx.x / NT, tid = P1 ? threadIdx.x : threadIdx.x % NT;
    unsigned sm_off = (unsigned)prob * (unsigned)((L.total + 1) & ~1) * 8u;
    if constexpr (!P1) asm volatile("" : "+r"(tid), "+r"(prob), "+r"(sm_off));   // kept in registers instead of being re-derived at every use
    const int bar = P1 ? 1 : 1 + prob;
    double *sm = reinterpret_cast<double *>(reinterpret_cast<char *>(sm_all) + sm_off);
    const int N = NC > 0 ? NC : c.N, n = 2 * N, ms = HS ? c.ms : 0;
    if (prob >= P) return;
    const TwMap<S> mp = tw_make_map<W, S>(LF(nb), tid);
    for (int b0 = blockIdx.x * P; b0 < a.B; b0 += gridDim.x * P) {
        const int b = b0 + prob;
        if (b >= a.B) break;
        tw_sync<W>(bar);
        if (tid < 6) sm[LF(x0) + tid] = a.x0[6 * (size_t)b + tid];
        if (tid < 2) sm[LF(uprev) + tid] = a.u_prev[2 * (size_t)b + tid];
        const double vx0 = a.x0[6 * (size_t)b + 3];
        for (int k = tid; k <= N; k += NT) {
            if (a.path_ref) {
                const double *pr = a.path_ref + 3 * ((size_t)b * (N + 1) + k);
                sm[LF(Xr) + k] = pr[0]; sm[LF(Yr) + k] = pr[1]; sm[LF(Pr) + k] = pr[2];
            } else { sm[LF(Xr) + k] = 0.0; sm[LF(Yr) + k] = 0.0; sm[LF(Pr) + k] = 0.0; }
            sm[LF(vref) + k] = a.vref ? a.vref[(size_t)b * (N + 1) + k] : vx0;
        }
        bool warm = false, warm_free = false;
        if (a.ws_valid && c.warm_start && a.ws_valid[b]) {
            warm = true; warm_free = a.ws_valid[b] == 2;
            const size_t m = 2 * (size_t)n + ms;
            for (int i = tid; i < n; i += NT) sm[LF(x) + i] = a.ws_x[(size_t)b * n + i];
            for (int i = tid; i < 2 * n; i += NT) sm[LF(y) + i] = a.ws_y[b * m + i];
            for (int i = tid; i < ms; i += NT) sm[L.ys + i] = a.ws_y[b * m + 2 * n + i];
        }
        tw_sync<W>(bar);
        StepTaps tap;
        tap.A = a.A ? a.A + (size_t)b * N * 36 : nullptr;
        tap.Bm = a.Bm ? a.Bm + (size_t)b * N * 12 : nullptr;
        tap.g = a.g ? a.g + (size_t)b * N * 6 : nullptr;
        tap.xbar = a.xbar ? a.xbar + (size_t)b * (N + 1) * 6 : nullptr;
        tap.H = a.H ? a.H + (size_t)b * n * n : nullptr;
        tap.q = a.q ? a.q + (size_t)b * n : nullptr;
        tap.c0 = a.c0 ? a.c0 + b : nullptr;
        tap.l = a.l ? a.l + (size_t)b * (2 * n + ms) : nullptr;
        tap.u = a.u ? a.u + (size_t)b * (2 * n + ms) : nullptr;
        tap.Gs = a.Gs ? a.Gs + (size_t)b * c.ms * n : nullptr;
        tap.stop = a.stop;
        const StepResult r = tw_step_body<W, S, NC, HS>(c, L, sm, mp, warm, warm_free, tap, nullptr, tid, bar);
        if (a.stop) continue;
        const bool ok = (r.status == TG_STATUS_OPTIMAL || r.status == TG_STATUS_OPTIMAL_INACCURATE);  // :261
        const double ud = sm[LF(uprev)], udel = sm[LF(uprev) + 1];
        if (tid == 0) {
            a.u_cmd[2 * (size_t)b] = ok ? ud + sm[LF(xt)] : ud;          // :265 / fallback :262
            a.u_cmd[2 * (size_t)b + 1] = ok ? udel + sm[LF(xt) + 1] : udel;
            if (a.status) a.status[b] = r.status;
            if (a.iters) a.iters[b] = r.iters;
            if (a.objective) a.objective[b] = ok ? r.objective : nan("");
        }
        if (a.U_opt)
            for (int i = tid; i < n; i += NT) a.U_opt[(size_t)b * n + i] = ok ? ((i & 1) ? udel : ud) + sm[LF(xt) + i] : nan("");
        if (a.y_opt) {
            const size_t m = 2 * (size_t)n + ms;
            for (int i = tid; i < 2 * n; i += NT) a.y_opt[b * m + i] = ok ? sm[LF(y) + i] : nan("");
            for (int i = tid; i < ms; i += NT) a.y_opt[b * m + 2 * n + i] = ok ? sm[L.ys + i] : nan("");
        }
        if (a.X_opt && tid == 0) {
            // X_k of the QP (x_{k+1} = A_k x_k + B_k u_k + g_k, :189-192).  The nominal rollout satisfies the same recursion
            // with u = u_prev, so X_k = xbar_k + dx_k with dx_{k+1} = A_k dx_k + B_k dU_k, dx_0 = 0.
            double *X = a.X_opt + (size_t)b * (N + 1) * 6;
            double dx[6] = {0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 6; ++i) X[i] = ok ? sm[LF(xbar) + i] : nan("");
            for (int k = 0; k < N; ++k) {
                const double *r_ = sm + LF(lin) + TG_LIN * k;
                const double d0 = sm[LF(xt) + 2 * k], d1 = sm[LF(xt) + 2 * k + 1];
                double nx[6];
                nx[0] = dx[0] + r_[0] * dx[2] + r_[1] * dx[3] + r_[2] * dx[4];
                nx[1] = dx[1] + r_[3] * dx[2] + r_[4] * dx[3] + r_[5] * dx[4];
                nx[2] = dx[2] + r_[6] * dx[5];
                nx[3] = r_[7] * dx[3] + r_[8] * dx[4] + r_[9] * dx[5] + r_[16] * d0 + r_[17] * d1;
                nx[4] = r_[10] * dx[3] + r_[11] * dx[4] + r_[12] * dx[5] + r_[18] * d1;
                nx[5] = r_[13] * dx[3] + r_[14] * dx[4] + r_[15] * dx[5] + r_[19] * d1;
                for (int i = 0; i < 6; ++i) { dx[i] = nx[i]; X[6 * (k + 1) + i] = ok ? sm[LF(xbar) + 6 * (k + 1) + i] + dx[i] : nan(""); }
            }
        }
        if (a.ws_valid && c.warm_start) {
            const size_t m = 2 * (size_t)n + ms;
            for (int i = tid; i < n; i += NT) a.ws_x[(size_t)b * n + i] = sm[LF(xt) + i];
            for (int i = tid; i < 2 * n; i += NT) a.ws_y[b * m + i] = sm[LF(y) + i];
            for (int i = tid; i < ms; i += NT) a.ws_y[b * m + 2 * n + i] = sm[L.ys + i];
            if (tid == 0) a.ws_valid[b] = ok ? (r.free_end ? 2 : 1) : 0;
        }
    }
}

template <int W, int S, int NC, bool P1 = false, bool HS = (NC == 0)>
__global__ void __launch_bounds__(TW_CTA_THREADS) __maxnreg__(TW_REGS(S))
tw_closed_loop_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ WLayout L, const __grid_constant__ LoopArgs a)
{
    extern __shared__ __align__(16) double sm_all[];
    constexpr int NT = 32 * W;
    const int P = P1 ? 1 : a.ppc;
    int prob = P1 ? 0 : threadIdx.x / NT, tid = P1 ? threadIdx.x : threadIdx.x % NT;
    unsigned sm_off = (unsigned)prob * (unsigned)((L.total + 1) & ~1) * 8u;
    if constexpr (!P1) asm volatile("" : "+r"(tid), "+r"(prob), "+r"(sm_off));
    const int bar = P1 ? 1 : 1 + prob;
    double *sm = reinterpret_cast<double *>(reinterpret_cast<char *>(sm_all) + sm_off);
    const int N = NC > 0 ? NC : c.N, n = 2 * N, ms = HS ? c.ms : 0, ns = HS ? c.ns : 0, T = a.T;
    if (prob >= P) return;
    const TwMap<S> mp = tw_make_map<W, S>(LF(nb), tid);
    const StepTaps tap = {};
    int *cnt = reinterpret_cast<int *>(sm + LF(misc) + M_CNT);                 // 6 status counters
    long long *itsum = reinterpret_cast<long long *>(sm + LF(misc) + M_CNT + 3);
    const bool my = tid < N;
    const int j0 = 2 * tid;
    // Trajectories are dealt round-robin over the CTAs: slot `prob` of CTA c takes b = c + G (round P + prob), G = gridDim.x;
    // `active` (the busy slots of this CTA in this round) is what the step barrier counts.
    const int G = gridDim.x;
    for (int round = 0; blockIdx.x + (long long)G * round * P < a.B; ++round) {
        const long long bfirst = blockIdx.x + (long long)G * round * P;
        const long long left = (a.B - 1 - bfirst) / G + 1;                   // slots with a trajectory, >= 1
        const int active = left < P ? (int)left : P;
        const bool lockstep = !P1 && active > 1;
        if (prob >= active) break;                                           // later rounds have no work for this slot either
        const int b = (int)(bfirst + (long long)G * prob);
        tw_sync<W>(bar);
        if (tid < 12) sm[LF(spec) + tid] = reinterpret_cast<const double *>(a.spec + b)[tid];   // scenario -> shared memory
        if (tid < 6) { const double v_ = a.x0[6 * (size_t)b + tid]; sm[LF(x0) + tid] = v_; a.clean[(size_t)b * (T + 1) * 6 + tid] = v_; }
        if (tid < 2) sm[LF(uprev) + tid] = a.u0[2 * (size_t)b + tid];
        if (tid < TG_NUM_STATUS) cnt[tid] = 0;
        if (tid == 0) *itsum = 0;
        bool warm = false, warm_free = false;
        tw_sync<W>(bar);
#pragma unroll 1
        for (int t = 0; t <= T; ++t) {
            FusedCtx fx;
            fx.brk = a.brk; fx.coef = a.coef;
            fx.noisy_row = a.noisy + ((size_t)b * (T + 1) + t) * 6;
            fx.seed = c.seed_base + (unsigned long long)(a.traj_id0 + b);
            fx.t_index = t;
            if (t == T) {   // last row: only its noisy copy remains to be written
                if (tid < 3) {
                    uint32_t r4[4];
                    tg_philox4x32_10((uint32_t)t, (uint32_t)(tid >> 1), 0u, 0u, (uint32_t)fx.seed, (uint32_t)(fx.seed >> 32), r4);
                    double n0, n1;
                    tg_box_muller(r4[(tid & 1) * 2], r4[(tid & 1) * 2 + 1], n0, n1);
                    fx.noisy_row[2 * tid] = sm[LF(x0) + 2 * tid] + c.noise_std[2 * tid] * n0;
                    fx.noisy_row[2 * tid + 1] = sm[LF(x0) + 2 * tid + 1] + c.noise_std[2 * tid + 1] * n1;
                }
                break;
            }
            const StepResult r = tw_step_body<W, S, NC, HS>(c, L, sm, mp, warm, warm_free, tap, &fx, tid, bar);
            const bool ok = (r.status == TG_STATUS_OPTIMAL || r.status == TG_STATUS_OPTIMAL_INACCURATE);
            if (tid == 0) { cnt[r.status] += 1; *itsum += r.iters; }
            // shifted warm start for the next step, in the next step's dU coordinates
            double2 nx = make_double2(0.0, 0.0), nyb = nx, nyr = nx;
            if (my) {
                const double2 d0 = tw_ld2(sm + LF(xt));
                const double2 xs_ = tw_ld2(sm + LF(xt) + ((tid + 1 < N) ? j0 + 2 : j0));
                nx = make_double2(xs_.x - d0.x, xs_.y - d0.y);
                if (tid + 1 < N) { nyb = tw_ld2(sm + LF(y) + j0 + 2); nyr = tw_ld2(sm + LF(y) + n + j0 + 2); }
            }
            if (tid < 32) {   // plant (MPC/main.py:97) + outputs, warp 0 (lane-parallel f_cont)
                const double ud = sm[LF(uprev)], udel = sm[LF(uprev) + 1];
                const double u0 = ok ? ud + sm[LF(xt)] : ud, u1 = ok ? udel + sm[LF(xt) + 1] : udel;
                double xs[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) xs[i] = sm[LF(x0) + i];
                tg_plant_step_lanes(c, xs, u0, u1, tid);
                if (tid == 0) {
                    double *clean = a.clean + ((size_t)b * (T + 1) + t + 1) * 6;
#pragma unroll
                    for (int i = 0; i < 6; ++i) { sm[LF(misc) + M_XNEXT + i] = xs[i]; clean[i] = xs[i]; }
                    a.U[((size_t)b * T + t) * 2] = u0; a.U[((size_t)b * T + t) * 2 + 1] = u1;
                    sm[LF(misc) + M_UCMD] = u0; sm[LF(misc) + M_UCMD + 1] = u1;
                }
            }
            tw_sync<W>(bar);
            // commit the new state / warm start
            if (tid < 6) sm[LF(x0) + tid] = sm[LF(misc) + M_XNEXT + tid];
            if (tid < 2) sm[LF(uprev) + tid] = sm[LF(misc) + M_UCMD + tid];
            if (my) { tw_st2(sm + LF(x) + j0, nx.x, nx.y); tw_st2(sm + LF(y) + j0, nyb.x, nyb.y); tw_st2(sm + LF(y) + n + j0, nyr.x, nyr.y); }
            if (ms > 0) {   // shift the state-row duals by one stage (ns rows), in place: ascending order, one thread
                if (tid == 0) {
                    for (int i = 0; i + ns < ms; ++i) sm[L.ys + i] = sm[L.ys + i + ns];
                    for (int i = (ms > ns ? ms - ns : 0); i < ms; ++i) sm[L.ys + i] = 0.0;
                }
            }
            warm = ok && c.warm_start;
            warm_free = warm && r.free_end;
            if constexpr (P1) tw_sync<W>(bar);
            else if (lockstep) tg_sync(15, NT * active); else tw_sync<W>(bar);
        }
        tw_sync<W>(bar);
        if (tid < TG_NUM_STATUS && a.status_counts) a.status_counts[(size_t)b * TG_NUM_STATUS + tid] = cnt[tid];
        if (tid == 0 && a.iters_total) a.iters_total[b] = *itsum;
    }
}

// a10 tap: the reference window of tw_ref_window_warp for B states, one warp per problem
__global__ void tg_ref_window_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ WLayout L, int B,
                                     const double *x0, const tg_ref_spec *spec, const double *brk, const double *coef,
                                     int t_index, double *path_ref, double *vref)
{
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x, N = c.N;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncwarp();
        if (lane < 6) sm[L.x0 + lane] = x0[6 * (size_t)b + lane];
        __syncwarp();
        const tg_ref_spec sp = spec[b];
        tw_ref_window_warp<0, 1>(c, L, sm, sp, brk, coef, t_index, lane);
        __syncwarp();
        for (int k = lane; k <= N; k += 32) {
            double *pr = path_ref + 3 * ((size_t)b * (N + 1) + k);
            pr[0] = sm[L.Xr + k]; pr[1] = sm[L.Yr + k]; pr[2] = sm[L.Pr + k];
            vref[(size_t)b * (N + 1) + k] = sm[L.vref + k];
        }
    }
}

__global__ void tg_plant_kernel(const __grid_constant__ DevCfg c, int B, int T, const double *x0, const double *U, double *X)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xs[6];
    double *Xo = X + (size_t)b * (T + 1) * 6;
    for (int i = 0; i < 6; ++i) { xs[i] = x0[6 * (size_t)b + i]; Xo[i] = xs[i]; }
    for (int t = 0; t < T; ++t) {
        if (c.plant == TG_PLANT_MPC) tg_plant_step(c, xs, U[((size_t)b * T + t) * 2], U[((size_t)b * T + t) * 2 + 1]);
        else tg_plant_step_gen(c, xs, U[((size_t)b * T + t) * 2], U[((size_t)b * T + t) * 2 + 1]);
        for (int i = 0; i < 6; ++i) Xo[6 * (size_t)(t + 1) + i] = xs[i];
    }
}

__global__ void tg_noise_kernel(unsigned long long seed0, int n_traj, int n_rows, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_traj * n_rows) return;
    const int tr = (int)(i / n_rows), row = (int)(i % n_rows);
    double nz[6];
    tg_noise_row(seed0 + (unsigned long long)tr, (uint32_t)row, nz);
    for (int k = 0; k < 6; ++k) out[6 * i + k] = nz[k];
}

__global__ void tg_philox_kernel(unsigned long long seed, uint32_t first, uint32_t block, int n, uint32_t *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[4];
    tg_philox4x32_10(first + (uint32_t)i, block, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int k = 0; k < 4; ++k) out[4 * (size_t)i + k] = r[k];
}

template <typename T>
__global__ void tg_fma_peak_kernel(T *out, int iters)
{
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3, a4 = a0 + (T)4, a5 = a0 + (T)5,
      a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T b = (T)0.999, cc = (T)1e-4;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * b + cc; a1 = a1 * b + cc; a2 = a2 * b + cc; a3 = a3 * b + cc;
        a4 = a4 * b + cc; a5 = a5 * b + cc; a6 = a6 * b + cc; a7 = a7 * b + cc;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------ host side
// Tyre-curve table (tg_device.cuh: tg_tyre_tab): per interval the degree-(NC-1) Chebyshev interpolant of
// g(alpha) = sin(C atan(B alpha)) in long double, converted to monomials of s in [-1, 1].  Returns the largest
// deviation from libm on a dense check grid (value and slope) so the caller can refuse a table that is not at
// rounding level for unusual B, C.
static double build_cheb_table(const std::function<long double(long double)> &fun, const std::function<long double(long double)> &dfun,
                               long double lo, long double hi, int NI, double *tab /*[NI + 1][NC]*/, double *slope_err)
{
    // NI + 1 intervals of width h = (hi - lo) / NI, interval i centred on lo + i h (tg_tab_row): the end intervals reach half
    // an interval beyond [lo, hi], where fun is still analytic
    const int NC = TG_TAB_NC;
    const long double PI = 3.141592653589793238462643383279502884L;
    long double T[TG_TAB_NC][TG_TAB_NC] = {};   // T[k][i] = coefficient of s^i in T_k(s)
    T[0][0] = 1.0L;
    if (NC > 1) T[1][1] = 1.0L;
    for (int k = 2; k < NC; ++k)
        for (int i = 0; i < NC; ++i) T[k][i] = (i > 0 ? 2.0L * T[k - 1][i - 1] : 0.0L) - T[k - 2][i];
    const long double hw = (hi - lo) / (2 * NI);
    double worst = 0.0, worst_d = 0.0;
    for (int it = 0; it <= NI; ++it) {
        const long double ctr = lo + 2 * it * hw;
        long double fv[TG_TAB_NC], a[TG_TAB_NC];
        for (int j = 0; j < NC; ++j) {
            const long double sj = cosl(PI * (j + 0.5L) / NC);
            fv[j] = fun(ctr + hw * sj);
        }
        for (int k = 0; k < NC; ++k) {
            long double acc = 0.0L;
            for (int j = 0; j < NC; ++j) acc += fv[j] * cosl(PI * k * (j + 0.5L) / NC);
            a[k] = acc * (k == 0 ? 1.0L : 2.0L) / NC;
        }
        for (int i = 0; i < NC; ++i) {
            long double m = 0.0L;
            for (int k = 0; k < NC; ++k) m += a[k] * T[k][i];
            tab[it * NC + i] = (double)m;
        }
        for (int q = 0; q <= 16; ++q) {   // check grid incl. the interval ends
            const double sq = -1.0 + q / 8.0;
            double v = tab[it * NC + NC - 1], dv = 0.0;
            for (int i = NC - 2; i >= 0; --i) { dv = dv * sq + v; v = v * sq + tab[it * NC + i]; }
            const long double al = ctr + hw * sq;
            worst = fmax(worst, fabs((double)(v - fun(al))));
            worst_d = fmax(worst_d, fabs((double)(dv / (double)hw - dfun(al))));
        }
    }
    if (slope_err) *slope_err = worst_d;
    return worst;
}

static double build_tyre_table(double B, double C, double ma, double *tab /*[TG_TAB_ROWS][NC]*/, double *slope_err)
{
    auto f = [=](long double al) { return sinl((long double)C * atanl((long double)B * al)); };
    auto df = [=](long double al) { const long double ba = (long double)B * al; return cosl((long double)C * atanl(ba)) * C * B / (1.0L + ba * ba); };
    return build_cheb_table(f, df, -(long double)ma, (long double)ma, TG_TAB_NI, tab, slope_err);
}

// atan on [-TG_ATAN_T0, TG_ATAN_T0] (the slip-angle argument n / vx_eff of the sequential nominal rollout): same form
static double build_atan_table(double *tab /*[TG_ATAN_ROWS][NC]*/)
{
    auto f = [](long double t) { return atanl(t); };
    auto df = [](long double t) { return 1.0L / (1.0L + t * t); };
    return build_cheb_table(f, df, -(long double)TG_ATAN_T0, (long double)TG_ATAN_T0, TG_ATAN_NI, tab, nullptr);
}

struct tg_handle {
    tg_config cfg;
    DevCfg dc;
    int device, num_sms, grid_cap;
    int ppc_max, ppc_env;     // closed-loop kernel: most problems per CTA that fit / TRAJGEN_PPC override (0 = automatic)
    // warp-per-problem kernels (tw_solver.cuh): W warps per problem, S tile blocks per thread
    int W, S;
    WLayout WL;
    WLayout Lref;             // layout of the reference-window tap kernel (no state rows, one warp)
    size_t smem_optin;
    size_t smem_bytes;
    cudaStream_t stream;
    long long launches;
    double *tyre_tab;   // device copy of the tyre-curve table, or null (fit not at rounding level -> atan/sin path)
    double tab_err[2];  // max |table - libm| of value and slope on the check grid
    // warm-start state for the step API
    double *ws_x, *ws_y; int *ws_valid; int ws_B;
    // staging for *_host entry points
    void *dstage; size_t dstage_bytes;
    void *hstage; size_t hstage_bytes;
};

static bool pick_tw_shape(int N, int &W, int &S)
{
    if (const char *e = getenv("TRAJGEN_SHAPE")) {   // development knob: "W,S" (must fit the horizon)
        int w = 0, s_ = 0;
        if (sscanf(e, "%d,%d", &w, &s_) == 2 && w >= 1 && s_ >= 1) {
            const int nb = (2 * N + 3) / 4;
            if (nb * (nb + 1) / 2 <= 32 * w * s_ && N <= 32 * w && 2 * N <= 64 * w) { W = w; S = s_; return true; }
        }
    }
    if (N <= 14) { W = 1; S = 1; }          // 28 blocks on one warp
    else if (N <= 20) { W = 2; S = 1; }     // 55 blocks on two warps
    else if (N <= 30) { W = 4; S = 1; }     // 120 blocks
    else if (N <= 44) { W = 8; S = 1; }     // 253 blocks
    else if (N <= 56) { W = 8; S = 2; }     // 406 blocks on 512 slots
    else return false;
    return true;
}
// kernel instances: (W, S) for a run-time horizon, plus the horizons of the reference / the BASELINE configurations compiled in
// (NC = N: shared-memory offsets become immediates, horizon loops get constant trip counts: 25-50 % faster).  The compile-time
// instances serve the standard controller only -- MPC tyre model from the tables, analytic Jacobians (tw_standard_controller) --
// and come without (HS = false) or with (HS = true) the code of the state-bound rows.
static bool tw_standard_controller(const DevCfg &d)
{
    return d.model == TG_MODEL_MPC && d.jacobian == TG_JAC_ANALYTIC && d.tyre_tab && d.atan_tab;
}
template <typename F>
static int dispatch_tw(int W, int S, const DevCfg &d, F &&f)
{
    using std::integral_constant;
    typedef integral_constant<bool, false> no_rows;
    typedef integral_constant<bool, true> rows;
    const int N = d.N;
    if (tw_standard_controller(d) && !getenv("TRAJGEN_DYNAMIC_N")) {   // mpc_step's default horizon (20), MPC/main.py's (40), BASELINE config 4's (10, 50)
        if (d.ms == 0) {
            if (W == 2 && S == 1 && N == 20) return f(integral_constant<int, 2>(), integral_constant<int, 1>(), integral_constant<int, 20>(), no_rows());
#ifndef TG_DEV_SHAPES_ONLY
            if (W == 1 && S == 1 && N == 10) return f(integral_constant<int, 1>(), integral_constant<int, 1>(), integral_constant<int, 10>(), no_rows());
            if (W == 8 && S == 1 && N == 40) return f(integral_constant<int, 8>(), integral_constant<int, 1>(), integral_constant<int, 40>(), no_rows());
            if (W == 8 && S == 2 && N == 50) return f(integral_constant<int, 8>(), integral_constant<int, 2>(), integral_constant<int, 50>(), no_rows());
#endif
        } else {
            if (W == 2 && S == 1 && N == 20) return f(integral_constant<int, 2>(), integral_constant<int, 1>(), integral_constant<int, 20>(), rows());
#ifndef TG_DEV_SHAPES_ONLY
            if (W == 1 && S == 1 && N == 10) return f(integral_constant<int, 1>(), integral_constant<int, 1>(), integral_constant<int, 10>(), rows());
            if (W == 8 && S == 2 && N == 50) return f(integral_constant<int, 8>(), integral_constant<int, 2>(), integral_constant<int, 50>(), rows());
#endif
        }
    }
    if (W == 1 && S == 1) return f(integral_constant<int, 1>(), integral_constant<int, 1>(), integral_constant<int, 0>(), rows());
    if (W == 2 && S == 1) return f(integral_constant<int, 2>(), integral_constant<int, 1>(), integral_constant<int, 0>(), rows());
#ifndef TG_DEV_SHAPES_ONLY
    if (W == 4 && S == 1) return f(integral_constant<int, 4>(), integral_constant<int, 1>(), integral_constant<int, 0>(), rows());
    if (W == 8 && S == 1) return f(integral_constant<int, 8>(), integral_constant<int, 1>(), integral_constant<int, 0>(), rows());
    if (W == 8 && S == 2) return f(integral_constant<int, 8>(), integral_constant<int, 2>(), integral_constant<int, 0>(), rows());
#endif
    return fail(TG_ERR_UNSUPPORTED, "no kernel shape for this horizon");
}
// instances that exist only in their single-problem-per-CTA form: compile-time horizon + state-bound rows on one or two warps
// (problems with state-bound rows always run one per CTA, choose_tw_ppc)
#define TW_P1_ONLY(W, NC, HS) ((HS) && (NC) > 0 && (W) <= 2)
static size_t tw_stride(const tg_handle *h) { return (((size_t)h->WL.total + 1) & ~(size_t)1) * sizeof(double); }

extern "C" {

const char *tg_last_error(void) { return g_err.c_str(); }
int tg_version(void) { return TG_VERSION; }

void tg_default_config(tg_config *c)
{
    memset(c, 0, sizeof(*c));
    c->N = 20; c->model = TG_MODEL_MPC; c->plant = TG_PLANT_MPC; c->jacobian = TG_JAC_ANALYTIC;
    c->Ts = 0.02;
    const double p[TG_NPARAMS] = {0.287, 0.0545, 0.0518, 0.00035, 3.3852, 1.2691, 0.1737, 2.579, 1.2, 0.192,
                                  0.041, 27.8e-6, 0.029, 0.033, 9.81, 0.6, 0.3};
    memcpy(c->params, p, sizeof(p));
    c->q_c = 6.0; c->q_phi = 0.5; c->q_vx = 0.5;
    c->R[0] = 0.02; c->R[3] = 2.0; c->Rd[0] = 0.01; c->Rd[3] = 5.0;
    c->u_lo[0] = -1.0; c->u_hi[0] = 1.0; c->u_lo[1] = -0.6; c->u_hi[1] = 0.6;
    c->du_lo[0] = -0.5; c->du_hi[0] = 0.5; c->du_lo[1] = -0.3; c->du_hi[1] = 0.3;
    for (int i = 0; i < 6; ++i) { c->x_lo[i] = -TG_INF; c->x_hi[i] = TG_INF; }
    c->rho = 0.1; c->sigma = 1e-6; c->alpha = 1.6; c->eps_abs = 1e-6; c->eps_rel = 1e-6; c->eps_prim_inf = 1e-4;   // eps: one decade tighter than CVXPY's OSQP setting (1e-5), DESIGN.md 2
    c->adaptive_rho_tol = 5.0; c->alpha_warm = 1.2; c->max_iter = 10000; c->check_every = 5; c->adaptive_rho = 1; c->adaptive_rho_min_iter = 20;
    c->warm_start = 0; c->vref_advance = 0;
    const double sd[6] = {0.05, 0.05, 0.003, 0.010, 0.003, 0.030};
    memcpy(c->noise_std, sd, sizeof(sd));
    c->noise_seed_base = 12345ull;
}

int tg_device_count(int *n) { CK(cudaGetDeviceCount(n)); return TG_OK; }

int tg_create(const tg_config *cfg, int device, tg_handle **out)
{
    if (!cfg || !out) return fail(TG_ERR_INVALID, "null argument");
    if (cfg->N < 1) return fail(TG_ERR_INVALID, "N must be >= 1");
    if (!(cfg->Ts > 0)) return fail(TG_ERR_INVALID, "Ts must be > 0");
    if (cfg->max_iter < 1 || cfg->check_every < 1) return fail(TG_ERR_INVALID, "max_iter and check_every must be >= 1");
    if (!(cfg->rho > 0) || !(cfg->sigma > 0) || !(cfg->alpha > 0 && cfg->alpha < 2)) return fail(TG_ERR_INVALID, "rho, sigma > 0 and 0 < alpha < 2 required");
    int W_ = 0, S_ = 0;
    if (!pick_tw_shape(cfg->N, W_, S_)) return fail(TG_ERR_UNSUPPORTED, "horizon N > 56 is not supported");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail(TG_ERR_CUDA, "no CUDA device: libtrajgen has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(TG_ERR_INVALID, "bad device index");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(TG_ERR_UNSUPPORTED, "libtrajgen is built for sm_100a (B200) only");

    tg_handle *h = new (std::nothrow) tg_handle();
    if (!h) return fail(TG_ERR_NOMEM, "out of host memory");
    memset(h, 0, sizeof(*h));
    struct Guard { tg_handle *p; ~Guard() { if (p) { cudaFree(p->tyre_tab); delete p; } } } guard{h};   // released on success
    h->cfg = *cfg; h->W = W_; h->S = S_; h->device = device; h->num_sms = prop.multiProcessorCount;
    DevCfg &d = h->dc;
    d.N = cfg->N; d.n = 2 * cfg->N; d.model = cfg->model; d.plant = cfg->plant; d.jacobian = cfg->jacobian;
    d.ns = 0;
    for (int i = 0; i < 6; ++i)
        if (cfg->x_lo[i] > -TG_INF || cfg->x_hi[i] < TG_INF) d.sidx[d.ns++] = i;
    d.ms = d.ns * d.N; d.m = 4 * d.N + d.ms;
    d.max_iter = cfg->max_iter; d.check_every = cfg->check_every; d.adaptive_rho = cfg->adaptive_rho;
    d.adaptive_rho_min_iter = cfg->adaptive_rho_min_iter; d.warm_start = cfg->warm_start; d.vref_advance = cfg->vref_advance;
    d.Ts = cfg->Ts;
    memcpy(d.p, cfg->params, sizeof(d.p));
    d.inv_m = 1.0 / d.p[P_m]; d.inv_Iz = 1.0 / d.p[P_Iz];
    d.tyre_tab = nullptr; d.tab_scale = 0.0;
    if (d.p[P_maxAlpha] > 0.0 && !getenv("TRAJGEN_NO_TYRE_TABLE")) {
        std::vector<double> tab(2 * TG_TAB_ROWS * TG_TAB_NC + TG_ATAN_ROWS * TG_TAB_NC);
        double se_f = 0.0, se_r = 0.0;
        const double e_f = build_tyre_table(d.p[P_Bf], d.p[P_Cf], d.p[P_maxAlpha], tab.data(), &se_f);
        const double e_r = build_tyre_table(d.p[P_Br], d.p[P_Cr], d.p[P_maxAlpha], tab.data() + TG_TAB_ROWS * TG_TAB_NC, &se_r);
        const double e_at = build_atan_table(tab.data() + 2 * TG_TAB_ROWS * TG_TAB_NC);
        h->tab_err[0] = fmax(e_f, e_r); h->tab_err[1] = fmax(se_f, se_r);
        if (h->tab_err[0] <= 4e-16 && h->tab_err[1] <= 1e-12) {
            CK(cudaMalloc(&h->tyre_tab, tab.size() * sizeof(double)));
            CK(cudaMemcpy(h->tyre_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
            d.tyre_tab = h->tyre_tab;
            d.tab_scale = TG_TAB_NI / (2.0 * d.p[P_maxAlpha]);
            if (e_at <= 4e-16) d.atan_tab = h->tyre_tab + 2 * TG_TAB_ROWS * TG_TAB_NC;
        }
    }
    d.q_c = cfg->q_c; d.q_phi = cfg->q_phi; d.q_vx = cfg->q_vx;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            d.Rs[i * 2 + j] = 0.5 * (cfg->R[i * 2 + j] + cfg->R[j * 2 + i]);
            d.Rds[i * 2 + j] = 0.5 * (cfg->Rd[i * 2 + j] + cfg->Rd[j * 2 + i]);
        }
    memcpy(d.u_lo, cfg->u_lo, 16); memcpy(d.u_hi, cfg->u_hi, 16); memcpy(d.du_lo, cfg->du_lo, 16); memcpy(d.du_hi, cfg->du_hi, 16);
    memcpy(d.x_lo, cfg->x_lo, 48); memcpy(d.x_hi, cfg->x_hi, 48);
    d.rho = cfg->rho; d.sigma = cfg->sigma; d.alpha = cfg->alpha;
    d.alpha_warm = (cfg->alpha_warm > 0.0 && cfg->alpha_warm < 2.0) ? cfg->alpha_warm : cfg->alpha; d.eps_abs = cfg->eps_abs; d.eps_rel = cfg->eps_rel;
    d.eps_pinf = cfg->eps_prim_inf; d.adapt_tol = cfg->adaptive_rho_tol > 1.0 ? cfg->adaptive_rho_tol : 5.0;
    memcpy(d.noise_std, cfg->noise_std, 48);
    d.seed_base = cfg->noise_seed_base;

    d.free_mode = (cfg->solver_flags & 1) ? 0 : 1;
    h->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
    h->Lref = tw_make_layout(d.N, 0, 1);
    CK(cudaFuncSetAttribute(tg_ref_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
    h->ppc_env = 0;   // TRAJGEN_PPC pins the number of problems per CTA (measurement knob)
    {
        h->WL = tw_make_layout(d.N, d.ms, h->W);
        const size_t stride = tw_stride(h);
        h->smem_bytes = stride;
        if (stride > h->smem_optin) return fail(TG_ERR_UNSUPPORTED, "state-bound rows x horizon exceed the 227 KB shared memory of one CTA");
        h->ppc_max = TW_CTA_THREADS / (32 * h->W);
        if (h->W > 1 && h->ppc_max > 14) h->ppc_max = 14;                  // named barriers 1..14 + the step barrier 15
        while (h->ppc_max > 1 && (size_t)h->ppc_max * stride > h->smem_optin) h->ppc_max -= 1;
        if (const char *e = getenv("TRAJGEN_PPC")) { const int v = atoi(e); if (v >= 1 && v <= h->ppc_max) h->ppc_env = v; }
        int occ = 0;
        int rc = dispatch_tw(h->W, h->S, d, [&](auto W_, auto S_, auto NC_, auto HS_) -> int {
            constexpr int W = decltype(W_)::value, S = decltype(S_)::value, NC = decltype(NC_)::value;
            constexpr bool HS = decltype(HS_)::value;
            // the attribute is per kernel function and process-wide: always raise it to the device limit so that handles with
            // different layouts can be used side by side
            if constexpr (!TW_P1_ONLY(W, NC, HS)) {
                CK(cudaFuncSetAttribute(tw_mpc_step_kernel<W, S, NC, false, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
                CK(cudaFuncSetAttribute(tw_closed_loop_kernel<W, S, NC, false, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
                CK(cudaFuncSetAttribute(tw_mpc_step_kernel<W, S, NC, false, HS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                CK(cudaFuncSetAttribute(tw_closed_loop_kernel<W, S, NC, false, HS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            }
            if constexpr (TW_HAS_P1(W, NC, HS)) {
                CK(cudaFuncSetAttribute(tw_mpc_step_kernel<W, S, NC, true, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
                CK(cudaFuncSetAttribute(tw_closed_loop_kernel<W, S, NC, true, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
                CK(cudaFuncSetAttribute(tw_mpc_step_kernel<W, S, NC, true, HS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                CK(cudaFuncSetAttribute(tw_closed_loop_kernel<W, S, NC, true, HS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tw_closed_loop_kernel<W, S, NC, true, HS>, 32 * W, stride));
            } else {
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tw_closed_loop_kernel<W, S, NC, false, HS>, 32 * W, stride));
            }
            return TG_OK;
        });
        if (rc != TG_OK) return rc;
        if (occ < 1) return fail(TG_ERR_UNSUPPORTED, "kernel does not fit on an SM with this configuration");
        h->grid_cap = occ * h->num_sms;     // resident problems with one problem per CTA
    }
    h->stream = 0;
    guard.p = nullptr;
    *out = h;
    return TG_OK;
}

int tg_destroy(tg_handle *h)
{
    if (!h) return TG_OK;
    cudaSetDevice(h->device);
    cudaFree(h->tyre_tab); cudaFree(h->ws_x); cudaFree(h->ws_y); cudaFree(h->ws_valid); cudaFree(h->dstage);
    if (h->hstage) cudaFreeHost(h->hstage);
    delete h;
    return TG_OK;
}

int tg_set_stream(tg_handle *h, void *s) { if (!h) return fail(TG_ERR_INVALID, "null handle"); h->stream = (cudaStream_t)s; return TG_OK; }
int tg_synchronize(tg_handle *h) { if (!h) return fail(TG_ERR_INVALID, "null handle"); CK(cudaStreamSynchronize(h->stream)); return TG_OK; }
int tg_info(tg_handle *h, int32_t *ctas_per_sm, int32_t *threads_per_cta, int32_t *smem_bytes, int32_t *num_sms)
{
    if (!h) return fail(TG_ERR_INVALID, "null handle");
    if (ctas_per_sm) *ctas_per_sm = h->grid_cap / h->num_sms;
    if (threads_per_cta) *threads_per_cta = 32 * h->W;
    if (smem_bytes) *smem_bytes = (int32_t)h->smem_bytes;
    if (num_sms) *num_sms = h->num_sms;
    return TG_OK;
}
#ifdef TG_PHASE_TIMING
int tg_debug_phases(long long *out16, int reset)
{
    if (out16) CK(cudaMemcpyFromSymbol(out16, g_tg_phase, sizeof(long long) * 16));
    if (reset) { long long z[16] = {0}; CK(cudaMemcpyToSymbol(g_tg_phase, z, sizeof(z))); }
    return TG_OK;
}
#endif
int tg_tyre_table_info(tg_handle *h, int32_t *in_use, double *max_value_err, double *max_slope_err)
{
    if (!h) return fail(TG_ERR_INVALID, "null handle");
    if (in_use) *in_use = (h->dc.tyre_tab ? 1 : 0) | (h->dc.atan_tab ? 2 : 0);   // bit 0: tyre curve, bit 1: slip-angle atan
    if (max_value_err) *max_value_err = h->tab_err[0];
    if (max_slope_err) *max_slope_err = h->tab_err[1];
    return TG_OK;
}
int tg_kernel_launches(tg_handle *h, int64_t *count) { if (!h || !count) return fail(TG_ERR_INVALID, "null argument"); *count = h->launches; return TG_OK; }

// warp-per-problem kernels: problems per CTA.  Resident problems per SM are bounded by registers (512 threads x 128) and
// shared memory either way; larger CTAs share more instruction fetches, smaller ones spread a small batch over more SMs.
static int choose_tw_ppc(const tg_handle *h, int B)
{
    if (h->ppc_env) return h->ppc_env;
    // problems with state-bound rows have long, unequal solves (tens to hundreds of ADMM iterations): free-running
    // single-problem CTAs beat problems that wait for one another at every step boundary
    if (h->dc.ms > 0) return 1;
    // measured at N = 20 (two warps per problem): 4 problems per CTA 2.65e7 steps/s (B = 2368), 8 per CTA 2.37e7 (B = 4736),
    // 2 per CTA 2.24e7 and 1 per CTA 1.43e7 (B = 1024): sharing instruction fetches matters, waiting for 7 neighbours costs more
    int p = h->ppc_max < 4 ? h->ppc_max : 4;
    while (p > 1 && (B + p - 1) / p < h->num_sms) p >>= 1;   // every SM gets a CTA before CTAs get wider
    return p < 1 ? 1 : p;
}

extern "C++" {
template <typename Kern, typename Args>
static int launch_tw(tg_handle *h, Kern kern, Args &a, int B, int W, int ppc_fixed = 0)
{
    const int ppc = ppc_fixed > 0 ? ppc_fixed : choose_tw_ppc(h, B);   // ppc_fixed = 1: a single-problem-per-CTA instance
    a.ppc = ppc;
    const size_t smem = (size_t)ppc * tw_stride(h);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * W * ppc, smem));
    if (per_sm < 1) return fail(TG_ERR_UNSUPPORTED, "kernel does not fit on an SM with this configuration");
    const int cap = per_sm * h->num_sms, ctas = (B + ppc - 1) / ppc;
    int grid = ctas < cap ? ctas : cap;
    if (const char *e = getenv("TRAJGEN_GRID")) { const int v = atoi(e); if (v >= 1 && v <= cap) grid = v; }   // measurement knob
    kern<<<grid, 32 * W * ppc, smem, h->stream>>>(h->dc, h->WL, a);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}
}  // extern "C++"

static int launch_step(tg_handle *h, StepArgs &a)
{
    if (a.B <= 0) return TG_OK;   // empty batch: nothing to do
    CK(cudaSetDevice(h->device));
    return dispatch_tw(h->W, h->S, h->dc, [&](auto W_, auto S_, auto NC_, auto HS_) -> int {
        constexpr int W = decltype(W_)::value, S = decltype(S_)::value, NC = decltype(NC_)::value;
        constexpr bool HS = decltype(HS_)::value;
        if constexpr (TW_P1_ONLY(W, NC, HS)) {
            return launch_tw(h, tw_mpc_step_kernel<W, S, NC, true, HS>, a, a.B, W, 1);
        } else {
            if constexpr (TW_HAS_P1(W, NC, HS))
                if (choose_tw_ppc(h, a.B) == 1) return launch_tw(h, tw_mpc_step_kernel<W, S, NC, true, HS>, a, a.B, W, 1);
            return launch_tw(h, tw_mpc_step_kernel<W, S, NC, false, HS>, a, a.B, W);
        }
    });
}

int tg_linearize(tg_handle *h, int B, const double *x0, const double *u_prev, double *A, double *Bm, double *g, double *xbar)
{
    if (!h || B < 0 || (B > 0 && (!x0 || !u_prev))) return fail(TG_ERR_INVALID, "bad argument");
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.x0 = x0; a.u_prev = u_prev; a.path_ref = nullptr; a.vref = nullptr;
    a.A = A; a.Bm = Bm; a.g = g; a.xbar = xbar; a.stop = 1;
    // the tap does not read the reference window; point it at x0 so loads stay in bounds
    a.path_ref = nullptr;
    return launch_step(h, a);
}

int tg_assemble(tg_handle *h, int B, const double *x0, const double *u_prev, const double *path_ref, const double *vref,
                double *H, double *q, double *c0, double *l, double *u, double *Gs)
{
    if (!h || B < 0 || (B > 0 && (!x0 || !u_prev || !path_ref))) return fail(TG_ERR_INVALID, "bad argument");
    if ((l == nullptr) != (u == nullptr)) return fail(TG_ERR_INVALID, "l and u must be given together");
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.x0 = x0; a.u_prev = u_prev; a.path_ref = path_ref; a.vref = vref;
    a.H = H; a.q = q; a.c0 = c0; a.l = l; a.u = u; a.Gs = Gs; a.stop = 2;
    return launch_step(h, a);
}

static int ensure_ws(tg_handle *h, int B)
{
    if (!h->cfg.warm_start || B <= h->ws_B) return TG_OK;
    cudaFree(h->ws_x); cudaFree(h->ws_y); cudaFree(h->ws_valid);
    h->ws_x = h->ws_y = nullptr; h->ws_valid = nullptr; h->ws_B = 0;
    CK(cudaMalloc(&h->ws_x, (size_t)B * h->dc.n * sizeof(double)));
    CK(cudaMalloc(&h->ws_y, (size_t)B * h->dc.m * sizeof(double)));
    CK(cudaMalloc(&h->ws_valid, (size_t)B * sizeof(int)));
    CK(cudaMemsetAsync(h->ws_valid, 0, (size_t)B * sizeof(int), h->stream));
    h->ws_B = B;
    return TG_OK;
}

int tg_mpc_step(tg_handle *h, int B, const double *x0, const double *u_prev, const double *path_ref, const double *vref,
                double *u_cmd, int32_t *status, int32_t *iters, double *objective, double *U_opt, double *X_opt, double *y_opt)
{
    if (!h || B < 0 || (B > 0 && (!x0 || !u_prev || !path_ref || !u_cmd))) return fail(TG_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    int rc = ensure_ws(h, B);
    if (rc != TG_OK) return rc;
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.x0 = x0; a.u_prev = u_prev; a.path_ref = path_ref; a.vref = vref;
    a.u_cmd = u_cmd; a.status = status; a.iters = iters; a.objective = objective; a.U_opt = U_opt; a.X_opt = X_opt; a.y_opt = y_opt;
    a.ws_x = h->ws_x; a.ws_y = h->ws_y; a.ws_valid = h->ws_valid;
    return launch_step(h, a);
}

int tg_ref_window(tg_handle *h, int B, const double *x0, const tg_ref_spec *spec, const double *brk, const double *coef,
                  int t_index, double *path_ref, double *vref)
{
    if (!h || B < 0 || (B > 0 && (!x0 || !spec || !path_ref || !vref))) return fail(TG_ERR_INVALID, "bad argument");
    if (B == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    const int grid = B < 8 * h->num_sms ? B : 8 * h->num_sms;
    tg_ref_window_kernel<<<grid, 32, (size_t)h->Lref.total * sizeof(double), h->stream>>>(h->dc, h->Lref, B, x0, spec, brk, coef, t_index, path_ref, vref);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

int tg_closed_loop(tg_handle *h, int B, int T, const double *x0, const double *u0, const tg_ref_spec *spec,
                   const double *brk, const double *coef, int64_t traj_id0, double *clean, double *noisy, double *U,
                   int32_t *status_counts, int64_t *iters_total)
{
    if (!h || B < 0 || T < 0 || (B > 0 && (!x0 || !u0 || !spec || !clean || !noisy || (T > 0 && !U)))) return fail(TG_ERR_INVALID, "bad argument");
    if (B == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    LoopArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.T = T; a.x0 = x0; a.u0 = u0; a.spec = spec; a.brk = brk; a.coef = coef; a.traj_id0 = traj_id0;
    a.clean = clean; a.noisy = noisy; a.U = U; a.status_counts = status_counts; a.iters_total = (long long *)iters_total;
    return dispatch_tw(h->W, h->S, h->dc, [&](auto W_, auto S_, auto NC_, auto HS_) -> int {
        constexpr int W = decltype(W_)::value, S = decltype(S_)::value, NC = decltype(NC_)::value;
        constexpr bool HS = decltype(HS_)::value;
        if constexpr (TW_P1_ONLY(W, NC, HS)) {
            return launch_tw(h, tw_closed_loop_kernel<W, S, NC, true, HS>, a, B, W, 1);
        } else {
            if constexpr (TW_HAS_P1(W, NC, HS))
                if (choose_tw_ppc(h, B) == 1) return launch_tw(h, tw_closed_loop_kernel<W, S, NC, true, HS>, a, B, W, 1);
            return launch_tw(h, tw_closed_loop_kernel<W, S, NC, false, HS>, a, B, W);
        }
    });
}

int tg_plant_rollout(tg_handle *h, int B, int T, const double *x0, const double *U, double *X)
{
    if (!h || B < 0 || T < 0 || (B > 0 && (!x0 || !X || (T > 0 && !U)))) return fail(TG_ERR_INVALID, "bad argument");
    if (B == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    tg_plant_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->dc, B, T, x0, U, X);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

int tg_sensor_noise(tg_handle *h, int64_t traj_id0, int n_traj, int n_rows, double *out)
{
    if (!h || n_traj < 0 || n_rows < 0 || ((long long)n_traj * n_rows > 0 && !out)) return fail(TG_ERR_INVALID, "bad argument");
    const long long tot = (long long)n_traj * n_rows;
    if (tot == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    tg_noise_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(h->dc.seed_base + (unsigned long long)traj_id0, n_traj, n_rows, out);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

// ---- scenario generation (tg_scenarios.cuh)
void tg_default_scenario_rules(tg_scenario_rules *r)
{
    memset(r, 0, sizeof(*r));
    const double lo[6] = {-2.0, 0.0, 0.0, 0.4, -0.05, -1.0}, hi[6] = {2.0, 0.0, 0.0, 1.5, 0.05, 1.0};   // generation_type1.py:260-265
    memcpy(r->x0_lo, lo, sizeof(lo)); memcpy(r->x0_hi, hi, sizeof(hi));
    r->lat_off[0] = -0.2; r->lat_off[1] = 0.2; r->head_off[0] = -0.2; r->head_off[1] = 0.2;
    r->vref0 = 0.8; r->vcruise[0] = 0.8; r->vcruise[1] = 2.0; r->t_ramp = 2.0;                          // MPC/main.py:87
    r->sine_A[0] = 0.2; r->sine_A[1] = 1.0; r->sine_k[0] = 0.3; r->sine_k[1] = 1.0; r->sine_psi[0] = 0.0; r->sine_psi[1] = 6.283185307179586;
    r->parab_c[0] = -0.2; r->parab_c[1] = 0.2;
    r->spl_x0 = -6.0; r->spl_dx[0] = 1.0; r->spl_dx[1] = 3.0; r->spl_sigma = 0.3; r->spl_knots = 27;
    r->n_cycle = 2; r->cycle[0] = TG_PATH_SPLINE; r->cycle[1] = TG_PATH_SINE;
    r->seed_base = 2025ull;
}

static int scenario_check(tg_handle *h, int B, const tg_scenario_rules *r)
{
    if (!h || !r || B < 0) return fail(TG_ERR_INVALID, "bad argument");
    if (r->spl_knots < 3 || r->spl_knots > TG_SCN_MAX_KNOTS) return fail(TG_ERR_INVALID, "scenario rules: 3 <= spl_knots <= 32");
    if (r->n_cycle < 1 || r->n_cycle > 4) return fail(TG_ERR_INVALID, "scenario rules: 1 <= n_cycle <= 4");
    for (int i = 0; i < r->n_cycle; ++i)
        if (r->cycle[i] < TG_PATH_PARABOLA || r->cycle[i] > TG_PATH_SPLINE) return fail(TG_ERR_INVALID, "scenario rules: cycle entries must be parabola / sine / spline");
    if (!(r->spl_dx[0] > 0) || r->spl_dx[1] < r->spl_dx[0]) return fail(TG_ERR_INVALID, "scenario rules: knot spacing must be positive");
    if ((long long)B * (r->spl_knots - 1) > 2147483647LL) return fail(TG_ERR_INVALID, "scenario rules: spline table index overflow");
    return TG_OK;
}

int tg_make_scenarios(tg_handle *h, int B, int64_t traj_id0, const tg_scenario_rules *rules, double *x0, double *u0,
                      tg_ref_spec *spec, double *spl_breaks, double *spl_coef)
{
    int rc = scenario_check(h, B, rules);
    if (rc != TG_OK) return rc;
    if (B == 0) return TG_OK;
    if (!x0 || !u0 || !spec || !spl_breaks || !spl_coef) return fail(TG_ERR_INVALID, "null output");
    CK(cudaSetDevice(h->device));
    tg_scenario_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->dc, *rules, B, (long long)traj_id0, x0, u0, spec, spl_breaks, spl_coef);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

// ---- open-loop generators (tg_openloop.cuh)
void tg_default_type1_rules(tg_type1_rules *r)
{
    memset(r, 0, sizeof(*r));
    r->d_mean = 0.2161; r->d_std = 0.1314; r->delta_mean = 0.0035; r->delta_std = 0.0338;    // generation_type1.py:250
    r->du_lo[0] = -0.1; r->du_hi[0] = 0.1; r->du_lo[1] = -0.04; r->du_hi[1] = 0.04;           // :251
    r->u_lo[0] = -1.0; r->u_hi[0] = 1.0; r->u_lo[1] = -0.6; r->u_hi[1] = 0.6;                 // :288-289
    r->transient_s[0] = 1.5; r->transient_s[1] = 3.0; r->checkpoint_s[0] = 3.0; r->checkpoint_s[1] = 5.0;
    r->period_s[0] = 4.0; r->period_s[1] = 8.0; r->amp_frac[0] = 0.5; r->amp_frac[1] = 1.5;
    r->p_straight = 0.5; r->tr_d_frac = 0.2; r->tr_delta_frac = 0.3; r->st_d_frac = 0.05;
    r->sin_noise_frac = 0.1; r->straight_frac = 0.01; r->ctrl_noise_frac = 0.1; r->mode = -1;
}

void tg_default_type2_rules(tg_type2_rules *r)
{
    memset(r, 0, sizeof(*r));
    r->v_turn_max = 1.2; r->v_high = 4.0; r->d_range[0] = 0.0; r->d_range[1] = 0.33;          // generation_type2.py:24-27
    r->delta_turn_range[0] = 0.015; r->delta_turn_range[1] = 0.04; r->delta_straight_noise = 0.004;
    r->delta_rate_max = 0.30; r->v_floor = 0.35; r->d_boost_min = 0.15;                         // :97,:100
    r->seg_s[0] = 0.4; r->seg_s[1] = 1.5;                                                       // :114
    r->p_modes[0] = 0.35; r->p_modes[1] = 0.35; r->p_modes[2] = 0.15; r->p_modes[3] = 0.15;     // :112
    r->p_after_turn[0] = 0.5; r->p_after_turn[1] = 0.5;                                         // :109
    r->acc_d_lo = 0.3; r->cruise_d[0] = -0.05; r->cruise_d[1] = 0.2;                            // :120,:122
    r->turn_d_fast[0] = 0.0; r->turn_d_fast[1] = 0.15; r->turn_d_slow[0] = 0.05; r->turn_d_slow[1] = 0.25;   // :124
    r->stall_v = 0.5; r->stall_d[0] = 0.5; r->stall_d[1] = 1.0; r->stall_min_s = 0.3;           // :131-133
    r->delta_clip = 0.6;                                                                        // :141
}

static int openloop_check(tg_handle *h, int B, int T, const double *x0, const void *rules, const double *clean,
                          const double *noisy, const double *U)
{
    if (!h || B < 0 || T < 1 || !rules || (B > 0 && !x0)) return fail(TG_ERR_INVALID, "bad argument (T >= 1 required)");
    if (((uintptr_t)clean | (uintptr_t)noisy | (uintptr_t)U) & 15) return fail(TG_ERR_INVALID, "clean / noisy / U must be 16-byte aligned");
    return TG_OK;
}

extern "C++" {
template <int KIND, typename Rules>
static int openloop_launch(tg_handle *h, int B, int T, const double *x0, const Rules &rules, uint64_t ctrl_seed_base,
                           int64_t traj_id0, double *clean, double *noisy, double *U, int8_t *modes)
{
    if (B == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    OpenLoopArgs a;
    a.B = B; a.T = T; a.x0 = x0; a.traj_id0 = traj_id0; a.ctrl_seed_base = ctrl_seed_base;
    a.clean = clean; a.noisy = noisy; a.U = U; a.modes = (signed char *)modes;
    const int per_cta = 32 * TG_OL_WARPS;
    tg_openloop_kernel<KIND, Rules><<<(B + per_cta - 1) / per_cta, per_cta, TG_OL_WARPS * TG_OL_TILE * sizeof(double), h->stream>>>(h->dc, rules, a);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}
}  // extern "C++"

int tg_openloop_type1(tg_handle *h, int B, int T, const double *x0, const tg_type1_rules *rules, uint64_t ctrl_seed_base,
                      int64_t traj_id0, double *clean, double *noisy, double *U, int8_t *modes)
{
    int rc = openloop_check(h, B, T, x0, rules, clean, noisy, U);
    if (rc != TG_OK) return rc;
    const tg_type1_rules &r = *rules;
    const double Ts = h->dc.Ts;
    if (!(r.transient_s[0] > 0) || r.transient_s[1] < r.transient_s[0] || !(r.checkpoint_s[0] > 0) || r.checkpoint_s[1] < r.checkpoint_s[0] ||
        !(r.period_s[0] > 0) || r.period_s[1] < r.period_s[0] || r.mode < -1 || r.mode > 1)
        return fail(TG_ERR_INVALID, "type-1 rules: bad range");
    if ((int)(r.transient_s[0] / Ts) < 1) return fail(TG_ERR_INVALID, "type-1 rules: the transient must last at least one step");
    {   // knots of the transient spline in the worst case (longest transient, densest checkpoints)
        const double n_tr = std::min((double)T, std::floor(r.transient_s[1] / Ts));
        const double every = std::max(1.0, std::nearbyint(r.checkpoint_s[0] / Ts));
        if (std::ceil(n_tr / every) + 1 > TG_OL_MAX_KNOTS) return fail(TG_ERR_UNSUPPORTED, "type-1 rules: more than 16 spline knots in the transient");
    }
    return openloop_launch<1>(h, B, T, x0, r, ctrl_seed_base, traj_id0, clean, noisy, U, modes);
}

int tg_openloop_type2(tg_handle *h, int B, int T, const double *x0, const tg_type2_rules *rules, uint64_t ctrl_seed_base,
                      int64_t traj_id0, double *clean, double *noisy, double *U, int8_t *modes)
{
    int rc = openloop_check(h, B, T, x0, rules, clean, noisy, U);
    if (rc != TG_OK) return rc;
    const tg_type2_rules &r = *rules;
    double tot4 = 0, tot2 = 0;
    for (int i = 0; i < 4; ++i) { if (!(r.p_modes[i] >= 0)) return fail(TG_ERR_INVALID, "type-2 rules: negative probability"); tot4 += r.p_modes[i]; }
    for (int i = 0; i < 2; ++i) { if (!(r.p_after_turn[i] >= 0)) return fail(TG_ERR_INVALID, "type-2 rules: negative probability"); tot2 += r.p_after_turn[i]; }
    if (!(tot4 > 0) || !(tot2 > 0) || !(r.seg_s[0] > 0) || r.seg_s[1] < r.seg_s[0]) return fail(TG_ERR_INVALID, "type-2 rules: bad range");
    return openloop_launch<2>(h, B, T, x0, r, ctrl_seed_base, traj_id0, clean, noisy, U, modes);
}

// ---- estimator physics (tg_estimator.cuh)
extern "C++" {
template <typename T>
static EstCfg<T> make_est_cfg(const tg_handle *h, const tg_state_limits *lim)
{
    EstCfg<T> e;
    const double *p = h->cfg.params;
    e.Ts = (T)h->cfg.Ts;
    for (int i = 0; i < 6; ++i) { e.lo[i] = (T)lim->lo[i]; e.hi[i] = (T)lim->hi[i]; }
    e.Cm1 = (T)p[P_Cm1]; e.Cm2 = (T)p[P_Cm2]; e.Cr0 = (T)p[P_Cr0]; e.Cr2 = (T)p[P_Cr2];
    e.Br = (T)p[P_Br]; e.Cr = (T)p[P_Cr]; e.Dr = (T)p[P_Dr]; e.Bf = (T)p[P_Bf]; e.Cf = (T)p[P_Cf]; e.Df = (T)p[P_Df];
    e.m = (T)p[P_m]; e.Iz = (T)p[P_Iz]; e.lf = (T)p[P_lf]; e.lr = (T)p[P_lr]; e.maxAlpha = (T)p[P_maxAlpha];
    // KalmanNet/vehicle_model.py:26 builds the threshold as a float32 tensor whatever the state dtype
    e.vx_zero = (T)(float)p[P_vx_zero];
    return e;
}
}  // extern "C++"

static bool est_aligned(int dtype, std::initializer_list<const void *> ptrs)
{
    const uintptr_t mask = dtype == 0 ? 15 : 7;      // the kernels move pairs of values
    for (const void *p : ptrs)
        if ((uintptr_t)p & mask) return false;
    return true;
}

static int est_check(tg_handle *h, int B, int dtype, const tg_state_limits *lim)
{
    if (!h || B < 0 || !lim || (dtype != 0 && dtype != 1)) return fail(TG_ERR_INVALID, "bad argument (dtype: 0 = fp64, 1 = fp32)");
    for (int i = 0; i < 6; ++i)
        if (!(lim->lo[i] <= lim->hi[i])) return fail(TG_ERR_INVALID, "state limits: lo > hi");
    return TG_OK;
}

int tg_estimator_step(tg_handle *h, int B, int dtype, const void *x, const void *u, const tg_state_limits *lim, void *x_next)
{
    int rc = est_check(h, B, dtype, lim);
    if (rc != TG_OK) return rc;
    if (B == 0) return TG_OK;
    if (!x || !u || !x_next) return fail(TG_ERR_INVALID, "null argument");
    if (!est_aligned(dtype, {x, u, x_next})) return fail(TG_ERR_INVALID, "buffers must be aligned to two elements (8 B fp32 / 16 B fp64)");
    CK(cudaSetDevice(h->device));
    const int blocks = (B + 127) / 128;
    if (dtype == 0) tg_estimator_step_kernel<double><<<blocks, 128, 0, h->stream>>>(make_est_cfg<double>(h, lim), B, (const double *)x, (const double *)u, (double *)x_next);
    else tg_estimator_step_kernel<float><<<blocks, 128, 0, h->stream>>>(make_est_cfg<float>(h, lim), B, (const float *)x, (const float *)u, (float *)x_next);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

int tg_estimator_step_vjp(tg_handle *h, int B, int dtype, const void *x, const void *u, const tg_state_limits *lim,
                          const void *grad_next, void *grad_x, void *grad_u)
{
    int rc = est_check(h, B, dtype, lim);
    if (rc != TG_OK) return rc;
    if (B == 0) return TG_OK;
    if (!x || !u || !grad_next) return fail(TG_ERR_INVALID, "null argument");
    if (!est_aligned(dtype, {x, u, grad_next, grad_x, grad_u})) return fail(TG_ERR_INVALID, "buffers must be aligned to two elements (8 B fp32 / 16 B fp64)");
    CK(cudaSetDevice(h->device));
    const int blocks = (B + 127) / 128;
    if (dtype == 0) tg_estimator_vjp_kernel<double><<<blocks, 128, 0, h->stream>>>(make_est_cfg<double>(h, lim), B, (const double *)x, (const double *)u, (const double *)grad_next, (double *)grad_x, (double *)grad_u);
    else tg_estimator_vjp_kernel<float><<<blocks, 128, 0, h->stream>>>(make_est_cfg<float>(h, lim), B, (const float *)x, (const float *)u, (const float *)grad_next, (float *)grad_x, (float *)grad_u);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

int tg_estimator_rollout(tg_handle *h, int B, int dtype, int T_u, int t_start, int H, const void *x0, const void *U,
                         const tg_state_limits *lim, void *preds, int32_t *H_out)
{
    int rc = est_check(h, B, dtype, lim);
    if (rc != TG_OK) return rc;
    if (T_u < 0 || t_start < 0 || H < 0) return fail(TG_ERR_INVALID, "bad argument");
    const int Hn = std::max(0, std::min(H, T_u - t_start));       // test_prediction.py:79: stop at the end of u
    if (H_out) *H_out = Hn;
    if (B == 0 || Hn == 0) return TG_OK;
    if (!x0 || !U || !preds) return fail(TG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    const int blocks = (B + 63) / 64;
    if (dtype == 0) tg_estimator_rollout_kernel<double><<<blocks, 64, 0, h->stream>>>(make_est_cfg<double>(h, lim), B, T_u, t_start, Hn, (const double *)x0, (const double *)U, (double *)preds);
    else tg_estimator_rollout_kernel<float><<<blocks, 64, 0, h->stream>>>(make_est_cfg<float>(h, lim), B, T_u, t_start, Hn, (const float *)x0, (const float *)U, (float *)preds);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

int tg_philox_u32(tg_handle *h, uint64_t seed, uint32_t first, uint32_t block, int n, uint32_t *out)
{
    if (!h || n < 0 || (n > 0 && !out)) return fail(TG_ERR_INVALID, "bad argument");
    if (n == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    tg_philox_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(seed, first, block, n, out);
    h->launches += 1;
    CK(cudaGetLastError());
    return TG_OK;
}

int tg_fma_peak(tg_handle *h, int dtype, double *tflops)
{
    if (!h || !tflops) return fail(TG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    const int blocks = h->num_sms * 8, threads = 256, iters = 1 << 15;
    void *buf = nullptr;
    CK(cudaMalloc(&buf, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, h->stream));
        if (dtype == 0) tg_fma_peak_kernel<double><<<blocks, threads, 0, h->stream>>>((double *)buf, iters);
        else tg_fma_peak_kernel<float><<<blocks, threads, 0, h->stream>>>((float *)buf, iters);
        CK(cudaEventRecord(e1, h->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
        h->launches += 1;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return TG_OK;
}

int tg_malloc(void **p, int64_t bytes) { if (!p || bytes < 0) return fail(TG_ERR_INVALID, "bad argument"); *p = nullptr; if (bytes == 0) return TG_OK; CK(cudaMalloc(p, (size_t)bytes)); return TG_OK; }
int tg_malloc_on(tg_handle *h, void **p, int64_t bytes)
{
    if (!h) return fail(TG_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    return tg_malloc(p, bytes);
}
int tg_free(void *p) { if (p) CK(cudaFree(p)); return TG_OK; }
int tg_malloc_host(void **p, int64_t bytes) { if (!p || bytes < 0) return fail(TG_ERR_INVALID, "bad argument"); *p = nullptr; if (bytes == 0) return TG_OK; CK(cudaMallocHost(p, (size_t)bytes)); return TG_OK; }
int tg_free_host(void *p) { if (p) CK(cudaFreeHost(p)); return TG_OK; }
int tg_memcpy_h2d(tg_handle *h, void *dst, const void *src, int64_t bytes)
{
    if (!h || bytes < 0) return fail(TG_ERR_INVALID, "bad argument");
    if (bytes == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return TG_OK;
}
int tg_memcpy_d2h(tg_handle *h, void *dst, const void *src, int64_t bytes)
{
    if (!h || bytes < 0) return fail(TG_ERR_INVALID, "bad argument");
    if (bytes == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return TG_OK;
}

// ---- host-buffer entry points: stage through one device arena, copies on the handle's stream
struct Arena {
    char *base; size_t off, cap;
    void *take(size_t bytes) { size_t o = (off + 255) & ~(size_t)255; off = o + bytes; return (off <= cap) ? base + o : nullptr; }
};
static int ensure_dstage(tg_handle *h, size_t bytes)
{
    if (bytes <= h->dstage_bytes) return TG_OK;
    if (h->dstage) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->dstage)); h->dstage = nullptr; h->dstage_bytes = 0; }
    CK(cudaMalloc(&h->dstage, bytes));
    h->dstage_bytes = bytes;
    return TG_OK;
}
#define PAD(x) (((size_t)(x) + 255) & ~(size_t)255)
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream))
#define D2H(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream))

int tg_mpc_step_host(tg_handle *h, int B, const double *x0, const double *u_prev, const double *path_ref, const double *vref,
                     double *u_cmd, int32_t *status, int32_t *iters, double *objective, double *U_opt, double *X_opt, double *y_opt)
{
    if (!h || B < 0 || (B > 0 && (!x0 || !u_prev || !path_ref || !u_cmd))) return fail(TG_ERR_INVALID, "bad argument");
    if (B == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    const int N = h->dc.N, n = h->dc.n, m = h->dc.m;
    const size_t b8 = sizeof(double);
    const size_t s_x0 = (size_t)B * 6 * b8, s_up = (size_t)B * 2 * b8, s_pr = (size_t)B * (N + 1) * 3 * b8, s_vr = (size_t)B * (N + 1) * b8;
    const size_t s_uc = (size_t)B * 2 * b8, s_i = (size_t)B * 4, s_ob = (size_t)B * b8, s_U = (size_t)B * n * b8, s_X = (size_t)B * (N + 1) * 6 * b8, s_y = (size_t)B * m * b8;
    const size_t total = PAD(s_x0) + PAD(s_up) + PAD(s_pr) + PAD(s_vr) + PAD(s_uc) + 2 * PAD(s_i) + PAD(s_ob) + PAD(s_U) + PAD(s_X) + PAD(s_y) + 4096;
    int rc = ensure_dstage(h, total);
    if (rc != TG_OK) return rc;
    Arena ar{(char *)h->dstage, 0, h->dstage_bytes};
    double *d_x0 = (double *)ar.take(s_x0), *d_up = (double *)ar.take(s_up), *d_pr = (double *)ar.take(s_pr);
    double *d_vr = vref ? (double *)ar.take(s_vr) : nullptr;
    double *d_uc = (double *)ar.take(s_uc);
    int32_t *d_st = (int32_t *)ar.take(s_i), *d_it = (int32_t *)ar.take(s_i);
    double *d_ob = (double *)ar.take(s_ob);
    double *d_U = U_opt ? (double *)ar.take(s_U) : nullptr, *d_X = X_opt ? (double *)ar.take(s_X) : nullptr, *d_y = y_opt ? (double *)ar.take(s_y) : nullptr;
    H2D(d_x0, x0, s_x0); H2D(d_up, u_prev, s_up); H2D(d_pr, path_ref, s_pr);
    if (vref) H2D(d_vr, vref, s_vr);
    rc = tg_mpc_step(h, B, d_x0, d_up, d_pr, d_vr, d_uc, d_st, d_it, d_ob, d_U, d_X, d_y);
    if (rc != TG_OK) return rc;
    D2H(u_cmd, d_uc, s_uc);
    if (status) D2H(status, d_st, s_i);
    if (iters) D2H(iters, d_it, s_i);
    if (objective) D2H(objective, d_ob, s_ob);
    if (U_opt) D2H(U_opt, d_U, s_U);
    if (X_opt) D2H(X_opt, d_X, s_X);
    if (y_opt) D2H(y_opt, d_y, s_y);
    CK(cudaStreamSynchronize(h->stream));
    return TG_OK;
}

// device alias of a pinned (cudaHostAlloc'ed / registered) host buffer, or null for pageable memory
static void *pinned_device_alias(const void *p)
{
    cudaPointerAttributes at;
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

int tg_closed_loop_host(tg_handle *h, int B, int T, const double *x0, const double *u0, const tg_ref_spec *spec,
                        const double *brk, int64_t n_breaks, const double *coef, int64_t n_coef, int64_t traj_id0,
                        double *clean, double *noisy, double *U, int32_t *status_counts, int64_t *iters_total)
{
    if (!h || B < 0 || T < 0 || (B > 0 && (!x0 || !u0 || !spec || !clean || !noisy || (T > 0 && !U)))) return fail(TG_ERR_INVALID, "bad argument");
    if (B == 0) return TG_OK;
    {   // the scenario table is host memory here: reject what would send the device out of bounds
        const int64_t n_tab = (brk && coef) ? (n_breaks < n_coef ? n_breaks : n_coef) : 0;
        for (int b = 0; b < B; ++b) {
            const tg_ref_spec &sp = spec[b];
            if (sp.path_kind < TG_PATH_PARABOLA || sp.path_kind > TG_PATH_ARC || sp.vref_kind < TG_VREF_HOLD || sp.vref_kind > TG_VREF_SINE)
                return fail(TG_ERR_INVALID, "scenario " + std::to_string(b) + ": unknown path_kind / vref_kind");
            if (sp.path_kind == TG_PATH_SPLINE || sp.path_kind == TG_PATH_ARC) {
                const int64_t need_c = (int64_t)sp.spline_first + (sp.path_kind == TG_PATH_ARC ? 2 : 1) * (int64_t)sp.spline_count;
                if (sp.spline_count < 1 || sp.spline_first < 0 || (int64_t)sp.spline_first + sp.spline_count > n_tab || need_c > n_coef || !coef)
                    return fail(TG_ERR_INVALID, "scenario " + std::to_string(b) + ": spline pieces outside the break / coefficient tables");
            }
        }
    }
    CK(cudaSetDevice(h->device));
    const size_t b8 = sizeof(double);
    const size_t s_x0 = (size_t)B * 6 * b8, s_u0 = (size_t)B * 2 * b8, s_sp = (size_t)B * sizeof(tg_ref_spec);
    const size_t s_bk = (size_t)(n_breaks > 0 ? n_breaks : 0) * b8, s_cf = (size_t)(n_coef > 0 ? n_coef : 0) * 4 * b8;
    const size_t s_cl = (size_t)B * (T + 1) * 6 * b8, s_U = (size_t)B * T * 2 * b8, s_sc = (size_t)B * TG_NUM_STATUS * 4, s_it = (size_t)B * 8;
    // Rows leave the SM at ~1 GB/s (112 B per MPC step), far below what PCIe carries, so when the caller's result buffers
    // are pinned the kernel stores straight into them (posted writes overlap the computation) and the device-to-host copy
    // of the rows after the kernel disappears.  Pageable buffers are staged through the device arena.
    // TRAJGEN_HOST_OUTPUT=staged forces the staged path (measurement knob).
    const char *mode = getenv("TRAJGEN_HOST_OUTPUT");
    const bool allow_direct = !(mode && strcmp(mode, "staged") == 0);
    double *a_cl = allow_direct ? (double *)pinned_device_alias(clean) : nullptr;
    double *a_no = allow_direct ? (double *)pinned_device_alias(noisy) : nullptr;
    double *a_U = allow_direct ? (double *)pinned_device_alias(U) : nullptr;
    const bool direct = a_cl && a_no && (a_U || !s_U);
    const size_t total = PAD(s_x0) + PAD(s_u0) + PAD(s_sp) + PAD(s_bk) + PAD(s_cf) + (direct ? 0 : 2 * PAD(s_cl) + PAD(s_U)) + PAD(s_sc) + PAD(s_it) + 4096;
    int rc = ensure_dstage(h, total);
    if (rc != TG_OK) return rc;
    Arena ar{(char *)h->dstage, 0, h->dstage_bytes};
    double *d_x0 = (double *)ar.take(s_x0), *d_u0 = (double *)ar.take(s_u0);
    tg_ref_spec *d_sp = (tg_ref_spec *)ar.take(s_sp);
    double *d_bk = s_bk ? (double *)ar.take(s_bk) : nullptr, *d_cf = s_cf ? (double *)ar.take(s_cf) : nullptr;
    double *d_cl = direct ? a_cl : (double *)ar.take(s_cl), *d_no = direct ? a_no : (double *)ar.take(s_cl);
    double *d_U = direct ? a_U : (double *)ar.take(s_U ? s_U : 8);
    int32_t *d_sc = (int32_t *)ar.take(s_sc);
    long long *d_it = (long long *)ar.take(s_it);
    H2D(d_x0, x0, s_x0); H2D(d_u0, u0, s_u0); H2D(d_sp, spec, s_sp);
    if (s_bk) H2D(d_bk, brk, s_bk);
    if (s_cf) H2D(d_cf, coef, s_cf);
    rc = tg_closed_loop(h, B, T, d_x0, d_u0, d_sp, d_bk, d_cf, traj_id0, d_cl, d_no, d_U, d_sc, (int64_t *)d_it);
    if (rc != TG_OK) return rc;
    if (!direct) {
        D2H(clean, d_cl, s_cl); D2H(noisy, d_no, s_cl);
        if (s_U) D2H(U, d_U, s_U);
    }
    if (status_counts) D2H(status_counts, d_sc, s_sc);
    if (iters_total) D2H(iters_total, d_it, s_it);
    CK(cudaStreamSynchronize(h->stream));
    return TG_OK;
}

int tg_make_scenarios_host(tg_handle *h, int B, int64_t traj_id0, const tg_scenario_rules *rules, double *x0, double *u0,
                           tg_ref_spec *spec, double *spl_breaks, double *spl_coef)
{
    int rc = scenario_check(h, B, rules);
    if (rc != TG_OK) return rc;
    if (B == 0) return TG_OK;
    if (!x0 || !u0 || !spec || !spl_breaks || !spl_coef) return fail(TG_ERR_INVALID, "null output");
    CK(cudaSetDevice(h->device));
    const size_t P = (size_t)rules->spl_knots - 1;
    const size_t s_x0 = (size_t)B * 48, s_u0 = (size_t)B * 16, s_sp = (size_t)B * sizeof(tg_ref_spec), s_bk = (size_t)B * P * 8, s_cf = s_bk * 4;
    rc = ensure_dstage(h, PAD(s_x0) + PAD(s_u0) + PAD(s_sp) + PAD(s_bk) + PAD(s_cf) + 4096);
    if (rc != TG_OK) return rc;
    Arena ar{(char *)h->dstage, 0, h->dstage_bytes};
    double *d_x0 = (double *)ar.take(s_x0), *d_u0 = (double *)ar.take(s_u0);
    tg_ref_spec *d_sp = (tg_ref_spec *)ar.take(s_sp);
    double *d_bk = (double *)ar.take(s_bk), *d_cf = (double *)ar.take(s_cf);
    rc = tg_make_scenarios(h, B, traj_id0, rules, d_x0, d_u0, d_sp, d_bk, d_cf);
    if (rc != TG_OK) return rc;
    D2H(x0, d_x0, s_x0); D2H(u0, d_u0, s_u0); D2H(spec, d_sp, s_sp); D2H(spl_breaks, d_bk, s_bk); D2H(spl_coef, d_cf, s_cf);
    CK(cudaStreamSynchronize(h->stream));
    return TG_OK;
}

extern "C++" {
template <typename Rules, typename Fn>
static int openloop_host(tg_handle *h, int B, int T, const double *x0, const Rules *rules, double *clean, double *noisy, double *U,
                         int8_t *modes, size_t n_modes, Fn &&run)
{
    if (!h || B < 0 || T < 1 || !rules || (B > 0 && !x0)) return fail(TG_ERR_INVALID, "bad argument (T >= 1 required)");
    if (B == 0) return TG_OK;
    CK(cudaSetDevice(h->device));
    const size_t b8 = sizeof(double);
    const size_t s_x0 = (size_t)B * 6 * b8, s_cl = (size_t)B * (T + 1) * 6 * b8, s_U = (size_t)B * T * 2 * b8;
    const size_t total = PAD(s_x0) + 2 * PAD(s_cl) + PAD(s_U) + PAD(n_modes) + 4096;
    int rc = ensure_dstage(h, total);
    if (rc != TG_OK) return rc;
    Arena ar{(char *)h->dstage, 0, h->dstage_bytes};
    double *d_x0 = (double *)ar.take(s_x0);
    double *d_cl = clean ? (double *)ar.take(s_cl) : nullptr, *d_no = noisy ? (double *)ar.take(s_cl) : nullptr;
    double *d_U = U ? (double *)ar.take(s_U) : nullptr;
    int8_t *d_m = modes ? (int8_t *)ar.take(n_modes) : nullptr;
    H2D(d_x0, x0, s_x0);
    rc = run(d_x0, d_cl, d_no, d_U, d_m);
    if (rc != TG_OK) return rc;
    if (clean) D2H(clean, d_cl, s_cl);
    if (noisy) D2H(noisy, d_no, s_cl);
    if (U) D2H(U, d_U, s_U);
    if (modes) D2H(modes, d_m, n_modes);
    CK(cudaStreamSynchronize(h->stream));
    return TG_OK;
}
}  // extern "C++"

int tg_openloop_type1_host(tg_handle *h, int B, int T, const double *x0, const tg_type1_rules *rules, uint64_t ctrl_seed_base,
                           int64_t traj_id0, double *clean, double *noisy, double *U, int8_t *modes)
{
    return openloop_host(h, B, T, x0, rules, clean, noisy, U, modes, (size_t)(B > 0 ? B : 0),
                         [&](double *dx, double *dc, double *dn, double *dU, int8_t *dm) {
                             return tg_openloop_type1(h, B, T, dx, rules, ctrl_seed_base, traj_id0, dc, dn, dU, dm);
                         });
}

int tg_openloop_type2_host(tg_handle *h, int B, int T, const double *x0, const tg_type2_rules *rules, uint64_t ctrl_seed_base,
                           int64_t traj_id0, double *clean, double *noisy, double *U, int8_t *modes)
{
    return openloop_host(h, B, T, x0, rules, clean, noisy, U, modes, (size_t)(B > 0 ? B : 0) * (size_t)(T > 0 ? T : 0),
                         [&](double *dx, double *dc, double *dn, double *dU, int8_t *dm) {
                             return tg_openloop_type2(h, B, T, dx, rules, ctrl_seed_base, traj_id0, dc, dn, dU, dm);
                         });
}

}  // extern "C"
