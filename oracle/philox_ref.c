/* Oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py): CPU restatement of the
 * sensor-noise generator of the B200 path.
 *
 * The reference draws its sensor noise with NumPy's PCG64 + ziggurat
 * (generation_type1.py:295-306, generation_type2.py:190-200: default_rng(12345 + i), six
 * column-wise .normal(0, sigma_c, T+1) draws).  A sequential generator cannot be evaluated
 * per (trajectory,row) on a GPU, so the B200 path keeps the reference's *contract*
 * (seed = base + trajectory_id, sigma per column, noise on all T+1 rows, never fed back)
 * and replaces the bit stream by the counter-based Philox4x32-10 of Salmon et al.,
 * "Parallel random numbers: as easy as 1, 2, 3" (SC'11), pinned by the Random123
 * known-answer vectors (tests/test_philox.py).
 *
 *   key     = (lo32(seed), hi32(seed)),   seed = seed_base + trajectory_id
 *   counter = (row, block, 0, 0),         block 0 -> columns X,Y,phi,vx ; block 1 -> vy,omega
 *   uniform = (r + 0.5) * 2^-32           (exact in fp64)
 *   normal  = Box-Muller on pairs (r0,r1) -> (n0,n1), (r2,r3) -> (n2,n3)
 *
 * log / sin / cos are evaluated with fixed polynomial code using only IEEE
 * add/mul/div/sqrt/fma in a fixed order, so the CUDA kernel (which uses the same
 * operations through __fma_rn/__dmul_rn/__dadd_rn) produces bit-identical doubles.
 * Build with -ffp-contract=off (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void tgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ln of the odd integer m = 2r+1 in [1, 2^33) scaled by 2^-33, i.e. ln((r+0.5) 2^-32). */
static double tgo_log_u(uint32_t r)
{
    uint64_t m = 2ull * r + 1ull;                 /* 33-bit odd integer, exact */
    int e = 63 - __builtin_clzll(m);              /* floor(log2 m) */
    double f = (double)m * ldexp(1.0, -e);        /* exact, in [1,2) */
    if (f > 1.4142135623730951) { f = f * 0.5; e += 1; }
    double s = (f - 1.0) / (f + 1.0);
    double s2 = s * s;
    /* 2 atanh(s) = 2 s (1 + s2/3 + s2^2/5 + ...), |s| <= 0.1716 */
    double p = 1.0 / 27.0;
    p = fma(p, s2, 1.0 / 25.0);
    p = fma(p, s2, 1.0 / 23.0);
    p = fma(p, s2, 1.0 / 21.0);
    p = fma(p, s2, 1.0 / 19.0);
    p = fma(p, s2, 1.0 / 17.0);
    p = fma(p, s2, 1.0 / 15.0);
    p = fma(p, s2, 1.0 / 13.0);
    p = fma(p, s2, 1.0 / 11.0);
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, 1.0 / 3.0);
    p = fma(p, s2, 1.0);
    double lf = 2.0 * s * p;
    return fma((double)(e - 33), 0.6931471805599453, lf);
}

/* sin and cos of 2 pi u, u = (r + 0.5) 2^-32: exact octant reduction, Taylor kernels. */
static void tgo_sincos_2pi_u(uint32_t r, double *sn, double *cs)
{
    uint64_t m = 2ull * r + 1ull;                 /* u = m 2^-33 ; 8u = m 2^-30 */
    uint32_t oct = (uint32_t)(m >> 30);           /* 0..7 */
    uint64_t frac = m & ((1ull << 30) - 1);       /* (8u - oct) 2^30 */
    double t = (double)frac * ldexp(1.0, -30);    /* in [0,1), angle = (oct + t) pi/4 */
    if (oct & 1u) t = 1.0 - t;                    /* reflect so the kernel angle is in [0,pi/4] */
    double a = t * 0.7853981633974483;
    double a2 = a * a;
    double ps = -1.0 / 1307674368000.0;           /* -1/15! */
    ps = fma(ps, a2, 1.0 / 6227020800.0);
    ps = fma(ps, a2, -1.0 / 39916800.0);
    ps = fma(ps, a2, 1.0 / 362880.0);
    ps = fma(ps, a2, -1.0 / 5040.0);
    ps = fma(ps, a2, 1.0 / 120.0);
    ps = fma(ps, a2, -1.0 / 6.0);
    ps = fma(ps, a2, 1.0);
    double sk = a * ps;
    double pc = 1.0 / 20922789888000.0;           /* 1/16! */
    pc = fma(pc, a2, -1.0 / 87178291200.0);
    pc = fma(pc, a2, 1.0 / 479001600.0);
    pc = fma(pc, a2, -1.0 / 3628800.0);
    pc = fma(pc, a2, 1.0 / 40320.0);
    pc = fma(pc, a2, -1.0 / 720.0);
    pc = fma(pc, a2, 1.0 / 24.0);
    pc = fma(pc, a2, -0.5);
    double ck = fma(pc, a2, 1.0);
    double s_, c_;
    switch (oct) {
        case 0: s_ = sk;  c_ = ck;  break;
        case 1: s_ = ck;  c_ = sk;  break;
        case 2: s_ = ck;  c_ = -sk; break;
        case 3: s_ = sk;  c_ = -ck; break;
        case 4: s_ = -sk; c_ = -ck; break;
        case 5: s_ = -ck; c_ = -sk; break;
        case 6: s_ = -ck; c_ = sk;  break;
        default: s_ = -sk; c_ = ck; break;
    }
    *sn = s_; *cs = c_;
}

void tgo_box_muller(uint32_t r0, uint32_t r1, double *n0, double *n1)
{
    double rad = sqrt(-2.0 * tgo_log_u(r0));
    double sn, cs;
    tgo_sincos_2pi_u(r1, &sn, &cs);
    *n0 = rad * cs;
    *n1 = rad * sn;
}

/* six standard normals of (seed, row): columns X, Y, phi, vx, vy, omega */
void tgo_noise_row(uint64_t seed, uint32_t row, double out[6])
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c0[4] = {row, 0u, 0u, 0u}, c1[4] = {row, 1u, 0u, 0u}, a[4], b[4];
    tgo_philox4x32_10(c0, key, a);
    tgo_philox4x32_10(c1, key, b);
    tgo_box_muller(a[0], a[1], &out[0], &out[1]);
    tgo_box_muller(a[2], a[3], &out[2], &out[3]);
    tgo_box_muller(b[0], b[1], &out[4], &out[5]);
}

/* standard normals for rows [0,n_rows) of one trajectory seed: out[n_rows][6] */
void tgo_noise_block(uint64_t seed, uint32_t n_rows, double *out)
{
    for (uint32_t t = 0; t < n_rows; ++t) tgo_noise_row(seed, t, out + 6 * (uint64_t)t);
}

/* raw stream tap: out[n][4] = philox(ctr = (first+i, block, 0, 0), key(seed)) */
void tgo_philox_stream(uint64_t seed, uint32_t first, uint32_t block, uint32_t n, uint32_t *out)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t c[4] = {first + i, block, 0u, 0u};
        tgo_philox4x32_10(c, key, out + 4 * (uint64_t)i);
    }
}

/* Box-Muller over arrays (open-loop control draws, oracle/openloop.py): pair i = (r0[i], r1[i]) */
void tgo_box_muller_vec(uint32_t n, const uint32_t *r0, const uint32_t *r1, double *n0, double *n1)
{
    for (uint32_t i = 0; i < n; ++i) tgo_box_muller(r0[i], r1[i], &n0[i], &n1[i]);
}
