// tg_csv.cpp -- host-side dataset writer: the clean / noisy CSV files of the generators, byte-identical to what
// pandas' DataFrame.to_csv(index=False) writes for the reference schema (generation_type1.py:315-339,
// generation_type2.py:202-218,309-322): columns t,X,Y,[phi],vx,vy,omega,d,delta,trajectory_id; T+1 rows per
// trajectory; t = k*Ts; the last row's d, delta are NaN -> empty fields; floats in Python's shortest repr.
// At ~1e7 MPC steps/s the pandas writer is >99 % of the wall time of a generation run; this one formats with
// std::to_chars (shortest round-trip digits) on all host cores.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trajgen.h"

namespace {

// Python float_repr_style 'short': shortest digits; fixed notation iff -4 < decpt <= 16, else d[.ddd]e+XX
inline char *put_pyfloat(char *p, double v)
{
    if (std::isnan(v)) return p;                       // pandas na_rep = ''
    if (std::isinf(v)) { if (v < 0) *p++ = '-'; memcpy(p, "inf", 3); return p + 3; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    // buf = [-]d[.ddd]e[+-]XX
    char *s = buf;
    if (*s == '-') { *p++ = '-'; ++s; }
    char digits[24];
    int nd = 0;
    digits[nd++] = *s++;
    if (*s == '.') { ++s; while (*s != 'e') digits[nd++] = *s++; }
    ++s;                                               // 'e'
    int ex = 0, sgn = 1;
    if (*s == '-') { sgn = -1; ++s; } else if (*s == '+') ++s;
    while (s < r.ptr) ex = ex * 10 + (*s++ - '0');
    ex *= sgn;
    const int decpt = ex + 1;                          // value = 0.d1d2.. * 10^decpt
    if (decpt > 16 || decpt < -3) {                    // exponential
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, digits + 1, nd - 1); p += nd - 1; }
        *p++ = 'e';
        *p++ = (ex < 0) ? '-' : '+';
        int ae = ex < 0 ? -ex : ex;
        if (ae >= 100) { *p++ = char('0' + ae / 100); ae %= 100; }
        *p++ = char('0' + ae / 10); *p++ = char('0' + ae % 10);
        return p;
    }
    if (decpt <= 0) {                                  // 0.000ddd
        *p++ = '0'; *p++ = '.';
        for (int i = 0; i < -decpt; ++i) *p++ = '0';
        memcpy(p, digits, nd); return p + nd;
    }
    if (decpt >= nd) {                                 // ddd000.0
        memcpy(p, digits, nd); p += nd;
        for (int i = nd; i < decpt; ++i) *p++ = '0';
        *p++ = '.'; *p++ = '0';
        return p;
    }
    memcpy(p, digits, decpt); p += decpt;              // dd.ddd
    *p++ = '.';
    memcpy(p, digits + decpt, nd - decpt);
    return p + (nd - decpt);
}

inline char *put_int(char *p, long long v)
{
    auto r = std::to_chars(p, p + 24, v);
    return r.ptr;
}

void format_traj(std::string &out, bool with_phi, int T, double Ts, long long id, const double *X, const double *U)
{
    out.resize((size_t)(T + 1) * 260);
    char *p = &out[0];
    for (int k = 0; k <= T; ++k) {
        const double *x = X + 6 * (size_t)k;
        p = put_pyfloat(p, (double)k * Ts); *p++ = ',';
        p = put_pyfloat(p, x[0]); *p++ = ',';
        p = put_pyfloat(p, x[1]); *p++ = ',';
        if (with_phi) { p = put_pyfloat(p, x[2]); *p++ = ','; }
        p = put_pyfloat(p, x[3]); *p++ = ',';
        p = put_pyfloat(p, x[4]); *p++ = ',';
        p = put_pyfloat(p, x[5]); *p++ = ',';
        if (k < T) p = put_pyfloat(p, U[2 * (size_t)k]);
        *p++ = ',';
        if (k < T) p = put_pyfloat(p, U[2 * (size_t)k + 1]);
        *p++ = ',';
        p = put_int(p, id);
        *p++ = '\n';
    }
    out.resize(p - &out[0]);
}

int write_one(const char *path, bool with_phi, int B, int T, double Ts, long long id0, const double *X, const double *U,
              int append, int nthreads)
{
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return TG_ERR_INVALID;
    if (!append) fputs(with_phi ? "t,X,Y,phi,vx,vy,omega,d,delta,trajectory_id\n" : "t,X,Y,vx,vy,omega,d,delta,trajectory_id\n", f);
    const int wave = nthreads * 8;
    std::vector<std::string> bufs(wave);
    for (int b0 = 0; b0 < B; b0 += wave) {
        const int nb = (B - b0 < wave) ? B - b0 : wave;
        std::vector<std::thread> th;
        for (int w = 0; w < nthreads; ++w)
            th.emplace_back([&, w]() {
                for (int i = w; i < nb; i += nthreads)
                    format_traj(bufs[i], with_phi, T, Ts, id0 + b0 + i, X + (size_t)(b0 + i) * (T + 1) * 6, U + (size_t)(b0 + i) * T * 2);
            });
        for (auto &t : th) t.join();
        for (int i = 0; i < nb; ++i)
            if (fwrite(bufs[i].data(), 1, bufs[i].size(), f) != bufs[i].size()) { fclose(f); return TG_ERR_INVALID; }
    }
    return fclose(f) == 0 ? TG_OK : TG_ERR_INVALID;
}

}  // namespace

extern "C" int tg_write_csv(const char *clean_path, const char *noisy_path, int B, int T, double Ts, int64_t traj_id0,
                            const double *clean, const double *noisy, const double *U, int append, int n_threads)
{
    if (B < 0 || T < 0 || !(Ts > 0) || (B > 0 && (!U && T > 0))) return TG_ERR_INVALID;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    int rc = TG_OK;
    if (clean_path && clean) rc = write_one(clean_path, true, B, T, Ts, traj_id0, clean, U, append, n_threads);
    if (rc == TG_OK && noisy_path && noisy) rc = write_one(noisy_path, false, B, T, Ts, traj_id0, noisy, U, append, n_threads);
    return rc;
}
