// tg_csv.cpp -- host-side dataset writer: the clean / noisy CSV files of the generators, byte-identical to what
// pandas' DataFrame.to_csv(index=False) writes for the reference schema (generation_type1.py:315-339,
// generation_type2.py:202-218,309-322): columns t,X,Y,[phi],vx,vy,omega,d,delta,trajectory_id; T+1 rows per
// trajectory; t = k*Ts; the last row's d, delta are NaN -> empty fields; floats in Python's shortest repr.
// At ~1e7 MPC steps/s the pandas writer is >99 % of the wall time of a generation run; this one formats with
// std::to_chars (shortest round-trip digits) on all host cores.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trajgen.h"

void tg_internal_set_error(const char *msg);   // trajgen.cu

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

static int io_fail(const char *what, const char *path)
{
    const std::string msg = std::string("tg_write_csv: ") + what + " '" + path + "': " + strerror(errno);
    tg_internal_set_error(msg.c_str());
    return TG_ERR_INVALID;
}

int write_one(const char *path, bool with_phi, int B, int T, double Ts, long long id0, const double *X, const double *U,
              int append, int nthreads)
{
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return io_fail("cannot open", path);
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
            if (fwrite(bufs[i].data(), 1, bufs[i].size(), f) != bufs[i].size()) { const int rc_ = io_fail("write failed on", path); fclose(f); return rc_; }
    }
    return fclose(f) == 0 ? TG_OK : io_fail("close failed on", path);
}

// ---- merge (generation_traj/merge_datasets.py)
struct LineReader {
    FILE *f;
    std::vector<char> buf;
    size_t lo = 0, hi = 0;
    bool eof = false;
    explicit LineReader(FILE *f_) : f(f_), buf(8u << 20) {}
    // next line without its terminator (handles a last line that lacks '\n'); false at end of file
    bool next(const char *&line, size_t &len)
    {
        for (;;) {
            if (lo < hi) {
                const char *nl = (const char *)memchr(buf.data() + lo, '\n', hi - lo);
                if (nl) { line = buf.data() + lo; len = (size_t)(nl - line); lo += len + 1; if (len && line[len - 1] == '\r') --len; return true; }
                if (eof) { line = buf.data() + lo; len = hi - lo; lo = hi; if (len && line[len - 1] == '\r') --len; return len > 0; }
            } else if (eof) return false;
            if (lo > 0) { memmove(buf.data(), buf.data() + lo, hi - lo); hi -= lo; lo = 0; }
            if (hi == buf.size()) buf.resize(buf.size() * 2);
            const size_t got = fread(buf.data() + hi, 1, buf.size() - hi, f);
            hi += got;
            if (got == 0) eof = true;
        }
    }
};

// [begin, end) of column `col` of a CSV line without quoting (the schema is numeric)
inline bool csv_field(const char *line, size_t len, int col, size_t &b, size_t &e)
{
    size_t pos = 0;
    for (int c = 0; c < col; ++c) {
        const char *k = (const char *)memchr(line + pos, ',', len - pos);
        if (!k) return false;
        pos = (size_t)(k - line) + 1;
    }
    const char *k = (const char *)memchr(line + pos, ',', len - pos);
    b = pos; e = k ? (size_t)(k - line) : len;
    return true;
}

inline bool parse_id(const char *s, size_t n, long long &v)
{
    auto r = std::from_chars(s, s + n, v);
    if (r.ec != std::errc()) return false;
    // pandas writes an integer column that went through NaN as "12.0": accept a zero fraction
    const char *q = r.ptr;
    if (q < s + n && *q == '.') { ++q; while (q < s + n && *q == '0') ++q; }
    return q == s + n;
}

}  // namespace

static int merge_fail(const std::string &msg) { tg_internal_set_error(msg.c_str()); return TG_ERR_INVALID; }

extern "C" int tg_merge_csv(const char *first_path, const char *second_path, const char *out_path, int64_t id_offset,
                            int64_t *id_offset_used, int64_t *rows_written)
{
    if (!first_path || !second_path || !out_path) return merge_fail("tg_merge_csv: null path");
    FILE *fa = fopen(first_path, "rb");
    if (!fa) return merge_fail(std::string("File not found: '") + first_path + "'");     // merge_datasets.py:26
    FILE *fb = fopen(second_path, "rb");
    if (!fb) { fclose(fa); return merge_fail(std::string("File not found: '") + second_path + "'"); }
    FILE *fo = fopen(out_path, "wb");
    if (!fo) { fclose(fa); fclose(fb); return merge_fail(std::string("cannot create '") + out_path + "'"); }
    std::vector<char> obuf(8u << 20);
    setvbuf(fo, obuf.data(), _IOFBF, obuf.size());
    struct Closer { FILE *a, *b, *o; ~Closer() { if (a) fclose(a); if (b) fclose(b); if (o) fclose(o); } } closer{fa, fb, fo};
    LineReader ra(fa), rb(fb);
    const char *line; size_t len;
    if (!ra.next(line, len)) return merge_fail("tg_merge_csv: first file is empty");
    const std::string header(line, len);
    int col = -1, ncol = 0;
    for (size_t pos = 0;; ++ncol) {
        const size_t k = header.find(',', pos);
        if (header.compare(pos, (k == std::string::npos ? header.size() : k) - pos, "trajectory_id") == 0) col = ncol;
        if (k == std::string::npos) { ++ncol; break; }
        pos = k + 1;
    }
    if (col < 0) return merge_fail("tg_merge_csv: no trajectory_id column");
    if (!rb.next(line, len) || header != std::string(line, len)) return merge_fail("tg_merge_csv: the two files have different columns");
    fwrite(header.data(), 1, header.size(), fo); fputc('\n', fo);
    long long max_id = -1, rows = 0;
    while (ra.next(line, len)) {          // first file verbatim (merge_datasets.py:37,50)
        size_t b, e; long long id;
        if (!csv_field(line, len, col, b, e) || !parse_id(line + b, e - b, id)) return merge_fail("tg_merge_csv: bad trajectory_id in the first file");
        if (id > max_id) max_id = id;
        fwrite(line, 1, len, fo); fputc('\n', fo);
        ++rows;
    }
    const long long off = id_offset >= 0 ? id_offset : max_id + 1;    // :42-45
    char num[32];
    while (rb.next(line, len)) {          // second file with re-indexed ids (:48)
        size_t b, e; long long id;
        if (!csv_field(line, len, col, b, e) || !parse_id(line + b, e - b, id)) return merge_fail("tg_merge_csv: bad trajectory_id in the second file");
        fwrite(line, 1, b, fo);
        char *q = put_int(num, id + off);
        fwrite(num, 1, (size_t)(q - num), fo);
        fwrite(line + e, 1, len - e, fo);
        fputc('\n', fo);
        ++rows;
    }
    closer.o = nullptr;
    if (fclose(fo) != 0) return merge_fail("tg_merge_csv: write failed");
    if (id_offset_used) *id_offset_used = off;
    if (rows_written) *rows_written = rows;
    return TG_OK;
}

extern "C" int tg_write_csv(const char *clean_path, const char *noisy_path, int B, int T, double Ts, int64_t traj_id0,
                            const double *clean, const double *noisy, const double *U, int append, int n_threads)
{
    if (B < 0 || T < 0 || !(Ts > 0) || (B > 0 && (!U && T > 0))) { tg_internal_set_error("tg_write_csv: bad argument (B, T >= 0, Ts > 0, U required when T > 0)"); return TG_ERR_INVALID; }
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    int rc = TG_OK;
    if (clean_path && clean) rc = write_one(clean_path, true, B, T, Ts, traj_id0, clean, U, append, n_threads);
    if (rc == TG_OK && noisy_path && noisy) rc = write_one(noisy_path, false, B, T, Ts, traj_id0, noisy, U, append, n_threads);
    return rc;
}
