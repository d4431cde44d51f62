// dsb_trajio.cpp -- one line of a trajectories file (host only, no GPU).
//
// The reference writes the positions of every time point as text, one value at a time
// (disimpy/simulations.py:1043-1048: `f.write(str(i) + " ")` for every element of
// positions.ravel(), then a newline); str() of a float64 is the shortest string that reads back to
// the same double, laid out like Python's repr.  That is what this file produces, for all values of
// a time point at once and on several threads: shortest digits from std::to_chars, then Python's
// layout rules (fixed notation for decimal exponents -4 ... 15, otherwise d[.ddd]e+XX with at least
// two exponent digits; "nan", "inf", "-inf"; "-0.0").
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>

#include "../../include/disimpy_b200.h"

namespace {

constexpr int kMaxChars = 26;   // "-1.2345678901234567e-308" + the separating space, with room to spare

// str(float64) + " " at p; returns the number of characters written
int format_value(double v, char *p)
{
    char *const p0 = p;
    if (v != v) {
        memcpy(p, "nan ", 4);
        return 4;
    }
    if (std::signbit(v)) {
        *p++ = '-';
        v = -v;
    }
    if (v > 1.7976931348623157e308) {
        memcpy(p, "inf ", 4);
        return (int)(p - p0) + 4;
    }
    // shortest round-trip digits: d[.ddd]e[+-]XX
    char sci[40];
    const auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);
    const char *e = sci;
    while (*e != 'e') ++e;
    char digits[20];
    int nd = 0;
    for (const char *c = sci; c < e; ++c)
        if (*c != '.') digits[nd++] = *c;
    int exp10 = 0;
    for (const char *c = e + 2; c < r.ptr; ++c) exp10 = exp10 * 10 + (*c - '0');
    if (e[1] == '-') exp10 = -exp10;
    if (v == 0.0) exp10 = 0;
    const int decpt = exp10 + 1;   // value = 0.d1d2... x 10^decpt
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            *p++ = '0';
            *p++ = '.';
            for (int k = 0; k < -decpt; ++k) *p++ = '0';
            memcpy(p, digits, nd);
            p += nd;
        } else if (decpt >= nd) {
            memcpy(p, digits, nd);
            p += nd;
            for (int k = nd; k < decpt; ++k) *p++ = '0';
            *p++ = '.';
            *p++ = '0';
        } else {
            memcpy(p, digits, decpt);
            p += decpt;
            *p++ = '.';
            memcpy(p, digits + decpt, nd - decpt);
            p += nd - decpt;
        }
    } else {
        *p++ = digits[0];
        if (nd > 1) {
            *p++ = '.';
            memcpy(p, digits + 1, nd - 1);
            p += nd - 1;
        }
        *p++ = 'e';
        *p++ = exp10 < 0 ? '-' : '+';
        const int a = exp10 < 0 ? -exp10 : exp10;
        if (a >= 100) *p++ = (char)('0' + a / 100);
        *p++ = (char)('0' + (a / 10) % 10);
        *p++ = (char)('0' + a % 10);
    }
    *p++ = ' ';
    return (int)(p - p0);
}

}  // namespace

extern "C" {

// The text of one time point: str(v) + " " for each of the n values, then "\n".  `out` must hold
// 26 * n + 1 characters; *len receives the number written (no terminating zero).
int dsb_format_traj_line(const double *values, int64_t n, char *out, int64_t capacity, int64_t *len)
{
    if (n < 0 || (n > 0 && !values) || !out || !len || capacity < (int64_t)kMaxChars * n + 1) return DSB_EINVAL;
    const int64_t grain = 1 << 15;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)hw, (n + grain - 1) / grain, 32}));
    if (n_threads > 1) {
        // every thread formats its range where the longest possible text of the ranges before it
        // would end, then the pieces are moved together
        int64_t used[32] = {0};
        std::thread pool[32];
        const int64_t per = (n + n_threads - 1) / n_threads;
        int started = 0;
        try {
            for (; started < n_threads; ++started) {
                const int t = started;
                pool[t] = std::thread([=, &used] {
                    const int64_t a = std::min<int64_t>(n, t * per), b = std::min<int64_t>(n, a + per);
                    char *p = out + (int64_t)kMaxChars * a;
                    for (int64_t i = a; i < b; ++i) p += format_value(values[i], p);
                    used[t] = p - (out + (int64_t)kMaxChars * a);
                });
            }
        } catch (...) {   // no more threads to be had: the serial loop below redoes the line
        }
        for (int t = 0; t < started; ++t) pool[t].join();
        if (started == n_threads) {
            char *p = out + used[0];
            for (int t = 1; t < n_threads; ++t) {
                const int64_t a = std::min<int64_t>(n, t * per);
                memmove(p, out + (int64_t)kMaxChars * a, (size_t)used[t]);
                p += used[t];
            }
            *p++ = '\n';
            *len = p - out;
            return DSB_OK;
        }
    }
    char *p = out;
    for (int64_t i = 0; i < n; ++i) p += format_value(values[i], p);
    *p++ = '\n';
    *len = p - out;
    return DSB_OK;
}

}  // extern "C"
