// ptk_csv.cuh -- host-side writer for rods_df_{color}.csv ("next" row).
//
// Replaces `df_out.to_csv(...)` at match2D.py:626-629: pandas spends ~1.2 s per colour of a
// 1000-frame experiment formatting floats, 300x the GPU time of the whole experiment.  Same file
// layout: unnamed RangeIndex first column, the given header, floats in Python's repr() form
// (shortest digits that round-trip; fixed notation for 1e-4 <= |x| < 1e16, else exponent form with
// a sign and at least two exponent digits), NaN as the empty field, integers as integers.
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace ptk {

// Python repr(float) into buf; returns the number of characters written.
inline int format_double_repr(double x, char* buf)
{
    if (std::isnan(x)) return 0;  // pandas na_rep=''
    if (std::isinf(x)) {
        const char* s = x > 0 ? "inf" : "-inf";
        const int n = (int)strlen(s);
        memcpy(buf, s, n);
        return n;
    }
    char tmp[40];
    auto res = std::to_chars(tmp, tmp + sizeof(tmp), x, std::chars_format::scientific);
    // tmp = [-]d[.ddd]e[+-]XX : split into sign, digits, exponent
    char* p = tmp;
    char* o = buf;
    if (*p == '-') { *o++ = '-'; p++; }
    char digits[24];
    int nd = 0;
    while (p < res.ptr && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        p++;
    }
    int e10 = 0;
    if (p < res.ptr && *p == 'e') {
        p++;
        int sign = 1;
        if (*p == '-') { sign = -1; p++; } else if (*p == '+') p++;
        while (p < res.ptr) e10 = e10 * 10 + (*p++ - '0');
        e10 *= sign;
    }
    if (nd == 1 && digits[0] == '0') {  // +-0.0
        memcpy(o, "0.0", 3);
        return (int)(o - buf) + 3;
    }
    if (e10 >= -4 && e10 < 16) {
        if (e10 < 0) {  // 0.000ddd
            *o++ = '0'; *o++ = '.';
            for (int k = 0; k < -e10 - 1; k++) *o++ = '0';
            memcpy(o, digits, nd); o += nd;
        } else if (nd <= e10 + 1) {  // integer valued: ddd000.0
            memcpy(o, digits, nd); o += nd;
            for (int k = 0; k < e10 + 1 - nd; k++) *o++ = '0';
            *o++ = '.'; *o++ = '0';
        } else {  // dd.ddd
            memcpy(o, digits, e10 + 1); o += e10 + 1;
            *o++ = '.';
            memcpy(o, digits + e10 + 1, nd - e10 - 1); o += nd - e10 - 1;
        }
    } else {  // d.ddde+XX  (repr: no ".0" when there is a single digit)
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, nd - 1); o += nd - 1; }
        *o++ = 'e';
        *o++ = e10 < 0 ? '-' : '+';
        int a = e10 < 0 ? -e10 : e10;
        char eb[8];
        int ne = 0;
        while (a) { eb[ne++] = (char)('0' + a % 10); a /= 10; }
        while (ne < 2) eb[ne++] = '0';
        while (ne) *o++ = eb[--ne];
    }
    return (int)(o - buf);
}

inline int format_int(long long v, char* buf)
{
    auto r = std::to_chars(buf, buf + 24, v);
    return (int)(r.ptr - buf);
}

}  // namespace ptk
