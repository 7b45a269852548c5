// smcb_tables.h -- the lookup tables of the step kernels' fp64 exp / log / sincos (smcb_math.cuh), built on
// the host in long double at context creation and staged into shared memory by every step-kernel CTA (one TMA
// bulk copy, 64 KB).  Plain C++ so that the CPU build of the math header (tests/math_host.cpp) uses the same
// function.  Layout (doubles):
//   [kTabExp, +4096)       T[j]   = 2^(j / 4096)
//   [kTabLog, +2 x 1024)   L[j]   = { 1/c_j, -log(1/c_j) },  c_j = centre of mantissa cell j of [1, 2)
//                                   (cells >= kLogSplit belong to m/2 in [0.707, 1): c_j is halved)
//   [kTabSc,  +2 x 1024)   S[j]   = { sin(2 pi j / 1024), cos(2 pi j / 1024) }   (exact symmetries)
#pragma once
#include <math.h>

namespace smcb {

constexpr int kExpTabBits = 12, kExpTabN = 1 << kExpTabBits;
constexpr int kLogTabBits = 10, kLogTabN = 1 << kLogTabBits;
constexpr int kLogSplit = 0x1A8;           // mantissa cells >= this hold values >= 1.4140625: treated as m/2
constexpr int kScTabBits = 10, kScTabN = 1 << kScTabBits;
constexpr int kTabExp = 0, kTabLog = kExpTabN, kTabSc = kTabLog + 2 * kLogTabN;
constexpr int kMathTabDoubles = kTabSc + 2 * kScTabN;                  // 8192 doubles = 64 KB
constexpr size_t kMathTabBytes = (size_t)kMathTabDoubles * sizeof(double);

inline void fill_math_tables(double *t) {
    for (int j = 0; j < kExpTabN; j++) t[kTabExp + j] = (double)exp2l((long double)j / kExpTabN);
    for (int j = 0; j < kLogTabN; j++) {
        long double c = 1.0L + ((long double)j + 0.5L) / kLogTabN;
        if (j >= kLogSplit) c *= 0.5L;
        const double inv = (double)(1.0L / c);
        t[kTabLog + 2 * j] = inv;
        t[kTabLog + 2 * j + 1] = (double)(-logl((long double)inv));
    }
    const long double two_pi = 6.283185307179586476925286766559L;
    const int quarter = kScTabN / 4;
    for (int j = 0; j < kScTabN; j++) {
        const int q = j / quarter, k = j % quarter;
        long double s0 = 0.0L, c0 = 1.0L;
        if (k != 0) { const long double a = two_pi * (long double)k / kScTabN; s0 = sinl(a); c0 = cosl(a); }
        long double s, c;
        switch (q) {
            case 0: s = s0; c = c0; break;
            case 1: s = c0; c = -s0; break;
            case 2: s = -s0; c = -c0; break;
            default: s = -c0; c = s0; break;
        }
        t[kTabSc + 2 * j] = (double)s;
        t[kTabSc + 2 * j + 1] = (double)c;
    }
}

}  // namespace smcb
