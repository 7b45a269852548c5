// smcb_math.cuh -- fp64 elementary functions for the step kernel, written for its
// restricted domains.  Why not the CUDA math library: in the first ncu capture of k_move
// (profiles/r01_*) 22% of all issued warp-instructions were UMOVs materialising the library's
// 64-bit polynomial immediates and 7% were branches around special cases; here the
// coefficients sit in the constant bank (DFMA takes a c[bank][offset] operand directly), the
// domains are restricted and the code is branch-free.  Accuracy: polynomial errors < 5e-18
// (gen_coeffs.py), total error <= ~1.5 ulp; checked against NumPy/mpmath in
// tests/test_gpu_kernels.py::test_device_math.
#pragma once
#ifndef SMCB_MATH_HOST_TEST      // tests/math_host.cpp compiles this header for the CPU with its own shim
#include "smcb_common.cuh"
#endif
#include "smcb_math_coeffs.inc"

// SMCB_TABLE_MATH=1 / 2: table-assisted exp and log (shorter polynomials, one L1-resident table load each):
//   exp: 1024-entry table of 2^(j/1024) + degree-4 polynomial  (11 -> 4 dependent DFMAs), or with
//        SMCB_TABLE_MATH=2 a 128-entry table (8 cache lines) + degree 5
//   log: 256-entry table of {1/c, -log(1/c)} + degree-6 log1p  (no fp64 division)
// Same accuracy class (<= ~1.5 ulp, tests/test_math_host.py checks both builds on the CPU).  Off by
// default until timed on the device against the polynomial-only build (profiles/build_variant.sh).
#ifndef SMCB_TABLE_MATH
#define SMCB_TABLE_MATH 0
#endif
#if SMCB_TABLE_MATH
#include "smcb_math_tables.inc"
#endif

namespace smcb {

constexpr double kRintMagic = 6755399441055744.0;  // 1.5 * 2^52: x + magic rounds x to nearest int

#ifndef SMCB_ESTRIN
#define SMCB_ESTRIN 0
#endif
// polynomial evaluation: Horner (N-1 dependent FMAs) or Estrin (depth ~log2 N, a few more
// multiplies) -- the latter shortens the dependent fp64 chains the step kernel waits on
template <int N>
__device__ __forceinline__ double horner(const double (&c)[N], double x) {
#if SMCB_ESTRIN
    constexpr int H = (N + 1) / 2;
    double q[H];
#pragma unroll
    for (int i = 0; i < N / 2; i++) q[i] = fma(c[2 * i + 1], x, c[2 * i]);
    if (N & 1) q[H - 1] = c[N - 1];
    double xp = x * x;
    int m = H;
#pragma unroll
    for (int level = 0; level < 5; level++) {
        if (m > 1) {
            const int h = (m + 1) / 2;
#pragma unroll
            for (int i = 0; i < H / 2 + 1; i++) {
                if (i < m / 2) q[i] = fma(q[2 * i + 1], xp, q[2 * i]);
            }
            if (m & 1) q[h - 1] = q[m - 1];
            xp = xp * xp;
            m = h;
        }
    }
    return q[0];
#else
    double p = c[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 0; i--) p = fma(p, x, c[i]);
    return p;
#endif
}

#if SMCB_TABLE_MATH
// exp(x) without range selects: x = (k/1024) ln2 + r, exp(x) = 2^(k >> 10) * T[k & 1023] * P4(r)
__device__ __forceinline__ double fexp_core(double x) {
#if SMCB_TABLE_MATH == 2      // 128-entry table (1 KB: a divergent lookup touches <= 8 L1 lines), degree 5
    constexpr int kBits = 7;
#else                         // 1024-entry table (8 KB), degree 4
    constexpr int kBits = 10;
#endif
    constexpr double kScale = (double)(1 << kBits);
    const double t = fma(x, kScale * SMCB_LOG2E, kRintMagic);
    const double kd = t - kRintMagic;
    const int k = __double2loint(t);                       // |k| < 2^20 for |x| <= 709
    double r = fma(kd, -(SMCB_LN2_HI / kScale), x);        // power-of-two scalings of the split: exact
    r = fma(kd, -(SMCB_LN2_LO / kScale), r);
    // exp(r) - 1 = r + r^2 (c2 + c3 r + ...); T + T * (exp(r) - 1) keeps the table value's half ulp
#if SMCB_TABLE_MATH == 2
    const double tj = __ldg(&kExp2Tab128[k & 127]);
    double h = fma(kExp5C[5], r, kExp5C[4]);
    h = fma(h, r, kExp5C[3]);
    h = fma(h, r, kExp5C[2]);
#else
    const double tj = __ldg(&kExp2Tab[k & 1023]);
    double h = fma(kExp4C[4], r, kExp4C[3]);
    h = fma(h, r, kExp4C[2]);
#endif
    const double em1 = fma(r * r, h, r);
    return fma(tj, em1, tj) * __hiloint2double(((k >> kBits) + 1023) << 20, 0);
}
__device__ __forceinline__ double fexp(double x) {
    double res = fexp_core(x);
    res = (x < -708.0) ? 0.0 : res;
    res = (x > 709.0) ? CUDART_INF : res;
    return res;
}
__device__ __forceinline__ double fexp_neg(double x) {
    const double res = fexp_core(x);
    return (x < -708.0) ? 0.0 : res;
}
__device__ __forceinline__ double fexp_mid(double x) { return fexp_core(x); }

// log(x) for positive NORMAL x: m in [sqrt(1/2), sqrt(2)) -> table interval j, r = m / c_j - 1, |r| <= 2^-7
__device__ __forceinline__ double flog_pos(double x) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) | 0x3FF00000;
    const bool big = hi > 0x3FF6A09E;
    hi = big ? hi - 0x00100000 : hi;
    e = big ? e + 1 : e;
    const double m = __hiloint2double(hi, lo);
    const double2 tc = __ldg(&kLogTab[(hi >> 13) & 0xFF]);
    const double r = fma(m, tc.x, -1.0);
    const double l1p = fma(r * r, horner(kLog1pC, r), r);
    const double ed = (double)e;
    return fma(ed, SMCB_LN2_HI, tc.y + fma(ed, SMCB_LN2_LO, l1p));
}
#else
// exp(x): x <= ~709; returns 0 for x < -708 (incl. -inf; the lost range is < 3e-308),
// +inf for x > 709, NaN for NaN.
__device__ __forceinline__ double fexp(double x) {
    const double t = fma(x, SMCB_LOG2E, kRintMagic);
    const double kd = t - kRintMagic;
    const int k = __double2loint(t);                       // low word of t holds the integer
    double r = fma(kd, -SMCB_LN2_HI, x);
    r = fma(kd, -SMCB_LN2_LO, r);
    const double p = horner(kExpC, r);
    const double scale = __hiloint2double((k + 1023) << 20, 0);   // 2^k, k in [-1022, 1023]
    double res = p * scale;
    res = (x < -708.0) ? 0.0 : res;
    res = (x > 709.0) ? CUDART_INF : res;
    return res;
}

// exp(x) for x <= 0 (weights relative to their maximum): no overflow branch
__device__ __forceinline__ double fexp_neg(double x) {
    const double t = fma(x, SMCB_LOG2E, kRintMagic);
    const double kd = t - kRintMagic;
    const int k = __double2loint(t);
    double r = fma(kd, -SMCB_LN2_HI, x);
    r = fma(kd, -SMCB_LN2_LO, r);
    const double p = horner(kExpC, r);
    const double res = p * __hiloint2double((k + 1023) << 20, 0);
    return (x < -708.0) ? 0.0 : res;
}

// exp(x) for |x| < 700 guaranteed by the caller's domain (no range selects at all); -inf -> NaN!
__device__ __forceinline__ double fexp_mid(double x) {
    const double t = fma(x, SMCB_LOG2E, kRintMagic);
    const double kd = t - kRintMagic;
    const int k = __double2loint(t);
    double r = fma(kd, -SMCB_LN2_HI, x);
    r = fma(kd, -SMCB_LN2_LO, r);
    return horner(kExpC, r) * __hiloint2double((k + 1023) << 20, 0);
}

// log(x) for positive NORMAL x (the 53-bit uniforms of box_muller are >= 2^-54)
__device__ __forceinline__ double flog_pos(double x) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) | 0x3FF00000;                   // mantissa in [1, 2)
    const bool big = hi > 0x3FF6A09E;                      // > sqrt(2) (top word compare is enough)
    hi = big ? hi - 0x00100000 : hi;                       // m/2
    e = big ? e + 1 : e;
    const double m = __hiloint2double(hi, lo);
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    const double p = horner(kLogC, z);                     // 2 atanh(s) / s
    const double ed = (double)e;
    return fma(ed, SMCB_LN2_HI, fma(s, p, ed * SMCB_LN2_LO));
}
#endif  // SMCB_TABLE_MATH

// (sin, cos)(2 pi u) for u in [0, 1)
__device__ __forceinline__ void fsincos2pi(double u, double &s, double &c) {
    const double t4 = u * 4.0;                             // quadrant index = rint(4u) in 0..4
    const double tm = t4 + kRintMagic;
    const double qd = tm - kRintMagic;
    const int q = __double2loint(tm);
    const double r = (t4 - qd) * 0.5;                      // angle = pi (q/2 + r), |r| <= 1/4
    const double z = r * r;
    const double sp = r * horner(kSinPiC, z);              // sin(pi r)
    const double cp = horner(kCosPiC, z);                  // cos(pi r)
    double ss = (q & 1) ? cp : sp;
    double cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// two N(0,1) from one Philox block (Box-Muller), fast fp64 path
__device__ __forceinline__ void box_muller_fast(const uint32_t r[4], double &z0, double &z1) {
    const double u1 = u53_open(r[0], r[1]);
    const double u2 = u53(r[2], r[3]);
    const double rad = sqrt(-2.0 * flog_pos(u1));
    double s, c;
    fsincos2pi(u2, s, c);
    z0 = rad * c;
    z1 = rad * s;
}

__device__ __forceinline__ void normal_pair_fast(const Philox &key, uint64_t pair, uint32_t t,
                                                 uint32_t comp, double &z0, double &z1) {
    uint32_t r[4];
    philox4x32_10k((uint32_t)pair, (uint32_t)(pair >> 32), t, (comp << 8) | kPurposeNormal, key, r);
    box_muller_fast(r, z0, z1);
}

// ---------------------------------------------------------------------------
// (max, sum exp, sum exp^2) accumulation of a small batch with ONE exp per value:
// the running shift m only moves when the batch maximum exceeds it (one extra exp per
// batch at most, amortised over NV values), instead of a rescale test per value.
// Values equal to -inf (or masked-out slots set to -inf) contribute exactly 0.
// ---------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void lse3_add_batch(Lse3 &a, const double (&v)[NV]) {
    double mb = v[0];
#pragma unroll
    for (int j = 1; j < NV; j++) mb = fmax(mb, v[j]);
    if (mb > a.m) {                       // also the first time (a.m = -inf): fexp(-inf) = 0
        const double r = fexp_neg(a.m - mb);
        a.s *= r;
        a.q *= r * r;
        a.m = mb;
    }
    if (a.m == -CUDART_INF) return;       // nothing but -inf so far
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const double e = fexp_neg(v[j] - a.m);
        a.s += e;
        a.q = fma(e, e, a.q);
    }
}

}  // namespace smcb
