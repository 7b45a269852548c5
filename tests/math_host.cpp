// Host build of particles_b200/csrc/smcb_math.cuh (the step kernel's fp64 exp / log / sincos / batch
// log-sum-exp) so that their ALGORITHMS can be checked on the CPU against NumPy / mpmath
// (tests/test_math_host.py).  CUDA intrinsics are restated bit-for-bit below; compile with
// -ffp-contract=off so that only the explicit fma() calls fuse, as nvcc -fmad=false does.
//   g++ -O2 -ffp-contract=off -shared -fPIC [-DSMCB_TABLE_MATH=1] -I particles_b200/csrc tests/math_host.cpp
#include <cmath>
#include <cstdint>
#include <cstring>

#define SMCB_MATH_HOST_TEST 1
#define __device__
#define __forceinline__ inline
#define __constant__ const
#define CUDART_INF INFINITY
#define CUDART_NAN NAN
struct double2 { double x, y; };
static inline int __double2hiint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(uint32_t)(b >> 32); }
static inline int __double2loint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(uint32_t)b; }
static inline double __hiloint2double(int hi, int lo) {
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double v; std::memcpy(&v, &b, 8); return v;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
using std::fma; using std::fmax; using std::sqrt;
namespace smcb {
struct Lse3 { double m, s, q; };
struct Philox { uint32_t k0, k1, rk[20]; };
constexpr uint32_t kPurposeNormal = 1;
// the two uniform constructions of smcb_common.cuh (53-bit mantissa from two 32-bit words)
static inline double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (b >> 6)) * 1.1102230246251565404e-16;
}
static inline double u53_open(uint32_t a, uint32_t b) {
    return ((double)(((uint64_t)(a >> 5) << 26) | (b >> 6)) + 0.5) * 1.1102230246251565404e-16;
}
static inline void philox4x32_10k(uint32_t, uint32_t, uint32_t, uint32_t, const Philox &, uint32_t *) {}
}  // namespace smcb
#include "smcb_math.cuh"

extern "C" {
int mh_table_math() { return SMCB_TABLE_MATH; }
void mh_exp(const double *x, double *y, long n, int kind) {
    for (long i = 0; i < n; i++)
        y[i] = kind == 0 ? smcb::fexp(x[i]) : (kind == 1 ? smcb::fexp_neg(x[i]) : smcb::fexp_mid(x[i]));
}
void mh_log(const double *x, double *y, long n) { for (long i = 0; i < n; i++) y[i] = smcb::flog_pos(x[i]); }
void mh_sincos2pi(const double *u, double *s, double *c, long n) { for (long i = 0; i < n; i++) smcb::fsincos2pi(u[i], s[i], c[i]); }
void mh_box_muller(const uint32_t *r, double *z, long npairs) {
    for (long i = 0; i < npairs; i++) smcb::box_muller_fast(r + 4 * i, z[2 * i], z[2 * i + 1]);
}
// (m, s, q) of v accumulated in batches of 4, as the step kernel does
void mh_lse3(const double *v, long n, double *out3) {
    smcb::Lse3 a{-INFINITY, 0.0, 0.0};
    long i = 0;
    for (; i + 4 <= n; i += 4) { const double b[4] = {v[i], v[i + 1], v[i + 2], v[i + 3]}; smcb::lse3_add_batch<4>(a, b); }
    for (; i < n; i++) { const double b[1] = {v[i]}; smcb::lse3_add_batch<1>(a, b); }
    out3[0] = a.m; out3[1] = a.s; out3[2] = a.q;
}
}
