// smcb_models.cuh -- device restatements of the stock state-space models of the
// reference (the closures that particles builds ProbDist objects from), as structs of
// constants + inline maps, so that PX / PY / proposal / logeta become register math
// inside the fused step kernel.
//
// Every model of the 1-D family exposes (all Normal kernels, as in the reference):
//   init (loc, scale)                     PX0()
//   trans(k, xp) -> (loc, scale)          PX(t, xp)
//   obs_logpdf(k, xp, x)                  PY(t, xp, x).logpdf(data[t])
//   prop0 / prop -> (loc, scale)          proposal0(data) / proposal(t, xp, data)
//   logeta(k, x)                          logeta(t, x, data)   (uses data[t+1])
// `k` carries the per-step scalars (t, y_t, y_{t+1}, host-computed step constants).
// Parameter layout of smcb_filter_desc.params is documented per model below.
#pragma once
#include "smcb_common.cuh"
#include "smcb_math.cuh"

namespace smcb {

struct StepK {
    double y;        // data[t]
    double y_next;   // data[t+1] (0 at the last step; only logeta reads it)
    double sc0;      // host-computed per-step constant (model specific)
    int64_t t;
};

// Normal.logpdf with log(scale) supplied (constant scales: computed once on the host
// with numpy, so the constant is bit-identical to the reference's np.log(scale))
__device__ __forceinline__ double normal_logpdf_ls(double x, double loc, double scale,
                                                   double logscale) {
    double z = (x - loc) / scale;
    return -z * z / 2.0 - kHalfLog2Pi - logscale;
}

// ---------------------------------------------------------------------------
// StochVol -- particles/state_space_models.py:446-498
// params: 0 mu, 1 rho, 2 sigma, 3 sig0, 4 (1-rho)*mu, 5 log(sigma), 6 log(sig0)
// ---------------------------------------------------------------------------
struct StochVolM {
    double mu, rho, sigma, sig0, c0, lsigma, lsig0;
    static constexpr bool has_proposal = true;
    __host__ void load(const double *p) {
        mu = p[0]; rho = p[1]; sigma = p[2]; sig0 = p[3]; c0 = p[4]; lsigma = p[5]; lsig0 = p[6];
    }
    __device__ __forceinline__ void init(double &loc, double &scale, double &ls) const {
        loc = mu; scale = sig0; ls = lsig0;                       // PX0, :462
    }
    __device__ __forceinline__ double ext(double xp) const { return c0 + rho * xp; }  // EXt, :465-467
    __device__ __forceinline__ void trans(const StepK &, double xp, double &loc, double &scale,
                                          double &ls) const {
        loc = ext(xp); scale = sigma; ls = lsigma;                // PX, :469-470
    }
    // PY = Normal(0, exp(x/2)), :472-473: z = y / exp(x/2), logpdf = -z^2/2 - log(2 pi)/2 - log(scale).
    // Evaluated with ONE exp: z^2 = y^2 exp(-x) and log(exp(x/2)) = x/2 (differences of a few ulp
    // of the individual terms, < 1e-15 absolute; tolerance of the parity tests is 1e-10).
    __device__ __forceinline__ double obs_logpdf(const StepK &k, double, double x) const {
        const double z2 = (k.y * k.y) * fexp(-x);
        return -0.5 * z2 - kHalfLog2Pi - 0.5 * x;
    }
    __device__ __forceinline__ double xhat(double xst, double sig, double yt) const {  // :475-476
        return xst + 0.5 * (sig * sig) * ((yt * yt) * fexp(-xst) - 1.0);
    }
    __device__ __forceinline__ void prop0(const StepK &k, double &loc, double &scale, double &ls) const {
        loc = xhat(0.0, sig0, k.y); scale = sig0; ls = lsig0;     // :478-482
    }
    __device__ __forceinline__ void prop(const StepK &k, double xp, double &loc, double &scale,
                                         double &ls) const {
        loc = xhat(ext(xp), sigma, k.y); scale = sigma; ls = lsigma;  // :484-488
    }
    __device__ __forceinline__ double logeta(const StepK &k, double x) const {  // :490-498
        double xst = ext(x);
        double xstmmu = xst - mu;
        double e = fexp(-xst);
        double xh = xst + 0.5 * (sigma * sigma) * ((k.y_next * k.y_next) * e - 1.0);
        double xhatmmu = xh - mu;
        return 0.5 / (sigma * sigma) * (xhatmmu * xhatmmu - xstmmu * xstmmu) -
               0.5 * (k.y_next * k.y_next) * e * (1.0 + xstmmu);
    }
};

// ---------------------------------------------------------------------------
// LinearGauss -- particles/kalman.py:397-452; README ToySSM = (rho=1, sigmaX=1, sigma0=1)
// params: 0 rho, 1 sigmaX, 2 sigmaY, 3 sigma0, 4 log sigmaX, 5 log sigmaY, 6 log sigma0,
//         7 sig2post0, 8 sqrt(sig2post0), 9 log sqrt(sig2post0), 10 sig2post, 11 sqrt(sig2post),
//         12 log sqrt(sig2post), 13 sqrt(sX^2+sY^2), 14 log of 13, 15 sigmaX^2, 16 sigmaY^2
// ---------------------------------------------------------------------------
struct LinGaussM {
    double rho, sX, sY, s0, lsX, lsY, ls0, s2p0, sp0, lsp0, s2p, sp, lsp, se, lse, sX2, sY2;
    static constexpr bool has_proposal = true;
    __host__ void load(const double *p) {
        rho = p[0]; sX = p[1]; sY = p[2]; s0 = p[3]; lsX = p[4]; lsY = p[5]; ls0 = p[6];
        s2p0 = p[7]; sp0 = p[8]; lsp0 = p[9]; s2p = p[10]; sp = p[11]; lsp = p[12];
        se = p[13]; lse = p[14]; sX2 = p[15]; sY2 = p[16];
    }
    __device__ __forceinline__ void init(double &loc, double &scale, double &ls) const {
        loc = 0.0; scale = s0; ls = ls0;                          // PX0, :426-427
    }
    __device__ __forceinline__ void trans(const StepK &, double xp, double &loc, double &scale,
                                          double &ls) const {
        loc = rho * xp; scale = sX; ls = lsX;                     // PX, :429-430
    }
    __device__ __forceinline__ double obs_logpdf(const StepK &k, double, double x) const {
        return normal_logpdf_ls(k.y, x, sY, lsY);                 // PY, :432-433
    }
    __device__ __forceinline__ void prop0(const StepK &k, double &loc, double &scale, double &ls) const {
        loc = s2p0 * (k.y / sY2); scale = sp0; ls = lsp0;         // :435-438
    }
    __device__ __forceinline__ void prop(const StepK &k, double xp, double &loc, double &scale,
                                         double &ls) const {
        loc = s2p * (rho * xp / sX2 + k.y / sY2); scale = sp; ls = lsp;  // :440-445
    }
    __device__ __forceinline__ double logeta(const StepK &k, double x) const {  // :447-451
        return normal_logpdf_ls(k.y_next, rho * x, se, lse);
    }
};

// ---------------------------------------------------------------------------
// Gordon et al -- particles/state_space_models.py:546-577 (Bootstrap only)
// params: 0 a, 1 b, 2 c, 3 sigmaX, 4 log sigmaX;  step constant sc0 = d*cos(e*(t-1))
// ---------------------------------------------------------------------------
struct GordonM {
    double a, b, c, sX, lsX;
    static constexpr bool has_proposal = false;
    __host__ void load(const double *p) { a = p[0]; b = p[1]; c = p[2]; sX = p[3]; lsX = p[4]; }
    __device__ __forceinline__ void init(double &loc, double &scale, double &ls) const {
        loc = 0.0; scale = 2.0; ls = 0.69314718055994530942;      // Normal(scale=2.), :563-564
    }
    __device__ __forceinline__ void trans(const StepK &k, double xp, double &loc, double &scale,
                                          double &ls) const {
        loc = b * xp + c * xp / (1.0 + xp * xp) + k.sc0;          // :566-572
        scale = sX; ls = lsX;
    }
    __device__ __forceinline__ double obs_logpdf(const StepK &k, double, double x) const {
        return normal_logpdf_ls(k.y, a * (x * x), 1.0, 0.0);      // Normal(loc=a x^2), :574-575
    }
    __device__ __forceinline__ void prop0(const StepK &, double &, double &, double &) const {}
    __device__ __forceinline__ void prop(const StepK &, double, double &, double &, double &) const {}
    __device__ __forceinline__ double logeta(const StepK &, double) const { return 0.0; }
};

// ---------------------------------------------------------------------------
// ThetaLogistic -- particles/state_space_models.py:657-689 (Bootstrap)
// params: 0 tau0, 1 tau1, 2 tau2, 3 sigmaX, 4 sigmaY, 5 log sigmaX, 6 log sigmaY
// ---------------------------------------------------------------------------
struct ThetaLogisticM {
    double tau0, tau1, tau2, sX, sY, lsX, lsY;
    static constexpr bool has_proposal = false;
    __host__ void load(const double *p) {
        tau0 = p[0]; tau1 = p[1]; tau2 = p[2]; sX = p[3]; sY = p[4]; lsX = p[5]; lsY = p[6];
    }
    __device__ __forceinline__ void init(double &loc, double &scale, double &ls) const {
        loc = 0.0; scale = 1.0; ls = 0.0;                         // :672-673
    }
    __device__ __forceinline__ void trans(const StepK &, double xp, double &loc, double &scale,
                                          double &ls) const {
        loc = xp + tau0 - tau1 * fexp(tau2 * xp); scale = sX; ls = lsX;  // :675-678
    }
    __device__ __forceinline__ double obs_logpdf(const StepK &k, double, double x) const {
        return normal_logpdf_ls(k.y, x, sY, lsY);                 // :680-681
    }
    __device__ __forceinline__ void prop0(const StepK &, double &, double &, double &) const {}
    __device__ __forceinline__ void prop(const StepK &, double, double &, double &, double &) const {}
    __device__ __forceinline__ double logeta(const StepK &, double) const { return 0.0; }
};

// ---------------------------------------------------------------------------
// Feynman-Kac adaptors -- particles/state_space_models.py:299-438
// ---------------------------------------------------------------------------
template <int FK> struct FkTraits {
    static constexpr bool guided = (FK == SMCB_FK_GUIDED || FK == SMCB_FK_APF);
    static constexpr bool apf = (FK == SMCB_FK_APF || FK == SMCB_FK_AUXBOOT);
};

// M0 + logG(0, None, x): Bootstrap :326-327, 332-333; GuidedPF :374-375, 381-386
template <class M, int FK>
__device__ __forceinline__ void fk_init(const M &m, const StepK &k, double z, double &x,
                                        double &delta) {
    double loc, scale, ls;
    if (FkTraits<FK>::guided) {
        m.prop0(k, loc, scale, ls);
        x = loc + scale * z;
        double l0, s0, ls0;
        m.init(l0, s0, ls0);
        delta = normal_logpdf_ls(x, l0, s0, ls0) + m.obs_logpdf(k, x, x) -
                normal_logpdf_ls(x, loc, scale, ls);
    } else {
        m.init(loc, scale, ls);
        x = loc + scale * z;
        delta = m.obs_logpdf(k, x, x);
    }
}

// M(t, xp) + logG(t, xp, x): Bootstrap :329-333; GuidedPF :377-392
template <class M, int FK>
__device__ __forceinline__ void fk_move(const M &m, const StepK &k, double xp, double z, double &x,
                                        double &delta) {
    double loc, scale, ls;
    if (FkTraits<FK>::guided) {
        m.prop(k, xp, loc, scale, ls);
        x = loc + scale * z;
        double lt, st, lst;
        m.trans(k, xp, lt, st, lst);
        delta = normal_logpdf_ls(x, lt, st, lst) + m.obs_logpdf(k, xp, x) -
                normal_logpdf_ls(x, loc, scale, ls);
    } else {
        m.trans(k, xp, loc, scale, ls);
        x = loc + scale * z;
        delta = m.obs_logpdf(k, xp, x);
    }
}

}  // namespace smcb
