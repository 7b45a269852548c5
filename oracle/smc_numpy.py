"""numpy restatement of the reference SMC step loop (TEST INFRASTRUCTURE, see
``oracle/__init__.py``).  Citations are relative to ``/root/reference``.

Two noise sources are supported:

* ``GlobalStream`` -- draws from the legacy global ``numpy.random`` stream in
  exactly the order the reference does, so a run
  with the same ``np.random.seed`` reproduces the reference bit-for-bit.
* ``InjectedNoise`` -- standard normals / uniforms supplied by the caller, so
  the CUDA path can be checked on identical inputs, independent of RNG choice.
"""
import math

import numpy as np

HALFLOG2PI = 0.5 * np.log(2.0 * np.pi)  # particles/distributions.py:212


# ----------------------------------------------------------------------------
# weights algebra -- particles/resampling.py:138-317
# ----------------------------------------------------------------------------
def exp_and_normalise(lw):
    """particles/resampling.py:138-163."""
    w = np.exp(lw - lw.max())
    return w / w.sum()


def essl(lw):
    """particles/resampling.py:166-188 (code returns (sum w)^2 / sum w^2)."""
    w = np.exp(lw - lw.max())
    return (w.sum()) ** 2 / np.sum(w ** 2)


def log_sum_exp(v):
    """particles/resampling.py:247-270."""
    m = v.max()
    return m + np.log(np.sum(np.exp(v - m)))


def log_sum_exp_ab(a, b):
    """particles/resampling.py:273-288."""
    if a > b:
        return a + np.log1p(np.exp(b - a))
    return b + np.log1p(np.exp(a - b))


def log_mean_exp(v, W=None):
    """particles/resampling.py:291-317."""
    m = v.max()
    V = np.exp(v - m)
    if W is None:
        return m + np.log(np.mean(V))
    return m + np.log(np.average(V, weights=W))


def wmean_and_var(W, x):
    """particles/resampling.py:320-338."""
    m = np.average(x, weights=W, axis=0)
    m2 = np.average(x ** 2, weights=W, axis=0)
    return {"mean": m, "var": m2 - m ** 2}


class Weights:
    """particles/resampling.py:191-244.  NaN -> -inf is written into the
    caller's array (line 220); ``Weights()`` has no W/ESS/log_mean."""

    def __init__(self, lw=None):
        self.lw = lw
        if lw is not None:
            self.lw[np.isnan(self.lw)] = -np.inf
            m = self.lw.max()
            w = np.exp(self.lw - m)
            s = w.sum()
            self.log_mean = m + np.log(s / self.N)
            self.W = w / s
            self.ESS = 1.0 / np.sum(self.W ** 2)

    @property
    def N(self):
        return 0 if self.lw is None else self.lw.shape[0]

    def add(self, delta):
        if self.lw is None:
            return Weights(lw=delta)
        return Weights(lw=self.lw + delta)


# ----------------------------------------------------------------------------
# resampling -- particles/resampling.py:445-627
# ----------------------------------------------------------------------------
def inverse_cdf(su, W):
    """particles/resampling.py:484-509 (numba loop).  The sequential loop
    ``while su[n] > s: j += 1; s += W[j]`` is bit-equal to
    ``searchsorted(cumsum(W), su, 'left')`` (sequential fp64 running sum);
    SURVEY.md section 9 item 7 verified it for N up to 1e7 and
    tests/test_oracle_golden.py re-checks it against the reference's output.
    The reference has no bounds check (reads past W when su > sum W by
    round-off); we clip to N-1, which is what a GPU kernel must do as well."""
    from . import cext

    if cext.available():
        return cext.inverse_cdf(su, W)
    A = np.searchsorted(np.cumsum(W), su, side="left")
    return np.minimum(A, W.shape[0] - 1).astype(np.int64)


def uniform_spacings(N, u=None):
    """particles/resampling.py:512-537; ``u`` = the N+1 uniforms (injected) or
    None to draw ``random.rand(N + 1)`` from the global stream."""
    if u is None:
        u = np.random.rand(N + 1)
    z = np.cumsum(-np.log(u))
    return z[:-1] / z[-1]


def su_systematic(M, u):
    """particles/resampling.py:609: ``(rand(1) + arange(M)) / M``."""
    return (u + np.arange(M)) / M


def su_stratified(M, u):
    """particles/resampling.py:602: ``(rand(M) + arange(M)) / M``."""
    return (u + np.arange(M)) / M


def multinomial(W, M, u=None):
    """particles/resampling.py:540-558."""
    return inverse_cdf(uniform_spacings(M, u), W)


def stratified(W, M, u=None):
    """particles/resampling.py:599-603."""
    if u is None:
        u = np.random.rand(M)
    return inverse_cdf(su_stratified(M, u), W)


def systematic(W, M, u=None):
    """particles/resampling.py:606-610."""
    if u is None:
        u = np.random.rand(1)
    return inverse_cdf(su_systematic(M, u), W)


def residual(W, M, u=None):
    """particles/resampling.py:613-627.  ``u`` = the ``M - sip + 1`` uniforms
    of the multinomial stage (only its first ``M - sip + 1`` entries are used)."""
    N = W.shape[0]
    A = np.empty(M, dtype=np.int64)
    MW = M * W
    intpart = np.floor(MW).astype(np.int64)
    sip = np.sum(intpart)
    res = MW - intpart
    sres = M - sip
    A[:sip] = np.arange(N).repeat(intpart)
    if sres > 0:
        uu = None if u is None else u[: sres + 1]
        A[sip:] = multinomial(res / sres, sres, uu)
    return A


def ssp(W, M, u=None):
    """particles/resampling.py:630-677 (Srinivasan sampling process).  State machine over two
    active particles (a, b): at step k the fractional parts are pushed towards each other until one
    of them reaches 0 (that particle retires with its integer part) or 1 (it retires with one more
    offspring), and particle k + 2 enters.  ``u`` = the N - 1 uniforms of the process."""
    N = W.shape[0]
    MW = M * W
    kids = np.floor(MW).astype(np.int64)
    frac = MW - kids
    if u is None:
        u = np.random.rand(N - 1)
    a, b = 0, 1
    for k in range(N - 1):
        up_a = min(frac[b], 1.0 - frac[a])      # mass a can take from b
        up_b = min(frac[a], 1.0 - frac[b])      # mass b can take from a
        tot = up_a + up_b
        p_swap = up_a / tot if tot > 0.0 else 0.0
        step = up_a
        if u[k] < p_swap:                       # relabel so that `a` is always the one that grows
            a, b = b, a
            step = up_b
        if frac[b] < 1.0 - frac[a]:             # b is emptied and retires
            frac[a] += step
            b = k + 2
        else:                                   # a is filled: one more offspring, a retires
            frac[b] -= step
            kids[a] += 1
            a = k + 2
    if N > 1 and kids.sum() == M - 1:           # accumulated round-off may have lost one particle
        last = a if b == N else b
        if frac[last] > 0.99:
            kids[last] += 1
    if kids.sum() != M:
        raise ValueError("ssp resampling: wrong size for output")
    return np.arange(N).repeat(kids)


def killing(W, M, u=None, u_multinomial=None):
    """particles/resampling.py:680-697: keep particle i with probability W[i] / max W, otherwise
    replace it by a multinomial draw; M = N only.  Draw order: rand(N), then the multinomial's."""
    N = W.shape[0]
    if M != N:
        raise ValueError("killing resampling defined only for M=N")
    if u is None:
        u = np.random.rand(N)
    dead = u * W.max() >= W
    A = np.arange(N)
    A[dead] = multinomial(W, int(dead.sum()), u_multinomial)
    return A


def multinomial_once(W, u=None):
    """particles/resampling.py:573-597."""
    if u is None:
        u = np.random.rand()
    return int(np.searchsorted(np.cumsum(W), u))


RS_FUNCS = {
    "multinomial": multinomial,
    "stratified": stratified,
    "systematic": systematic,
    "residual": residual,
    "ssp": ssp,
    "killing": killing,
}


def n_uniforms(scheme, M):
    """How many uniforms a scheme consumes at most."""
    return {"systematic": 1, "stratified": M, "multinomial": M + 1,
            "residual": M + 1}[scheme]


def resampling(scheme, W, M=None, u=None):
    """particles/resampling.py:464-481 (M defaults to N; unknown -> ValueError)."""
    M = W.shape[0] if M is None else M
    try:
        f = RS_FUNCS[scheme]
    except KeyError:
        raise ValueError(f"{scheme} is not a valid resampling scheme")
    return f(W, M, u)


# ----------------------------------------------------------------------------
# distributions -- particles/distributions.py
# ----------------------------------------------------------------------------
def normal_logpdf(x, loc, scale):
    """particles/distributions.py:273-274 -> scipy.stats.norm.logpdf; the
    operation order below is bit-identical to scipy for finite inputs
    (SURVEY.md section 9 item 2)."""
    z = (x - loc) / scale
    return -z * z / 2.0 - HALFLOG2PI - np.log(scale)


class Normal:
    """particles/distributions.py:267-285."""
    dim = 1

    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = loc, scale

    def rvs(self, size, z=None):
        if z is None:  # distributions.py:270-271: random.normal(loc, scale, size)
            return np.random.normal(loc=self.loc, scale=self.scale, size=size)
        return self.loc + self.scale * z  # legacy normal == loc + scale*gauss

    def logpdf(self, x):
        return normal_logpdf(x, self.loc, self.scale)


class Poisson:
    """particles/distributions.py:519-532; logpmf as scipy.stats.poisson evaluates it:
    xlogy(k, mu) - gammaln(k + 1) - mu."""
    dim = 1

    def __init__(self, rate=1.0):
        self.rate = rate

    def rvs(self, size, z=None):
        return np.random.poisson(self.rate, size=size)

    def logpdf(self, x):
        k = np.asarray(x, dtype=np.float64)
        lg = np.vectorize(math.lgamma)(k + 1.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            xl = np.where(k == 0, 0.0, k * np.log(self.rate))
        return xl - lg - self.rate


class Dirac:
    """particles/distributions.py:454-472."""
    dim = 1

    def __init__(self, loc=0.0):
        self.loc = loc

    def rvs(self, size, z=None):
        if isinstance(self.loc, np.ndarray):
            return self.loc.copy()
        return np.full(size, self.loc)

    def logpdf(self, x):
        return np.where(x == self.loc, 0.0, -np.inf)


class IndepProd:
    """particles/distributions.py:1066-1109.  Component i draws all its N
    variates before component i+1 (1105-1106); Dirac draws nothing."""

    def __init__(self, *dists):
        self.dists = dists
        self.dim = len(dists)

    def rvs(self, size, z=None):
        cols, k = [], 0
        for d in self.dists:
            if isinstance(d, Dirac):
                cols.append(d.rvs(size))
            else:
                cols.append(d.rvs(size, None if z is None else z[:, k]))
                k += 1
        return np.stack(cols, axis=1)

    def logpdf(self, x):
        return sum([d.logpdf(x[..., i]) for i, d in enumerate(self.dists)])

    @property
    def n_noise(self):
        return sum(0 if isinstance(d, Dirac) else 1 for d in self.dists)


def solve_lower(L, b):
    """forward substitution, b is (d, N) -- scipy.linalg.solve_triangular
    (distributions.py:952) restated so the oracle needs numpy only."""
    d = L.shape[0]
    z = np.empty_like(b)
    for i in range(d):
        acc = b[i].copy()
        for j in range(i):
            acc -= L[i, j] * z[j]
        z[i] = acc / L[i, i]
    return z


class MvNormal:
    """particles/distributions.py:888-982."""

    def __init__(self, loc=0.0, scale=1.0, cov=None):
        self.loc, self.scale = loc, scale
        self.cov = np.eye(loc.shape[-1]) if cov is None else cov
        self.L = np.linalg.cholesky(self.cov)  # distributions.py:937

    @property
    def dim(self):
        return self.cov.shape[-1]

    def linear_transform(self, z):  # distributions.py:946-947
        return self.loc + self.scale * np.dot(z, self.L.T)

    def logpdf(self, x):  # distributions.py:949-959
        halflogdetcor = np.sum(np.log(np.diag(self.L)))
        xc = (x - self.loc) / self.scale
        z = solve_lower(self.L, np.transpose(np.atleast_2d(xc)))
        if np.asarray(self.scale).ndim == 0:
            logdet = self.dim * np.log(self.scale)
        else:
            logdet = np.sum(np.log(self.scale), axis=-1)
        logdet = logdet + halflogdetcor
        out = -0.5 * np.sum(z * z, axis=0) - logdet - self.dim * HALFLOG2PI
        return out if np.ndim(xc) > 1 else out[0]

    def rvs(self, size, z=None):  # distributions.py:961-969
        if z is None:
            # stats.norm.rvs(size=(N, d)) == global-stream standard normals, row-major
            z = np.random.standard_normal(size=(size, self.dim))
        return self.linear_transform(z)


# ----------------------------------------------------------------------------
# state-space models -- particles/state_space_models.py, particles/kalman.py
# ----------------------------------------------------------------------------
class SSM:
    """particles/state_space_models.py:172-296 (only what the filter needs)."""
    default_params = {}

    def __init__(self, **kwargs):
        self.__dict__.update(self.default_params)
        self.__dict__.update(kwargs)

    def simulate(self, T):
        """state_space_models.py:272-296; global stream, same draw order."""
        x = []
        for t in range(T):
            law_x = self.PX0() if t == 0 else self.PX(t, x[-1])
            x.append(law_x.rvs(size=1))
        lag_x = [None] + x[:-1]
        y = [self.PY(t, xp, xx).rvs(size=1) for t, (xp, xx) in enumerate(zip(lag_x, x))]
        return x, y


class StochVol(SSM):
    """particles/state_space_models.py:446-498."""
    default_params = {"mu": -1.02, "rho": 0.9702, "sigma": 0.178}

    def sig0(self):
        return self.sigma / np.sqrt(1.0 - self.rho ** 2)

    def PX0(self):
        return Normal(loc=self.mu, scale=self.sig0())

    def EXt(self, xp):
        return (1.0 - self.rho) * self.mu + self.rho * xp

    def PX(self, t, xp):
        return Normal(loc=self.EXt(xp), scale=self.sigma)

    def PY(self, t, xp, x):
        return Normal(loc=0.0, scale=np.exp(0.5 * x))

    def _xhat(self, xst, sig, yt):
        return xst + 0.5 * sig ** 2 * (yt ** 2 * np.exp(-xst) - 1.0)

    def proposal0(self, data):
        return Normal(loc=self._xhat(0.0, self.sig0(), data[0]), scale=self.sig0())

    def proposal(self, t, xp, data):
        return Normal(loc=self._xhat(self.EXt(xp), self.sigma, data[t]), scale=self.sigma)

    def logeta(self, t, x, data):
        xst = self.EXt(x)
        xstmmu = xst - self.mu
        xhat = self._xhat(xst, self.sigma, data[t + 1])
        xhatmmu = xhat - self.mu
        return 0.5 / self.sigma ** 2 * (xhatmmu ** 2 - xstmmu ** 2) - 0.5 * data[
            t + 1] ** 2 * np.exp(-xst) * (1.0 + xstmmu)


class StochVolLeverage(StochVol):
    """particles/state_space_models.py:501-543."""
    default_params = {"mu": -1.02, "rho": 0.9702, "sigma": 0.178, "phi": 0.0}

    def PY(self, t, xp, x):
        if t == 0:
            u = (x - self.mu) / self.sig0()
        else:
            u = (x - self.EXt(xp)) / self.sigma
        std_x = np.exp(0.5 * x)
        return Normal(loc=std_x * self.phi * u, scale=std_x * np.sqrt(1.0 - self.phi ** 2))


class DiscreteCox(SSM):
    """particles/state_space_models.py:611-630."""
    default_params = {"mu": 0.0, "sigma": 1.0, "phi": 0.95}

    def PX0(self):
        return Normal(loc=self.mu, scale=self.sigma / np.sqrt(1.0 - self.phi ** 2))

    def PX(self, t, xp):
        return Normal(loc=self.mu + self.phi * (xp - self.mu), scale=self.sigma)

    def PY(self, t, xp, x):
        return Poisson(rate=np.exp(x))


class LinearGauss(SSM):
    """particles/kalman.py:397-452."""
    default_params = {"sigmaY": 0.2, "rho": 0.9, "sigmaX": 1.0, "sigma0": None}

    def __init__(self, **kwargs):
        SSM.__init__(self, **kwargs)
        if self.sigma0 is None:
            self.sigma0 = self.sigmaX / np.sqrt(1.0 - self.rho ** 2)

    def PX0(self):
        return Normal(scale=self.sigma0)

    def PX(self, t, xp):
        return Normal(loc=self.rho * xp, scale=self.sigmaX)

    def PY(self, t, xp, x):
        return Normal(loc=x, scale=self.sigmaY)

    def proposal0(self, data):
        sig2post = 1.0 / (1.0 / self.sigma0 ** 2 + 1.0 / self.sigmaY ** 2)
        mupost = sig2post * (data[0] / self.sigmaY ** 2)
        return Normal(loc=mupost, scale=np.sqrt(sig2post))

    def proposal(self, t, xp, data):
        sig2post = 1.0 / (1.0 / self.sigmaX ** 2 + 1.0 / self.sigmaY ** 2)
        mupost = sig2post * (self.rho * xp / self.sigmaX ** 2 + data[t] / self.sigmaY ** 2)
        return Normal(loc=mupost, scale=np.sqrt(sig2post))

    def logeta(self, t, x, data):
        law = Normal(loc=self.rho * x, scale=np.sqrt(self.sigmaX ** 2 + self.sigmaY ** 2))
        return law.logpdf(data[t + 1])

    def kalman_loglik(self, data):
        """Exact log-likelihood, scalar Kalman recursion -- kalman.py:169-229,
        483-505 specialised to dx = dy = 1."""
        ll = []
        pm, pc = 0.0, self.sigma0 ** 2
        for t, yt in enumerate(data):
            yt = float(np.asarray(yt).reshape(-1)[0])
            if t > 0:
                pm, pc = self.rho * fm, self.rho ** 2 * fc + self.sigmaX ** 2
            dc = pc + self.sigmaY ** 2
            ll.append(float(normal_logpdf(yt, pm, np.sqrt(dc))))
            gain = pc / dc
            fm, fc = pm + gain * (yt - pm), pc - gain * pc
        return np.array(ll)


class ToySSM(SSM):
    """README.md:58-66 of the reference."""
    default_params = {"sigma": 0.2}

    def PX0(self):
        return Normal()

    def PX(self, t, xp):
        return Normal(loc=xp)

    def PY(self, t, xp, x):
        return Normal(loc=x, scale=self.sigma)


class Gordon_etal(SSM):
    """particles/state_space_models.py:546-577."""
    default_params = {"a": 0.05, "b": 0.5, "c": 25.0, "d": 8.0, "e": 1.2,
                      "sigmaX": 3.162278}

    def PX0(self):
        return Normal(scale=2.0)

    def PX(self, t, xp):
        return Normal(loc=self.b * xp + self.c * xp / (1.0 + xp ** 2)
                      + self.d * np.cos(self.e * (t - 1)), scale=self.sigmaX)

    def PY(self, t, xp, x):
        return Normal(loc=self.a * x ** 2)


class ThetaLogistic(SSM):
    """particles/state_space_models.py:657-689 (PX0/PX/PY only)."""
    default_params = {"tau0": 0.15, "tau1": 0.12, "tau2": 0.1, "sigmaX": 0.47,
                      "sigmaY": 0.39}

    def PX0(self):
        return Normal(loc=0.0, scale=1.0)

    def PX(self, t, xp):
        return Normal(loc=xp + self.tau0 - self.tau1 * np.exp(self.tau2 * xp),
                      scale=self.sigmaX)

    def PY(self, t, xp, x):
        return Normal(loc=x, scale=self.sigmaY)


class BearingsOnly(SSM):
    """particles/state_space_models.py:580-608."""
    default_params = {"sigmaX": 2.0e-4, "sigmaY": 1e-3,
                      "x0": np.array([3e-3, -3e-3, 1.0, 1.0])}

    def PX0(self):
        return IndepProd(Normal(loc=self.x0[0], scale=self.sigmaX),
                         Normal(loc=self.x0[1], scale=self.sigmaX),
                         Dirac(loc=self.x0[2]), Dirac(loc=self.x0[3]))

    def PX(self, t, xp):
        return IndepProd(Normal(loc=xp[:, 0], scale=self.sigmaX),
                         Normal(loc=xp[:, 1], scale=self.sigmaX),
                         Dirac(loc=xp[:, 0] + xp[:, 2]),
                         Dirac(loc=xp[:, 1] + xp[:, 3]))

    def PY(self, t, xp, x):
        angle = np.arctan(x[:, 3] / x[:, 2])
        angle[x[:, 2] < 0.0] += np.pi
        return Normal(loc=angle, scale=self.sigmaY)


class MVStochVol(SSM):
    """particles/state_space_models.py:633-654 (mu, covX, corY, F have no defaults in the reference: pass them)."""
    default_params = {"mu": 0.0, "covX": None, "corY": None, "F": None}

    def offset(self):
        return self.mu - np.dot(self.F, self.mu)

    def PX0(self):
        return MvNormal(loc=self.mu, cov=self.covX)

    def PX(self, t, xp):
        return MvNormal(loc=np.dot(xp, self.F.T) + self.offset(), cov=self.covX)

    def PY(self, t, xp, x):
        return MvNormal(scale=np.exp(0.5 * x), cov=self.corY)


class MVLinearGauss(SSM):
    """particles/kalman.py:296-361."""

    def __init__(self, F=None, G=None, covX=None, covY=None, mu0=None, cov0=None):
        self.covX, self.covY = np.atleast_2d(covX), np.atleast_2d(covY)
        self.dx, self.dy = self.covX.shape[0], self.covY.shape[0]
        self.mu0 = np.zeros(self.dx) if mu0 is None else mu0
        self.cov0 = self.covX if cov0 is None else np.atleast_2d(cov0)
        self.F = np.eye(self.dx) if F is None else np.atleast_2d(F)
        self.G = np.eye(self.dy, self.dx) if G is None else np.atleast_2d(G)

    def PX0(self):
        return MvNormal(loc=self.mu0, cov=self.cov0)

    def PX(self, t, xp):
        return MvNormal(loc=np.dot(xp, self.F.T), cov=self.covX)

    def PY(self, t, xp, x):
        return MvNormal(loc=np.dot(x, self.G.T), cov=self.covY)

    # Kalman algebra: kalman.py:169-229
    def _filter_step(self, pred_mean, pred_cov, yt):
        G, covY = self.G, self.covY
        dpm = np.matmul(pred_mean, G.T)
        dpc = G @ pred_cov @ G.T + covY
        if covY.shape[0] == 1:
            logpyt = normal_logpdf(yt, dpm, np.sqrt(dpc))
        else:
            logpyt = MvNormal(loc=dpm, cov=dpc).logpdf(yt)
        resid = yt - dpm
        gain = np.linalg.solve(dpc, (pred_cov @ G.T).T).T
        fmean = pred_mean + np.matmul(resid, gain.T)
        fcov = pred_cov - gain @ G @ pred_cov
        return fmean, fcov, logpyt

    def proposal0(self, data):
        fm, fc, _ = self._filter_step(self.mu0, self.cov0, data[0])
        return MvNormal(loc=fm, cov=fc)

    def proposal(self, t, xp, data):
        fm, fc, _ = self._filter_step(np.matmul(xp, self.F.T), self.covX, data[t])
        return MvNormal(loc=fm, cov=fc)

    def logeta(self, t, x, data):
        _, _, lp = self._filter_step(np.matmul(x, self.F.T), self.covX, data[t + 1])
        return lp

    def kalman_loglik(self, data):
        """kalman.py:483-505."""
        ll = []
        for t, yt in enumerate(data):
            if t == 0:
                pm, pc = self.mu0, self.cov0
            else:
                pm, pc = np.matmul(fm, self.F.T), self.F @ fc @ self.F.T + self.covX
            fm, fc, lp = self._filter_step(pm, pc, np.asarray(yt))
            ll.append(float(np.asarray(lp).reshape(-1)[0]))
        return np.array(ll)


class MVLinearGauss_Guarniero_etal(MVLinearGauss):
    """particles/kalman.py:364-394."""

    def __init__(self, alpha=0.4, dx=2):
        F = np.empty((dx, dx))
        for i in range(dx):
            for j in range(dx):
                F[i, j] = alpha ** (1 + abs(i - j))
        MVLinearGauss.__init__(self, F=F, G=np.eye(dx), covX=np.eye(dx), covY=np.eye(dx))


# ----------------------------------------------------------------------------
# Feynman-Kac adaptors -- particles/state_space_models.py:299-438
# ----------------------------------------------------------------------------
class Bootstrap:
    isAPF = False

    def __init__(self, ssm, data):
        self.ssm, self.data = ssm, data

    @property
    def T(self):
        return 0 if self.data is None else len(self.data)

    def M0(self, N, z=None):
        return self.ssm.PX0().rvs(N, z)

    def M(self, t, xp, z=None):
        return self.ssm.PX(t, xp).rvs(xp.shape[0], z)

    def logG(self, t, xp, x):
        return self.ssm.PY(t, xp, x).logpdf(self.data[t])


class GuidedPF(Bootstrap):
    def M0(self, N, z=None):
        return self.ssm.proposal0(self.data).rvs(N, z)

    def M(self, t, xp, z=None):
        return self.ssm.proposal(t, xp, self.data).rvs(xp.shape[0], z)

    def logG(self, t, xp, x):
        if t == 0:
            return (self.ssm.PX0().logpdf(x)
                    + self.ssm.PY(0, xp, x).logpdf(self.data[0])
                    - self.ssm.proposal0(self.data).logpdf(x))
        return (self.ssm.PX(t, xp).logpdf(x)
                + self.ssm.PY(t, xp, x).logpdf(self.data[t])
                - self.ssm.proposal(t, xp, self.data).logpdf(x))


class AuxiliaryPF(GuidedPF):
    isAPF = True

    def logeta(self, t, x):
        return self.ssm.logeta(t, x, self.data)


class AuxiliaryBootstrap(Bootstrap):
    isAPF = True

    def logeta(self, t, x):
        return self.ssm.logeta(t, x, self.data)


# ----------------------------------------------------------------------------
# noise sources
# ----------------------------------------------------------------------------
class GlobalStream:
    """Legacy global numpy.random stream, consumed as the reference does."""

    def normals(self, t, shape):
        return None  # -> dists draw from the global stream themselves

    def uniforms(self, t, scheme, M):
        return None


class InjectedNoise:
    """z[t]: (N,) or (N, k) standard normals for step t; u[t]: uniforms for the
    resampling of step t (shape per ``n_uniforms``)."""

    def __init__(self, z, u):
        self.z, self.u = z, u

    def normals(self, t, shape):
        return self.z[t]

    def uniforms(self, t, scheme, M):
        return self.u[t]


# ----------------------------------------------------------------------------
# the step loop -- particles/core.py:299-383
# ----------------------------------------------------------------------------
class SMC:
    """Restatement of ``particles.core.SMC`` (non-QMC branch, no history)."""

    def __init__(self, fk, N=100, resampling="systematic", ESSrmin=0.5, noise=None,
                 keep=False):
        self.fk, self.N = fk, N
        self.resampling, self.ESSrmin = resampling, ESSrmin
        self.noise = GlobalStream() if noise is None else noise
        self.t, self.rs_flag, self.logLt = 0, False, 0.0
        self.wgts, self.aux = Weights(), None
        self.X = self.Xp = self.A = None
        self.ESSs, self.logLts, self.rs_flags = [], [], []
        self.keep = keep
        self.trace = []

    @property
    def W(self):
        return self.wgts.W

    def reset_weights(self):  # core.py:299-305
        if self.fk.isAPF:
            lw = log_mean_exp(self.logetat, W=self.W) - self.logetat[self.A]
            self.wgts = Weights(lw=lw)
        else:
            self.wgts = Weights()

    def setup_auxiliary_weights(self):  # core.py:307-313
        if self.fk.isAPF:
            self.logetat = self.fk.logeta(self.t - 1, self.X)
            self.aux = self.wgts.add(self.logetat)
        else:
            self.aux = self.wgts

    def resample_move(self):  # core.py:326-337
        self.rs_flag = bool(self.aux.ESS < self.N * self.ESSrmin)  # core.py:183
        if self.rs_flag:
            u = self.noise.uniforms(self.t, self.resampling, self.N)
            self.A = resampling(self.resampling, self.aux.W, M=self.N, u=u)
            self.Xp = self.X[self.A]
            self.reset_weights()
        else:
            self.A = np.arange(self.N)
            self.Xp = self.X
        self.X = self.fk.M(self.t, self.Xp, self.noise.normals(self.t, None))

    def compute_summaries(self):  # core.py:351-367
        if self.t > 0:
            prec_log_mean_w = self.log_mean_w
        self.log_mean_w = self.wgts.log_mean
        if self.t == 0 or self.rs_flag:
            self.loglt = self.log_mean_w
        else:
            self.loglt = self.log_mean_w - prec_log_mean_w
        self.logLt += self.loglt
        self.ESSs.append(self.wgts.ESS)
        self.logLts.append(self.logLt)
        self.rs_flags.append(self.rs_flag)
        if self.keep:
            self.trace.append({"X": self.X, "A": self.A, "lw": self.wgts.lw,
                               "W": self.wgts.W})

    def step(self):  # core.py:369-383
        if self.t >= self.fk.T:
            raise StopIteration
        if self.t == 0:
            self.X = self.fk.M0(self.N, self.noise.normals(0, None))
        else:
            self.setup_auxiliary_weights()
            self.resample_move()
        self.wgts = self.wgts.add(self.fk.logG(self.t, self.Xp, self.X))
        self.compute_summaries()
        self.t += 1

    def run(self, nsteps=None):
        n = 0
        while self.t < self.fk.T and (nsteps is None or n < nsteps):
            self.step()
            n += 1
        return self


def config2_data(T=1000, seed=1):
    """``np.random.seed(1); StochVol().simulate(T)``."""
    state = np.random.get_state()
    np.random.seed(seed)
    _, ys = StochVol().simulate(T)
    np.random.set_state(state)
    return np.array([float(y[0]) for y in ys])
