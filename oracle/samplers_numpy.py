"""numpy restatement of the waste-free adaptive-tempering SMC sampler of the reference
(TEST INFRASTRUCTURE, see oracle/__init__.py) -- BASELINE config 5 / SURVEY.md section 8 row a23.
Citations are relative to /root/reference.  Third-party pieces the reference itself calls on this
path: scipy.optimize.brentq (smc_samplers.py:893), numpy.cov / numpy.linalg (617-622).

Random numbers come from the legacy global numpy.random stream in the reference's order:
prior draws (N*P x d normals), then per SMC step one resampling uniform (systematic) and, per
Metropolis step, the (M x d) proposal normals followed by M acceptance uniforms.
"""
import numpy as np
from scipy import optimize

from . import smc_numpy as orc


class LogisticModel:
    """book/smc_samplers/logistic_reg.py:60-67: prior beta ~ MvNormal(scale=5, cov=I_d);
    logpyt(theta, t) = -logaddexp(0, -theta . data[t]) with `data` the sign-flipped predictors
    (datasets.py:286-292); loglik accumulates t = 0..T-1 in order (smc_samplers.py:263-284)."""

    def __init__(self, data, prior_scale=5.0):
        self.data = np.asarray(data, dtype=np.float64)
        self.d = self.data.shape[1]
        self.prior = orc.MvNormal(loc=0.0, scale=prior_scale, cov=np.eye(self.d))

    @property
    def T(self):
        return self.data.shape[0]

    def loglik(self, theta):
        ll = np.zeros(theta.shape[0])
        for s in range(self.T):
            ll += -np.logaddexp(0.0, -np.matmul(theta, self.data[s, :]))
        np.nan_to_num(ll, copy=False, nan=-np.inf)
        return ll


class ThetaParticles:
    """smc_samplers.py:401-500 (fields theta, lprior, llik, lpost + a shared dict)."""

    def __init__(self, shared=None, **fields):
        self.shared = {} if shared is None else shared
        self.__dict__.update(fields)

    @property
    def fields(self):
        return {k: v for k, v in self.__dict__.items() if k != "shared"}

    @property
    def N(self):
        return len(next(iter(self.fields.values())))

    def __getitem__(self, key):
        return ThetaParticles(shared=self.shared.copy(), **{k: v[key] for k, v in self.fields.items()})

    def copy(self):
        return ThetaParticles(shared=self.shared.copy(), **{k: v.copy() for k, v in self.fields.items()})

    @staticmethod
    def concatenate(*xs):
        f = {k: np.concatenate([getattr(x, k) for x in xs]) for k in xs[0].fields}
        return ThetaParticles(shared=xs[0].shared.copy(), **f)

    def copyto(self, src, where):
        for k, v in self.fields.items():
            wh = np.expand_dims(where, tuple(range(1, v.ndim)))
            np.copyto(v, getattr(src, k), where=wh)


def next_annealing_epn(epn, alpha, lw):
    """smc_samplers.py:876-895."""
    N = lw.shape[0]

    def f(e):
        ess = orc.essl(e * lw) if e > 0.0 else N
        return ess - alpha * N

    if f(1.0 - epn) < 0.0:
        return epn + optimize.brentq(f, 0.0, 1.0 - epn)
    return 1.0


def wmean_and_cov(W, x):
    """resampling.py:341-358."""
    m = np.average(x, weights=W, axis=0)
    cov = np.cov(x.T, aweights=W, ddof=0)
    return m, cov


class AdaptiveTemperingWF:
    """smc_samplers.py:714-769, 797-874, 897-936 with wastefree=True and the default
    MCMCSequenceWF(ArrayRandomWalk) move (596-629, 669-683)."""

    def __init__(self, model, len_chain=10, ESSrmin=0.5, max_iter=1000):
        self.model, self.len_chain, self.ESSrmin, self.max_iter = model, len_chain, ESSrmin, max_iter

    def target(self, epn):                                   # Tempering.current_target, 836-845
        def func(x):
            x.lprior = self.model.prior.logpdf(x.theta)
            x.llik = self.model.loglik(x.theta)
            x.lpost = x.lprior + epn * x.llik if epn > 0.0 else x.lprior.copy()
        return func

    def M0(self, N):                                         # 767-769, 847-852
        x0 = ThetaParticles(theta=self.model.prior.rvs(N * self.len_chain))
        x0.shared["exponents"] = [0.0]
        x0.shared["path_sampling"] = [0.0]
        self.target(0.0)(x0)
        return x0

    def calibrate(self, W, x):                               # ArrayRandomWalk.calibrate, 617-622
        N, d = x.theta.shape
        m, cov = wmean_and_cov(W, x.theta)
        x.shared["chol_cov"] = (2.38 / np.sqrt(d)) * np.linalg.cholesky(cov)

    def mh_step(self, x, target):                            # ArrayMetropolis.step, 601-611
        xprop = ThetaParticles(theta=np.empty_like(x.theta))
        L = x.shared["chol_cov"]
        xprop.theta[:, :] = x.theta + np.random.standard_normal(x.theta.shape) @ L.T   # 624-629
        target(xprop)
        lp_acc = xprop.lpost - x.lpost + 0.0
        pb_acc = np.exp(np.clip(lp_acc, None, 0.0))
        mean_acc = np.mean(pb_acc)
        accept = np.random.rand(x.N) < pb_acc
        x.copyto(xprop, where=accept)
        return mean_acc

    def move(self, x, target):                               # MCMCSequenceWF.__call__, 672-683
        xs, ars = [x], []
        for _ in range(self.len_chain - 1):
            x = x.copy()
            ars.append(self.mh_step(x, target))
            xs.append(x)
        xout = ThetaParticles.concatenate(*xs)
        xout.shared["acc_rates"] = x.shared.get("acc_rates", []) + [ars]
        return xout

    def logG(self, x):                                       # 929-933 + logG_tempering 826-830
        epn = x.shared["exponents"][-1]
        new_epn = next_annealing_epn(epn, self.ESSrmin, x.llik)
        x.shared["exponents"].append(new_epn)
        delta = new_epn - epn
        dl = delta * x.llik
        x.lpost += dl
        # update_path_sampling_est, 812-824
        grid_size = 10
        binwidth = delta / (grid_size - 1)
        new_ps = x.shared["path_sampling"][-1]
        for i, e in enumerate(np.linspace(0.0, delta, grid_size)):
            mult = 0.5 if i == 0 or i == grid_size - 1 else 1.0
            new_ps += mult * binwidth * np.average(x.llik, weights=orc.exp_and_normalise(e * x.llik))
            x.shared["path_sampling"].append(new_ps)
        return dl


def run_tempering(model, N, len_chain=10, ESSrmin=0.5, resampling="systematic", max_iter=1000):
    """particles.SMC(fk=AdaptiveTempering(model, wastefree=True, len_chain=P), N=N).run():
    the loop of core.py:369-383 with X a ThetaParticles of N*P particles, resampling M = N of
    them (core.py:329-331), always resampling (smc_samplers.py:917-919)."""
    fk = AdaptiveTemperingWF(model, len_chain, ESSrmin, max_iter)
    out = {"ESSs": [], "logLts": [], "exponents": None}
    X = fk.M0(N)
    wgts = orc.Weights().add(fk.logG(X))
    logLt = wgts.log_mean
    out["ESSs"].append(wgts.ESS)
    out["logLts"].append(logLt)
    t = 1
    while not (t >= max_iter or X.shared["exponents"][-1] >= 1.0):
        fk.calibrate(wgts.W, X)
        A = orc.resampling(resampling, wgts.W, M=N)
        Xp = X[A]
        X = fk.move(Xp, fk.target(Xp.shared["exponents"][-1]))
        wgts = orc.Weights().add(fk.logG(X))
        logLt += wgts.log_mean                              # rs_flag is always True: loglt = log_mean_w
        out["ESSs"].append(wgts.ESS)
        out["logLts"].append(logLt)
        t += 1
    out.update(logLt=logLt, X=X, W=wgts.W, exponents=list(X.shared["exponents"]), t=t,
               path_sampling=X.shared["path_sampling"][-1])
    return out


def synthetic_logistic(n_data=1000, d=20, seed=0):
    """predictors N(0, 0.5^2) + intercept (datasets.py:153-181),
    beta* ~ N(0, 1), response sign folded into the predictors (datasets.py:286-292)."""
    r = np.random.RandomState(seed)
    preds = r.randn(n_data, d - 1)
    preds = 0.5 * (preds - preds.mean(axis=0)) / preds.std(axis=0)
    X = np.empty((n_data, d))
    X[:, 0] = 1.0
    X[:, 1:] = preds
    beta = r.randn(d)
    p = 1.0 / (1.0 + np.exp(-X @ beta))
    resp = 2.0 * (r.rand(n_data) < p) - 1.0
    return X * resp[:, None]
