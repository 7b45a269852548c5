"""SMC samplers on the device: tempering / adaptive tempering with standard or waste-free MCMC
moves -- the part of ``particles/smc_samplers.py`` that BASELINE config 5 exercises (file:line
cited per class).  They are Feynman-Kac models for ``particles_b200.SMC`` (plugin path):

    model = LogisticRegression(data=flipped_predictors, prior_scale=5.)
    fk = AdaptiveTempering(model=model, wastefree=True, len_chain=100)
    pf = particles_b200.SMC(fk=fk, N=10_000, ESSrmin=1.)      # N * len_chain particles
    pf.run();  pf.logLt;  pf.X.theta;  pf.X.shared["exponents"]

Particles are a ``ThetaParticles`` of CUDA tensors; the per-particle work (tempered target of
the model, random-walk proposal, Metropolis accept/copy, resampling, weights) runs in libsmcb
kernels; the O(d^2) calibration of the proposal (weighted covariance + Cholesky) and the scalar
root-find for the next exponent (scipy.optimize.brentq, as in the reference) run on the host.
"""
import ctypes as C

import numpy as np
import torch
from scipy import optimize

from . import _lib
from . import resampling as rs
from .core import FeynmanKac
from .device import as_device, context, empty, ptr


class ThetaParticles:
    """smc_samplers.py:401-500: N particles packed as named CUDA tensors (``theta`` (N, d),
    ``lprior``, ``llik``, ``lpost`` (N,)) plus a ``shared`` dict; fancy indexing by an int64
    ancestor tensor returns a new object (gather kernels), as the reference's class does."""

    def __init__(self, shared=None, **fields):
        self.shared = {} if shared is None else shared
        self.__dict__.update(fields)

    @property
    def dict_fields(self):
        return {k: v for k, v in self.__dict__.items() if k != "shared"}

    @property
    def N(self):
        return len(next(iter(self.dict_fields.values())))

    def __getitem__(self, key):
        if not (isinstance(key, torch.Tensor) and key.dtype == torch.int64):
            return self.__class__(shared=self.shared.copy(),
                                  **{k: v[key] for k, v in self.dict_fields.items()})
        ctx = context(key.device)
        out = {}
        for k, v in self.dict_fields.items():
            d = 1 if v.ndim == 1 else v.shape[1]
            o = torch.empty((key.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            _lib.check(ctx.lib.smcb_gather_rows(ctx.handle, ptr(v), v.shape[0], ptr(key), key.shape[0], d,
                                                ptr(o)))
            out[k] = o
        return self.__class__(shared=self.shared.copy(), **out)

    def copy(self):
        return self.__class__(shared=self.shared.copy(),
                              **{k: v.clone() for k, v in self.dict_fields.items()})

    @classmethod
    def concatenate(cls, *xs):
        fields = {k: torch.cat([getattr(x, k) for x in xs]) for k in xs[0].dict_fields}
        return cls(shared=xs[0].shared.copy(), **fields)


class LogisticRegression:
    """Bayesian logistic regression static model (book/smc_samplers/logistic_reg.py:60-67):
    ``data`` = (n_data, d) predictors with the response sign folded in (datasets.py:286-292),
    prior beta ~ MvNormal(scale=prior_scale, cov=I_d).  ``target(x, epn)`` fills
    ``x.lprior / x.llik / x.lpost`` (Tempering.current_target, smc_samplers.py:836-845)."""

    def __init__(self, data=None, prior_scale=5.0):
        self.data_host = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
        self.data = as_device(self.data_host)
        self.prior_scale = float(prior_scale)
        self.d = self.data_host.shape[1]

    @property
    def T(self):
        return self.data_host.shape[0]

    def prior_rvs(self, size):
        ctx = context()
        z = empty(size * self.d)
        _lib.check(ctx.lib.smcb_standard_normal(ctx.handle, ptr(z), size * self.d))
        return (self.prior_scale * z).reshape(size, self.d)     # loc + scale * (z @ I)

    def wf_move(self, x, epn, P, noise=None):
        """Fused waste-free move (smcb_logistic_wf_move): x = the M resampled particles with their
        lprior / llik / lpost at exponent ``epn`` and ``shared['chol_cov']``; returns P*M particles."""
        ctx = context()
        M, d = x.theta.shape
        out = x.__class__(shared=x.shared.copy(), theta=empty((P * M, d)), lprior=empty(P * M),
                          llik=empty(P * M), lpost=empty(P * M))
        pb = empty((P - 1, M))
        z = u = None
        if noise is not None:
            z, u = as_device(noise[0]), as_device(noise[1])
        _lib.check(ctx.lib.smcb_logistic_wf_move(
            ctx.handle, M, d, P, ptr(x.theta), ptr(x.lprior), ptr(x.llik), ptr(x.lpost), ptr(self.data), self.T,
            self.prior_scale, float(epn), ptr(x.shared["chol_cov"]), ptr(z), ptr(u), ptr(out.theta),
            ptr(out.lprior), ptr(out.llik), ptr(out.lpost), ptr(pb)))
        out.shared["acc_rates"] = x.shared.get("acc_rates", []) + [pb.mean(dim=1)]
        return out

    def target(self, x, epn):
        ctx = context()
        n = x.theta.shape[0]
        x.lprior, x.llik, x.lpost = empty(n), empty(n), empty(n)
        _lib.check(ctx.lib.smcb_logistic_target(ctx.handle, ptr(x.theta), n, self.d, ptr(self.data), self.T,
                                                self.prior_scale, float(epn), ptr(x.lprior), ptr(x.llik),
                                                ptr(x.lpost)))


class ArrayRandomWalk:
    """Gaussian random-walk Metropolis, smc_samplers.py:596-629."""

    def calibrate(self, W, x):
        """smc_samplers.py:617-622: L = 2.38 / sqrt(d) * chol(wcov(W, theta)) -- weighted mean, covariance and the
        Cholesky factor all on the device (smcb_rw_calibrate), nothing comes back to the host."""
        theta = x.theta
        n, d = theta.shape
        ctx = context()
        L = empty(d * d).reshape(d, d)
        _lib.check(ctx.lib.smcb_rw_calibrate(ctx.handle, ptr(as_device(W)), ptr(theta), n, d, 2.38 / np.sqrt(d), ptr(L)))
        x.shared["chol_cov"] = L

    def step(self, x, target, noise=None):
        """ArrayMetropolis.step (601-611); returns the mean acceptance probability (device scalar)."""
        ctx = context()
        n, d = x.theta.shape
        xprop = x.__class__(theta=torch.empty_like(x.theta))
        z = None if noise is None else as_device(noise[0])
        _lib.check(ctx.lib.smcb_rw_propose(ctx.handle, ptr(x.theta), n, d, ptr(x.shared["chol_cov"]), ptr(z),
                                           ptr(xprop.theta)))
        target(xprop)
        u = None if noise is None else as_device(noise[1])
        acc = empty(1)
        _lib.check(ctx.lib.smcb_mh_accept(ctx.handle, n, d, ptr(x.theta), ptr(x.lprior), ptr(x.llik),
                                          ptr(x.lpost), ptr(xprop.theta), ptr(xprop.lprior), ptr(xprop.llik),
                                          ptr(xprop.lpost), ptr(u), ptr(acc)))
        return acc


class MCMCSequence:
    """smc_samplers.py:651-663."""

    def __init__(self, mcmc=None, len_chain=10):
        self.mcmc = ArrayRandomWalk() if mcmc is None else mcmc
        self.nsteps = len_chain - 1

    def calibrate(self, W, x):
        self.mcmc.calibrate(W, x)


class MCMCSequenceWF(MCMCSequence):
    """Waste-free: keep every intermediate state, smc_samplers.py:669-683."""

    def __call__(self, x, target, noise=None):
        fused = getattr(target, "fused_wf", None)
        if fused is not None and isinstance(self.mcmc, ArrayRandomWalk) and self.nsteps >= 1:
            return fused(x, self.nsteps + 1, noise)     # all chains, all steps: one kernel launch
        xs, ars = [x], []
        for _ in range(self.nsteps):
            x = x.copy()
            ars.append(self.mcmc.step(x, target))
            xs.append(x)
        xout = x.concatenate(*xs)
        xout.shared["acc_rates"] = x.shared.get("acc_rates", []) + [ars]
        return xout


class AdaptiveMCMCSequence(MCMCSequence):
    """Standard SMC sampler move: keep only the final states, smc_samplers.py:686-709
    (fixed number of steps; the adaptive stopping rule of the reference is not implemented)."""

    def __call__(self, x, target):
        xout, ars = x.copy(), []
        for _ in range(self.nsteps):
            ars.append(self.mcmc.step(xout, target))
        xout.shared["acc_rates"] = x.shared.get("acc_rates", []) + [ars]
        return xout


class FKSMCsampler(FeynmanKac):
    """smc_samplers.py:714-769."""

    def __init__(self, model=None, wastefree=True, len_chain=10, move=None):
        self.model, self.wastefree, self.len_chain = model, wastefree, len_chain
        if move is None:
            move = MCMCSequenceWF(len_chain=len_chain) if wastefree else AdaptiveMCMCSequence(len_chain=len_chain)
        self.move = move

    @property
    def T(self):
        return self.model.T

    def default_moments(self, W, x):
        return rs.wmean_and_var(W, x.theta)

    def summary_format(self, smc):
        return "t=%i, ESS=%.2f" % (smc.t, smc.wgts.ESS)

    def time_to_resample(self, smc):
        rs_flag = smc.aux.ESS < smc.X.N * smc.ESSrmin
        smc.X.shared["rs_flag"] = rs_flag
        if rs_flag:
            self.move.calibrate(smc.W, smc.X)
        return rs_flag

    def M0(self, N):
        return self._M0(N * self.len_chain if self.wastefree else N)


class Tempering(FKSMCsampler):
    """smc_samplers.py:797-874."""

    def __init__(self, model=None, wastefree=True, len_chain=10, move=None, exponents=None):
        super().__init__(model=model, wastefree=wastefree, len_chain=len_chain, move=move)
        self.exponents = exponents
        self.deltas = None if exponents is None else np.diff(exponents, prepend=0.0)

    @property
    def T(self):
        return len(self.exponents)

    def logG_tempering(self, x, delta):
        dl = delta * x.llik
        x.lpost = x.lpost + dl
        return dl

    def logG(self, t, xp, x):
        x.shared["exponents"].append(self.exponents[t])
        return self.logG_tempering(x, self.deltas[t])

    def current_target(self, epn):
        def func(x):
            self.model.target(x, epn)
        if hasattr(self.model, "wf_move"):      # lets MCMCSequenceWF run the whole move in one kernel
            func.fused_wf = lambda x, P, noise=None: self.model.wf_move(x, epn, P, noise)
        return func

    def _M0(self, N):
        x0 = ThetaParticles(theta=self.model.prior_rvs(N))
        x0.shared["exponents"] = [0.0]
        self.current_target(0.0)(x0)
        return x0

    def _M(self, t, xp, epn):
        return self.move(xp, self.current_target(epn))

    def M(self, t, xp):
        if xp.shared["rs_flag"]:
            return self._M(t, xp, self.exponents[t - 1])
        return xp


def next_annealing_epn(epn, alpha, lw):
    """smc_samplers.py:876-895: the exponent at which ESS(e * lw) = alpha * N.  The whole bracketing root-find runs on
    the device (smcb_next_annealing_epn: 16 candidate exponents per pass, 11 passes, bracket < 1e-13); the host reads
    the one resulting scalar, because the reference keeps the exponents as Python floats in ``shared``."""
    lw = as_device(lw)
    ctx = context()
    out = empty(1)
    _lib.check(ctx.lib.smcb_next_annealing_epn(ctx.handle, ptr(lw), lw.shape[0], float(epn), float(alpha), ptr(out)))
    return float(out.item())


def next_annealing_epn_host(epn, alpha, lw):
    """The reference's own formulation (brentq on the host, essl on the device): kept for the parity test."""
    N = lw.shape[0]

    def f(e):
        ess = rs.essl(e * lw) if e > 0.0 else N
        return ess - alpha * N

    if f(1.0 - epn) < 0.0:
        return epn + optimize.brentq(f, 0.0, 1.0 - epn)
    return 1.0


class AdaptiveTempering(Tempering):
    """smc_samplers.py:897-936."""

    def __init__(self, model=None, wastefree=True, len_chain=10, move=None, ESSrmin=0.5, max_iter=1000):
        FKSMCsampler.__init__(self, model=model, wastefree=wastefree, len_chain=len_chain, move=move)
        self.ESSrmin, self.max_iter = ESSrmin, max_iter

    @property
    def T(self):
        return self.max_iter

    def time_to_resample(self, smc):
        self.move.calibrate(smc.W, smc.X)
        return True

    def done(self, smc):
        if smc.t >= self.max_iter:
            return True
        if smc.X is None:
            return False
        return smc.X.shared["exponents"][-1] >= 1.0

    def logG(self, t, xp, x):
        epn = x.shared["exponents"][-1]
        new_epn = next_annealing_epn(epn, self.ESSrmin, x.llik)
        x.shared["exponents"].append(new_epn)
        return self.logG_tempering(x, new_epn - epn)

    def M(self, t, xp):
        xp.shared["rs_flag"] = True
        return self._M(t, xp, xp.shared["exponents"][-1])
