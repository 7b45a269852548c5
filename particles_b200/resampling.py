"""Weights algebra and resampling on the device -- same names, arguments and error
behaviour as ``particles/resampling.py`` of the reference (file:line cited per
function), with hand-written sm_100a kernels underneath (libsmcb.so).

Arrays are CUDA fp64 tensors (numpy inputs are copied to the device); ancestor
indices come back as CUDA int64 tensors; scalars come back as Python floats, which
costs one device->host read -- the fused filter (``core.SMC``) never does that on
its hot path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import as_device, context, empty, ptr

__all__ = ["Weights", "exp_and_normalise", "essl", "log_sum_exp", "log_sum_exp_ab",
           "log_mean_exp", "wmean_and_var", "resampling", "rs_funcs", "inverse_cdf",
           "uniform_spacings", "multinomial", "stratified", "systematic", "residual", "cumsum",
           "ssp", "killing", "multinomial_iid", "multinomial_once", "idiotic"]


def _scalar(ctx, fn, *args):
    out = empty(1)
    _lib.check(fn(ctx.handle, *args, ptr(out)))
    return float(out.item())


def exp_and_normalise(lw):
    """particles/resampling.py:138-163."""
    lw = as_device(lw)
    ctx = context(lw.device)
    W = torch.empty_like(lw)
    _lib.check(ctx.lib.smcb_exp_and_normalise(ctx.handle, ptr(lw), lw.shape[0], ptr(W)))
    return W


def essl(lw):
    """particles/resampling.py:166-188 (returns (sum w)^2 / sum w^2)."""
    lw = as_device(lw)
    ctx = context(lw.device)
    return _scalar(ctx, ctx.lib.smcb_lse, _lib.LSE_ESSL, ptr(lw), ptr(None), lw.shape[0])


def log_sum_exp(v):
    """particles/resampling.py:247-270."""
    v = as_device(v)
    ctx = context(v.device)
    return _scalar(ctx, ctx.lib.smcb_lse, _lib.LSE_SUM, ptr(v), ptr(None), v.shape[0])


def log_sum_exp_ab(a, b):
    """particles/resampling.py:273-288 (two scalars: host arithmetic)."""
    if a > b:
        return a + np.log1p(np.exp(b - a))
    return b + np.log1p(np.exp(a - b))


def log_mean_exp(v, W=None):
    """particles/resampling.py:291-317."""
    v = as_device(v)
    ctx = context(v.device)
    Wd = None if W is None else as_device(W)
    return _scalar(ctx, ctx.lib.smcb_lse, _lib.LSE_MEAN, ptr(v), ptr(Wd), v.shape[0])


def wmean_and_var(W, x):
    """particles/resampling.py:320-338; x is (N,) or (N, d)."""
    W, x = as_device(W), as_device(x)
    ctx = context(W.device)
    n = W.shape[0]
    if x.ndim == 1:
        d, xs = 1, x
    else:
        d, xs = x.shape[1], x.t().contiguous()     # kernels take SoA (d, n)
    out = empty(2 * d, like=W)
    _lib.check(ctx.lib.smcb_wmean_and_var(ctx.handle, ptr(W), ptr(xs), n, d, ptr(out)))
    o = out.cpu().numpy()
    if x.ndim == 1:
        return {"mean": float(o[0]), "var": float(o[1])}
    return {"mean": o[:d].copy(), "var": o[d:].copy()}


class Weights:
    """particles/resampling.py:191-244.  ``lw`` is a CUDA tensor; NaN entries are
    rewritten to -inf IN PLACE (line 220); ``Weights()`` (lw=None) has no
    ``W`` / ``ESS`` / ``log_mean`` attributes and ``N == 0``; ``add`` returns a new
    object.  ``W`` is materialised lazily (one exp pass) the first time it is read;
    the scalars are read from the device lazily as well."""

    def __init__(self, lw=None):
        self.lw = None if lw is None else as_device(lw)
        if self.lw is not None:
            ctx = context(self.lw.device)
            self._stats = empty(4, like=self.lw)
            _lib.check(ctx.lib.smcb_normalise(ctx.handle, ptr(self.lw), self.lw.shape[0], ptr(None),
                                              ptr(self._stats)))
            self._host = None
            self._W = None

    @classmethod
    def _from_device_stats(cls, lw, stats):
        """Wrap weights whose (max, log_mean, ESS, sum) the fused kernels already hold."""
        self = cls.__new__(cls)
        self.lw, self._stats, self._host, self._W = lw, stats, None, None
        return self

    def _scalars(self):
        if self._host is None:
            self._host = self._stats.cpu().numpy()
        return self._host

    def __getattr__(self, name):
        # only reached when normal lookup fails: lazy scalars / W of a non-empty set
        if name in ("log_mean", "ESS", "W") and self.__dict__.get("lw") is not None:
            if name == "log_mean":
                return float(self._scalars()[1])
            if name == "ESS":
                return float(self._scalars()[2])
            if self._W is None:
                ctx = context(self.lw.device)
                W = torch.empty_like(self.lw)
                # W = exp(lw - m) / s with the (m, s) already on the device (the normalise pass above, or the fused
                # filter's own state -- never recomputed, so a sharded filter's global (m, s) stay global)
                _lib.check(ctx.lib.smcb_weights_from_stats(ctx.handle, ptr(self.lw), self.lw.shape[0],
                                                           ptr(self._stats), ptr(W)))
                self._W = W
            return self._W
        raise AttributeError(name)

    @property
    def N(self):
        return 0 if self.lw is None else self.lw.shape[0]

    def add(self, delta):
        """resampling.py:232-244."""
        delta = as_device(delta)
        if self.lw is None:
            return self.__class__(lw=delta)
        return self.__class__(lw=self.lw + delta)


# ---------------------------------------------------------------------------
# resampling schemes -- particles/resampling.py:445-627
# ---------------------------------------------------------------------------
rs_funcs = {}


def cumsum(W):
    """Deterministic, non-decreasing inclusive prefix sum (the CDF inverse_cdf walks)."""
    W = as_device(W)
    ctx = context(W.device)
    out = torch.empty_like(W)
    _lib.check(ctx.lib.smcb_cumsum(ctx.handle, ptr(W), W.shape[0], ptr(out)))
    return out


def inverse_cdf(su, W):
    """particles/resampling.py:484-509: ``su`` sorted; returns int64 ancestors."""
    su, W = as_device(su), as_device(W)
    ctx = context(W.device)
    cdf = cumsum(W)
    A = empty(su.shape[0], dtype=torch.int64, like=W)
    _lib.check(ctx.lib.smcb_searchsorted(ctx.handle, ptr(cdf), W.shape[0], ptr(su), su.shape[0],
                                         ptr(A)))
    return A


def uniform_spacings(N):
    """particles/resampling.py:512-537: N ordered uniforms in O(N) (device Philox)."""
    ctx = context()
    u = empty(N + 1)
    _lib.check(ctx.lib.smcb_uniform(ctx.handle, ptr(u), N + 1))
    z = cumsum(-torch.log(u))
    return z[:-1] / z[-1]


def _resample(scheme, W, M, u=None, return_scratch=False):
    W = as_device(W)
    ctx = context(W.device)
    n = W.shape[0]
    M = n if M is None else int(M)
    A = empty(M, dtype=torch.int64, like=W)
    scratch = empty(int(ctx.lib.smcb_resample_scratch_doubles(n, M)), like=W)
    ud = None
    if u is not None:                      # injected uniforms, in the reference's draw order
        ud = torch.ones(max(M, n) + 2, dtype=torch.float64, device=W.device)
        uu = as_device(u).reshape(-1)
        ud[: uu.shape[0]] = uu
    _lib.check(ctx.lib.smcb_resample(ctx.handle, _lib.RS_CODES[scheme], ptr(W), n, M, ptr(A),
                                     ptr(ud), ptr(scratch)))
    if return_scratch:
        return A, scratch
    return A


def _scheme(name):
    def f(W, M=None, u=None):
        return _resample(name, W, M, u)
    f.__name__ = name
    f.__doc__ = f"{name} resampling on the device (particles/resampling.py:540-627)."
    rs_funcs[name] = f
    return f


multinomial = _scheme("multinomial")
stratified = _scheme("stratified")
systematic = _scheme("systematic")
residual = _scheme("residual")
ssp = _scheme("ssp")                       # resampling.py:630-677 (sequential on the device, see smcb.h)


def _register(f):
    def g(W, M=None, **kw):
        W = as_device(W)
        return f(W, W.shape[0] if M is None else int(M), **kw)
    g.__name__, g.__doc__ = f.__name__, f.__doc__
    rs_funcs[f.__name__] = g
    return g


def _uniforms(n, like):
    ctx = context(like.device)
    u = empty(n, like=like)
    _lib.check(ctx.lib.smcb_uniform(ctx.handle, ptr(u), n))
    return u


@_register
def multinomial_iid(W, M, u=None, u_perm=None):
    """particles/resampling.py:560-570: multinomial resampling followed by a uniformly random
    permutation (argsort of M device uniforms), so that the indices are IID."""
    A = _resample("multinomial", W, M, u)
    keys = _uniforms(M, W) if u_perm is None else as_device(u_perm)
    return A[torch.argsort(keys)]


def multinomial_once(W, u=None):
    """particles/resampling.py:573-597: one draw, ``searchsorted(cumsum(W), rand())``."""
    W = as_device(W)
    su = _uniforms(1, W) if u is None else as_device(np.atleast_1d(u))
    ctx = context(W.device)
    cdf = cumsum(W)
    A = empty(1, dtype=torch.int64, like=W)
    # the reference does not clip: a draw above cdf[-1] returns N; searchsorted here clips to N - 1
    _lib.check(ctx.lib.smcb_searchsorted(ctx.handle, ptr(cdf), W.shape[0], ptr(su), 1, ptr(A)))
    return int(A.item())


@_register
def killing(W, M, u=None, u_multinomial=None):
    """particles/resampling.py:680-697: particle i survives with probability W[i] / max(W), otherwise it
    is replaced by a multinomial draw.  Defined only for M = N (ValueError otherwise, as the reference)."""
    n = W.shape[0]
    if M != n:
        raise ValueError("killing resampling defined only for M=N")
    uu = _uniforms(n, W) if u is None else as_device(u)
    killed = uu * W.max() >= W
    nkilled = int(killed.sum().item())              # sizes the multinomial draw: one host read
    A = torch.arange(n, dtype=torch.int64, device=W.device)
    if nkilled:
        A[killed] = _resample("multinomial", W, nkilled, u_multinomial)
    return A


@_register
def idiotic(W, M, u=None):
    """particles/resampling.py:700-707 (testing only): every offspring is the same single draw."""
    return torch.full((M,), multinomial_once(W, u), dtype=torch.int64, device=W.device)


def resampling(scheme, W, M=None):
    """particles/resampling.py:477-481."""
    try:
        return rs_funcs[scheme](W, M=M)
    except KeyError:
        raise ValueError(f"{scheme} is not a valid resampling scheme")
