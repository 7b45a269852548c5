"""Per-step summaries -- the part of ``particles/collectors.py`` that sits on the
boundary of the hot path: the three default
collectors and ``Moments``.  A collector reads attributes of the running ``SMC``
object; on the fused path the three defaults are filled lazily from the (T, 4)
device table the kernels write, so they cost no per-step host sync.
"""
from . import resampling as rs


class Collector:
    """particles/collectors.py:234-271: subclass and define ``fetch(smc)``."""
    signature = {}

    @property
    def summary_name(self):
        cn = self.__class__.__name__
        return cn[0].lower() + cn[1:]          # Moments -> moments, LogLts -> logLts

    def __init__(self, **kwargs):
        self.summary = []
        for k, v in self.signature.items():
            setattr(self, k, v)
        for k, v in kwargs.items():
            if k in self.signature:
                setattr(self, k, v)
            else:
                raise ValueError(f"Collector {self.__class__.__name__}: unknown parameter {k}")
        self._kwargs = kwargs

    def __call__(self):
        # a collector instance is a template: calling it clones it (collectors.py:263-266)
        return self.__class__(**self._kwargs)

    def collect(self, smc):
        self.summary.append(self.fetch(smc))


class ESSs(Collector):      # collectors.py:278-282
    summary_name = "ESSs"

    def fetch(self, smc):
        return smc.wgts.ESS


class LogLts(Collector):    # collectors.py:285-287
    def fetch(self, smc):
        return smc.logLt


class Rs_flags(Collector):  # collectors.py:290-292
    def fetch(self, smc):
        return smc.rs_flag


default_collector_cls = [ESSs, LogLts, Rs_flags]


class Moments(Collector):
    """particles/collectors.py:301-317: ``mom_func(W, X)`` or the model's
    ``default_moments`` (weighted mean and variance, resampling.py:320-338)."""
    signature = {"mom_func": None}

    def fetch(self, smc):
        f = smc.fk.default_moments if self.mom_func is None else self.mom_func
        return f(smc.W, smc.X)


class Summaries:
    """particles/collectors.py:215-231."""

    def __init__(self, cols):
        self._collectors = [cls() for cls in default_collector_cls]
        self._n_default = len(self._collectors)
        if cols is not None:
            self._collectors.extend(col() for col in cols)
        for col in self._collectors:
            setattr(self, col.summary_name, col.summary)

    @property
    def only_defaults(self):
        return len(self._collectors) == self._n_default

    def device_moments(self, fk):
        """True when every non-default collector is ``Moments()`` with the default ``mom_func`` and the model keeps
        ``FeynmanKac.default_moments`` (resampling.wmean_and_var): exactly what the fused step kernel accumulates."""
        extra = self._collectors[self._n_default:]
        if not extra or not all(type(c) is Moments and c.mom_func is None for c in extra):
            return False
        dm = getattr(type(fk), "default_moments", None)
        return getattr(dm, "__qualname__", "") == "FeynmanKac.default_moments"

    def collect(self, smc):
        for col in self._collectors:
            if type(col) is Moments and col.mom_func is None and getattr(smc, "_dev_moments", False):
                col.summary.append(_moments_row(smc._engine.mom[smc._done - 1].cpu().numpy(), smc._engine.dim))
            else:
                col.collect(smc)

    def _extend_moments(self, table, dim):
        """Bulk fill of every ``Moments`` collector from the (T, 8) device table (fused ``run()``)."""
        rows = [_moments_row(r, dim) for r in table]
        for col in self._collectors[self._n_default:]:
            col.summary.extend(dict(r) for r in rows)

    def _extend_defaults(self, ess, loglt, rs):
        """Bulk fill from the device table (fused ``run()``)."""
        self.ESSs.extend(ess)
        self.logLts.extend(loglt)
        self.rs_flags.extend(rs)


def _moments_row(r, dim):
    """One row of the device table -> what resampling.wmean_and_var returns (resampling.py:320-338)."""
    if dim == 1:
        return {"mean": float(r[0]), "var": float(r[4])}
    return {"mean": r[:dim].copy(), "var": r[4:4 + dim].copy()}


def default_moments(W, X):
    return rs.wmean_and_var(W, X)
