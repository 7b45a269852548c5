"""Linear-Gaussian state-space models of ``particles/kalman.py`` (the model classes
only: 296-452) over device arrays.  The exact Kalman filter of the reference is the
known-answer oracle of this path and lives with the test infrastructure
(``oracle/smc_numpy.py``), not here.
"""
import numpy as np
import torch

from . import distributions as dists
from . import state_space_models as ssms
from .device import as_device


def _h(y):
    return np.asarray(y.cpu() if isinstance(y, torch.Tensor) else y, dtype=np.float64).reshape(-1)


class MVLinearGauss(ssms.StateSpaceModel):
    """kalman.py:296-361: X_0 ~ N(mu0, cov0); X_t = F X_{t-1} + U_t; Y_t = G X_t + V_t."""

    def __init__(self, F=None, G=None, covX=None, covY=None, mu0=None, cov0=None):
        self.covX, self.covY = np.atleast_2d(covX), np.atleast_2d(covY)
        self.dx, self.dy = self.covX.shape[0], self.covY.shape[0]
        self.mu0 = np.zeros(self.dx) if mu0 is None else mu0
        self.cov0 = self.covX if cov0 is None else np.atleast_2d(cov0)
        self.F = np.eye(self.dx) if F is None else np.atleast_2d(F)
        self.G = np.eye(self.dy, self.dx) if G is None else np.atleast_2d(G)
        assert self.F.shape == (self.dx, self.dx) and self.G.shape == (self.dy, self.dx)

    def _dev(self, M):
        return as_device(np.ascontiguousarray(M))

    def PX0(self):
        return dists.MvNormal(loc=self.mu0, cov=self.cov0)

    def PX(self, t, xp):
        return dists.MvNormal(loc=xp @ self._dev(self.F.T), cov=self.covX)

    def PY(self, t, xp, x):
        return dists.MvNormal(loc=x @ self._dev(self.G.T), cov=self.covY)

    # Kalman update with a common predictive covariance (kalman.py:196-229, 232-262)
    def _gain(self, pred_cov):
        dpc = self.G @ pred_cov @ self.G.T + self.covY
        gain = np.linalg.solve(dpc, (pred_cov @ self.G.T).T).T
        fcov = pred_cov - gain @ self.G @ pred_cov
        return dpc, gain, fcov

    def proposal0(self, data):
        dpc, gain, fcov = self._gain(self.cov0)
        resid = _h(data[0]) - self.mu0 @ self.G.T
        return dists.MvNormal(loc=self.mu0 + resid @ gain.T, cov=fcov)

    def proposal(self, t, xp, data):
        dpc, gain, fcov = self._gain(self.covX)
        pm = xp @ self._dev(self.F.T)
        resid = as_device(_h(data[t])) - pm @ self._dev(self.G.T)
        return dists.MvNormal(loc=pm + resid @ self._dev(gain.T), cov=fcov)

    def logeta(self, t, x, data):
        dpc, _, _ = self._gain(self.covX)
        pm = x @ self._dev(self.F.T)
        return dists.MvNormal(loc=pm @ self._dev(self.G.T), cov=dpc).logpdf(_h(data[t + 1]))


class MVLinearGauss_Guarniero_etal(MVLinearGauss):
    """kalman.py:364-394: F_ij = alpha^(1 + |i - j|), G = covX = covY = cov0 = I."""

    def __init__(self, alpha=0.4, dx=2):
        F = np.empty((dx, dx))
        for i in range(dx):
            for j in range(dx):
                F[i, j] = alpha ** (1 + abs(i - j))
        MVLinearGauss.__init__(self, F=F, G=np.eye(dx), covX=np.eye(dx), covY=np.eye(dx))


class LinearGauss(MVLinearGauss):
    """kalman.py:397-452."""
    default_params = {"sigmaY": 0.2, "rho": 0.9, "sigmaX": 1.0, "sigma0": None}

    def __init__(self, **kwargs):
        ssms.StateSpaceModel.__init__(self, **kwargs)
        if self.sigma0 is None:
            self.sigma0 = self.sigmaX / np.sqrt(1.0 - self.rho ** 2)
        MVLinearGauss.__init__(self, F=self.rho, G=1.0, covX=self.sigmaX ** 2,
                               covY=self.sigmaY ** 2, cov0=self.sigma0 ** 2)

    def PX0(self):
        return dists.Normal(scale=self.sigma0)

    def PX(self, t, xp):
        return dists.Normal(loc=self.rho * xp, scale=self.sigmaX)

    def PY(self, t, xp, x):
        return dists.Normal(loc=x, scale=self.sigmaY)

    def proposal0(self, data):
        sig2post = 1.0 / (1.0 / self.sigma0 ** 2 + 1.0 / self.sigmaY ** 2)
        mupost = sig2post * (_h(data[0])[0] / self.sigmaY ** 2)
        return dists.Normal(loc=mupost, scale=np.sqrt(sig2post))

    def proposal(self, t, xp, data):
        sig2post = 1.0 / (1.0 / self.sigmaX ** 2 + 1.0 / self.sigmaY ** 2)
        mupost = sig2post * (self.rho * xp / self.sigmaX ** 2 + _h(data[t])[0] / self.sigmaY ** 2)
        return dists.Normal(loc=mupost, scale=np.sqrt(sig2post))

    def logeta(self, t, x, data):
        law = dists.Normal(loc=self.rho * x, scale=np.sqrt(self.sigmaX ** 2 + self.sigmaY ** 2))
        return law.logpdf(_h(data[t + 1])[0])
