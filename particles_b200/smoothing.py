"""Particle history containers -- the part of ``particles/smoothing.py`` that touches the
step loop (``hist.save(smc)``, core.py:362-363; classes at smoothing.py:151-270).  The off-line
smoothing algorithms that consume a history are outside the accelerated path.

The reference stores references to ``smc.X / A / wgts`` (it allocates new arrays every step).
The device loop ping-pongs two buffers instead, so a history OWNS what it saves: ``save`` clones
the device arrays (8(d+2) bytes per particle per saved step -- at N = 1e7 keep the window short).
"""
from collections import deque

import torch

from . import resampling as rs


def _own(x):
    return x.clone() if isinstance(x, torch.Tensor) else x


def _own_weights(w):
    if w.lw is None:
        return rs.Weights()
    return rs.Weights._from_device_stats(w.lw.clone(), w._stats.clone())


def generate_hist_obj(option, smc):
    """smoothing.py:151-161."""
    if option is True:
        return ParticleHistory(smc.fk, smc.qmc)
    elif option is False:
        return None
    elif callable(option):
        return PartialParticleHistory(option)
    elif isinstance(option, int) and option >= 0:
        return RollingParticleHistory(option)
    raise ValueError("store_history: invalid option")


class PartialParticleHistory:
    """smoothing.py:164-178: records the particle system at the times ``func(t)`` selects."""

    def __init__(self, func):
        self.is_save_time = func
        self.X, self.wgts = {}, {}

    def save(self, smc):
        t = smc.t
        if self.is_save_time(t):
            self.X[t] = _own(smc.X)
            self.wgts[t] = _own_weights(smc.wgts)


class RollingParticleHistory:
    """smoothing.py:181-219: keeps the k most recent particle systems."""

    def __init__(self, length):
        self.X = deque([], length)
        self.A = deque([], length)
        self.wgts = deque([], length)

    @property
    def N(self):
        return self.X[0].shape[0]

    @property
    def T(self):
        return len(self.X)

    def save(self, smc):
        self.X.append(_own(smc.X))
        self.A.append(_own(smc.A))
        self.wgts.append(_own_weights(smc.wgts))

    def compute_trajectories(self):
        """(T, N) int64 tensor B with B[t, n] = index at time t of the ancestor of X_T^n
        (smoothing.py:209-219); iterated gathers ``A[B]`` on the device."""
        Bs = [torch.arange(self.N, device=self.X[0].device)]
        for A in list(self.A)[-1:0:-1]:
            Bs.append(A[Bs[-1]])
        Bs.reverse()
        return torch.stack(Bs)


class ParticleHistory(RollingParticleHistory):
    """smoothing.py:222-270 (storage + ``extract_one_trajectory``)."""

    def __init__(self, fk, qmc):
        self.X, self.A, self.wgts = [], [], []
        self.fk = fk

    def extract_one_trajectory(self):
        traj, n = [], None
        for t in reversed(range(self.T)):
            if t == self.T - 1:
                n = int(rs.multinomial(self.wgts[-1].W, M=1)[0])
            else:
                n = int(self.A[t + 1][n])
            traj.append(self.X[t][n])
        return traj[::-1]
