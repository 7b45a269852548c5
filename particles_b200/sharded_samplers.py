"""Waste-free adaptive tempering sharded over the GPUs of one box (BASELINE config 5; one process per GPU,
``torch.distributed`` backend nccl) -- ``particles/smc_samplers.py:596-629, 669-683, 876-936`` and the resampling
of ``core.py:329-331`` with the population split over ranks.

What shards: the chains.  Between two resamplings every chain is independent (all P - 1 Metropolis steps of all
local chains run in ONE launch of the fused waste-free kernel, as on a single GPU), so rank r owns M_loc chains and
their M_loc * P particles.  What couples the ranks, once per tempering step:

* the next exponent -- ESS(delta * llik) over ALL particles: every rank evaluates the 16-point ESS grid of the
  device root-find on its shard (``smcb_essl_grid``), one NCCL all-reduce of 32 doubles per pass sums them, and
  every rank takes the same bracket decision (11 passes);
* the normalising constant -- an all-reduce (max, then sum) of the shard's log-sum-exp;
* the proposal calibration -- weighted mean and covariance from all-reduced raw sums (``smcb_wcov_sums``), the
  Cholesky factor on the device (``smcb_chol_from_sums``);
* the resampling -- ONE systematic resampling of M = world * M_loc starting points out of all world * M_loc * P
  weighted particles, the reference's global scheme: shard offsets of the global CDF come from an all-gather of the
  shard masses; every rank finds, in its OWN local CDF, the ancestors of the grid points that fall into its share
  (they are consecutive), gathers those rows locally and sends them to the ranks that own the corresponding chain
  slots with one ``all_to_all_single`` (the split sizes follow from the gathered offsets, so no size exchange).

The particle exchange is real data movement (NCCL); everything else a rank computes uses the same kernels as
``smc_samplers``.  With world = 1 the run reduces to the single-GPU algorithm.
"""
import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import resampling as rs
from .device import as_device, context, empty, ptr
from .smc_samplers import LogisticRegression, ThetaParticles


def plan_global_resample(mass, u, M, world, rank):
    """Host plan of one global systematic resampling (resampling.py:606-610 on the concatenation of the shards) seen
    from ``rank``: ``mass`` = the shards' shares of the total weight (identical on every rank), ``u`` the common
    uniform, ``M`` chain slots per rank.  Returns (mine, v, send_counts, recv_counts): the global slots whose grid
    point falls into this rank's share of the global CDF (consecutive), their positions ``v`` in this shard's OWN
    normalised CDF, and the ``all_to_all_single`` split sizes (slot s lives on rank s // M)."""
    mass = np.asarray(mass, dtype=np.float64)
    Mg = world * M
    goff = np.concatenate([[0.0], np.cumsum(mass)])
    goff[-1] = max(goff[-1], 1.0)
    su = (u + np.arange(Mg)) / Mg                                      # global grid, the same on every rank
    owner = np.minimum(np.searchsorted(goff, su, side="right") - 1, world - 1)
    for _ in range(world):                                             # a grid point on an empty shard's edge: step down
        owner = np.where((mass[owner] > 0.0) | (owner == 0), owner, owner - 1)
    for _ in range(world):                                             # (rank 0 itself empty: step up)
        owner = np.where((mass[owner] > 0.0) | (owner == world - 1), owner, owner + 1)
    mine = np.flatnonzero(owner == rank)
    v = np.minimum((su[mine] - goff[rank]) / mass[rank], 1.0) if mine.size else np.zeros(0)
    v = np.maximum(v, 0.0)
    slot_rank = np.arange(Mg) // M
    send_counts = [int(np.sum(slot_rank[mine] == r)) for r in range(world)]
    recv_counts = [int(np.sum((owner == r) & (slot_rank == rank))) for r in range(world)]
    return mine, v, send_counts, recv_counts


class ShardedAdaptiveTempering:
    """``ShardedAdaptiveTempering(model, M_local, len_chain).run()`` on every rank of an NCCL group.
    ``M_local`` resampled starting points (chains) and ``M_local * len_chain`` particles per rank."""

    def __init__(self, model=None, M_local=1000, len_chain=100, ESSrmin=0.5, seed=0, group=None, max_iter=1000):
        if not isinstance(model, LogisticRegression):
            raise NotImplementedError("the sharded sampler runs the fused waste-free move of LogisticRegression")
        self.model, self.M, self.P = model, int(M_local), int(len_chain)
        self.alpha, self.group, self.max_iter = float(ESSrmin), group, int(max_iter)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ctx = context()
        self.ctx.seed((int(seed) * 1000003 + 7919 * self.rank) & (2 ** 63 - 1))     # independent streams per rank
        self.seed = int(seed)
        self.exponents, self.logLt, self.cpu_time = [0.0], 0.0, None
        self.X, self.W = None, None

    # ------------------------------------------------------------------ collectives on tiny device tensors
    def _allreduce(self, t, op):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)
        return t

    def _next_exponent(self, llik, epn):
        """next_annealing_epn over the union of the shards (smc_samplers.py:876-895)."""
        lib, h = self.ctx.lib, self.ctx.handle
        n_glob = llik.shape[0] * self.world
        st = empty(4)
        _lib.check(lib.smcb_normalise(h, ptr(llik), llik.shape[0], ptr(None), ptr(st)))     # st[0] = max llik
        mx = self._allreduce(st[:1].clone(), dist.ReduceOp.MAX)
        lo, hi = 0.0, 1.0 - epn
        out = empty(32)
        target = self.alpha * n_glob
        for p in range(11):
            _lib.check(lib.smcb_essl_grid(h, ptr(llik), llik.shape[0], lo, hi, ptr(mx), ptr(out)))
            tot = self._allreduce(out.clone(), dist.ReduceOp.SUM).cpu().numpy()
            ess = tot[0::2] ** 2 / tot[1::2]
            below = np.flatnonzero(ess - target < 0.0)
            if below.size == 0:
                if p == 0:
                    return 1.0
                j = 15
            else:
                j = int(below[0])
            lo, hi = lo + (hi - lo) * (j / 16.0), lo + (hi - lo) * ((j + 1) / 16.0)
        return epn + 0.5 * (lo + hi)

    def _calibrate(self, W_loc, share, theta):
        """ArrayRandomWalk.calibrate over all shards: L = 2.38 / sqrt(d) chol(wcov) (smc_samplers.py:617-622).
        ``W_loc`` sums to one on this shard, ``share`` is the shard's part of the total mass: the global weighted sums
        are the share-weighted sums of the local ones, so no per-particle rescaling is needed."""
        lib, h = self.ctx.lib, self.ctx.handle
        n, d = theta.shape
        s0 = empty(d + 1)
        _lib.check(lib.smcb_wcov_sums(h, ptr(W_loc), ptr(theta), n, d, ptr(None), ptr(s0)))
        s0 *= share
        self._allreduce(s0, dist.ReduceOp.SUM)
        mean = (s0[:d] / s0[d]).contiguous()
        tri = empty(d * (d + 1) // 2)
        _lib.check(lib.smcb_wcov_sums(h, ptr(W_loc), ptr(theta), n, d, ptr(mean), ptr(tri)))
        tri *= share
        self._allreduce(tri, dist.ReduceOp.SUM)
        L = empty(d * d).reshape(d, d)
        sw = s0[d:].contiguous()
        _lib.check(lib.smcb_chol_from_sums(h, ptr(tri), ptr(sw), d, 2.38 / np.sqrt(d), ptr(L)))
        return L

    def _global_resample(self, x, W_loc, mass, u):
        """The M = world * M_loc starting points of the next generation: systematic resampling of the global
        population (resampling.py:606-610 applied to the concatenation of the shards), rows delivered to the rank
        that owns each chain slot.  ``W_loc``: this shard's weights normalised to sum to 1, ``mass``: the shards'
        shares of the total (host array, identical on every rank), ``u``: the common uniform."""
        world, M, rank = self.world, self.M, self.rank
        mine, v, send_counts, recv_counts = plan_global_resample(mass, u, M, world, rank)
        fields = ("theta", "lprior", "llik", "lpost")
        send = {}
        if mine.size:
            A = rs.inverse_cdf(as_device(v), W_loc)                    # cumsum + searchsorted kernels
            sel = x[A]                                                 # gather kernels
            send = {k: getattr(sel, k) for k in fields}
        d = x.theta.shape[1]
        if world == 1:
            return ThetaParticles(shared=x.shared.copy(), **send)
        out = {}
        for k in fields:
            w_ = d if k == "theta" else 1
            src = send[k].reshape(-1, w_) if mine.size else torch.empty((0, w_), dtype=torch.float64, device="cuda")
            dst = torch.empty((M, w_), dtype=torch.float64, device="cuda")
            dist.all_to_all_single(dst, src.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts,
                                   group=self.group)
            out[k] = dst if k == "theta" else dst.reshape(-1)
        return ThetaParticles(shared=x.shared.copy(), **out)

    # ------------------------------------------------------------------ the run
    def run(self):
        t0 = time.perf_counter()
        model, M, P, world = self.model, self.M, self.P, self.world
        n_loc = M * P
        x = ThetaParticles(theta=model.prior_rvs(n_loc))
        model.target(x, 0.0)
        epn, logLt, it = 0.0, 0.0, 0
        g = torch.Generator().manual_seed(self.seed)                   # the common uniforms (same on every rank)
        while epn < 1.0 and it < self.max_iter:
            new_epn = self._next_exponent(x.llik, epn)
            lw = (new_epn - epn) * x.llik                              # logG_tempering, smc_samplers.py:847-850
            W_loc, st = empty(n_loc), empty(4)
            _lib.check(self.ctx.lib.smcb_normalise(self.ctx.handle, ptr(lw), n_loc, ptr(W_loc), ptr(st)))
            ms = st[[0, 3]].contiguous()                               # (max lw, sum exp(lw - max)) of this shard
            allms = torch.zeros(2 * world, dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_gather_into_tensor(allms, ms, group=self.group)
            else:
                allms = ms
            hm = allms.cpu().numpy().reshape(world, 2)                 # the one device->host read of the step
            Mx = hm[:, 0].max()
            Sr = hm[:, 1] * np.exp(hm[:, 0] - Mx)
            S = float(Sr.sum())
            logLt += float(Mx) + np.log(S / (n_loc * world))           # log mean weight (weights restart every step)
            mass = Sr / S
            L = self._calibrate(W_loc, float(mass[self.rank]), x.theta)
            u = float(torch.rand(1, generator=g, dtype=torch.float64).item())
            x0 = self._global_resample(x, W_loc, mass, u)
            x0.shared["chol_cov"] = L
            model.target(x0, new_epn)                                  # lpost at the new exponent (lprior / llik unchanged)
            x = model.wf_move(x0, new_epn, P)                          # ONE launch: all chains, all P - 1 steps
            epn = new_epn
            self.exponents.append(epn)
            it += 1
        torch.cuda.synchronize()
        self.X, self.logLt = x, logLt
        self.cpu_time = time.perf_counter() - t0
        return self

    def posterior_mean(self):
        """Mean of theta over all shards (final weights are uniform: the last step ends with a move)."""
        s = self.X.theta.sum(0)
        self._allreduce(s, dist.ReduceOp.SUM)
        return (s / (self.X.theta.shape[0] * self.world)).cpu().numpy()
