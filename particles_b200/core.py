"""``SMC`` and ``FeynmanKac``: the step loop of ``particles/core.py:108-409`` on a B200.

``SMC(fk=..., N=..., resampling=..., ESSrmin=..., collect=...)`` keeps the reference's
constructor, iterator protocol and attributes (``t, X, Xp, A, wgts, aux, W, logLt,
loglt, log_mean_w, rs_flag, cpu_time, summaries``).  Two execution paths:

* FUSED (stock models recognised by ``state_space_models.fused_spec``): the whole
  step -- ESS test, scan + search + gather, propagate, log-weight, max-shifted
  normalisation, logLt recursion -- runs in libsmcb's kernels with the decision taken
  on the device; ``run()`` enqueues all T steps without a host sync and reads the
  (T, 4) summary table once.
* PLUGIN (any other ``FeynmanKac``): the reference's loop (core.py:299-383) with the
  model's ``M0 / M / logG / logeta`` called on CUDA tensors and our
  ``resampling`` / ``Weights`` kernels underneath.

There is no CPU path: without the CUDA library / device the constructor raises.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from . import collectors
from . import resampling as rs
from .device import as_device, context, empty, ptr, require_cuda, tensor_from_ptr


class FeynmanKac:
    """Abstract Feynman-Kac model -- particles/core.py:108-197."""

    def __init__(self, T):
        self.T = T

    def _error_msg(self, meth):
        return f"method/property {meth} missing in class {self.__class__.__name__}"

    def M0(self, N):
        raise NotImplementedError(self._error_msg("M0"))

    def M(self, t, xp):
        raise NotImplementedError(self._error_msg("M"))

    def logG(self, t, xp, x):
        raise NotImplementedError(self._error_msg("logG"))

    @property
    def isAPF(self):
        return "logeta" in dir(self)

    def done(self, smc):
        return smc.t >= self.T

    def time_to_resample(self, smc):
        return smc.aux.ESS < smc.N * smc.ESSrmin          # strict <, core.py:183

    def default_moments(self, W, X):
        return rs.wmean_and_var(W, X)

    def summary_format(self, smc):
        return "t=%i: resample:%s, ESS (end of iter)=%.2f" % (smc.t, smc.rs_flag, smc.wgts.ESS)


def _is_apf(fk):
    return fk.isAPF if hasattr(fk, "isAPF") else ("logeta" in dir(fk))


class _FusedEngine:
    """Owns the device buffers of one fused filter and the smcb_filter handle."""

    def __init__(self, spec, N, scheme, ESSrmin, seed, noise=None, n_global=None, index_offset=0,
                 world=1, rank=0, group=None, p2p=False, global_rs=False, moments=False):
        self.world, self.rank, self.group = int(world), int(rank), group
        self.p2p = bool(p2p) and self.world > 1
        self.global_rs = bool(global_rs) and self.world > 1
        if self.global_rs and not self.p2p:
            raise ValueError("global resampling over shards needs the peer-memory exchange (exchange='p2p')")
        self._pool = None
        self.ctx = context()
        self.lib = self.ctx.lib
        self.N, self.T = int(N), int(spec["data"].shape[0])
        n, T = self.N, self.T
        self.dim, self.dy = int(spec.get("dim", 1)), int(spec.get("dy", 1))
        dev = self.ctx.device
        f64 = dict(dtype=torch.float64, device=dev)
        xshape = (n,) if self.dim == 1 else (self.dim, n)      # SoA: component-major
        if self.p2p:
            self._pool = _p2p_pool(self.ctx, self.world, self.rank, group)
        if self.global_rs:      # particles and CDF in peer-mapped memory: peers pull ancestors from it
            arena = self._pool.arena((2 * self.dim + 1) * n * 8)
            nd = self.dim * n
            self.X = [tensor_from_ptr(arena, xshape, owner=self._pool), tensor_from_ptr(arena + nd * 8, xshape, owner=self._pool)]
            self.cdf = tensor_from_ptr(arena + 2 * nd * 8, (n,), owner=self._pool)
        else:
            self.X = [torch.empty(xshape, **f64), torch.empty(xshape, **f64)]
            self.cdf = torch.empty(n, **f64)
        self.lw = [torch.empty(n, **f64), torch.empty(n, **f64)]
        self.A = torch.empty(n, dtype=torch.int64, device=dev)
        self.summ = torch.zeros((T, _lib.SUMMARY_STRIDE), **f64)
        # collectors.Moments on the device: per step the weighted mean / variance of every component
        self.mom = torch.zeros((T, 8), **f64) if moments else None
        # the observations are the only per-run host input of this path: pinned -> device
        self.data_host = torch.from_numpy(spec["data"].reshape(-1)).pin_memory()
        self.data = self.data_host.to(dev, non_blocking=True)
        self.sc = None
        if spec.get("step_consts") is not None:
            self.sc = as_device(spec["step_consts"])
        self.scratch = torch.empty(n + 2, **f64) if scheme == "multinomial" else None
        self.z_in = self.u_in = None
        if noise is not None:
            z, u = noise
            self.z_in = None if z is None else as_device(z)
            self.u_in = None if u is None else as_device(u)
        d = _lib.FilterDesc()
        d.model, d.fk, d.scheme, d.dim = spec["model"], spec["fk"], _lib.RS_CODES[scheme], self.dim
        d.dy, d.n_params = self.dy, len(spec["params"])
        d.n, d.n_global = n, int(n_global or n)
        d.index_offset, d.T = int(index_offset), T
        d.essrmin, d.seed = float(ESSrmin), int(seed) & (2 ** 64 - 1)
        for i, v in enumerate(spec["params"]):
            d.params[i] = float(v)
        d.X[0], d.X[1] = self.X[0].data_ptr(), self.X[1].data_ptr()
        d.lw[0], d.lw[1] = self.lw[0].data_ptr(), self.lw[1].data_ptr()
        d.A, d.cdf = self.A.data_ptr(), self.cdf.data_ptr()
        d.data, d.summaries = self.data.data_ptr(), self.summ.data_ptr()
        d.z_in = self.z_in.data_ptr() if self.z_in is not None else None
        d.u_in = self.u_in.data_ptr() if self.u_in is not None else None
        d.scratch = self.scratch.data_ptr() if self.scratch is not None else None
        d.step_consts = self.sc.data_ptr() if self.sc is not None else None
        d.moments = self.mom.data_ptr() if self.mom is not None else None
        d.world, d.rank = self.world, self.rank
        if self.world > 1:      # per-step exchange buffers of the sharded filter (16 doubles / rank)
            self.local_stats = torch.zeros(16, **f64)
            self.gathered = torch.zeros(16 * self.world, **f64)
            d.local_stats, d.gathered = self.local_stats.data_ptr(), self.gathered.data_ptr()
            if self.p2p:
                mail = self._pool.mailbox()
                d.mail_local = mail[self.rank]
                for r in range(self.world):
                    d.mail_peer[r] = mail[r]
            if self.global_rs:
                nd = self.dim * n * 8
                d.rs_global = 1
                for r in range(self.world):
                    base = self._pool.arena_of(r)
                    d.peer_X0[r], d.peer_X1[r], d.peer_cdf[r] = base, base + nd, base + 2 * nd
        self.desc = d
        h = C.c_void_p()
        _lib.check(self.lib.smcb_filter_create(self.ctx.handle, C.byref(d), C.byref(h)))
        self.handle = h

    def step(self, nsteps=1):
        self.ctx.bind_stream()
        if self.world == 1 or self.p2p:      # the whole loop is enqueued by the C side
            _lib.check(self.lib.smcb_filter_step(self.handle, int(nsteps)))
            return
        import torch.distributed as dist
        for _ in range(int(nsteps)):    # kernels -> one tiny all-gather (NCCL, stream-ordered) -> finish
            _lib.check(self.lib.smcb_filter_step_local(self.handle))
            dist.all_gather_into_tensor(self.gathered, self.local_stats, group=self.group)
            _lib.check(self.lib.smcb_filter_step_finish(self.handle))

    def step_timed(self, nsteps):
        """smcb_filter_step_timed: per-kernel device milliseconds (CUDA events)."""
        self.ctx.bind_stream()
        out = (C.c_double * 8)()
        _lib.check(self.lib.smcb_filter_step_timed(self.handle, int(nsteps), out))
        ms = dict(zip(("init", "step_rs", "tail", "step"), out[0:4]))
        cnt = dict(zip(("init", "step_rs", "tail", "step"), (int(v) for v in out[4:8])))
        return ms, cnt

    def state(self):
        out = (C.c_double * 8)()
        _lib.check(self.lib.smcb_filter_state(self.handle, out))
        return list(out)

    def close(self):
        """Free the filter handle.  Peer-mapped memory (mailboxes, arenas) belongs to the per-(process, group)
        pool and stays mapped for the next sharded filter, so closing needs no collective call."""
        if getattr(self, "handle", None):
            self.lib.smcb_filter_destroy(self.handle)      # synchronises this rank's stream
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _P2PPool:
    """Peer-mapped memory of one (process, process group): ONE mailbox and one (growing) arena per rank, allocated
    and exchanged once -- a single tensor all-gather of the 64-byte CUDA IPC handles -- and then reused by every
    sharded filter of the group: constructing the second ``ShardedSMC`` costs no collective, no
    ``cudaIpcOpenMemHandle`` and no barrier.

    Reuse is made safe by generations.  The mailbox has four slabs; filter number g of the group (all ranks make
    their filters in the same order) uses slab g % 4 and zeroes -- locally, stream-ordered -- slab (g + 2) % 4.
    Epochs inside a slab are step indices.  Nobody can still be writing into the slab being zeroed: it last served
    generation g - 2, and this rank could only finish generation g - 1 after every peer had sent the statistics of
    that filter's last step, i.e. after all their kernels of generation g - 2 had retired.  The arena (particles and
    CDF of the exact global resampling) needs no tag: peers read it only inside a resampling step of the current
    filter, and a rank finishes a filter only after every peer's reads of its last step are done."""

    def __init__(self, ctx, world, rank, group):
        self.ctx, self.lib, self.world, self.rank, self.group = ctx, ctx.lib, world, rank, group
        self._mail = None          # [ptr per rank]
        self._arena = None         # [ptr per rank]
        self._arena_bytes = 0
        self._gen = 0

    def _exchange(self, nbytes):
        """Allocate nbytes here (zeroed, synchronously), hand the IPC handle to every rank, map theirs."""
        import torch.distributed as dist
        ptr_ = C.c_void_p()
        hbuf = C.create_string_buffer(64)
        _lib.check(self.lib.smcb_p2p_alloc(self.ctx.handle, int(nbytes), C.byref(ptr_), hbuf))
        mine = torch.frombuffer(bytearray(hbuf.raw), dtype=torch.uint8).to(self.ctx.device)
        allh = torch.empty(64 * self.world, dtype=torch.uint8, device=self.ctx.device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        allh = allh.cpu().numpy().tobytes()
        out = []
        for r in range(self.world):
            if r == self.rank:
                out.append(ptr_.value)
                continue
            pp = C.c_void_p()
            _lib.check(self.lib.smcb_p2p_open(self.ctx.handle, allh[64 * r:64 * (r + 1)], C.byref(pp)))
            out.append(pp.value)
        return out

    def mailbox(self):
        """[mailbox of rank r as mapped here] for the next filter of this group."""
        nslab, slab = 4, 2 * self.world * 32 * 8
        if self._mail is None:
            self._mail = self._exchange(nslab * slab)
        g = self._gen
        self._gen += 1
        tensor_from_ptr(self._mail[self.rank] + ((g + 2) % nslab) * slab, (slab // 8,), owner=self).zero_()
        return [p + (g % nslab) * slab for p in self._mail]

    def arena(self, nbytes):
        if self._arena is None or nbytes > self._arena_bytes:
            self._arena = self._exchange(nbytes)      # a smaller, earlier arena stays mapped until process exit
            self._arena_bytes = nbytes
        return self._arena[self.rank]

    def arena_of(self, r):
        return self._arena[r]


_pools = {}


def _p2p_pool(ctx, world, rank, group):
    key = (ctx.device.index, id(group) if group is not None else None, world, rank)
    if key not in _pools:
        _pools[key] = _P2PPool(ctx, world, rank, group)
    return _pools[key]


class SMC:
    """Drop-in for ``particles.SMC`` (particles/core.py:200-409).

    Extra keyword arguments (all optional, defaults keep the reference's behaviour):
    ``seed`` re-keys the device generator for this run; ``fused=False`` forces the
    plugin path; ``noise=(z, u)`` injects standard normals / uniforms (parity tests).
    """

    def __init__(self, fk=None, N=100, qmc=False, resampling="systematic", ESSrmin=0.5,
                 store_history=False, verbose=False, collect=None, seed=None, fused=None,
                 noise=None):
        require_cuda()
        _lib.load()
        if qmc:
            raise NotImplementedError("SQMC (qmc=True) is outside the accelerated path")
        if resampling not in rs.rs_funcs:
            raise ValueError(f"{resampling} is not a valid resampling scheme")
        self.fk, self.N, self.qmc = fk, N, qmc
        self.resampling, self.ESSrmin, self.verbose = resampling, ESSrmin, verbose
        self.t = 0
        self._done = 0        # completed steps (== t outside of a step; collectors see t = index)
        self.cpu_time = None
        self.summaries = None if collect == "off" else collectors.Summaries(collect)
        self._seed = np.random.randint(0, 2 ** 31 - 1) if seed is None else int(seed)
        self._engine = None
        self._dev_moments = False
        self._noise = noise
        spec = None
        if fused is not False:
            from .state_space_models import fused_spec
            spec = fused_spec(fk)
            if spec is not None and resampling not in _lib.FUSED_SCHEMES:
                spec = None                      # residual, ssp, killing, ...: plugin path (stand-alone kernels)
            if spec is None and fused is True:
                raise NotImplementedError("this Feynman-Kac model has no fused kernel")
        if spec is not None:
            # collect=[Moments()] with the default mom_func: the step kernel accumulates sum w x / sum w x^2 next
            # to its log-sum-exp triple and writes a (T, 8) table -- run() keeps its sync-free fast path
            self._dev_moments = self.summaries is not None and self.summaries.device_moments(fk)
            self._engine = _FusedEngine(spec, N, resampling, ESSrmin, self._seed, noise,
                                        moments=self._dev_moments)
            self._row_cache = {}
        else:
            context().seed(self._seed)
            self._p = {"rs_flag": False, "logLt": 0.0, "wgts": rs.Weights(), "aux": None,
                       "X": None, "Xp": None, "A": None}
        from . import smoothing
        self.hist = smoothing.generate_hist_obj(store_history, self)   # smoothing.py:151-161

    # ------------------------------------------------------------------ fused
    @property
    def fused(self):
        return self._engine is not None

    def _row(self, t):
        """(ESS, logLt, rs_flag, log_mean_w) of step t from the device table."""
        if t not in self._row_cache:
            self._row_cache = {t: self._engine.summ[t].cpu().numpy()}
        return self._row_cache[t]

    def _cur(self):
        return (self._done - 1) & 1      # step s writes buffers [s & 1]

    # ------------------------------------------------------------- attributes
    @property
    def X(self):
        if not self.fused:
            return self._p["X"]
        if self._done == 0:
            return None
        x = self._engine.X[self._cur()]
        return x if x.ndim == 1 else x.t()          # (N, d) view of the SoA buffer

    @X.setter
    def X(self, v):
        self._p["X"] = v

    @property
    def rs_flag(self):
        if not self.fused:
            return self._p["rs_flag"]
        return False if self._done == 0 else bool(self._row(self._done - 1)[2])

    @property
    def logLt(self):
        if not self.fused:
            return self._p["logLt"]
        return 0.0 if self._done == 0 else float(self._row(self._done - 1)[1])

    @property
    def log_mean_w(self):
        if not self.fused:
            return self._p["log_mean_w"]
        return float(self._row(self._done - 1)[3])

    @property
    def loglt(self):
        if not self.fused:
            return self._p["loglt"]
        t = self._done - 1
        if t == 0 or self.rs_flag:
            return self.log_mean_w
        return self.log_mean_w - float(self._engine.summ[t - 1, 3].item())

    @property
    def A(self):
        if not self.fused:
            return self._p["A"]
        if self._done <= 1:
            return None
        if self.rs_flag:
            return self._engine.A
        return torch.arange(self.N, device=self._engine.A.device)      # core.py:335

    @property
    def Xp(self):
        if not self.fused:
            return self._p["Xp"]
        if self._done <= 1:
            return None
        prev = self._engine.X[self._cur() ^ 1]
        if not self.rs_flag:
            return prev if prev.ndim == 1 else prev.t()
        out = torch.empty_like(prev)
        ctx = self._engine.ctx
        _lib.check(ctx.lib.smcb_gather(ctx.handle, ptr(prev), self.N, ptr(self._engine.A), self.N,
                                       self._engine.dim, ptr(out)))
        return out if out.ndim == 1 else out.t()

    @property
    def wgts(self):
        if not self.fused:
            return self._p["wgts"]
        if self._done == 0:
            return rs.Weights()
        st = self._engine.state()
        stats = torch.tensor([st[6], st[5], st[4], st[7]], dtype=torch.float64,
                             device=self._engine.lw[0].device)
        return rs.Weights._from_device_stats(self._engine.lw[self._cur()], stats)

    @property
    def aux(self):
        if not self.fused:
            return self._p["aux"]
        return self.wgts

    @property
    def W(self):
        return self.wgts.W

    def __str__(self):
        return self.fk.summary_format(self)

    # ----------------------------------------------------------- plugin path
    def reset_weights(self):                                  # core.py:299-305
        p = self._p
        if _is_apf(self.fk):
            lw = rs.log_mean_exp(self.logetat, W=p["wgts"].W) - self._gather(self.logetat, p["A"])
            p["wgts"] = rs.Weights(lw=lw)
        else:
            p["wgts"] = rs.Weights()

    def setup_auxiliary_weights(self):                        # core.py:307-313
        p = self._p
        if _is_apf(self.fk):
            self.logetat = as_device(self.fk.logeta(self.t - 1, p["X"]))
            p["aux"] = p["wgts"].add(self.logetat)
        else:
            p["aux"] = p["wgts"]

    def _gather(self, X, A):
        if not isinstance(X, (torch.Tensor, np.ndarray)):
            return X[A]          # particle containers (smc_samplers.ThetaParticles) index themselves
        ctx = context()
        X = as_device(X)
        d = 1 if X.ndim == 1 else X.shape[1]
        out = torch.empty((A.shape[0],) + tuple(X.shape[1:]), dtype=X.dtype, device=X.device)
        _lib.check(ctx.lib.smcb_gather_rows(ctx.handle, ptr(X), X.shape[0], ptr(A), A.shape[0], d,
                                            ptr(out)))
        return out

    def generate_particles(self):                             # core.py:315-321
        self._p["X"] = self.fk.M0(self.N)

    def reweight_particles(self):                             # core.py:323-324
        p = self._p
        p["wgts"] = p["wgts"].add(self.fk.logG(self.t, p["Xp"], p["X"]))

    def resample_move(self):                                  # core.py:326-337
        p = self._p
        p["rs_flag"] = bool(self.fk.time_to_resample(self))
        if p["rs_flag"]:
            p["A"] = rs.resampling(self.resampling, p["aux"].W, M=self.N)
            p["Xp"] = self._gather(p["X"], p["A"])
            self.reset_weights()
        else:
            p["A"] = torch.arange(self.N, device="cuda")
            p["Xp"] = p["X"]
        p["X"] = self.fk.M(self.t, p["Xp"])

    def compute_summaries(self):                              # core.py:351-367
        p = self._p
        if self.t > 0:
            prec = p["log_mean_w"]
        p["log_mean_w"] = p["wgts"].log_mean
        if self.t == 0 or p["rs_flag"]:
            p["loglt"] = p["log_mean_w"]
        else:
            p["loglt"] = p["log_mean_w"] - prec
        p["logLt"] += p["loglt"]
        if self.verbose:
            print(self)
        if self.hist:
            self.hist.save(self)
        if self.summaries:
            self.summaries.collect(self)

    # -------------------------------------------------------------- iterator
    def __next__(self):
        """One step of a particle filter (core.py:369-383)."""
        if self.fk.done(self):
            raise StopIteration
        if self.fused:
            self._engine.step(1)
            self._done += 1
            if self.verbose:
                print(self)
            if self.hist:
                self.hist.save(self)              # core.py:362-363 (before the collectors)
            if self.summaries:
                self.summaries.collect(self)      # smc.t is still the index of this step
            self.t += 1
            return
        if self.t == 0:
            self.generate_particles()
        else:
            self.setup_auxiliary_weights()
            self.resample_move()
        self.reweight_particles()
        self.compute_summaries()
        self.t += 1
        self._done = self.t

    def next(self):
        return self.__next__()

    def __iter__(self):
        return self

    def run(self):
        """Run until completion (core.py:391-409); ``cpu_time`` is the wall time of this
        call, device work included (utils.timer semantics, utils.py:81-89)."""
        t0 = time.perf_counter()
        if self.fused and not self.verbose and not self.hist \
                and (self.summaries is None or self.summaries.only_defaults or self._dev_moments) \
                and getattr(getattr(type(self.fk), "done", None), "__qualname__", "") == "FeynmanKac.done":
            T = self._engine.T
            first = self.t
            if first < T:
                self._engine.step(T - first)
                self.t = self._done = T
            table = self._engine.summ.cpu().numpy()       # the one device->host read of the run
            self._row_cache = {T - 1: table[T - 1]}
            if self.summaries is not None:
                self.summaries._extend_defaults([float(v) for v in table[first:T, 0]],
                                                [float(v) for v in table[first:T, 1]],
                                                [bool(v) for v in table[first:T, 2]])
                if self._dev_moments:
                    self.summaries._extend_moments(self._engine.mom.cpu().numpy()[first:T], self._engine.dim)
        else:
            for _ in self:
                pass
            if self.fused:
                torch.cuda.synchronize()
        self.cpu_time = time.perf_counter() - t0
