"""world_size-2 gloo tests (CPU) of the sharded-filter protocol of particles_b200/parallel.py:
the per-step all-gather of (max, sum exp, sum exp^2) triples, the rank-order merge, the
island restart log-weight, and the two-level CDF search of the exact global resampling mode.  The shard arithmetic is done by the oracle (NumPy) here; the
device kernels implement the same algebra (csrc/smcb_filter.cu: k_finish / finalize_step)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import smc_numpy as orc
from particles_b200 import parallel as par


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def island_filter(rank, world, n_local, y, seed, essrmin=0.5):
    """One rank of the island bootstrap filter for StochVol, stats exchanged through gloo."""
    rng = np.random.RandomState(seed * 1000 + rank)
    m = orc.StochVol()
    N = n_local * world
    logLt, prev_lm, out = 0.0, None, []
    x = lw = None
    for t in range(len(y)):
        if t == 0:
            px0 = m.PX0()
            x = px0.loc + px0.scale * rng.standard_normal(n_local)
            lw = m.PY(0, None, x).logpdf(y[0])
            rs = False
        else:
            rs = bool(ess < N * essrmin)
            base = lw
            if rs:
                W = orc.exp_and_normalise(lw)
                A = orc.systematic(W, n_local, u=rng.rand(1))
                x = x[A]
                base = par.island_restart(mine, gw, world)           # shard mass carried
            x = m.EXt(x) + m.sigma * rng.standard_normal(n_local)
            lw = base + m.PY(t, None, x).logpdf(y[t])
        mine = par.lse3_of(lw)
        buf = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(buf, torch.tensor(mine, dtype=torch.float64))
        triples = [tuple(b.tolist()) for b in buf]
        gw = par.merge_lse3(triples)
        lm, ess = par.global_stats(triples, N)
        logLt += lm if (t == 0 or rs) else lm - prev_lm
        prev_lm = lm
        out.append((ess, logLt, rs))
    return out


def _worker(rank, world, port, n_local, y, seeds, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = [island_filter(rank, world, n_local, y, s) for s in seeds]
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_merge_algebra_matches_direct():
    r = np.random.RandomState(0)
    v = r.randn(1001) * 7 - 300
    parts = np.array_split(v, 5)
    m, s, q = par.merge_lse3([par.lse3_of(p) for p in parts])
    ref = orc.Weights(lw=v.copy())
    lm, ess = par.global_stats([par.lse3_of(p) for p in parts], v.size)
    assert m == v.max()
    np.testing.assert_allclose([lm, ess], [ref.log_mean, ref.ESS], rtol=1e-13)
    # restart weights conserve the mean weight: mean over all particles of exp(restart) == 1
    w = par.lse3_of(v)
    tot = sum(np.exp(par.island_restart(par.lse3_of(p), w, 5)) * (v.size / 5) for p in parts) / v.size
    np.testing.assert_allclose(tot, 1.0, rtol=1e-12)
    assert par.merge_lse3([(-np.inf, 0.0, 0.0), par.lse3_of(v)]) == par.lse3_of(v)


def test_island_filter_world2_gloo(golden):
    """2 ranks x 4000 particles: every rank holds the same summaries, and the likelihood
    estimate agrees with the single-process (global resampling) oracle within Monte-Carlo error."""
    y = golden["data/sv_seed1_T1000"][:120]
    world, n_local, seeds = 2, 4000, list(range(6))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_local, y, seeds, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]                                  # identical decisions and logLt bits
    isl = np.array([run[-1][1] for run in got[0]])
    assert any(flag for run in got[0] for (_, _, flag) in run)
    ref = []
    for s in seeds:
        np.random.seed(500 + s)
        pf = orc.SMC(orc.Bootstrap(orc.StochVol(), [np.atleast_1d(v) for v in y]), N=world * n_local)
        pf.run()
        ref.append(pf.logLt)
    ref = np.array(ref)
    sd = max(ref.std(ddof=1), isl.std(ddof=1))
    assert abs(isl.mean() - ref.mean()) < 4 * sd * np.sqrt(2 / len(seeds)) + 0.05, (isl, ref)


def _global_worker(rank, world, port, lw_all, u, q):
    """One rank of a global systematic resampling: statistics and (standing in for the NVLink reads of
    the device kernel) the shards' CDFs travel through gloo; each rank resolves its own N/G grid points."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = lw_all.shape[0] // world
    lw = lw_all[rank * n:(rank + 1) * n]
    mine = par.lse3_of(lw)
    buf = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(buf, torch.tensor(mine, dtype=torch.float64))
    triples = [tuple(b.tolist()) for b in buf]
    goff, gpi = par.shard_shares(triples)
    cdf = np.cumsum(np.exp(lw - mine[0]) / mine[1])            # the shard's own normalised CDF (k_scan_w)
    cdfs = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(cdfs, torch.from_numpy(cdf))
    N = n * world
    su = (u + np.arange(rank * n, (rank + 1) * n)) / N         # this rank's slice of the global grid
    A = par.global_ancestors(su, goff, gpi, [c.numpy() for c in cdfs])
    q.put((rank, A, goff, gpi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("skew", [0.0, 6.0])
def test_global_resampling_two_level_search_world2_gloo(skew):
    """The union of the ranks' ancestors equals ONE systematic resampling of all particles (reference
    semantics, resampling.py:606-610) up to ties within rounding of the two-level CDF; shards with very
    different masses (skew) make many grid points cross the shard boundary."""
    world, n = 2, 5000
    r = np.random.RandomState(4)
    lw = r.randn(world * n) * 2.0
    lw[:n] += skew                                              # shard 0 holds almost all the mass
    u = r.rand(1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_global_worker, args=(k, world, port, lw, u, q)) for k in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        k, A, goff, gpi = q.get(timeout=300)
        got[k] = (A, goff, gpi)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0][1], got[1][1]) and np.array_equal(got[0][2], got[1][2])   # same offsets everywhere
    assert abs(got[0][1][-1] - 1.0) < 1e-12
    A = np.concatenate([got[0][0], got[1][0]])
    ref = orc.systematic(orc.exp_and_normalise(lw), world * n, u=u)
    assert np.all(np.diff(A) >= 0) and A.min() >= 0 and A.max() < world * n
    bad = np.flatnonzero(A != ref)
    assert bad.size <= 2 and np.all(np.abs(A[bad] - ref[bad]) <= 1), (bad.size,)
    if skew:
        assert (got[1][0] < n).mean() > 0.9                     # rank 1 pulls most ancestors from shard 0


def test_global_ancestors_skips_empty_shards():
    goff, gpi = par.shard_shares([par.lse3_of(np.zeros(4)), (-np.inf, 0.0, 0.0), par.lse3_of(np.zeros(4))])
    assert gpi[1] == 0.0 and goff[1] == goff[2] == 0.5
    cdfs = [np.arange(1, 5) / 4.0, np.full(4, np.nan), np.arange(1, 5) / 4.0]
    su = (0.3 + np.arange(12)) / 12
    A = par.global_ancestors(su, goff, gpi, cdfs)
    W = np.concatenate([np.full(4, 0.125), np.zeros(4), np.full(4, 0.125)])
    assert np.array_equal(A, orc.systematic(W, 12, u=np.array([0.3])))


def test_global_resampling_plan_of_the_sharded_sampler():
    """Host plan of ShardedAdaptiveTempering's global systematic resampling: every chain slot is served by exactly
    one rank, send / receive counts of the all_to_all are transposes of each other, and the two-level inverse CDF
    picks the same ancestors as one searchsorted over the concatenated weights."""
    import importlib.util
    import sys
    import types
    # the module imports torch-backed helpers at import time; only the NumPy planner is exercised here
    spec = importlib.util.find_spec("particles_b200.sharded_samplers")
    src = open(spec.origin).read()
    start, stop = src.index("def plan_global_resample"), src.index("class ShardedAdaptiveTempering")
    ns = {"np": np}
    exec(src[start:stop], ns)
    plan = ns["plan_global_resample"]
    rng = np.random.default_rng(5)
    for world, M, n_loc in ((2, 50, 400), (3, 33, 257), (4, 16, 128)):
        for trial in range(4):
            w = [rng.random(n_loc) ** (3 + 4 * trial) for _ in range(world)]
            if trial == 3:
                w[1][:] = 0.0                                     # an empty shard
            tot = sum(x.sum() for x in w)
            mass = np.array([x.sum() / tot for x in w])
            u = rng.random()
            Mg = world * M
            plans = [plan(mass, u, M, world, r) for r in range(world)]
            served = np.concatenate([p[0] for p in plans])
            assert np.array_equal(np.sort(served), np.arange(Mg))
            S = np.array([p[2] for p in plans]); R = np.array([p[3] for p in plans])
            assert np.array_equal(S, R.T) and S.sum() == Mg and np.all(R.sum(axis=1) == M)
            # ancestors: two-level vs flat
            flat = np.concatenate(w) / tot
            ref = np.minimum(np.searchsorted(np.cumsum(flat), (u + np.arange(Mg)) / Mg, "left"), world * n_loc - 1)
            got = np.empty(Mg, dtype=np.int64)
            for r, (mine, v, _s, _r) in enumerate(plans):
                if mine.size:
                    cdf = np.cumsum(w[r] / w[r].sum())
                    got[mine] = r * n_loc + np.minimum(np.searchsorted(cdf, v, "left"), n_loc - 1)
            assert np.mean(got != ref) < 0.02 and np.max(np.abs(got - ref)) <= 2   # rounding of the two-level CDF only
