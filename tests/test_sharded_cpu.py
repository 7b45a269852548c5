"""world_size-2 gloo tests (CPU) of the sharded-filter protocol of particles_b200/parallel.py:
the per-step all-gather of (max, sum exp, sum exp^2) triples, the rank-order merge, the
island restart log-weight.  The shard arithmetic is done by the oracle (NumPy) here; the
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
