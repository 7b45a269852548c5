"""profiles/bench_configs.py -- wall-clock of the other BASELINE configs through the public API
(not bench lines: config 2 is the bench; these numbers are for DESIGN.md).  Each GPU figure is the
second of two runs (the first warms the allocator / JIT-free kernels); the CPU figure is the oracle's
NumPy restatement on a reduced problem, scaled per particle-step (or per likelihood evaluation)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import particles_b200 as pb  # noqa: E402
from particles_b200 import kalman, smc_samplers as ssp, state_space_models as ssm  # noqa: E402
from oracle import samplers_numpy as osp  # noqa: E402
from oracle import smc_numpy as orc  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "golden_stats.npz"))
out = {}


def timed(make, reps=2):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pf = make()
        pf.run()
        _ = pf.logLt
        torch.cuda.synchronize()
        best = time.perf_counter() - t0
    return best, pf


def cpu(make_fk, N, nsteps):
    np.random.seed(0)
    pf = orc.SMC(make_fk(), N=N, resampling="stratified")
    pf.step()
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        pf.run(nsteps=nsteps)
    return N * nsteps / (time.perf_counter() - t0)


# C1: ToySSM-as-LinearGauss, N = 1000, T = 200 (launch-bound regime)
gx = np.load(os.path.join(ROOT, "tests", "golden", "golden_exact.npz"))
yt = [np.atleast_1d(v) for v in gx["data/toy_seed0_T200"]]
dt, pf = timed(lambda: pb.SMC(fk=ssm.Bootstrap(ssm=kalman.LinearGauss(rho=1.0, sigmaX=1.0, sigmaY=0.2, sigma0=1.0),
                                                data=yt), N=1000, seed=1))
out["C1 toy N=1e3 T=200"] = {"seconds": dt, "us_per_step": 1e6 * dt / 200, "logLt": pf.logLt}

# C3 (i): BearingsOnly bootstrap, N = 1e6, T = 500 (data cycled), stratified
yb = list(np.tile(g["data/bearings_seed0_T40"], 13)[:500].reshape(-1, 1))
dt, pf = timed(lambda: pb.SMC(fk=ssm.Bootstrap(ssm=ssm.BearingsOnly(), data=yb), N=1_000_000,
                              resampling="stratified", seed=2))
c = cpu(lambda: orc.Bootstrap(orc.BearingsOnly(), yb[:40]), 100_000, 10)
out["C3i bearings boot N=1e6 T=500"] = {"seconds": dt, "particle_steps_per_s": 1e6 * 500 / dt,
                                         "cpu_port_particle_steps_per_s": c, "resamplings": int(sum(pf.summaries.rs_flags))}

# C3 (ii): 4-D MvNormal guided filter (Guarniero et al), N = 1e6, T = 500, stratified
ym = list(np.tile(g["data/mvlg_seed5_T30"], (17, 1))[:500])
mv = kalman.MVLinearGauss_Guarniero_etal(alpha=0.4, dx=4)
for name, cls, ocls in [("guided", ssm.GuidedPF, orc.GuidedPF), ("apf", ssm.AuxiliaryPF, orc.AuxiliaryPF)]:
    dt, pf = timed(lambda: pb.SMC(fk=cls(ssm=mv, data=ym), N=1_000_000, resampling="stratified", seed=3))
    c = cpu(lambda: ocls(orc.MVLinearGauss_Guarniero_etal(0.4, 4), ym[:30]), 100_000, 8)
    out[f"C3ii mvlg {name} N=1e6 T=500"] = {"seconds": dt, "particle_steps_per_s": 1e6 * 500 / dt,
                                            "cpu_port_particle_steps_per_s": c,
                                            "resamplings": int(sum(pf.summaries.rs_flags))}

# C5: waste-free adaptive tempering, 20-D logistic regression, n_data = 1000, M x P = 1e4 x 100 = 1e6
data = osp.synthetic_logistic(1000, 20, seed=0)
model = ssp.LogisticRegression(data=data, prior_scale=5.0)
dt, pf = timed(lambda: pb.SMC(fk=ssp.AdaptiveTempering(model=model, ESSrmin=0.5, wastefree=True, len_chain=100),
                              N=10_000, ESSrmin=1.0, seed=4), reps=2)
nsteps = len(pf.summaries.ESSs)
evals = 1e6 * nsteps                      # likelihood evaluations (each = 1000 logistic terms)
th = np.random.RandomState(0).randn(20_000, 20)
t0 = time.perf_counter()
osp.LogisticModel(data).loglik(th)
cpu_eval = 20_000 / (time.perf_counter() - t0)
out["C5 tempering d=20 n=1000 MxP=1e4x100"] = {"seconds": dt, "tempering_steps": nsteps, "logLt": pf.logLt,
                                               "likelihood_evals_per_s": evals / dt,
                                               "cpu_port_likelihood_evals_per_s": cpu_eval}
print(json.dumps(out, indent=1))
