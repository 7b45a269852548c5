"""Golden vectors for the waste-free adaptive-tempering sampler (config 5), from the LIVE reference.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden_tempering.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import particles  # noqa: E402
from particles import distributions as dists  # noqa: E402
from particles import smc_samplers as ssp  # noqa: E402
from oracle.samplers_numpy import synthetic_logistic  # noqa: E402  (data generator only)

HERE = os.path.dirname(os.path.abspath(__file__))


def make_model(data):
    d = data.shape[1]
    prior = dists.StructDist({"beta": dists.MvNormal(scale=5.0, cov=np.eye(d))})

    class LogisticRegression(ssp.StaticModel):          # book/smc_samplers/logistic_reg.py:63-67
        def logpyt(self, theta, t):
            lin = np.matmul(theta["beta"], data[t, :])
            return -np.logaddexp(0.0, -lin)

    return LogisticRegression(data=data, prior=prior)


def run(data, N, P, seed):
    np.random.seed(seed)
    fk = ssp.AdaptiveTempering(model=make_model(data), ESSrmin=0.5, wastefree=True, len_chain=P)
    pf = particles.SMC(fk=fk, N=N, ESSrmin=1.0)
    pf.run()
    return pf


if __name__ == "__main__":
    g = {}
    data = synthetic_logistic(150, 4, seed=3)
    g["exact/data"] = data
    pf = run(data, 100, 8, 17)
    g["exact/logLt"] = np.array([pf.logLt])
    g["exact/logLts"] = np.array(pf.summaries.logLts)
    g["exact/ESSs"] = np.array(pf.summaries.ESSs)
    g["exact/exponents"] = np.array(pf.X.shared["exponents"])
    g["exact/theta"] = pf.X.theta["beta"]
    g["exact/lpost"] = pf.X.lpost
    g["exact/path_sampling"] = np.array([pf.X.shared["path_sampling"][-1]])
    g["exact/meta"] = np.array([100, 8, 17])
    # Monte-Carlo anchors: d = 6, n_data = 300, N = 400 chains x P = 25 (1e4 particles)
    data2 = synthetic_logistic(300, 6, seed=4)
    g["stat/data"] = data2
    lls, means, nsteps = [], [], []
    for r in range(12):
        pf = run(data2, 400, 25, 100 + r)
        lls.append(pf.logLt)
        means.append(np.average(pf.X.theta["beta"], weights=pf.W, axis=0))
        nsteps.append(len(pf.summaries.ESSs))
    g["stat/logLt"] = np.array(lls)
    g["stat/post_mean"] = np.array(means)
    g["stat/nsteps"] = np.array(nsteps)
    g["stat/meta"] = np.array([400, 25])
    np.savez_compressed(os.path.join(HERE, "golden_tempering.npz"), **g)
    print("exact logLt", g["exact/logLt"], "steps", len(g["exact/ESSs"]))
    print("stat logLt mean/sd", np.mean(lls), np.std(lls, ddof=1), "steps", nsteps)
