"""MVStochVol (particles/state_space_models.py:633-654) through the LIVE reference: the model ships without default
parameters, so the fixture fixes a 3-dimensional instance.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden_mvsv.py

Writes tests/golden/golden_mvsv.npz: the parameters, data simulated by the reference (seed 11, T = 40), one seeded
bootstrap-filter run at N = 300 (per-step logLt / ESS / rs flags: the oracle must reproduce it from the same
numpy.random stream) and logLt of 30 runs at N = 2000 (the 3-sigma anchor for the GPU plugin path)."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import particles  # noqa: E402
from particles import state_space_models as ssm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
d = 3
mu = np.array([-1.0, -0.5, 0.2])
F = np.array([[0.9, 0.05, 0.0], [0.0, 0.8, 0.1], [0.05, 0.0, 0.7]])
covX = 0.09 * (0.6 * np.eye(d) + 0.4 * np.ones((d, d)))
corY = np.array([[1.0, 0.3, -0.2], [0.3, 1.0, 0.1], [-0.2, 0.1, 1.0]])
T = 40
model = ssm.MVStochVol(mu=mu, covX=covX, corY=corY, F=F)
np.random.seed(11)
_, ys = model.simulate(T)
y = np.array([np.asarray(v).reshape(-1) for v in ys])
fk = ssm.Bootstrap(ssm=model, data=ys)
np.random.seed(21)
pf = particles.SMC(fk=fk, N=300)
pf.run()
seeded = dict(logLts=np.array(pf.summaries.logLts), ESSs=np.array(pf.summaries.ESSs),
              rs=np.array(pf.summaries.rs_flags, dtype=np.int64), X=pf.X.copy())
lls = []
for r in range(30):
    np.random.seed(500 + r)
    p2 = particles.SMC(fk=ssm.Bootstrap(ssm=model, data=ys), N=2000)
    p2.run()
    lls.append(p2.logLt)
np.savez_compressed(os.path.join(HERE, "golden_mvsv.npz"), mu=mu, F=F, covX=covX, corY=corY, y=y,
                    seeded_logLts=seeded["logLts"], seeded_ESSs=seeded["ESSs"], seeded_rs=seeded["rs"],
                    seeded_X=seeded["X"], stat_logLt_N2000=np.array(lls))
print("wrote golden_mvsv.npz", y.shape, seeded["logLts"][-1], np.mean(lls), np.std(lls, ddof=1))
