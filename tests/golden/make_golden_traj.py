"""Per-step log-likelihood trajectories of the LIVE reference on BASELINE config 2's data (bootstrap filter of
StochVol, systematic resampling, ESSrmin 0.5, N = 1e5, the 8 seeds of golden_stats.npz): the Monte-Carlo anchor
for the 3-sigma parity statement of bench.py / tests at ANY number of steps K <= 1000.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden_traj.py

Writes tests/golden/golden_sv_traj.npz: logLts (8, 1000), rs_cum (8, 1000) cumulative resampling counts.  The
final column must equal golden_stats.npz stat/sv_T1000_N100000/logLt (checked here)."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import particles  # noqa: E402
from particles import state_space_models as ssm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
g = np.load(os.path.join(HERE, "golden_stats.npz"))
ys = g["data/sv_seed1_T1000"]
N, nrep = 100000, 8
ll, rc = [], []
for r in range(nrep):
    np.random.seed(1000 + r)
    pf = particles.SMC(fk=ssm.Bootstrap(ssm=ssm.StochVol(), data=ys), N=N)
    pf.run()
    ll.append(np.array(pf.summaries.logLts))
    rc.append(np.cumsum(np.array(pf.summaries.rs_flags, dtype=np.int64)))
    print(r, ll[-1][-1], rc[-1][-1], flush=True)
ll, rc = np.array(ll), np.array(rc)
assert np.array_equal(ll[:, -1], g["stat/sv_T1000_N100000/logLt"]), "not the runs of golden_stats.npz"
np.savez_compressed(os.path.join(HERE, "golden_sv_traj.npz"), logLts=ll, rs_cum=rc, N=np.array([N]))
print("wrote golden_sv_traj.npz", ll.shape)
