"""profiles/bench_c5.py [out.json] -- BASELINE config 5 on ONE GPU: waste-free adaptive tempering of a 20-D logistic
regression posterior (n_data = 1000), M = 1e4 chains x P = 100 = 1e6 particles, through the public API
(particles_b200.SMC(fk=AdaptiveTempering(...))).  Reports tempering steps, wall time (median of 3), likelihood
evaluations per second and the fraction of the measured fp64 peak they correspond to (one evaluation = 1000 data
rows x (20-term dot product + softplus); fp64 instructions per row counted from the kernel's SASS: see DESIGN.md)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import particles_b200 as pb  # noqa: E402
from particles_b200 import smc_samplers as ssp  # noqa: E402
from oracle import samplers_numpy as osp  # noqa: E402  (synthetic data generator only)

data = osp.synthetic_logistic(1000, 20, seed=0)
model = ssp.LogisticRegression(data=data, prior_scale=5.0)
runs = []
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pf = pb.SMC(fk=ssp.AdaptiveTempering(model=model, ESSrmin=0.5, wastefree=True, len_chain=100), N=10_000, ESSrmin=1.0,
                seed=4 + rep)
    pf.run()
    torch.cuda.synchronize()
    runs.append((time.perf_counter() - t0, len(pf.summaries.ESSs), pf.logLt))
runs = runs[1:]                                   # first call: warm-up
dt = float(np.median([r[0] for r in runs]))
nsteps = runs[0][1]
evals = 1e6 * nsteps
FP64_PER_ROW = 2 * 20 + 60                        # mul+add per feature (-fmad=false) + softplus (exp, log1p)
peak = None
try:
    import ctypes as C
    from particles_b200.device import context
    o = (C.c_double * 3)()
    context().lib.smcb_measure_fp64_peak(context().handle, 0.0, o)
    peak = o[0]
except Exception:
    pass
out = {"seconds_median_of_3": dt, "seconds_all": [r[0] for r in runs], "tempering_steps": nsteps, "logLt": [r[2] for r in runs],
       "likelihood_evals_per_s": evals / dt,
       "fp64_tflops_equiv": evals / dt * 1000 * FP64_PER_ROW * 2 / 1e12, "fp64_peak_tflops_measured": peak}
if peak:
    out["frac_of_fp64_peak"] = out["fp64_tflops_equiv"] / peak
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
