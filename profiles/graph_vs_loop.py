"""profiles/graph_vs_loop.py -- us per filter step of the CUDA-graph step loop (WHILE/IF nodes) against
the launch-per-step loop, over N (StochVol bootstrap, T = 1000, ESSrmin 0.5).  Run twice per setting."""
import os, subprocess, sys, json
code = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import particles_b200 as pb
from particles_b200 import state_space_models as ssm
g = np.load("tests/golden/golden_stats.npz"); y = [np.atleast_1d(v) for v in g["data/sv_seed1_T1000"]]
out = {}
for N in (1000, 10_000, 100_000, 1_000_000, 10_000_000):
    best = 1e9
    for rep in range(3):
        pf = pb.SMC(fk=ssm.Bootstrap(ssm=ssm.StochVol(), data=y), N=N, seed=rep)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pf._engine.step(1000); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    out[N] = 1e6 * best / 1000
print(out)
'''
for env in ({}, {"SMCB_NO_GRAPH": "1"}):
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env))
    print("graph" if not env else "loop ", r.stdout.strip(), r.stderr[-300:])
