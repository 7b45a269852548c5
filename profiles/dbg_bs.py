import sys, os, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from particles_b200 import state_space_models as ssm
from particles_b200.core import _FusedEngine
from bench import load_data
K, n = 400, 10_000_000
y = load_data(K)
fk = ssm.Bootstrap(ssm=ssm.StochVol(), data=[np.atleast_1d(v) for v in y])
sp = dict(ssm.fused_spec(fk)); sp["data"] = y.reshape(-1, 1).copy()
eng = _FusedEngine(sp, n, "systematic", 0.5, 2024)
out = {"rows": [], "xsum": [], "x0": []}
for t in range(K):
    eng.step(1)
    torch.cuda.synchronize()
    X = eng.X[t & 1]
    out["xsum"].append(float(X.sum().item()))
    out["x0"].append([float(v) for v in X[:4].cpu()] + [float(v) for v in X[-4:].cpu()])


out["rows"] = eng.summ.cpu().numpy().tolist()
json.dump(out, open(sys.argv[1], "w"))
print("done", out["rows"][K-1])
