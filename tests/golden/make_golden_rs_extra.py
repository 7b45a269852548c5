"""Golden vectors for the resampling schemes outside the fused path (ssp, killing), generated from the
live reference:  PYTHONPATH=/root/reference python tests/golden/make_golden_rs_extra.py
Weights come from golden_exact.npz (same cases as the other schemes); each entry stores the ancestors
the reference returned under np.random.seed(99) and the uniforms that call consumed, in draw order.
`ssp` is numba-compiled in the reference (its own RNG state); the pure-Python body of the same function
(`.py_func`) draws from the seeded NumPy stream and is what is recorded here."""
import os
import warnings

import numpy as np

warnings.filterwarnings("ignore")
from particles import resampling as rs  # noqa: E402  (the reference)

HERE = os.path.dirname(os.path.abspath(__file__))
g = np.load(os.path.join(HERE, "golden_exact.npz"))
names = sorted({k.split("/")[1] for k in g.files if k.startswith("rs/") and k.endswith("/W")})
out = {}
ssp_py = rs.ssp.__wrapped__.py_func
for name in names:
    W, M = g[f"rs/{name}/W"], int(g[f"rs/{name}/M"][0])
    N = W.shape[0]
    np.random.seed(99)
    try:
        out[f"rs/{name}/ssp/A"] = ssp_py(W, M)
    except ValueError as e:
        out[f"rs/{name}/ssp/error"] = np.frombuffer(str(e).encode(), dtype=np.uint8)
    np.random.seed(99)
    out[f"rs/{name}/ssp/u"] = np.random.rand(N - 1)
    if M == N:
        np.random.seed(99)
        out[f"rs/{name}/killing/A"] = rs.killing(W, M)
        np.random.seed(99)
        u = np.random.rand(N)
        nk = int((u * W.max() >= W).sum())
        out[f"rs/{name}/killing/u"] = u
        out[f"rs/{name}/killing/u_multinomial"] = np.random.rand(nk + 1)
try:
    rs.killing(g["rs/M_lt_N/W"], 10)
except ValueError as e:
    out["rs/killing_error"] = np.frombuffer(str(e).encode(), dtype=np.uint8)
np.savez_compressed(os.path.join(HERE, "golden_rs_extra.npz"), **out)
print(len(out), "arrays;", [k for k in out if k.endswith("error")])
