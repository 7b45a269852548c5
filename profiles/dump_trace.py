"""profiles/dump_trace.py out.json -- per-CTA timeline of the last step-kernel launch (trace build only:
profiles/build_variant.sh trace -DSMCB_TRACE; SMCB_LIB=particles_b200/variants/libsmcb_trace.so)."""
import ctypes as C
import json
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from particles_b200 import _lib, state_space_models as ssm
from particles_b200.core import _FusedEngine
from bench import load_data

K, n = 60, 10_000_000
y = load_data(K)
fk = ssm.Bootstrap(ssm=ssm.StochVol(), data=[np.atleast_1d(v) for v in y])
sp = dict(ssm.fused_spec(fk)); sp["data"] = y.reshape(-1, 1).copy()
eng = _FusedEngine(sp, n, "systematic", 0.5, 2024)
eng.step(K)
torch.cuda.synchronize()
lib = C.CDLL(_lib.SO_PATH)
NW = 8 * 256 + 32 * 256
buf = (C.c_ulonglong * NW)()
lib.smcb_debug_trace(buf, NW)
raw = np.array(buf[:], dtype=np.uint64).astype(np.int64)
G = int((raw[:8 * 256].reshape(256, 8)[:, 0] > 0).sum())
a = raw[:8 * 256].reshape(256, 8)[:G]
w = raw[8 * 256:].reshape(256, 32)[:G]
t0 = a[:, 1].min()                                   # first CTA past the grid dependency
nw = int((w[0] > 0).sum())
names = ["start", "dep_resolved", "prologue_done", "loop_done", "exit"]
rec = {"grid": G, "warps": nw, "smid": a[:, 5].tolist(), "rs_flag_last_step": float(eng.summ.cpu().numpy()[K - 1, 2])}
for i, nm in enumerate(names):
    rec[nm + "_ns"] = (a[:, i] - t0).tolist()
rec["warp_done_ns"] = (w[:, :nw] - t0).tolist()
wd = np.array(rec["warp_done_ns"])
rec["summary"] = {"grid": G}
for nm in names:
    v = np.array(rec[nm + "_ns"])
    rec["summary"][nm] = [int(v.min()), int(np.median(v)), int(v.max())]
rec["summary"]["prologue_ns_p50"] = int(np.median(np.array(rec["prologue_done_ns"]) - np.array(rec["dep_resolved_ns"])))
rec["summary"]["warp_done"] = [int(wd.min()), int(np.median(wd)), int(wd.max())]
rec["summary"]["within_cta_warp_spread_p50"] = int(np.median(wd.max(axis=1) - wd.min(axis=1)))
print(json.dumps(rec["summary"]))
json.dump(rec, open(sys.argv[1], "w"))
