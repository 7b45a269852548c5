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
NW = 4 * 256 + 32 * 256
buf = (C.c_ulonglong * NW)()
lib.smcb_debug_trace(buf, NW)
raw = np.array(buf[:], dtype=np.uint64).astype(np.int64)
G = int((raw[:4 * 256].reshape(256, 4)[:, 0] > 0).sum())
a = raw[:4 * 256].reshape(256, 4)[:G]
w = raw[4 * 256:].reshape(256, 32)[:G]
t0 = a[:, 0].min()
nw = int((w[0] > 0).sum())
rec = {"grid": G, "warps": nw, "start_ns": (a[:, 0] - t0).tolist(), "prologue_done_ns": (a[:, 1] - t0).tolist(),
       "loop_done_ns": (a[:, 2] - t0).tolist(), "smid": a[:, 3].tolist(),
       "warp_done_ns": (w[:, :nw] - t0).tolist(), "rs_flag_last_step": float(eng.summ.cpu().numpy()[K - 1, 2])}
ld = np.array(rec["loop_done_ns"]); pr = np.array(rec["prologue_done_ns"]); wd = np.array(rec["warp_done_ns"])
rec["summary"] = {"grid": G, "start_spread_ns": int(max(rec["start_ns"])), "prologue_done_p50": int(np.median(pr)), "prologue_done_max": int(pr.max()),
                  "loop_done_min": int(ld.min()), "loop_done_p50": int(np.median(ld)), "loop_done_max": int(ld.max()),
                  "warp_done_min": int(wd.min()), "warp_done_p50": int(np.median(wd)), "warp_done_max": int(wd.max()),
                  "within_cta_warp_spread_p50": int(np.median(wd.max(axis=1) - wd.min(axis=1)))}
print(json.dumps(rec["summary"]))
json.dump(rec, open(sys.argv[1], "w"))
