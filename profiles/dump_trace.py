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
buf = (C.c_ulonglong * (4 * 2048))()
lib.smcb_debug_trace(buf, 4 * 2048)
a = np.array(buf[:], dtype=np.uint64).reshape(2048, 4)[:444].astype(np.int64)
t0 = a[:, 0].min()
rec = {"start_ns": (a[:, 0] - t0).tolist(), "loop_done_ns": (a[:, 1] - t0).tolist(), "exit_ns": (a[:, 2] - t0).tolist(),
       "smid": a[:, 3].tolist(), "rs_flag_last_step": float(eng.summ.cpu().numpy()[K - 1, 2])}
ld = np.array(rec["loop_done_ns"]); ex = np.array(rec["exit_ns"])
rec["summary"] = {"start_spread_ns": int(max(rec["start_ns"])), "loop_done_min": int(ld.min()), "loop_done_p50": int(np.median(ld)),
                  "loop_done_max": int(ld.max()), "exit_max": int(ex.max())}
print(json.dumps(rec["summary"]))
json.dump(rec, open(sys.argv[1], "w"))
