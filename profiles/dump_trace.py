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

K, n = int(os.environ.get("TRACE_K", "60")), int(os.environ.get("TRACE_N", "10000000"))
ESS = float(os.environ.get("TRACE_ESSRMIN", "0.5"))
CFG = os.environ.get("TRACE_CONFIG", "c2")
if CFG == "c2":
    y = load_data(K)
    fk = ssm.Bootstrap(ssm=ssm.StochVol(), data=[np.atleast_1d(v) for v in y])
    sp = dict(ssm.fused_spec(fk)); sp["data"] = y.reshape(-1, 1).copy()
    eng = _FusedEngine(sp, n, os.environ.get("TRACE_SCHEME", "systematic"), ESS, 2024)
else:
    from particles_b200 import kalman, device
    device.seed(12345)
    m = ssm.BearingsOnly() if CFG == "c3i" else kalman.MVLinearGauss_Guarniero_etal(alpha=0.4, dx=4)
    _, ys = m.simulate(K)
    y = np.array([np.asarray(v.cpu()).reshape(-1) for v in ys])
    yl = [np.asarray(v).reshape(-1) for v in y]
    fk = ssm.Bootstrap(ssm=m, data=yl) if CFG == "c3i" else ssm.GuidedPF(ssm=m, data=yl)
    sp = dict(ssm.fused_spec(fk)); sp["data"] = np.ascontiguousarray(y)
    eng = _FusedEngine(sp, n, "stratified", ESS, 2024)
eng.step(K)
torch.cuda.synchronize()
lib = C.CDLL(_lib.SO_PATH)
NW = 8 * 256 + 32 * 256 + 64 * 256
buf = (C.c_ulonglong * NW)()
lib.smcb_debug_trace(buf, NW)
raw = np.array(buf[:], dtype=np.uint64).astype(np.int64)
G = int((raw[:8 * 256].reshape(256, 8)[:, 0] > 0).sum())
a = raw[:8 * 256].reshape(256, 8)[:G]
w = raw[8 * 256:40 * 256].reshape(256, 32)[:G]
parts = raw[40 * 256:].reshape(256, 16, 4)[:G]
t0 = a[:, 1].min()                                   # first CTA past the grid dependency
rs_last = bool(eng.summ.cpu().numpy()[K - 1, 2])
nw = 16 if rs_last else max(1, int((w[0] > 0).sum()))
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
if rec["rs_flag_last_step"]:      # resampling step: slots 6 / 7 = scan + scatter done / grid barrier passed
    rec["summary"]["rs_parts_ns_p50"] = {"scan_scatter": int(np.median(a[:, 6] - a[:, 2])), "barrier_wait": int(np.median(a[:, 7] - a[:, 6])),
                                         "move": int(np.median(a[:, 3] - a[:, 7])), "scan_min_max": [int((a[:, 6] - a[:, 2]).min()), int((a[:, 6] - a[:, 2]).max())],
                                         "move_min_max": [int((a[:, 3] - a[:, 7]).min()), int((a[:, 3] - a[:, 7]).max())]}
else:
    rec["summary"]["prologue_parts_ns_p50"] = {"shard_totals": int(np.median(a[:, 6] - a[:, 1])), "scalars": int(np.median(a[:, 7] - a[:, 6])),
                                                "table_wait_and_rest": int(np.median(a[:, 2] - a[:, 7]))}
rec["summary"]["warp_done"] = [int(wd.min()), int(np.median(wd)), int(wd.max())]
rec["summary"]["within_cta_warp_spread_p50"] = int(np.median(wd.max(axis=1) - wd.min(axis=1)))
if rs_last:                      # words 16.. of a CTA's warp area: (settle calls << 32) | largest hint error, per warp
    stat = w[:, 16:32]
    calls, far = (stat >> 32).sum(axis=1), (stat & 0xffffffff).max(axis=1)
    rec["settle_calls_per_cta"], rec["max_hint_error_per_cta"] = calls.tolist(), far.tolist()
    slow = np.argsort(a[:, 3])[-4:]
    rec["summary"]["settle"] = {"calls_total": int(calls.sum()), "calls_per_cta_p50_max": [int(np.median(calls)), int(calls.max())],
                                "max_hint_error": int(far.max()),
                                "slowest_ctas": [[int(c), int(a[c, 3] - a[c, 7]), int(calls[c]), int(far[c])] for c in slow]}
if rs_last:
    rec["round_parts_max_ns"] = parts[:, :, :3].tolist()
    late = np.argwhere(wd > np.median(wd) + 20000)
    rec["summary"]["late_warps"] = [[int(c), int(wi), int(wd[c, wi])] + parts[c, wi, :3].tolist() for c, wi in late[:12]]
    rec["summary"]["round_parts_max_ns_p50"] = np.median(parts[:, :, :3].reshape(-1, 3), axis=0).tolist()
print(json.dumps(rec["summary"]))
json.dump(rec, open(sys.argv[1], "w"))
