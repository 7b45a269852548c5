"""profiles/measure_peaks.py [out.json] -- builder-side ceilings of the device the bench runs on:
fp64 FMA issue peak (smcb_measure_fp64_peak) and a 16-byte read+write streaming probe, next to the
driver's MEASURED_PEAKS.json numbers.  Run under gpurun; bench.py calls the same entry points."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sm_clock_now():
    try:
        o = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                     "-i", "0"], text=True)
        a, b = [float(x) for x in o.strip().split(",")]
        return a, b
    except Exception:
        return None, None


def measure():
    import torch
    from particles_b200 import _lib, device
    ctx = device.context()
    lib = _lib.load()
    out3 = (C.c_double * 3)()
    _lib.check(lib.smcb_measure_fp64_peak(ctx.handle, 0.0, out3))          # warm the clocks
    sm, smax = sm_clock_now()
    _lib.check(lib.smcb_measure_fp64_peak(ctx.handle, float(smax or 0.0), out3))
    n = 1 << 27                                                            # 1 GiB each way
    a = torch.zeros(n, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    out1 = (C.c_double * 1)()
    _lib.check(lib.smcb_measure_stream_peak(ctx.handle, device.ptr(a), device.ptr(b), n, out1))
    res = {"fp64_tflops": out3[0], "dfma_warp_inst_per_cycle_per_sm_at_max_clock": out3[1], "dfma_kernel_ms": out3[2],
           "stream16_gbs": out1[0], "sm_mhz_idle_sample": sm, "sm_max_mhz": smax,
           "how": "k_dfma: 8 independent DFMA chains/thread, 148x8 CTAs x 256 threads, 20000 links, best of 4; "
                  "k_stream: 16-byte read + write over 2 x 1 GiB, grid 148x8, best of 5"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        res["driver_measured"] = {k: v for k, v in json.load(open(p)).items() if k in ("hbm_gbs", "sm_max_mhz")}
    return res


if __name__ == "__main__":
    r = measure()
    s = json.dumps(r, indent=1)
    print(s)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)
