"""profiles/bench_standalone.py [out.json] -- the stand-alone kernels the PLUGIN path runs on (user-defined
Feynman-Kac models): device time (CUDA events, best of 5 after warm-up) at N = 1e7 against their algorithmic
bytes and the measured HBM peak; plus a plugin-path end-to-end line (the README ToySSM written against
particles_b200.distributions, not a fused model).  north_star names the weight kernel and the prefix sum."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import particles_b200 as pb  # noqa: E402
from particles_b200 import distributions as dists, resampling as rs, state_space_models as ssm  # noqa: E402
from particles_b200 import _lib  # noqa: E402
from particles_b200.device import context, empty, ptr  # noqa: E402

N = 10_000_000
peak = 6650.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
ctx = context()
lib = ctx.lib
g = torch.Generator(device="cuda").manual_seed(0)
lw = torch.randn(N, dtype=torch.float64, device="cuda", generator=g) * 2.0
W = torch.softmax(lw, 0)
x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
x4 = torch.randn((4, N), dtype=torch.float64, device="cuda", generator=g)
A = torch.sort(torch.randint(0, N, (N,), device="cuda", generator=g)).values
out = {}


def timeit(name, fn, nbytes, reps=5, inner=8):
    """Device time per call: `inner` calls between one pair of events (a single short kernel between two events also
    measures the host's launch latency -- the GPU idles until the launch arrives), best of `reps`; the arrays are
    80 MB each, so successive calls do not find their input in the 126 MB L2 once two or more arrays are touched."""
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _i in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    gbs = nbytes / (best * 1e-3) / 1e9
    out[name] = {"us": 1e3 * best, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak}
    print(name, out[name], flush=True)


st = empty(4); Wo = empty(N); cdf = empty(N); o = empty(N); Ao = torch.empty(N, dtype=torch.int64, device="cuda")
scratch = empty(int(lib.smcb_resample_scratch_doubles(N, N)))
lw2 = lw.clone()
_flip = [0]


def _stats_only():          # alternate two 80 MB inputs: one alone would sit in the 126 MB L2 between calls
    _flip[0] ^= 1
    _lib.check(lib.smcb_normalise(ctx.handle, ptr(lw2 if _flip[0] else lw), N, ptr(None), ptr(st)))


timeit("smcb_normalise (stats only: the weight kernel)", _stats_only, 8 * N)
timeit("smcb_normalise (+ W)", lambda: _lib.check(lib.smcb_normalise(ctx.handle, ptr(lw), N, ptr(Wo), ptr(st))), 24 * N)
timeit("smcb_exp_and_normalise", lambda: _lib.check(lib.smcb_exp_and_normalise(ctx.handle, ptr(lw), N, ptr(Wo))), 24 * N)
timeit("smcb_cumsum (the prefix sum)", lambda: _lib.check(lib.smcb_cumsum(ctx.handle, ptr(W), N, ptr(cdf))), 16 * N)
timeit("smcb_resample systematic (scan + search)", lambda: _lib.check(lib.smcb_resample(ctx.handle, 2, ptr(W), N, N, ptr(Ao), ptr(None), ptr(scratch))), 32 * N)
timeit("smcb_resample stratified", lambda: _lib.check(lib.smcb_resample(ctx.handle, 1, ptr(W), N, N, ptr(Ao), ptr(None), ptr(scratch))), 32 * N)
timeit("smcb_resample multinomial", lambda: _lib.check(lib.smcb_resample(ctx.handle, 0, ptr(W), N, N, ptr(Ao), ptr(None), ptr(scratch))), 48 * N)
timeit("smcb_gather d=1 (sorted ancestors)", lambda: _lib.check(lib.smcb_gather(ctx.handle, ptr(x), N, ptr(A), N, 1, ptr(o))), 24 * N)
o4 = torch.empty_like(x4)
timeit("smcb_gather d=4", lambda: _lib.check(lib.smcb_gather(ctx.handle, ptr(x4), N, ptr(A), N, 4, ptr(o4))), (8 + 64) * N)
timeit("smcb_normal_rvs (loc array)", lambda: _lib.check(lib.smcb_normal_rvs(ctx.handle, ptr(x), 0.0, ptr(None), 0.5, ptr(None), ptr(o), N)), 16 * N)
timeit("smcb_normal_logpdf (loc array)", lambda: _lib.check(lib.smcb_normal_logpdf(ctx.handle, ptr(x), 0.0, ptr(lw), 0.0, ptr(None), 0.5, ptr(o), N)), 24 * N)
timeit("smcb_wmean_and_var d=1", lambda: _lib.check(lib.smcb_wmean_and_var(ctx.handle, ptr(W), ptr(x), N, 1, ptr(st))), 16 * N)
mvn = dists.MvNormal(loc=x4.t().contiguous(), cov=np.eye(4) * 0.25 + 0.05)
timeit("MvNormal.rvs d=4 (Python call)", lambda: mvn.rvs(size=N), 64 * N)
xs = mvn.rvs(size=N)
timeit("MvNormal.logpdf d=4 (Python call)", lambda: mvn.logpdf(xs), 72 * N)


# plugin path end to end: the README's ToySSM, a USER model (no fused kernel)
class ToySSM(ssm.StateSpaceModel):
    def PX0(self):
        return dists.Normal()

    def PX(self, t, xp):
        return dists.Normal(loc=xp)

    def PY(self, t, xp, x):
        return dists.Normal(loc=x, scale=self.sigma)


m = ToySSM(sigma=0.2)
pb.seed(1)
_, ys = m.simulate(100)
for n_ in (1_000_000, 10_000_000):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pf = pb.SMC(fk=ssm.Bootstrap(ssm=m, data=ys), N=n_, seed=3)
        pf.run(); ll = pf.logLt
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[f"plugin path ToySSM bootstrap N={n_} T=100"] = {"seconds": dt, "particle_steps_per_s": n_ * 100 / dt, "fused": pf.fused,
                                                          "resamplings": int(sum(pf.summaries.rs_flags)), "logLt": ll}
    print(out[f"plugin path ToySSM bootstrap N={n_} T=100"], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
