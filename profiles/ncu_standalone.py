"""One launch of each stand-alone weight / prefix-sum / search kernel at N = 1e7 (profiles/ncu_standalone.sh wraps
this in `ncu --set full`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from particles_b200 import _lib  # noqa: E402
from particles_b200.device import context, empty, ptr  # noqa: E402

N = 10_000_000
ctx = context()
lib = ctx.lib
g = torch.Generator(device="cuda").manual_seed(0)
lw = torch.randn(N, dtype=torch.float64, device="cuda", generator=g) * 2.0
W = torch.softmax(lw, 0)
st, Wo, cdf = empty(4), empty(N), empty(N)
Ao = torch.empty(N, dtype=torch.int64, device="cuda")
scratch = empty(int(lib.smcb_resample_scratch_doubles(N, N)))
x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
o2 = empty(2)
for _ in range(2):       # the second round is the one to read (first: cold instruction caches / clocks)
    _lib.check(lib.smcb_normalise(ctx.handle, ptr(lw), N, ptr(Wo), ptr(st)))        # k_lse + k_exp_normalise
    _lib.check(lib.smcb_cumsum(ctx.handle, ptr(W), N, ptr(cdf)))                    # k_scan_sums + k_scan_chunks
    _lib.check(lib.smcb_resample(ctx.handle, 2, ptr(W), N, N, ptr(Ao), ptr(None), ptr(scratch)))   # + k_search_*
    _lib.check(lib.smcb_wmean_and_var(ctx.handle, ptr(W), ptr(x), N, 1, ptr(o2)))   # k_wmoments
torch.cuda.synchronize()
print("done", float(st[1]), float(cdf[-1]), int(Ao[-1]))
