import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import particles_b200 as pb
from particles_b200 import state_space_models as ssm
from oracle import smc_numpy as orc
g = np.load(os.path.join(ROOT, "tests", "golden", "golden_exact.npz"))
N = 2000
y = list(g["data/bearings_seed0_T40"].reshape(-1, 1)); T = len(y)
r = np.random.RandomState(5); z = r.standard_normal((T, N, 2)); u = r.rand(T, N + 1)
pf = pb.SMC(fk=ssm.Bootstrap(ssm=ssm.BearingsOnly(), data=y), N=N, resampling="stratified", ESSrmin=0.5,
            noise=(np.ascontiguousarray(z.transpose(0, 2, 1)), u), fused=True)
with np.errstate(all="ignore"):
    ref = orc.SMC(orc.Bootstrap(orc.BearingsOnly(), y), N=N, resampling="stratified", ESSrmin=0.5,
                  noise=orc.InjectedNoise(z, [row[:N] for row in u]))
    for t in range(T):
        next(pf); ref.step()
        if ref.rs_flag:
            A = pf.A.cpu().numpy(); bad = np.flatnonzero(A != ref.A)
            cdf = pf._engine.cdf.cpu().numpy()
            su = (u[t][:N] + np.arange(N)) / N
            own = np.minimum(np.searchsorted(cdf, su, "left"), N - 1)
            print("t", t, "rs; bad vs oracle", bad.size, "bad vs own searchsorted", int((A != own).sum()), "oracle vs own", int((ref.A != own).sum()))
            if bad.size:
                b = bad[:8]; print(" idx", b, "A", A[b], "ref", ref.A[b], "own", own[b])
                W = ref.aux.W if hasattr(ref, "aux") and ref.aux is not None else None
                print(" max offspring", np.bincount(ref.A).max(), "cdf[-3:]", cdf[-3:], "su[-1]", su[-1])
                break
