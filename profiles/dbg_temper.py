import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import particles_b200 as pb
from particles_b200 import smc_samplers as ssp
from oracle import samplers_numpy as osp
data = osp.synthetic_logistic(1000, 20, seed=0)
for fused in (True, False):
    model = ssp.LogisticRegression(data=data, prior_scale=5.0)
    if not fused:
        model_wf = model.wf_move
        ssp.LogisticRegression.wf_move_disabled = True
        fk = ssp.AdaptiveTempering(model=model, ESSrmin=0.5, wastefree=True, len_chain=100)
        orig = fk.current_target
        def ct(epn, orig=orig):
            f = orig(epn)
            if hasattr(f, "fused_wf"): del f.fused_wf
            return f
        fk.current_target = ct
    else:
        fk = ssp.AdaptiveTempering(model=model, ESSrmin=0.5, wastefree=True, len_chain=100)
    pf = pb.SMC(fk=fk, N=10000, ESSrmin=1.0, seed=4)
    pf.run()
    ars = pf.X.shared["acc_rates"]
    print("fused" if fused else "unfused", "logLt", pf.logLt, "steps", len(pf.summaries.ESSs),
          "acc first/last", [float(torch.as_tensor(a if not isinstance(a, list) else torch.stack([t.reshape(()) for t in a])).mean()) for a in (ars[0], ars[-1])],
          "exponents", [round(e, 4) for e in pf.X.shared["exponents"][:6]])
