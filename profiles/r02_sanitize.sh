#!/bin/bash
# compute-sanitizer runs of the hot path (one B200): memcheck + racecheck + synccheck of smoke() (a fused filter with
# resampling steps, checked against the oracle) and of a short stratified / multinomial / APF mix.
set -u
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_mix.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import particles_b200 as pb
from particles_b200 import state_space_models as ssm, kalman
g = np.load("tests/golden/golden_stats.npz")
y = [np.atleast_1d(v) for v in g["data/sv_seed1_T1000"][:25]]
for fk, scheme in ((ssm.Bootstrap(ssm=ssm.StochVol(), data=y), "stratified"), (ssm.Bootstrap(ssm=ssm.StochVol(), data=y), "multinomial"),
                   (ssm.AuxiliaryPF(ssm=ssm.StochVol(), data=y), "systematic")):
    pf = pb.SMC(fk=fk, N=20001, resampling=scheme, ESSrmin=0.9, seed=3, collect=[pb.collectors.Moments()]); pf.run()
    print(type(fk).__name__, scheme, pf.logLt, sum(pf.summaries.rs_flags))
ym = [np.asarray(v) for v in g["data/mvlg_seed5_T30"][:12]]
pf = pb.SMC(fk=ssm.GuidedPF(ssm=kalman.MVLinearGauss_Guarniero_etal(alpha=0.4, dx=4), data=ym), N=30000, resampling="stratified", ESSrmin=0.9, seed=4); pf.run()
print("guided d=4", pf.logLt, sum(pf.summaries.rs_flags))
P
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02_sanitizer_${tool}_smoke.log 2>&1
  echo "$tool smoke: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' $OUT/r02_sanitizer_${tool}_smoke.log | tr '\n' ' ')"
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_mix.py > $OUT/r02_sanitizer_${tool}_mix.log 2>&1
  echo "$tool mix: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/r02_sanitizer_${tool}_mix.log | tr '\n' ' ')"; grep -E "Bootstrap|Auxiliary|guided" $OUT/r02_sanitizer_${tool}_mix.log | head -4
done
