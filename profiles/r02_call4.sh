#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02d_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/r02d_smoke.log; tail -2 $OUT/r02d_smoke.log
timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02d_bench.json 2> $OUT/r02d_bench.err
timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02d_bench_rs.json 2> $OUT/r02d_bench_rs.err
for v in sched1 sched0t2; do
  SMCB_LIB=$PWD/particles_b200/variants/libsmcb_$v.so timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02d_bench_$v.json 2> $OUT/r02d_bench_$v.err
done
SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 200 python profiles/dump_trace.py $OUT/r02d_trace.json > $OUT/r02d_trace.log 2>&1
for f in "" _rs _sched1 _sched0t2; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02d_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "ms/step", round(d["ms_per_step"],4), "step_us", round(r["avg_launch_us"],1), "frac", round(r["frac"],3), "rs_us", r["resampling_steps"]["avg_launch_us"], "logLt", d["config"]["logLt"], "nrs", d["config"]["resampling_steps"], "e2e", d["e2e"]["value"], "parity", d.get("parity",{}).get("n_sigma"))
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02d_bench$f.err").read()[-1500:])
P
done
tail -3 $OUT/r02d_trace.log
timeout 900 python -m pytest tests -m gpu -q --timeout 200 > $OUT/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02d_pytest.log
tail -25 $OUT/r02d_pytest.log | cut -c1-220
bash profiles/ncu_capture.sh r02d > $OUT/r02d_ncu.log 2>&1; tail -2 $OUT/r02d_ncu.log
