#!/bin/bash
# round-2 second GPU call: the one-kernel-per-step design: tests, bench, timeline, throttle A/B
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02b_pytest.log
tail -5 $OUT/r02b_pytest.log
timeout 300 python bench.py --no-cpu --steps 400 > $OUT/r02b_bench.json 2> $OUT/r02b_bench.err
for v in nothrottle thr2; do
  SMCB_LIB=$PWD/particles_b200/variants/libsmcb_$v.so timeout 300 python bench.py --no-cpu --steps 400 > $OUT/r02b_bench_$v.json 2> $OUT/r02b_bench_$v.err
done
SMCB_PDL=0 timeout 300 python bench.py --no-cpu --steps 400 > $OUT/r02b_bench_nopdl.json 2> $OUT/r02b_bench_nopdl.err
SMCB_COOP=1 timeout 300 python bench.py --no-cpu --steps 400 > $OUT/r02b_bench_coop.json 2> $OUT/r02b_bench_coop.err
SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 300 python profiles/dump_trace.py $OUT/r02b_trace.json > $OUT/r02b_trace.log 2>&1
for f in "" _nothrottle _thr2 _nopdl _coop; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02b_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "ms/step", round(d["ms_per_step"],4), "step_us", round(r["avg_launch_us"],1), "frac", round(r["frac"],3), "rs_us", r["resampling_steps"]["avg_launch_us"], "logLt", d["config"]["logLt"], "nrs", d["config"]["resampling_steps"], "e2e", d["e2e"]["value"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02b_bench$f.err").read()[-1500:])
P
done
tail -3 $OUT/r02b_trace.log
