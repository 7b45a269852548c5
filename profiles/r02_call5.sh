#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r02f_smoke.log
timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02f_bench.json 2> $OUT/r02f_bench.err
for v in rec208 bs512 bs512p sched1; do
  SMCB_LIB=$PWD/particles_b200/variants/libsmcb_$v.so timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02f_bench_$v.json 2> $OUT/r02f_bench_$v.err
done
SMCB_SLAB_IT=4 SMCB_LIB=$PWD/particles_b200/variants/libsmcb_rec208.so timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02f_bench_rec208s4.json 2> $OUT/r02f_bench_rec208s4.err
timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02f_bench_rs.json 2> $OUT/r02f_bench_rs.err
SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 200 python profiles/dump_trace.py $OUT/r02f_trace.json > $OUT/r02f_trace.log 2>&1
for f in "" _rec208 _rec208s4 _bs512 _bs512p _sched1 _rs; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02f_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "ms/step", round(d["ms_per_step"],4), "step_us", round(r["avg_launch_us"],1), "frac", round(r["frac"],3), "rs_us", round(r["resampling_steps"]["avg_launch_us"],1), "logLt", d["config"]["logLt"], "e2e", d["e2e"]["value"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02f_bench$f.err").read()[-800:])
P
done
tail -3 $OUT/r02f_trace.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 > $OUT/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02f_pytest.log
tail -15 $OUT/r02f_pytest.log | cut -c1-200
