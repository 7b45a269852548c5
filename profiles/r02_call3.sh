#!/bin/bash
# round-2 third GPU call: table math + scheduling variants of the streaming branch, then the GPU tests
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02c_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/r02c_smoke.log
timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02c_bench.json 2> $OUT/r02c_bench.err
SMCB_SLAB_IT=4 timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02c_bench_slab4.json 2> $OUT/r02c_bench_slab4.err
for v in sched0 sched0t2 sched1; do
  SMCB_LIB=$PWD/particles_b200/variants/libsmcb_$v.so timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02c_bench_$v.json 2> $OUT/r02c_bench_$v.err
done
SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 200 python profiles/dump_trace.py $OUT/r02c_trace.json > $OUT/r02c_trace.log 2>&1
for f in "" _slab4 _sched0 _sched0t2 _sched1; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02c_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "ms/step", round(d["ms_per_step"],4), "step_us", round(r["avg_launch_us"],1), "frac", round(r["frac"],3), "rs_us", round(r["resampling_steps"]["avg_launch_us"],1), "logLt", d["config"]["logLt"], "nrs", d["config"]["resampling_steps"], "e2e", d["e2e"]["value"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02c_bench$f.err").read()[-1500:])
P
done
tail -3 $OUT/r02c_trace.log
timeout 1000 python -m pytest tests -m gpu -x -q --timeout 150 > $OUT/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02c_pytest.log
tail -12 $OUT/r02c_pytest.log
