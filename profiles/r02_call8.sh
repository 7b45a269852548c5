#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02k_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r02k_smoke.log
timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02k_bench.json 2> $OUT/r02k_bench.err
timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02k_bench_rs.json 2> $OUT/r02k_bench_rs.err
timeout 200 python bench.py --no-cpu --config c3i > $OUT/r02k_bench_c3i.json 2> $OUT/r02k_bench_c3i.err
timeout 200 python bench.py --no-cpu --config c3ii > $OUT/r02k_bench_c3ii.json 2> $OUT/r02k_bench_c3ii.err
TRACE_ESSRMIN=1.0 SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 200 python profiles/dump_trace.py $OUT/r02k_trace_rs.json > $OUT/r02k_trace_rs.log 2>&1
for f in "" _rs _c3i _c3ii; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02k_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "value", d["value"], "ms/step", round(d["ms_per_step"],4), "step_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "rs_frac", r["resampling_steps"]["frac"], "nrs", d["config"]["resampling_steps"], "logLt", d["config"]["logLt"], "e2e", d["e2e"]["value"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02k_bench$f.err").read()[-1200:])
P
done
tail -1 $OUT/r02k_trace_rs.log | cut -c1-900
timeout 800 python -m pytest tests -m gpu -q --timeout 200 > $OUT/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02k_pytest.log
tail -12 $OUT/r02k_pytest.log | cut -c1-220
