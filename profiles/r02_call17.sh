#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
T=$PWD/particles_b200/variants/libsmcb_trace.so
tr() { name=$1; shift; env "$@" SMCB_LIB=$T timeout 200 python profiles/dump_trace.py $OUT/r02q_trace_$name.json > $OUT/r02q_trace_$name.log 2>&1; echo "trace $name: $(tail -1 $OUT/r02q_trace_$name.log | cut -c1-2500)"; }
tr c3ii_k12 TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=0.5 TRACE_K=12
tr c3i_k9 TRACE_CONFIG=c3i TRACE_N=1000000 TRACE_ESSRMIN=0.5 TRACE_K=9
tr c3ii_rs1 TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=1.0
timeout 200 python bench.py --no-cpu --config c3i > $OUT/r02q_bench_c3i.json 2> $OUT/r02q_bench_c3i.err
timeout 200 python bench.py --no-cpu --config c3ii > $OUT/r02q_bench_c3ii.json 2> $OUT/r02q_bench_c3ii.err
timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02q_bench_rs.json 2> $OUT/r02q_bench_rs.err
for f in _c3i _c3ii _rs; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02q_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "value %.3e"%d["value"], "ms/step", round(d["ms_per_step"],4), "st_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "nrs", d["config"]["resampling_steps"], "logLt", d["config"]["logLt"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02q_bench$f.err").read()[-1200:])
P
done
timeout 400 python -m pytest tests/test_gpu_filter.py -m gpu -q -x --timeout 200 > $OUT/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02q_pytest.log
tail -4 $OUT/r02q_pytest.log | cut -c1-220
