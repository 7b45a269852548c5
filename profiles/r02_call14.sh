#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
T=$PWD/particles_b200/variants/libsmcb_trace.so
tr() { name=$1; shift; env "$@" SMCB_LIB=$T timeout 200 python profiles/dump_trace.py $OUT/r02n_trace_$name.json > $OUT/r02n_trace_$name.log 2>&1; echo "trace $name: $(tail -1 $OUT/r02n_trace_$name.log | cut -c1-900)"; }
tr c3ii_rs TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=1.0
tr c3i_rs TRACE_CONFIG=c3i TRACE_N=1000000 TRACE_ESSRMIN=1.0
tr c2_1e6_sys TRACE_N=1000000 TRACE_ESSRMIN=1.0
tr c2_1e6_strat TRACE_N=1000000 TRACE_ESSRMIN=1.0 TRACE_SCHEME=stratified
tr c2_1e7_strat TRACE_ESSRMIN=1.0 TRACE_SCHEME=stratified
tr c3ii_stream TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=0.0
timeout 200 python bench.py --no-cpu --config c3i > $OUT/r02n_bench_c3i.json 2> $OUT/r02n_bench_c3i.err
timeout 200 python bench.py --no-cpu --config c3ii > $OUT/r02n_bench_c3ii.json 2> $OUT/r02n_bench_c3ii.err
for f in _c3i _c3ii; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02n_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "value %.3e"%d["value"], "ms/step", round(d["ms_per_step"],4), "st_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "nrs", d["config"]["resampling_steps"], "logLt", d["config"]["logLt"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02n_bench$f.err").read()[-1200:])
P
done
timeout 300 python -m pytest tests/test_gpu_filter.py -m gpu -q -x --timeout 200 -k "mv or MV or guided or kalman or nd or bearings" > $OUT/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02n_pytest.log
tail -4 $OUT/r02n_pytest.log | cut -c1-220
