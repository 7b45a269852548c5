#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r02g_smoke.log
timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02g_bench.json 2> $OUT/r02g_bench.err
timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02g_bench_rs.json 2> $OUT/r02g_bench_rs.err
for v in bs768 l2p bs384; do
  SMCB_LIB=$PWD/particles_b200/variants/libsmcb_$v.so timeout 200 python bench.py --no-cpu --steps 400 > $OUT/r02g_bench_$v.json 2> $OUT/r02g_bench_$v.err
done
SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 200 python profiles/dump_trace.py $OUT/r02g_trace.json > $OUT/r02g_trace.log 2>&1
TRACE_ESSRMIN=1.0 SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 200 python profiles/dump_trace.py $OUT/r02g_trace_rs.json > $OUT/r02g_trace_rs.log 2>&1
for f in "" _rs _bs768 _l2p _bs384; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02g_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "ms/step", round(d["ms_per_step"],4), "step_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "logLt", d["config"]["logLt"], "e2e", d["e2e"]["value"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02g_bench$f.err").read()[-800:])
P
done
tail -2 $OUT/r02g_trace.log; tail -2 $OUT/r02g_trace_rs.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 > $OUT/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02g_pytest.log
tail -8 $OUT/r02g_pytest.log | cut -c1-200
