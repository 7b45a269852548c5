#!/bin/bash
# round-2 first GPU call: ceilings, baseline, table-math variants, per-CTA timeline
set -u
OUT=gpurun_out; mkdir -p $OUT
python profiles/measure_peaks.py $OUT/r02_peaks.json > $OUT/r02_peaks.log 2>&1
timeout 300 python bench.py --no-cpu --steps 400 > $OUT/r02a_bench_base.json 2> $OUT/r02a_bench_base.err
for v in table1 table2; do
  SMCB_LIB=$PWD/particles_b200/variants/libsmcb_$v.so timeout 300 python bench.py --no-cpu --steps 400 > $OUT/r02a_bench_$v.json 2> $OUT/r02a_bench_$v.err
done
SMCB_LIB=$PWD/particles_b200/variants/libsmcb_trace.so timeout 300 python profiles/dump_trace.py $OUT/r02a_trace.json > $OUT/r02a_trace.log 2>&1
for f in base table1 table2; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02a_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "ms/step", round(d["ms_per_step"],4), "move_us", round(d["roofline"]["avg_launch_us"],1), "frac", round(d["roofline"]["frac"],3), "logLt", d["config"]["logLt"], "e2e", d["e2e"]["value"])
except Exception as e: print("$f", "failed", e)
P
done
cat $OUT/r02_peaks.log | tail -15
tail -3 $OUT/r02a_trace.log
