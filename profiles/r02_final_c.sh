#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python profiles/bench_standalone.py $OUT/r02y_standalone.json > $OUT/r02y_standalone.log 2>&1; grep -E "normalise|exp_and|cumsum|systematic|wmean" $OUT/r02y_standalone.log | cut -c1-170
for v in "" _pipe _bs640; do
  L=""; [ -n "$v" ] && L="SMCB_LIB=$PWD/particles_b200/variants/libsmcb${v}.so"
  env $L timeout 200 python bench.py --no-cpu --steps 500 > $OUT/r02y_bench_n1$v.json 2> $OUT/r02y_bench_n1$v.err
  env $L timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02y_bench_rs$v.json 2> $OUT/r02y_bench_rs$v.err
  python -c "
import json
try:
    d=json.loads(open('$OUT/r02y_bench_n1$v.json').read().strip().splitlines()[-1]); r=json.loads(open('$OUT/r02y_bench_rs$v.json').read().strip().splitlines()[-1])
    print('n1$v ms/step', round(d['ms_per_step'],4), 'st_us', round(d['roofline']['avg_launch_us'],2), 'rs_us', round(d['roofline']['resampling_steps']['avg_launch_us'],1), 'e2e %.4e'%d['e2e']['value'], '| essrmin=1: ms/step', round(r['ms_per_step'],4), 'value %.3e'%r['value'], 'logLt', d['config']['logLt'])
except Exception as e: print('n1$v failed', e); print(open('$OUT/r02y_bench_n1$v.err').read()[-800:])"
done
bash profiles/r02_sanitize.sh 2>&1 | grep -E "mix|Bootstrap|Auxiliary|guided"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02y_bench_reference.json 2> $OUT/r02y_bench_reference.err; tail -1 $OUT/r02y_bench_reference.json | cut -c1-900
