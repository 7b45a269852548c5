#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python bench.py --no-cpu --config c3i > $OUT/r02r_bench_c3i.json 2> $OUT/r02r_bench_c3i.err
timeout 200 python bench.py --no-cpu --config c3ii > $OUT/r02r_bench_c3ii.json 2> $OUT/r02r_bench_c3ii.err
for f in _c3i _c3ii; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02r_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "value %.3e"%d["value"], "ms/step", round(d["ms_per_step"],4), "st_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "nrs", d["config"]["resampling_steps"], "logLt", d["config"]["logLt"])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02r_bench$f.err").read()[-1200:])
P
done
B="python bench.py --no-cpu --warmup 3 --config c3ii"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 1 \
    -o $OUT/c3ii_move_r02r $B --steps 30 --essrmin 0.0 > $OUT/ncu_c3ii_move.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 1 \
    -o $OUT/c3ii_movers_r02r $B --steps 30 --essrmin 1.0 > $OUT/ncu_c3ii_movers.log 2>&1
for k in c3ii_move c3ii_movers; do
  ncu -i $OUT/${k}_r02r.ncu-rep --page raw --csv > $OUT/${k}_r02r_raw.csv 2>/dev/null
  ncu -i $OUT/${k}_r02r.ncu-rep --page details --csv > $OUT/${k}_r02r_details.csv 2>/dev/null
  ncu -i $OUT/${k}_r02r.ncu-rep --page source --csv > $OUT/${k}_r02r_source.csv 2>/dev/null
done
rm -f $OUT/*.ncu-rep
tail -3 $OUT/ncu_c3ii_move.log
timeout 300 python -m pytest tests/test_gpu_filter.py -m gpu -q -x --timeout 200 -k "mv or MV or guided or kalman or nd or bearings" > $OUT/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02r_pytest.log
tail -3 $OUT/r02r_pytest.log | cut -c1-220
