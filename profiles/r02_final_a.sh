#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r02v_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02v_pytest.log; tail -4 $OUT/r02v_pytest.log | cut -c1-220
timeout 300 python profiles/bench_standalone.py $OUT/r02v_standalone.json > $OUT/r02v_standalone.log 2>&1; tail -30 $OUT/r02v_standalone.log | cut -c1-200
timeout 400 python bench.py > $OUT/r02v_bench.json 2> $OUT/r02v_bench.err
timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02v_bench_rs.json 2> $OUT/r02v_bench_rs.err
timeout 200 python bench.py --no-cpu --config c3i > $OUT/r02v_bench_c3i.json 2> $OUT/r02v_bench_c3i.err
timeout 200 python bench.py --no-cpu --config c3ii > $OUT/r02v_bench_c3ii.json 2> $OUT/r02v_bench_c3ii.err
for f in "" _rs _c3i _c3ii; do python - <<P
import json
try:
    d=json.loads(open("$OUT/r02v_bench$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench$f", "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "ms/step", round(d["ms_per_step"],4), "st_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "nrs", d["config"]["resampling_steps"], "cpu", json.dumps(d.get("cpu_baseline"))[:300])
except Exception as e: print("bench$f", "failed", e); print(open("$OUT/r02v_bench$f.err").read()[-1200:])
P
done
bash profiles/ncu_capture.sh r02v > $OUT/r02v_ncu.log 2>&1; tail -2 $OUT/r02v_ncu.log
