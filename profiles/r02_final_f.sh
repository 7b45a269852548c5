#!/bin/bash
# last call of the round: the committed build end to end
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/r02ab_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02ab_pytest.log; tail -3 $OUT/r02ab_pytest.log | cut -c1-220
timeout 300 python profiles/bench_standalone.py $OUT/r02ab_standalone.json > $OUT/r02ab_standalone.log 2>&1; grep -E "normalise|exp_and|cumsum|resample|wmean|plugin" $OUT/r02ab_standalone.log | cut -c1-170
timeout 400 python bench.py > $OUT/r02ab_bench.json 2> $OUT/r02ab_bench.err
python - <<P
import json
try:
    d=json.loads(open("$OUT/r02ab_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench", "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "ms/step", round(d["ms_per_step"],4), "st_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "nrs", d["config"]["resampling_steps"], "parity", json.dumps(d.get("parity"))[:160])
except Exception as e: print("bench failed", e); print(open("$OUT/r02ab_bench.err").read()[-1200:])
P
