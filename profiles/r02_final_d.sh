#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r02z_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02z_pytest.log; tail -4 $OUT/r02z_pytest.log | cut -c1-220
timeout 400 python bench.py > $OUT/r02z_bench.json 2> $OUT/r02z_bench.err
python - <<P
import json
try:
    d=json.loads(open("$OUT/r02z_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bench", "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "ms/step", round(d["ms_per_step"],4), "st_us", r["avg_launch_us"], "frac", r["frac"], "rs_us", r["resampling_steps"]["avg_launch_us"], "nrs", d["config"]["resampling_steps"], "sec", json.dumps(r.get("secondary"))[:300], "parity", json.dumps(d.get("parity"))[:200], "clocks", d["clocks"])
except Exception as e: print("bench failed", e); print(open("$OUT/r02z_bench.err").read()[-1200:])
P
