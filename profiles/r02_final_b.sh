#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/r02w_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02w_pytest.log; tail -4 $OUT/r02w_pytest.log | cut -c1-220
timeout 300 python profiles/bench_standalone.py $OUT/r02w_standalone.json > $OUT/r02w_standalone.log 2>&1; tail -20 $OUT/r02w_standalone.log | cut -c1-200
SMCB_SCAN_LOOKBACK=1 timeout 300 python profiles/bench_standalone.py $OUT/r02w_standalone_lookback.json 2>&1 | grep -E "cumsum|resample" | cut -c1-200
bash profiles/r02_sanitize.sh
