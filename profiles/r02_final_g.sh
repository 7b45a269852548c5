#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_filter.py -m gpu -q -x --timeout 300 > $OUT/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02ac_pytest.log; tail -3 $OUT/r02ac_pytest.log | cut -c1-220
timeout 300 python profiles/bench_standalone.py $OUT/r02ac_standalone.json > $OUT/r02ac_standalone.log 2>&1; grep -E "normalise|exp_and|cumsum|resample|wmean|plugin" $OUT/r02ac_standalone.log | cut -c1-170
timeout 300 python profiles/bench_c5.py $OUT/r02ac_c5.json > $OUT/r02ac_c5.log 2>&1; tail -3 $OUT/r02ac_c5.log | cut -c1-700
