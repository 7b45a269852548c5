#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q --timeout 850 > $OUT/r02t_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02t_pytest.log; tail -5 $OUT/r02t_pytest.log | cut -c1-250
