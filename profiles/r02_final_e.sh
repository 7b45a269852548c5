#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/r02aa_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02aa_pytest.log; tail -4 $OUT/r02aa_pytest.log | cut -c1-220
timeout 300 python profiles/bench_standalone.py $OUT/r02aa_standalone.json > $OUT/r02aa_standalone.log 2>&1; grep -E "normalise|exp_and|cumsum|resample|wmean|plugin" $OUT/r02aa_standalone.log | cut -c1-170
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
