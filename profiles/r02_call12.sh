#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 120 --csv --log-file $OUT/launches_r02l_c3ii.csv python bench.py --no-cpu --config c3ii --steps 100 > $OUT/ncu_r02l_c3ii.log 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open("$OUT/launches_r02l_c3ii.csv")) if len(r)>10]
hdr=rows[0]; iN=hdr.index("Kernel Name"); iV=hdr.index("Metric Value")
d=[(r[iN][:60], float(r[iV].replace(",",""))) for r in rows[1:] if r[iV].replace(",","").replace(".","").isdigit()]
import collections
for name,v in d[:80]: print(name, v)
P
timeout 300 python profiles/bench_standalone.py $OUT/r02l_standalone.json > $OUT/r02l_standalone.log 2>&1; tail -22 $OUT/r02l_standalone.log | cut -c1-230
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02l_pytest.log
tail -8 $OUT/r02l_pytest.log | cut -c1-220
