#!/bin/bash
# profiles/ncu_capture_rs.sh <tag> -- lean capture (one B200, under gpurun): launch list of a short bench run
# and ONE `--set full` profile of the resampling-branch step kernel (ESSrmin = 1 makes every step resample).
set -u
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu --warmup 3"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv $B --steps 120 > $OUT/ncu_launches_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_move -s 12 -c 1 \
    -o $OUT/movers_$TAG $B --steps 30 --essrmin 1.0 > $OUT/ncu_movers_$TAG.log 2>&1
for p in raw details source; do
  ncu -i $OUT/movers_$TAG.ncu-rep --page $p --csv > $OUT/movers_${TAG}_$p.csv 2>/dev/null
done
rm -f $OUT/movers_$TAG.ncu-rep
du -sh $OUT
