#!/bin/bash
# profiles/ncu_capture.sh <tag> -- run under gpurun on ONE B200.  Captures the launch list of a short bench run and
# one `--set full` profile per kind of step kernel launch, then exports the pages we read (raw metrics + per-line
# source counters) as CSV so that only small files travel back.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv $B --steps 120 > $OUT/ncu_launches_$TAG.log 2>&1
# non-resampling step (config 2 mostly runs this one): launch 12 of a run without resampling so far
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 1 \
    -o $OUT/move_$TAG $B --steps 30 > $OUT/ncu_move_$TAG.log 2>&1
# resampling step: weight scan + grid barrier + search/gather/move in ONE launch (ESSrmin = 1: every step resamples)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 1 \
    -o $OUT/movers_$TAG $B --steps 30 --essrmin 1.0 > $OUT/ncu_movers_$TAG.log 2>&1
for k in move movers; do
  ncu -i $OUT/${k}_$TAG.ncu-rep --page raw --csv > $OUT/${k}_${TAG}_raw.csv 2>/dev/null
  ncu -i $OUT/${k}_$TAG.ncu-rep --page details --csv > $OUT/${k}_${TAG}_details.csv 2>/dev/null
  ncu -i $OUT/${k}_$TAG.ncu-rep --page source --csv > $OUT/${k}_${TAG}_source.csv 2>/dev/null
done
rm -f $OUT/movers_$TAG.ncu-rep
find $OUT -name "*.ncu-rep" -size +30M -delete
du -sh $OUT
