#!/bin/bash
# profiles/ncu_capture.sh <tag> -- run under gpurun on ONE B200.  Captures the launch list of a
# short bench run and one `--set full` profile per hot kernel, then exports the pages we read
# (raw metrics + per-line source counters) as CSV so that only small files travel back.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-cpu --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv $B --steps 120 > $OUT/ncu_launches_$TAG.log 2>&1
# non-resampling step kernel (config 2 mostly runs this one)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_move -s 12 -c 1 \
    -o $OUT/move_$TAG $B --steps 30 > $OUT/ncu_move_$TAG.log 2>&1
# resampling step: weight scan + search/gather/move (ESSrmin = 1 makes every step resample)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_scan_w -s 12 -c 1 \
    -o $OUT/scan_$TAG $B --steps 30 --essrmin 1.0 > $OUT/ncu_scan_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_move -s 12 -c 1 \
    -o $OUT/movers_$TAG $B --steps 30 --essrmin 1.0 > $OUT/ncu_movers_$TAG.log 2>&1
for k in move scan movers; do
  ncu -i $OUT/${k}_$TAG.ncu-rep --page raw --csv > $OUT/${k}_${TAG}_raw.csv 2>/dev/null
  ncu -i $OUT/${k}_$TAG.ncu-rep --page details --csv > $OUT/${k}_${TAG}_details.csv 2>/dev/null
  ncu -i $OUT/${k}_$TAG.ncu-rep --page source --csv > $OUT/${k}_${TAG}_source.csv 2>/dev/null
done
# a .ncu-rep with imported sources is ~32 MB and gpurun brings back at most 64 MiB:
# keep the CSV exports of all three and the binary report of the dominant kernel only
rm -f $OUT/scan_$TAG.ncu-rep $OUT/movers_$TAG.ncu-rep
# ... and even that one only while it is small enough to travel (the embedded cubin grows with the library)
find $OUT -name "*.ncu-rep" -size +30M -delete
ls -la $OUT
du -sh $OUT
