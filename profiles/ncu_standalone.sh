#!/bin/bash
# ncu --set full of the stand-alone weight / prefix-sum / search kernels (one B200); CSV pages only travel back
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none -k regex:"k_lse|k_exp_normalise|k_scan_sums|k_scan_chunks|k_search|k_wmoments" -s 9 -c 9 \
    -o $OUT/standalone_r02 python profiles/ncu_standalone.py > $OUT/ncu_standalone.log 2>&1
tail -2 $OUT/ncu_standalone.log
ncu -i $OUT/standalone_r02.ncu-rep --page raw --csv > $OUT/standalone_r02_raw.csv 2>/dev/null
rm -f $OUT/standalone_r02.ncu-rep
wc -l $OUT/standalone_r02_raw.csv
