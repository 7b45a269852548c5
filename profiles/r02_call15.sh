#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
T=$PWD/particles_b200/variants/libsmcb_trace.so
tr() { name=$1; shift; env "$@" SMCB_LIB=$T timeout 200 python profiles/dump_trace.py $OUT/r02o_trace_$name.json > $OUT/r02o_trace_$name.log 2>&1; echo "trace $name: $(tail -1 $OUT/r02o_trace_$name.log | cut -c1-1200)"; }
tr c3ii_rs1 TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=1.0
tr c3i_rs1 TRACE_CONFIG=c3i TRACE_N=1000000 TRACE_ESSRMIN=1.0
for k in 12 17 23 31 38 44 52 60; do
tr c3ii_k$k TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=0.5 TRACE_K=$k
done
for k in 5 9 14 20 26; do
tr c3i_k$k TRACE_CONFIG=c3i TRACE_N=1000000 TRACE_ESSRMIN=0.5 TRACE_K=$k
done
