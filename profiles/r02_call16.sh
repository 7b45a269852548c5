#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
T=$PWD/particles_b200/variants/libsmcb_trace.so
tr() { name=$1; shift; env "$@" SMCB_LIB=$T timeout 200 python profiles/dump_trace.py $OUT/r02p_trace_$name.json > $OUT/r02p_trace_$name.log 2>&1; echo "trace $name: $(tail -1 $OUT/r02p_trace_$name.log | cut -c1-2500)"; }
tr c3ii_k12 TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=0.5 TRACE_K=12
tr c3i_k9 TRACE_CONFIG=c3i TRACE_N=1000000 TRACE_ESSRMIN=0.5 TRACE_K=9
tr c3ii_rs1 TRACE_CONFIG=c3ii TRACE_N=1000000 TRACE_ESSRMIN=1.0
