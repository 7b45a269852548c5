#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: kernel-variant library for A/B timing.  Recompiles only smcb_filter.cu
# (config-2 instantiation, -DSMCB_BENCH_ONLY) with the extra flags and links it with the objects of the last
# full build -> particles_b200/variants/libsmcb_NAME.so; select it with SMCB_LIB=... python bench.py
# The directory is git-ignored but travels with the gpurun snapshot (~28 MB per library): delete it after the run.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=particles_b200/csrc
mkdir -p particles_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC \
     -DSMCB_BENCH_ONLY "$@" -c $C/smcb_filter.cu -o /tmp/smcb_filter_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o particles_b200/variants/libsmcb_$name.so \
     /tmp/smcb_filter_$name.o $C/smcb_api.o $C/smcb_filter_1d.o $C/smcb_filter_nd.o $C/smcb_sampler.o $C/smcb_peaks.o \
     -lcudart_static -lpthread -ldl -lrt
echo particles_b200/variants/libsmcb_$name.so
