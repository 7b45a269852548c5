#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: kernel-variant library for A/B timing.  Recompiles only smcb_filter.cu
# (config-2 instantiation, -DSMCB_BENCH_ONLY) with the extra flags and links it with stub builds of the other
# model translation units and the objects of the last full build -> particles_b200/variants/libsmcb_NAME.so
# (~10 MB); select it with SMCB_LIB=... python bench.py.  The directory is git-ignored but travels with the gpurun
# snapshot: delete it after the run.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=particles_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -DSMCB_BENCH_ONLY"
mkdir -p particles_b200/variants
nvcc $F "$@" -c $C/smcb_filter.cu -o /tmp/smcb_filter_$name.o
nd=/tmp/stub_smcb_filter_nd.o
if [ -n "$VARIANT_ND" ]; then      # VARIANT_ND=1: also compile the two config-3 kernels with the extra flags
  nd=/tmp/smcb_filter_nd_$name.o
  nvcc $F -DSMCB_BENCH_ND "$@" -c $C/smcb_filter_nd.cu -o $nd
fi
for u in smcb_filter_1d smcb_filter_nd; do
  [ /tmp/stub_$u.o -nt $C/smcb_step.cuh ] || nvcc $F -c $C/$u.cu -o /tmp/stub_$u.o
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o particles_b200/variants/libsmcb_$name.so \
     /tmp/smcb_filter_$name.o $C/smcb_api.o /tmp/stub_smcb_filter_1d.o $nd $C/smcb_sampler.o $C/smcb_peaks.o \
     -lcudart_static -lpthread -ldl -lrt
echo particles_b200/variants/libsmcb_$name.so
