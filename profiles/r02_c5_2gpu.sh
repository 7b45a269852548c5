#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for n in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 profiles/bench_c5_sharded.py > $OUT/r02_c5_sharded_n$n.log 2>&1
grep '"world"' $OUT/r02_c5_sharded_n$n.log | tail -1 | cut -c1-600 || tail -5 $OUT/r02_c5_sharded_n$n.log
done
