#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 5 > $OUT/r02ae_bench_n2.json 2> $OUT/r02ae_bench_n2.err
timeout 300 python bench.py --steps 1000 --warmup 10 > $OUT/r02ae_bench_n1.json 2> $OUT/r02ae_bench_n1.err
python - <<P
import json
for f in ("n2","n1"):
    try:
        d=json.loads(open("$OUT/r02ae_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "value %.4e"%d["value"], "e2e %.4e"%d["e2e"]["value"], d["e2e"]["seconds_all_runs"], "nrs", d["config"]["resampling_steps"])
    except Exception as e: print(f, "failed", e); print(open("$OUT/r02ae_bench_%s.err"%f).read()[-1500:])
P
