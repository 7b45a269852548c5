#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q --timeout 850 > $OUT/r02ad_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02ad_pytest.log; tail -3 $OUT/r02ad_pytest.log | cut -c1-250
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 5 --resampling-mode island > $OUT/r02ad_bench_n2_island.json 2> $OUT/r02ad_bench_n2_island.err
timeout 200 python bench.py --no-cpu --steps 500 > $OUT/r02ad_bench_n1.json 2> $OUT/r02ad_bench_n1.err
python - <<P
import json
for f in ("n2_island","n1"):
    try:
        d=json.loads(open("$OUT/r02ad_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"],4), "value %.4e"%d["value"], "e2e %.4e"%d["e2e"]["value"], "nrs", d["config"]["resampling_steps"], "logLt", d["config"]["logLt"])
    except Exception as e: print(f, "failed", e); print(open("$OUT/r02ad_bench_%s.err"%f).read()[-1500:])
P
