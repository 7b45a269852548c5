#!/bin/bash
# 8-GPU call: bench lines (island, global) + a 4-GPU island line; every command under its own timeout
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
run() { n=$1; name=$2; shift; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --warmup 5 "$@" > $OUT/r02u_bench_n${n}_$name.json 2> $OUT/r02u_bench_n${n}_$name.err
python - <<P
import json
try:
    d=json.loads(open("$OUT/r02u_bench_n${n}_$name.json").read().strip().splitlines()[-1])
    print("n=$n $name", "ms/step", round(d["ms_per_step"],4), "value %.4e"%d["value"], "e2e %.4e"%d["e2e"]["value"], d["e2e"]["seconds_all_runs"], "nrs", d["config"]["resampling_steps"], "parity", json.dumps(d.get("parity"))[:700])
except Exception as e: print("n=$n $name failed", e); print(open("$OUT/r02u_bench_n${n}_$name.err").read()[-2500:])
P
}
run 8 island --steps 500 --resampling-mode island
run 8 global --steps 300 --resampling-mode global
run 4 island --steps 500 --resampling-mode island
timeout 200 python bench.py --no-cpu --steps 500 > $OUT/r02u_bench_n1.json 2> $OUT/r02u_bench_n1.err; python -c "
import json; d=json.loads(open('$OUT/r02u_bench_n1.json').read().strip().splitlines()[-1]); print('n1 ms/step', d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'], d['e2e']['seconds_all_runs'])"
