#!/bin/bash
# 2-GPU call: sharded worker, bench island (speculative / not) and global, single-GPU reference line, scan-group variant
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q --timeout 850 > $OUT/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02s_pytest.log; tail -5 $OUT/r02s_pytest.log | cut -c1-250
run2() { name=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 5 $EXTRA > $OUT/r02s_bench_n2_$name.json 2> $OUT/r02s_bench_n2_$name.err
python - <<P
import json
try:
    d=json.loads(open("$OUT/r02s_bench_n2_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(d["ms_per_step"],4), "value %.4e"%d["value"], "e2e %.4e"%d["e2e"]["value"], "nrs", d["config"]["resampling_steps"], "parity", json.dumps(d.get("parity"))[:260])
except Exception as e: print("$name failed", e); print(open("$OUT/r02s_bench_n2_$name.err").read()[-2500:])
P
}
EXTRA="--resampling-mode island" run2 island A=1
EXTRA="--resampling-mode island" run2 island_nospec SMCB_LIB=$PWD/particles_b200/variants/libsmcb_nospec.so
EXTRA="--resampling-mode global" run2 global A=1
EXTRA="--resampling-mode global" run2 global_nospec SMCB_LIB=$PWD/particles_b200/variants/libsmcb_nospec.so
for v in "" _g1; do
  L=""; [ -n "$v" ] && L="SMCB_LIB=$PWD/particles_b200/variants/libsmcb${v}.so"
  env $L timeout 200 python bench.py --no-cpu --steps 500 > $OUT/r02s_bench_n1$v.json 2> $OUT/r02s_bench_n1$v.err
  env $L timeout 200 python bench.py --no-cpu --steps 300 --essrmin 1.0 > $OUT/r02s_bench_rs$v.json 2> $OUT/r02s_bench_rs$v.err
  python -c "
import json
d=json.loads(open('$OUT/r02s_bench_n1$v.json').read().strip().splitlines()[-1]); r=json.loads(open('$OUT/r02s_bench_rs$v.json').read().strip().splitlines()[-1])
print('n1$v ms/step', round(d['ms_per_step'],4), 'st_us', round(d['roofline']['avg_launch_us'],2), 'rs_us', round(d['roofline']['resampling_steps']['avg_launch_us'],1), 'e2e %.4e'%d['e2e']['value'], '| essrmin=1: ms/step', round(r['ms_per_step'],4), 'value %.3e'%r['value'])"
done
