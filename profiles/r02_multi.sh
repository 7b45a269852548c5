#!/bin/bash
# 2-GPU call: sharded worker (island / NCCL == P2P / exact global incl. APF), then the bench lines
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q --timeout 850 > $OUT/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02m_pytest.log; tail -25 $OUT/r02m_pytest.log | cut -c1-250
for mode in island global; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 5 --resampling-mode $mode > $OUT/r02m_bench_n2_$mode.json 2> $OUT/r02m_bench_n2_$mode.err
python - <<P
import json
try:
    d=json.loads(open("$OUT/r02m_bench_n2_$mode.json").read().strip().splitlines()[-1])
    print("$mode", "ms/step", round(d["ms_per_step"],4), "value", d["value"], "e2e", d["e2e"]["value"], "e2e_s", d["e2e"]["seconds_all_runs"], "parity", json.dumps(d.get("parity"))[:600])
except Exception as e: print("$mode failed", e); print(open("$OUT/r02m_bench_n2_$mode.err").read()[-2500:])
P
done
timeout 200 python bench.py --no-cpu --steps 500 > $OUT/r02m_bench_n1.json 2> $OUT/r02m_bench_n1.err; python -c "
import json; d=json.loads(open('$OUT/r02m_bench_n1.json').read().strip().splitlines()[-1]); print('n1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['seconds_all_runs'])"
