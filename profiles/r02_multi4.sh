#!/bin/bash
# 2-GPU call: sharded worker (incl. the globally normalised ShardedSMC.W) + memcheck of a tiny 2-rank filter
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q --timeout 850 > $OUT/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02x_pytest.log; tail -4 $OUT/r02x_pytest.log | cut -c1-250
cat > /tmp/san2.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from particles_b200 import state_space_models as ssm
from particles_b200.parallel import ShardedFilter
g = np.load("tests/golden/golden_stats.npz")
fk = ssm.Bootstrap(ssm=ssm.StochVol(), data=[np.atleast_1d(v) for v in g["data/sv_seed1_T1000"][:25]])
for mode in ("island", "global"):
    f = ShardedFilter(ssm.fused_spec(fk), 20000, "systematic", 0.9, 3, rank, world, resampling_mode=mode)
    f.step(25); f.state()
    if rank == 0: print(mode, "logLt", float(f.summ[24, 1]), "resamplings", int(f.summ[:, 2].sum()), flush=True)
    f.close()
dist.barrier(); dist.destroy_process_group()
P
timeout 600 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 /tmp/san2.py > $OUT/r02_sanitizer_memcheck_2gpu.log 2>&1
echo "memcheck 2gpu rc=$?: $(grep -E 'ERROR SUMMARY|logLt' $OUT/r02_sanitizer_memcheck_2gpu.log | tr '\n' ' ' | cut -c1-600)"
