"""torchrun worker: BASELINE config 5 over the GPUs of one box -- ShardedAdaptiveTempering of the 20-D logistic
regression posterior (n_data = 1000), 1e4 chains x 100 in total (1e4 / world chains per rank).  Rank 0 prints one
JSON line: tempering steps, seconds (median of 3 runs after a warm-up), likelihood evaluations per second."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from particles_b200 import smc_samplers as ssp
    from particles_b200.sharded_samplers import ShardedAdaptiveTempering
    from oracle import samplers_numpy as osp
    data = osp.synthetic_logistic(1000, 20, seed=0)
    runs = []
    for rep in range(4):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        sm = ShardedAdaptiveTempering(model=ssp.LogisticRegression(data=data, prior_scale=5.0), M_local=10_000 // world,
                                      len_chain=100, ESSrmin=0.5, seed=40 + rep).run()
        torch.cuda.synchronize(); dist.barrier()
        runs.append((time.perf_counter() - t0, len(sm.exponents) - 1, sm.logLt))
    runs = runs[1:]
    if rank == 0:
        dt = float(np.median([r[0] for r in runs]))
        print(json.dumps({"world": world, "seconds_median_of_3": dt, "seconds_all": [r[0] for r in runs],
                          "tempering_steps": [r[1] for r in runs], "logLt": [r[2] for r in runs],
                          "likelihood_evals_per_s": 1e6 * runs[0][1] / dt}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
