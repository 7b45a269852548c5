"""Multi-GPU check of the sharded filter (needs >= 2 GPUs; skipped on a 1-GPU box).
Launches `torchrun --nproc-per-node 2 tests/sharded_worker.py`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharded_filter():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tests", "sharded_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
